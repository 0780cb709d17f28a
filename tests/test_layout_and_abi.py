"""CPU tests: libhvla.so builds, loads, exports every symbol include/hvla.h declares, and its
blob layouts agree with the Python packers.  No compute call is made (no GPU here)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hvla import _native as N
    return N.lib()


def test_library_exports_every_declared_symbol(lib):
    from hvla import _native as N
    hdr = open(os.path.join(ROOT, "include", "hvla.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hvla_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hvla_version() == 1


def test_sizes_agree_with_python(lib):
    from hvla import metadata as M, params as P
    assert lib.hvla_generated_elems() == M.N_GENERATED == 201500
    assert lib.hvla_generated_row_stride() == M.N_GENERATED_PADDED
    assert lib.hvla_hn_blob_elems() == P.hn_blob_size()
    assert lib.hvla_dino_vec_elems() == P.dino_vec_layout()["__total__"][0]
    assert lib.hvla_dino_mat_elems() == P.dino_mat_layout(True)["__total__"][0] == P.dino_mat_layout(False)["__total__"][0]


def test_generated_row_offsets_agree(lib):
    from hvla import _native as N, metadata as M
    tbl = M.packed_offsets()
    A = "MultiHeadDotProductAttention_0"
    names = {
        "gen.proj_w": ("encoder", "image_embedding_projection", "kernel"),
        "gen.proj_b": ("encoder", "image_embedding_projection", "bias"),
        "gen.pos": ("encoder", "pos_embedding"),
        "gen.encn_s": ("encoder", "Transformer_0", "encoder_norm", "scale"),
        "gen.encn_b": ("encoder", "Transformer_0", "encoder_norm", "bias"),
        "gen.wc": ("action_head", "continuous_head", "kernel"), "gen.bc": ("action_head", "continuous_head", "bias"),
        "gen.wd": ("action_head", "discrete_head", "kernel"), "gen.bd": ("action_head", "discrete_head", "bias"),
    }
    for l in range(4):
        blk = ("encoder", "Transformer_0", f"encoderblock_{l}")
        names.update({
            f"gen.l{l}.ln0_s": blk + ("LayerNorm_0", "scale"), f"gen.l{l}.ln0_b": blk + ("LayerNorm_0", "bias"),
            f"gen.l{l}.wq": blk + (A, "query", "kernel"), f"gen.l{l}.bq": blk + (A, "query", "bias"),
            f"gen.l{l}.wk": blk + (A, "key", "kernel"), f"gen.l{l}.bk": blk + (A, "key", "bias"),
            f"gen.l{l}.wv": blk + (A, "value", "kernel"), f"gen.l{l}.bv": blk + (A, "value", "bias"),
            f"gen.l{l}.wo": blk + (A, "out", "kernel"), f"gen.l{l}.bo": blk + (A, "out", "bias"),
            f"gen.l{l}.ln1_s": blk + ("LayerNorm_1", "scale"), f"gen.l{l}.ln1_b": blk + ("LayerNorm_1", "bias"),
            f"gen.l{l}.w0": blk + ("MlpBlock_0", "Dense_0", "kernel"), f"gen.l{l}.b0": blk + ("MlpBlock_0", "Dense_0", "bias"),
            f"gen.l{l}.w1": blk + ("MlpBlock_0", "Dense_1", "kernel"), f"gen.l{l}.b1": blk + ("MlpBlock_0", "Dense_1", "bias"),
        })
    assert len(names) == 73
    for cname, path in names.items():
        assert N.layout_offset(cname) == tbl[path][0], cname
    assert N.layout_offset("gen.total") == 201500
    assert N.layout_offset("gen.nonsense") == -1


def test_dino_and_hn_offsets_agree(lib):
    from hvla import _native as N, params as P
    vl = P.dino_vec_layout()
    for name, (off, _) in vl.items():
        if name == "__total__":
            assert N.layout_offset("dvec.total") == off
        else:
            assert N.layout_offset("dvec." + name) == off, name
    ml = P.dino_mat_layout(True)
    for name, (off, _) in ml.items():
        if name == "__total__":
            assert N.layout_offset("dmat.total") == off
        else:
            assert N.layout_offset("dmat." + name) == off, name
    # HN blob: recompute the running offsets the packer uses
    d, m = 128, 512
    off = 0
    expect = {}
    for nm, n in (("tok_w", 768 * d), ("tok_b", d), ("img_w", 768 * d), ("img_b", d), ("task_pos", 32 * d), ("img_pos", d), ("layer_pos", d)):
        expect["hn." + nm] = off
        off += n
    for l in range(6):
        for nm, n in (("ln0_s", d), ("ln0_b", d), ("wqkv", d * 3 * d), ("bqkv", 3 * d), ("wo", d * d), ("bo", d), ("ln1_s", d),
                      ("ln1_b", d), ("w0", d * m), ("b0", m), ("w1", m * d), ("b1", d)):
            expect[f"hn.l{l}.{nm}"] = off
            off += n
    expect["hn.encn_s"] = off
    expect["hn.encn_b"] = off + d
    expect["hn.total"] = off + 2 * d
    for k, v in expect.items():
        assert N.layout_offset(k) == v, k


def test_workspace_is_monotonic_and_host_only(lib):
    w1 = lib.hvla_workspace_bytes(1, 1, 1)
    w64 = lib.hvla_workspace_bytes(64, 64, 1)
    assert 0 < w1 < w64
    assert lib.hvla_workspace_bytes(64, 64, 0) > w64      # fp32 scratch is larger than bf16


def test_argument_errors_do_not_need_a_gpu(lib):
    from hvla import _native as N
    assert lib.hvla_act(None, None, None, None, None, None, 1, 1, None, None, None, 0, 1) == -1
    assert b"NULL" in lib.hvla_last_error()
    assert lib.hvla_generate(None, None, None, None, None, None, None, None, None, 1, None, None, None, 0, 0) == -1


def test_xla_status_wrappers_report_a_short_opaque_without_a_gpu(lib):
    """include/hvla.h: the status-returning custom-call form must report failures through the XlaCustomCallStatus hook
    (here the explicitly registered setter) instead of returning silently with untouched outputs."""
    import ctypes as C
    seen = []
    SETTER = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_size_t)

    @SETTER
    def setter(status, msg, n):
        seen.append((status, msg[:n].decode()))
    lib.hvla_xla_register_status_setter(C.cast(setter, C.c_void_p))
    try:
        bufs = (C.c_void_p * 11)()
        token = C.c_int(0)
        for fn, name in ((lib.hvla_xla_act_status, "hvla_xla_act"), (lib.hvla_xla_generate_status, "hvla_xla_generate")):
            fn(None, bufs, b"\0" * 8, 8, C.addressof(token))
            assert seen and seen[-1][0] == C.addressof(token) and name in seen[-1][1] and "opaque" in seen[-1][1]
        n = len(seen)
        lib.hvla_xla_act(None, bufs, b"\0" * 8, 8)          # original form: no status channel, must not call the setter
        assert len(seen) == n
    finally:
        lib.hvla_xla_register_status_setter(None)
