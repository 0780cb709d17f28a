"""CPU tests of the host-side mirror: config envelope, metadata rule, packing, API errors, sharding."""
import copy

import numpy as np
import pytest

from hvla import config as C, metadata as M, params as P, parallel as PL


def test_metadata_counts_and_rule():
    meta = M.build_base_net_metadata(C.default_config())
    gen = M.generated_leaves_canonical()
    assert len(gen) == 73 and len(M.shared_leaves()) == 223
    assert sum(int(np.prod(s)) for _, s in gen) == 201_500
    assert sum(int(np.prod(s)) for _, s in M.shared_leaves()) == 86_580_480
    assert meta["total_param_num"] == 201_500 + 86_580_480 and meta["block_num"] == 1
    name = "encoder_Transformer_0_encoderblock_0_MultiHeadDotProductAttention_0_query_kernel"
    assert meta["output_head_info"][name] == dict(output_dim=4096, generation_flag=True, init_strategy=0, init_variance=0.0)
    assert not meta["output_head_info"]["encoder_image_encoder_layernorm_scale"]["generation_flag"]


def test_canonical_flatten_order_matches_survey_appendix_a3():
    off, table = 0, {}
    for path, shape in M.generated_leaves_canonical():
        table["/".join(path)] = off
        off += int(np.prod(shape))
    assert table["action_head/continuous_head/bias"] == 0
    assert table["action_head/discrete_head/kernel"] == 1564
    assert table["encoder/Transformer_0/encoder_norm/bias"] == 1820
    assert table["encoder/Transformer_0/encoderblock_0/MlpBlock_0/Dense_0/bias"] == 2204
    assert table["encoder/Transformer_0/encoderblock_0/MultiHeadDotProductAttention_0/key/bias"] == 18780
    assert table["encoder/Transformer_0/encoderblock_1/LayerNorm_0/bias"] == 35420
    assert table["encoder/image_embedding_projection/bias"] == 135836
    assert table["encoder/pos_embedding"] == 185052 and off == 201_500


def test_pack_unpack_roundtrip(params_p1):
    W, b = P.pack_heads(params_p1)
    assert W.shape == (128, M.N_GENERATED_PADDED) and not W[:, M.N_GENERATED:].any()
    tree = P.unpack_generated(b[:M.N_GENERATED])
    k = tree["encoder"]["Transformer_0"]["encoderblock_3"]["MlpBlock_0"]["Dense_1"]["kernel"]
    name = "output_head_encoder_Transformer_0_encoderblock_3_MlpBlock_0_Dense_1_kernel"
    assert k.shape == (128, 64) and np.array_equal(k.ravel(), params_p1[name]["bias"])
    assert P.pack_hn_blob(params_p1).size == P.hn_blob_size()
    dino = P.dino_tree_from_params(params_p1)
    assert dino["encoder"]["layer"]["11"]["mlp"]["fc2"]["kernel"].shape == (3072, 768)


def test_dino_packing_layouts(params_p1):
    vec, mat_t = P.pack_dino(params_p1, transposed=True)
    _, mat = P.pack_dino(params_p1, transposed=False)
    lt, ln = P.dino_mat_layout(True), P.dino_mat_layout(False)
    o, (r, c) = lt["l3.w1"]
    o2, (r2, c2) = ln["l3.w1"]
    assert (r, c) == (3072, 768) and (r2, c2) == (768, 3072) and o == o2
    dino = P.dino_tree_from_params(params_p1)
    L3 = dino["encoder"]["layer"]["3"]
    # tensor-core layout: transposed, with LayerNorm-2's gamma folded into fc1 (LN(x) W + b = xhat (gamma*W) + (beta W + b))
    assert np.array_equal(mat_t[o:o + r * c].reshape(r, c).T, L3["norm2"]["scale"].astype(np.float32)[:, None] * mat[o2:o2 + r2 * c2].reshape(r2, c2))
    vl = P.dino_vec_layout()
    b1f = vec[vl["l3.b1_f"][0]:][:3072]
    ref = L3["norm2"]["bias"].astype(np.float64) @ L3["mlp"]["fc1"]["kernel"].astype(np.float64) + L3["mlp"]["fc1"]["bias"]
    assert np.allclose(b1f, ref, rtol=1e-6, atol=1e-7)
    cs = vec[vl["l3.cs_1"][0]:][:3072]
    assert np.allclose(cs, P.bf16_round(L3["norm2"]["scale"].astype(np.float32)[:, None] * L3["mlp"]["fc1"]["kernel"]).sum(0), rtol=1e-5, atol=1e-5)
    o, (r, c) = lt["l3.w2"]
    assert np.array_equal(mat_t[o:o + r * c].reshape(r, c).T, mat[o:o + r * c].reshape(c, r))        # fc2 / wo are not folded
    o, (r, c) = ln["patch_w"]
    assert not mat[o:o + r * c].reshape(r, c)[588:].any()          # K padding rows are zero
    vo, vn = P.dino_vec_layout()["pos"]
    pos = vec[vo:vo + vn].reshape(257, 768)
    assert np.array_equal(pos[0], dino["embeddings"]["position_embeddings"][0, 0])   # CLS position is copied
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415927, 65504.0], np.float32)
    import torch
    assert np.array_equal(P.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def test_pos_table_interpolation_agrees_with_oracle_and_is_a_partition_of_unity(params_p1):
    from oracle import hypervla_oracle as O
    pe = P.dino_tree_from_params(params_p1)["embeddings"]["position_embeddings"]
    assert np.array_equal(P.interpolate_pos_table(pe), O.interpolate_pos_table(pe))
    ones = np.ones_like(pe)
    assert np.allclose(P.interpolate_pos_table(ones), 1.0, atol=1e-5)


@pytest.mark.parametrize("path,value", [
    (("hypernet_kwargs", "context_embedding_dim"), 256), (("hypernet_kwargs", "generation_strategy"), "full"),
    (("hypernet_kwargs", "shared_modules"), ()), (("base_net_kwargs", "action_head_type"), "diffusion"),
    (("base_net_kwargs", "vit_kwargs", "encoder_type"), "SmallStem"), (("base_net_kwargs", "vit_kwargs", "num_layers"), 6),
    (("base_net_kwargs", "vit_kwargs", "use_language_token"), True), (("base_net_kwargs", "action_head_kwargs", "token_per_horizon"), True),
])
def test_unsupported_config_raises_before_any_launch(path, value):
    cfg = copy.deepcopy(C.default_config())
    d = cfg
    for k in path[:-1]:
        d = d[k]
    d[path[-1]] = value
    with pytest.raises(ValueError):
        C.validate_config(cfg)
    C.validate_config(C.default_config())


def test_model_api_errors_without_gpu(params_p1):
    """Constructing the model and validating arguments needs no GPU; compute fails loudly (no CPU fallback)."""
    import torch
    from hvla import _native as N
    from hvla.model import HyperVLA
    m = HyperVLA.from_config(C.default_config(), precision="bf16", params=params_p1)
    assert m.base_net.action_horizon == 4 and m.hypernet.layer_token_num == 1
    with pytest.raises(ValueError):
        m.create_tasks(instruction_dict=None)
    with pytest.raises(TypeError):
        m.sample_actions(np.zeros((1, 1, 224, 224, 3), np.uint8), None, None, None, base_params={"not": "ours"})
    if not torch.cuda.is_available():
        from hvla import synthetic as S
        inp = S.make_inputs(1, 1, 1)
        with pytest.raises(N.HvlaError):
            m.create_tasks(instruction_dict=inp["instruction_dict"], initial_state=inp["initial_state"])


def test_save_and_load_pretrained_roundtrip(tmp_path, params_p0):
    from hvla.model import HyperVLA
    small = {k: v for k, v in params_p0.items()}
    m = HyperVLA.from_config(C.default_config(), params=small, dataset_statistics={"bridge": {"action": {"mean": [0.0] * 7}}})
    m.save_pretrained(100, str(tmp_path))
    m2 = HyperVLA.load_pretrained(str(tmp_path), 100)
    assert m2.config["hypernet_kwargs"]["shared_modules"] == ("image_encoder",)
    a = m.params["context_encoder"]["encoderblock_2"]["MlpBlock_0"]["Dense_0"]["kernel"]
    b = m2.params["context_encoder"]["encoderblock_2"]["MlpBlock_0"]["Dense_0"]["kernel"]
    assert np.array_equal(a, b) and m2.dataset_statistics["bridge"]["action"]["mean"] == [0.0] * 7


def test_shard_range_covers_every_env_once():
    for n, w in ((1024, 8), (1000, 8), (7, 4), (3, 8), (0, 2)):
        spans = [PL.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    ti = np.repeat(np.arange(10), 100)
    uniq, local = PL.local_task_table(ti, 250, 375)
    assert uniq.tolist() == [2, 3] and np.array_equal(uniq[local], ti[250:375])


def test_bench_kernel_roofline_table_arithmetic():
    """bench.py's kernel_rooflines: algorithmic work (SURVEY.md 8(d) figures x units of the run) / live time against the measured
    peaks; a tensor-bound class is reported in TFLOP/s, an HBM-bound class in GB/s, frac = achieved / peak."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("hvla_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    act = {"im2col": (1, 0.024), "cls_rows": (1, 0.019), "gemm_tc": (49, 2.471), "layernorm": (25, 0.3683),
           "dino_attention": (12, 0.4245), "base_fused": (1, 0.0947), "unknown_class": (3, 1.0)}
    gen = {"ctx_fused": (1, 0.1528), "heads_gemm": (1, 0.028), "idle": (1, 0.0)}
    peaks = {"hbm_gbs": 6542.1, "bf16_tflops_sustained": 1386.3}
    peaks["bf16_tflops"] = 1648.0
    # the tensor denominator follows the SM clock sampled while the kernels were timed (a burst-clock time must not
    # be divided by the sustained peak)
    assert bench.tensor_peak(peaks, {"sm_mhz": 1965.0, "sm_max_mhz": 1965, "reasons": []})[0] == 1648.0
    assert bench.tensor_peak(peaks, {"sm_mhz": 1400.0, "sm_max_mhz": 1965, "reasons": ["sw_power_cap"]})[0] == 1386.3
    assert bench.tensor_peak({}, {"sm_mhz": None, "sm_max_mhz": None, "reasons": []})[0] == 1400.0
    assert bench.default_batch(1) == 64 and [bench.default_batch(n) for n in (2, 4, 8)] == [512, 256, 128]
    assert "configs[2]" in bench.workload_config(128, 8)["workload"] and "configs[1]" in bench.workload_config(64, 1)["workload"]
    t = bench.kernel_roofline_table(act, gen, 64, 64, peaks, "bf16", 1386.3)
    assert set(t) == {"im2col", "cls_rows", "gemm_tc", "layernorm", "dino_attention", "base_fused", "ctx_fused", "heads_gemm"}
    for name, r in t.items():
        assert r["unit"] == ("TFLOP/s" if r["bound"] == "tensor" else "GB/s")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert t["gemm_tc"]["bound"] == "tensor" and abs(t["gemm_tc"]["achieved"] - (46_322_454_528 - 12 * 4 * 257 * 257 * 768) * 64 / 2.471e-3 / 1e12) < 0.5
    # base net: 796,328 algorithmic bytes per env when every env has its own weights
    assert abs(t["base_fused"]["achieved"] - 796_328 * 64 / 0.0947e-3 / 1e9) < 0.5
    # fallback peaks of the profiling recipe when MEASURED_PEAKS.json is absent
    t2 = bench.kernel_roofline_table(act, gen, 64, 64, {}, "bf16", bench.tensor_peak({}, {})[0])
    assert t2["layernorm"]["peak"] == 6650.0 and t2["gemm_tc"]["peak"] == 1400.0


def test_action_head_scalars_outside_the_kernel_constants_are_rejected():
    """The head epilogues hard-code tanh(x / 5) * 5 (MixActionHead defaults, action_heads.py:439, 445)."""
    for key, val in (("tanh_scaling_factor", 2.0), ("max_action", 1.0)):
        cfg = C.default_config()
        cfg["base_net_kwargs"]["action_head_kwargs"][key] = val
        with pytest.raises(ValueError, match=key):
            C.validate_config(cfg)
    cfg = C.default_config()
    del cfg["base_net_kwargs"]["action_head_kwargs"]["tanh_scaling_factor"]      # absent -> reference default 5.0
    C.validate_config(cfg)


def test_replace_params_drops_the_device_runtime(params_p1, params_p0):
    """model.replace(params=ema) is the reference's EMA swap (data/simpler/evaluate.py:443): the copy must not keep device
    blobs of the old parameters."""
    from hvla.model import HyperVLA
    m = HyperVLA.from_config(C.default_config(), params=params_p1)
    m._runtime = object()                        # stands for an uploaded runtime (no GPU here)
    m2 = m.replace(params=params_p0)
    assert m2.params is params_p0 and m2._runtime is None and m._runtime is not None
    assert m.replace(precision="fp32")._runtime is None and m.replace(config=m.config)._runtime is m._runtime


def test_split_operand_matrix_packing():
    """hvla.params.split_matrices_x3 (dtype HVLA_BF16X3, csrc/dino_x3.cuh): every matrix [N, 3K] = [hi | hi | lo] of its transpose at 3x its
    offset, hi + lo reproducing the fp32 weight to 2^-17."""
    rng = np.random.default_rng(3)
    lay = P.dino_mat_layout(False)
    mat = (rng.standard_normal(lay["__total__"][0]) * 0.02).astype(np.float32)
    x3 = P.split_matrices_x3(mat)
    assert x3.dtype == np.uint16 and x3.size == 3 * mat.size
    f = lambda u: (u.astype(np.uint32) << 16).view(np.float32)
    for name in ("patch_w", "l0.wqkv", "l11.w2"):
        off, (k, n) = lay[name]
        blk = x3[3 * off:3 * off + 3 * k * n].reshape(n, 3 * k)
        w = mat[off:off + k * n].reshape(k, n).T
        hi, hi2, lo = f(blk[:, :k]), f(blk[:, k:2 * k]), f(blk[:, 2 * k:])
        assert np.array_equal(hi, hi2) and np.array_equal(hi, P.bf16_round(w))
        assert np.abs(hi + lo - w).max() <= 2.0 ** -17 * np.abs(w).max()


def test_discrete_head_configuration_and_layout():
    """action_head_type='discrete' (base_network.py:22-33): 4 or 28 readout tokens, 71 generated leaves, row sizes agree with the C library."""
    from hvla import _native as N
    for tok, A, V, total in (("action_horizon", 4, 1792, 316352), ("action_dim_and_action_horizon", 28, 256, 218048)):
        cfg = C.default_config()
        cfg["base_net_kwargs"]["action_head_type"] = "discrete"
        cfg["base_net_kwargs"]["action_head_kwargs"] = {"discrete_token_type": tok}
        cfg = C.validate_config(cfg)
        spec = M.HeadSpec.from_config(cfg)
        assert (spec.kind, spec.n_action_tokens, spec.vocab_out, spec.tokens) == ("discrete", A, V, 256 + A)
        assert M.n_generated(spec) == total == N.lib().hvla_discrete_generated_elems(A)
        assert M.n_generated_padded(spec) == N.lib().hvla_discrete_row_stride(A)
        tbl = M.packed_offsets(spec)
        assert tbl[("encoder", "pos_embedding")] == (49216, (1, 256 + A, 64))
        assert tbl[("action_head", "vocab_proj", "kernel")][0] == 49216 + (256 + A) * 64 + 4 * 33472 + 128
        meta = M.build_base_net_metadata(cfg)
        assert meta["output_head_info"]["action_head_vocab_proj_kernel"]["output_dim"] == 64 * V
    bad = C.default_config()
    bad["base_net_kwargs"]["action_head_type"] = "discrete"
    bad["base_net_kwargs"]["action_head_kwargs"] = {"discrete_token_type": ""}
    with pytest.raises(ValueError):
        C.validate_config(bad)
    assert N.lib().hvla_discrete_row_stride(5) == -1


# ---- GEMM chain unit list (csrc/gemm_chain.cuh: units_before / decode_unit), restated: the invariants the kernel's liveness rests on ----
def _chain_units(ntm, ntn, lags):
    """The ordered unit list of a chain of len(ntn) GEMMs over ntm row blocks: slot s holds the column tiles of row block s - lag_g of GEMM g."""
    lag = [0]
    for l in lags[:len(ntn) - 1]:
        lag.append(lag[-1] + min(l, ntm))
    n_slots = ntm + lag[-1]

    def before(s):
        return sum(min(max(s - lag[g], 0), ntm) * ntn[g] for g in range(len(ntn)))
    units = []
    for s in range(n_slots):
        assert before(s) == len(units)
        for g in range(len(ntn)):
            m = s - lag[g]
            if 0 <= m < ntm:
                units += [(g, m, n) for n in range(ntn[g])]
    assert before(n_slots) == len(units)
    return units


@pytest.mark.parametrize("ntm", [1, 3, 26, 65, 514])
@pytest.mark.parametrize("lags", [(0, 0, 0), (3, 3, 3), (12, 10, 24), (1 << 20, 1 << 20, 1 << 20)])
@pytest.mark.parametrize("ntn", [(3, 12, 3, 9), (3, 12, 3)])
def test_gemm_chain_unit_list_is_a_permutation_with_dependencies_first(ntm, lags, ntn):
    units = _chain_units(ntm, ntn, lags)
    assert len(units) == len(set(units)) == ntm * sum(ntn)                     # every (GEMM, row block, column tile) exactly once
    pos = {u: i for i, u in enumerate(units)}
    for (g, m, n), i in pos.items():
        if g > 0:                                                              # a unit reads row block m of GEMM g - 1: all of its tiles come earlier
            assert all(pos[(g - 1, m, k)] < i for k in range(ntn[g - 1]))
    if min(lags) >= ntm:                                                       # default lags: GEMM after GEMM
        assert [u[0] for u in units] == sorted(u[0] for u in units)
