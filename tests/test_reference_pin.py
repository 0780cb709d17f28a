"""Pins the oracle (and through it the CUDA path) to the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz were produced by tests/golden/make_ref_golden.py, which imports the unmodified
/root/reference/hypervla/model.py + components and executes them in float64 through oracle/refshim
(NumPy stand-ins for the jax/flax primitives; HF torch DINOv2 for the un-vendored FlaxDinov2Model).

* always (also on the GPU box, where /root/reference does not exist): the fp64 / fp32 oracle against the
  committed reference-run fixtures;
* when /root/reference is present: the reference is re-run here and must reproduce the fixtures, its
  `from_config` must produce the same parameter pytree (names and shapes) as hvla.params, and its
  `init_base_net` metadata must agree with hvla.metadata.
"""
import os

import numpy as np
import pytest

REF_ROOT = "/root/reference"
CASES = {"ref_c1_b1_t1": (1, 1, 1), "ref_c2_b3_t3": (2, 3, 3), "ref_c5_b6_t2": (5, 6, 2)}
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "hypervla")),
                                     reason="reference checkout not present (GPU box): fixtures only")


def rel(x, ref):
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


def oracle_case(params, ci, B, T, dtype):
    from hvla import metadata as M, params as P, synthetic as S
    from oracle import hypervla_oracle as O
    inp = S.make_inputs(ci, B, T)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, emb = O.generate(params, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                          dtype=dtype, generated_paths=M.generated_leaves_canonical())
    per_env = O.to_tree(O.take_tasks(gen, inp["task_index"]))
    act, logit, hidden, h = O.sample_actions(P.dino_tree_from_params(params), per_env, inp["images"][:, 0], dtype=dtype,
                                             return_all=True)
    rows = np.zeros((T, M.N_GENERATED), np.float64)
    for path, (off, shape) in M.packed_offsets().items():
        n = int(np.prod(shape))
        rows[:, off:off + n] = gen[path].reshape(T, n)
    return dict(rows=rows, ctx=emb[:, 0], action=act, logit=logit, h=h)


@pytest.mark.parametrize("case", list(CASES))
def test_fp64_oracle_matches_reference_run(params_p1, golden, case):
    """The restatement and the reference's own code agree to float64 round-off."""
    ci, B, T = CASES[case]
    g, r = golden[case], oracle_case(params_p1, *CASES[case], np.float64)
    assert "reference code executed" in str(g["source"])
    assert rel(r["ctx"], g["ctx"]) < 1e-12
    assert rel(r["rows"][:, ::97], g["rows_sample"]) < 1e-12
    assert rel(np.abs(r["rows"]).sum(1), g["rows_abs"]) < 1e-12
    assert rel(r["h"].reshape(B, -1), g["h"]) < 1e-9
    assert rel(r["logit"].reshape(B, -1), g["logit"]) < 1e-9
    assert rel(r["action"][..., :6], g["action"][..., :6]) < 2e-7          # the oracle returns float32 actions
    sure = np.abs(g["logit"].reshape(B, 4)) > 1e-9
    assert np.array_equal(np.asarray(r["action"])[..., 6][sure], g["action"][..., 6][sure])


def test_fp64_oracle_matches_the_full_size_reference_run_on_sampled_envs(params_p1, golden):
    """BASELINE configs[1] at its full size (64 envs, one task each): the reference's own code was run for all 64 environments
    (ref_c2_b64_t64); the restatement is checked on all 64 generated weight sets and on a spread of environments."""
    from hvla import metadata as M, params as P, synthetic as S
    from oracle import hypervla_oracle as O
    g = golden["ref_c2_b64_t64"]
    ci, B, T = int(g["config_index"]), int(g["B"]), int(g["T"])
    assert (B, T) == (64, 64) and "reference code executed" in str(g["source"])
    inp = S.make_inputs(ci, B, T)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, emb = O.generate(params_p1, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                          dtype=np.float64, generated_paths=M.generated_leaves_canonical())
    rows = np.zeros((T, M.N_GENERATED), np.float64)
    for path, (off, shape) in M.packed_offsets().items():
        n = int(np.prod(shape))
        rows[:, off:off + n] = gen[path].reshape(T, n)
    assert rel(emb[:, 0], g["ctx"]) < 1e-12
    assert rel(rows[:, ::97], g["rows_sample"]) < 1e-12
    assert rel(np.abs(rows).sum(1), g["rows_abs"]) < 1e-12
    envs = np.array([0, 21, 42, 63])
    per_env = O.to_tree(O.take_tasks(gen, inp["task_index"][envs]))
    act, logit, _, h = O.sample_actions(P.dino_tree_from_params(params_p1), per_env, inp["images"][envs, 0], dtype=np.float64, return_all=True)
    assert rel(h.reshape(len(envs), -1), g["h"][envs]) < 1e-9
    assert rel(logit.reshape(len(envs), -1), g["logit"][envs]) < 1e-9
    assert rel(act[..., :6], g["action"][envs][..., :6]) < 2e-7


def test_fp32_oracle_within_north_star_tolerance_of_reference_run(params_p1, golden):
    """The fp32 restatement (the JAX-CPU stand-in timed by bench.py) is within 1e-5 of the reference run."""
    case = "ref_c2_b3_t3"
    ci, B, T = CASES[case]
    g, r = golden[case], oracle_case(params_p1, ci, B, T, np.float32)
    assert rel(r["rows"][:, ::97], g["rows_sample"]) < 1e-5
    assert rel(r["action"][..., :6], g["action"][..., :6]) < 1e-5


def test_oracle_golden_and_reference_golden_agree(golden):
    """The older fixtures (fp64 oracle) and the reference-run fixtures describe the same numbers."""
    for case in CASES:
        g, o = golden[case], golden[case[4:]]
        assert rel(o["rows_sample"], g["rows_sample"]) < 1e-12
        assert rel(o["ctx"], g["ctx"]) < 1e-12
        assert rel(o["logit"].reshape(g["logit"].shape), g["logit"]) < 1e-9
        assert rel(o["action"][..., :6], g["action"][..., :6]) < 1e-6       # the older fixture stores float32 actions


@pytest.fixture(scope="module")
def reference_model(params_p1):
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location(
        "make_ref_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["make_ref_golden"] = mod
    spec.loader.exec_module(mod)
    model_init, RM = mod.build_reference_model(None)             # reference-initialised parameters
    return mod, model_init, model_init.replace(params=mod.f64(params_p1)), RM


@needs_reference
def test_reference_rerun_reproduces_fixture(reference_model, golden):
    mod, _, model, RM = reference_model
    case = "ref_c5_b6_t2"
    r, g = mod.run_reference(model, RM, *CASES[case]), golden[case]
    for k in ("ctx", "action", "logit", "h"):
        assert rel(r[k], g[k]) < 1e-12, k
    assert rel(r["rows"][:, ::97], g["rows_sample"]) < 1e-12


@needs_reference
def test_reference_from_config_pytree_equals_ours(reference_model, params_p1):
    """HyperVLA.from_config (hypervla/model.py:286-368) run through the shim yields the pytree hvla.params builds."""
    _, model_init, _, _ = reference_model

    def flat(d, pre=()):
        out = {}
        for k, v in d.items():
            out.update(flat(v, pre + (k,)) if isinstance(v, dict) else {pre + (k,): tuple(np.shape(v))})
        return out
    ref, ours = flat(model_init.params), flat(params_p1)
    assert ref == ours
    assert len(ref) == 474
    # BIAS_INIT (model.py:328-346): head kernels are zero, head biases hold the base net's own init draw
    k = "output_head_encoder_Transformer_0_encoderblock_0_MlpBlock_0_Dense_0_kernel"
    assert not np.any(model_init.params[k]["kernel"]) and np.any(model_init.params[k]["bias"])


@needs_reference
def test_reference_metadata_equals_ours(reference_model):
    """init_base_net's generated-vs-shared rule, head names and sizes (model.py:390-515) vs hvla.metadata."""
    from hvla import config as C, metadata as M
    _, model_init, _, _ = reference_model
    ref = model_init.base_net_metadata
    ours = M.build_base_net_metadata(C.default_config())
    assert ref["block_num"] == ours["block_num"] == 1
    assert ref["total_param_num"] == ours["total_param_num"] == 86_781_980
    assert list(np.asarray(ref["layer_token_mask"])) == list(np.asarray(ours["layer_token_mask"]))
    rinfo, oinfo = ref["output_head_info"], ours["output_head_info"]
    assert set(rinfo) == set(oinfo) and len(rinfo) == 296
    for name, info in rinfo.items():
        assert int(info["output_dim"]) == int(oinfo[name]["output_dim"]), name
        assert bool(info["generation_flag"]) == bool(oinfo[name]["generation_flag"]), name
        assert int(info["init_strategy"]) == int(oinfo[name]["init_strategy"]), name
    generated = sorted(n for n, i in rinfo.items() if i["generation_flag"])
    assert len(generated) == 73 and sum(int(rinfo[n]["output_dim"]) for n in generated) == 201_500
    for path, shape in M.generated_leaves_canonical():
        node = ref["param_shape"]
        for k in path:
            node = node[k]
        assert tuple(int(v) for v in node) == tuple(shape), path


@needs_reference
def test_reference_initialised_params_degenerate_case(reference_model):
    """with the reference's own from_config initialisation (BIAS_INIT) every head kernel is zero, so the weights
    the reference generates are its head biases whatever the context -- executed with the reference's code and parameters,
    and reproduced by the oracle on the same parameters."""
    from hvla import metadata as M, synthetic as S
    from oracle import hypervla_oracle as O
    mod, model_init, _, RM = reference_model
    inp = S.make_inputs(7, 1, 1)
    idict, istate = mod.f64(inp["instruction_dict"]), mod.f64(inp["initial_state"])
    base_params, _, _ = model_init.create_tasks(instruction_dict=idict, initial_state=istate)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, _ = O.generate(model_init.params, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                        dtype=np.float64, generated_paths=M.generated_leaves_canonical())
    for path, shape in M.generated_leaves_canonical():
        leaf = base_params
        for k in path:
            leaf = leaf[k]
        bias = np.asarray(model_init.params["output_head_" + "_".join(path)]["bias"], np.float64).reshape(shape)
        assert np.array_equal(np.asarray(leaf), bias), path                 # reference: generated leaf == head bias, exactly
        assert np.abs(gen[path][0] - bias).max() < 1e-15, path              # oracle on the same parameters


@pytest.mark.parametrize("name", ["ref_disc4_b2_t2", "ref_disc28_b2_t2"])
def test_discrete_head_oracle_matches_the_reference_run_fixture(golden, name):
    """DiscreteActionHead configuration: the oracle's multi-readout-token base ViT, vocab projection, argmax and
    BinTokenizer.decode against fixtures produced by the reference's own code (tests/golden/make_ref_discrete_golden.py)."""
    from hvla import metadata as M, params as P, synthetic as S
    from oracle import hypervla_oracle as O
    g = golden[name]
    A = int(g["n_action_tokens"])
    spec = M.HeadSpec("discrete", A)
    assert M.n_generated(spec) == {4: 316352, 28: 218048}[A]
    params = P.init_params(2025, "P1", spec)
    inp = S.make_inputs(int(g["config_index"]), int(g["B"]), int(g["T"]))
    lang = inp["instruction_dict"]["language_instruction"]
    gen, ctx = O.generate(params, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                          generated_paths=M.generated_leaves_canonical(spec), dtype=np.float64)
    rows = np.concatenate([np.asarray(gen[p]).reshape(int(g["T"]), -1) for p, _ in M.generated_leaves_packed(spec)], axis=1)
    assert np.abs(rows[:, ::97] - g["rows_sample"]).max() <= 1e-12 * np.abs(g["rows_sample"]).max()
    tree = O.to_tree(O.take_tasks(gen, inp["task_index"]))
    act, tokens, logits, h = O.sample_actions_discrete(P.dino_tree_from_params(params), tree, inp["images"][:, 0], A, dtype=np.float64)
    assert np.abs(h - g["h"]).max() <= 1e-8 * np.abs(g["h"]).max()
    assert np.abs(logits - g["logits"]).max() <= 1e-8 * np.abs(g["logits"]).max()
    assert np.array_equal(tokens, g["tokens"]) and np.array_equal(act.astype(np.float64), g["action"])
    # bin centres of 256 uniform bins over [-1, 1]
    assert np.allclose(g["action"], -1.0 + (2.0 * g["tokens"] + 1.0) / 256.0, atol=0)
