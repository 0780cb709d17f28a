"""CPU tests of the oracle: fp32 restatement vs the committed fp64 golden vectors, and the
size-independent properties the GPU tests rely on."""
import numpy as np
import pytest

from hvla import metadata as M, params as P, synthetic as S
from oracle import hypervla_oracle as O

CASES = {"c1_b1_t1": (1, 1, 1), "c2_b3_t3": (2, 3, 3)}


def rel(x, ref):
    return float(np.abs(np.asarray(x, np.float64) - ref).max() / np.abs(ref).max())


def _run(params, ci, B, T, dtype=np.float32):
    inp = S.make_inputs(ci, B, T)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, emb = O.generate(params, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                          dtype=dtype, generated_paths=M.generated_leaves_canonical())
    tree = O.to_tree(O.take_tasks(gen, inp["task_index"]))
    act, logit, hidden, h = O.sample_actions(P.dino_tree_from_params(params), tree, inp["images"][:, 0], dtype=dtype, return_all=True)
    return inp, gen, emb, act, logit, hidden


@pytest.mark.parametrize("case", list(CASES))
def test_fp32_oracle_matches_fp64_golden(params_p1, golden, case):
    ci, B, T = CASES[case]
    g = golden[case]
    inp, gen, emb, act, logit, hidden = _run(params_p1, ci, B, T)
    assert rel(emb[:, 0], g["ctx"]) < 1e-5
    rows = np.zeros((T, M.N_GENERATED))
    for path, (off, shape) in M.packed_offsets().items():
        rows[:, off:off + int(np.prod(shape))] = gen[path].reshape(T, -1)
    assert rel(rows[:, ::97], g["rows_sample"]) < 1e-5
    assert rel(hidden[:, ::16, ::48], g["hidden_sample"]) < 1e-5
    assert rel(act[..., :6], g["action"][..., :6].astype(np.float64)) < 1e-5
    sure = np.abs(g["logit"]) > 1e-3
    assert np.array_equal(act[..., 6][sure], g["action"][..., 6][sure])


def test_p0_degenerate_init_reproduces_head_biases(params_p0):
    """with BIAS_INIT the head kernels are zero, so generated weights == head biases whatever
    the context is (hypernetwork.py:72-77, model.py:328-346)."""
    inp = S.make_inputs(9, 1, 2)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, _ = O.generate(params_p0, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                        generated_paths=M.generated_leaves_canonical())
    for path, v in gen.items():
        bias = params_p0["output_head_" + "_".join(path)]["bias"]
        assert np.array_equal(v[0].ravel(), bias) and np.array_equal(v[1].ravel(), bias)


def test_patches_do_not_see_the_action_token(params_p1):
    """base_vit.py:209-214: changing the action token's position embedding must not change any patch row;
    and the last block only needs the action-token query (the kernel's legal shortcut)."""
    rng = np.random.default_rng(3)
    inp = S.make_inputs(4, 1, 1)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, _ = O.generate(params_p1, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                        generated_paths=M.generated_leaves_canonical())
    tree = O.to_tree(gen)
    emb = rng.standard_normal((1, 256, 768)).astype(np.float32)

    def tokens(t):
        enc = t["encoder"]
        x = np.matmul(emb, enc["image_embedding_projection"]["kernel"]) + enc["image_embedding_projection"]["bias"][:, None]
        x = np.concatenate([x, np.zeros((1, 1, 64), np.float32)], 1) + enc["pos_embedding"].reshape(1, 257, 64)
        return O.transformer(x, enc["Transformer_0"], O.base_mask(1, 257), 4, per_sample=True)

    a = tokens(tree)
    gen2 = dict(gen)
    pe = gen[("encoder", "pos_embedding")].copy()
    pe[:, :, 256] += 1.0
    gen2[("encoder", "pos_embedding")] = pe
    b = tokens(O.to_tree(gen2))
    assert np.array_equal(a[:, :256], b[:, :256])
    assert not np.allclose(a[:, 256], b[:, 256])


def test_grouped_by_task_equals_per_sample_loop(params_p1):
    """Batched semantics (scripts/train.py:559-579): envs sharing a task == each env evaluated alone."""
    inp = S.make_inputs(5, 4, 2)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, _ = O.generate(params_p1, lang["token_embedding"], lang["attention_mask"], inp["initial_state"]["patch_embeddings"][:, 0],
                        generated_paths=M.generated_leaves_canonical())
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((4, 256, 768)).astype(np.float32)
    ti = inp["task_index"]
    h_all = O.base_vit_forward(O.to_tree(O.take_tasks(gen, ti)), emb)
    for b in range(4):
        h_b = O.base_vit_forward(O.to_tree(O.take_tasks(gen, ti[b:b + 1])), emb[b:b + 1])
        assert np.allclose(h_all[b], h_b[0], rtol=0, atol=2e-6)


def test_context_mask_blocks(params_p1):
    am = np.zeros((2, 32), np.int32); am[0, :4] = 1; am[1, :9] = 1
    m = O.context_mask(am, np.array([True, False]))
    assert m.shape == (2, 1, 34, 34)
    assert m[0, 0, :, :4].all() and not m[0, 0, :, 4:32].any()
    assert not m[1, 0, :, :32].any()                 # lang_pad False masks every language column
    assert m[:, 0, :, 32].all()                      # image column always visible
    assert m[:, 0, 33, 33].all() and not m[:, 0, :33, 33].any()


def test_mix_head_semantics():
    """tanh(x/5)*5 for 6 continuous dims x 4 steps; gripper = logit >= 0 (action_heads.py:464-470, 536)."""
    h = np.zeros((1, 64), np.float32); h[0, 0] = 1.0
    wc = np.zeros((1, 64, 24), np.float32); wc[0, 0] = np.arange(24) - 10.0
    wd = np.zeros((1, 64, 4), np.float32); wd[0, 0] = [-1.0, 0.0, 1e-4, 3.0]
    gen = {"action_head": {"continuous_head": {"kernel": wc, "bias": np.zeros((1, 24), np.float32)},
                           "discrete_head": {"kernel": wd, "bias": np.zeros((1, 4), np.float32)}}}
    act, logit = O.mix_head(gen, h)
    assert act.shape == (1, 4, 7)
    assert np.allclose(act[0, :, :6].ravel(), np.tanh((np.arange(24) - 10.0) / 5) * 5, atol=1e-6)
    assert act[0, :, 6].tolist() == [0.0, 1.0, 1.0, 1.0]


def test_postprocess_oracle_euler_matches_scipy_and_ensemble_rule():
    """The transforms3d.euler2axangle restatement is pinned against scipy (extrinsic xyz == 'sxyz'); the ensemble rule
    of data/utils/action_ensemble.py (older predictions, later horizon step, more weight) on a hand-checkable case."""
    from scipy.spatial.transform import Rotation as R
    from oracle.postprocess_oracle import EnvPostprocessor, euler2axangle
    rng = np.random.default_rng(0)
    for e in rng.uniform(-3, 3, (50, 3)):
        ax, ang = euler2axangle(*e)
        rv = R.from_euler("xyz", e).as_rotvec()
        # same rotation (axis-angle is unique up to the 2*pi wrap transforms3d does not apply)
        assert np.allclose(R.from_rotvec(ax * ang).as_matrix(), R.from_rotvec(rv).as_matrix(), atol=1e-12)
    stats = {"mean": np.zeros(7), "std": np.ones(7)}
    env = EnvPostprocessor("libero", "normal", stats, True, 0.0)
    a0 = np.arange(28, dtype=np.float32).reshape(4, 7)
    r0, _ = env.step(a0)
    assert np.allclose(r0, a0[0])
    r1, act = env.step(a0 + 100)
    assert np.allclose(r1, 0.5 * (a0[1] + (a0 + 100)[0]))          # oldest prediction contributes its step 1
    assert np.isclose(act[6], 2 * r1[6] - 1)


def test_position_table_resize_machinery_against_pil_bicubic():
    """The 37x37 -> 16x16 position-table interpolation restates jax.image.scale_and_translate(method='bicubic',
    antialias=False) (un-vendored FlaxDinov2Embeddings.interpolate_pos_encoding; no jax here).  Its machinery -- Keys cubic
    kernel a=-0.5, half-pixel sample positions, renormalisation over the in-range taps -- is pinned against an independent
    implementation: Pillow's BICUBIC uses the same kernel and the same border rule, and when UPsampling its support is not
    widened, i.e. it computes exactly the antialias=False formula.  (The down-scaling call of the path differs from this
    only through the scale argument; that the reference passes antialias=False and scale (16+0.1)/37 is taken from the
    transformers 4.50 source as recalled in SURVEY.md Appendix B and stays unpinned.)"""
    PIL = pytest.importorskip("PIL.Image")
    from hvla import params as P
    from oracle import hypervla_oracle as O
    rng = np.random.default_rng(0)
    for n_in, n_out in ((16, 37), (10, 23), (37, 37), (7, 8)):
        a = rng.standard_normal((n_in, n_in)).astype(np.float32)
        ref = np.asarray(PIL.fromarray(a, mode="F").resize((n_out, n_out), PIL.BICUBIC))
        w = P._resize_weights(n_in, n_out, np.float32(n_out / n_in))
        assert w.shape == (n_in, n_out) and np.allclose(w.sum(0), 1.0, atol=1e-6)
        got = np.einsum("hw,ha,wb->ab", a, w, w)
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    # the path's own call: product and oracle restatements agree bit for bit, rows of the weight matrix sum to one, and a
    # constant table stays constant (partition of unity under the +0.1 scale trick)
    pe = rng.standard_normal((1, 1370, 768)).astype(np.float32)
    assert np.array_equal(P.interpolate_pos_table(pe), O.interpolate_pos_table(pe))
    w = P._resize_weights(37, 16, np.float32(16.1 / 37))
    assert np.allclose(w.sum(0), 1.0, atol=1e-6) and (np.count_nonzero(w, axis=0) <= 4).all()
