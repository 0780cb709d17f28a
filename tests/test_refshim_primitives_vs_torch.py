"""The flax primitives restated in oracle/refshim/flaxlite.py (used to execute the reference's own code) against the
independent torch implementations of the same published definitions: Dense, LayerNorm (eps 1e-6), tanh-GELU,
multi-head dot-product attention with a boolean mask, plus the Module scoping / auto-naming rules the reference's
parameter pytree depends on."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def F():
    from oracle.refshim import flaxlite
    return flaxlite


def test_dense_layernorm_gelu_match_torch(F):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 7, 64))
    k, b = rng.standard_normal((64, 24)), rng.standard_normal(24)
    y = F.Dense(24).apply({"params": {"kernel": k, "bias": b}}, x)
    ref = torch.nn.functional.linear(torch.from_numpy(x), torch.from_numpy(k.T.copy()), torch.from_numpy(b)).numpy()
    assert np.abs(np.asarray(y) - ref).max() < 1e-12
    sc, bi = rng.standard_normal(64), rng.standard_normal(64)
    y = F.LayerNorm().apply({"params": {"scale": sc, "bias": bi}}, x)
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (64,), torch.from_numpy(sc), torch.from_numpy(bi), eps=1e-6).numpy()
    assert np.abs(np.asarray(y) - ref).max() < 1e-10          # fast variance E[x^2]-E[x]^2 vs two-pass: rounding only
    ref = torch.nn.functional.gelu(torch.from_numpy(x), approximate="tanh").numpy()
    assert np.abs(np.asarray(F.gelu(x)) - ref).max() < 1e-12
    ref = torch.nn.functional.gelu(torch.from_numpy(x)).numpy()
    assert np.abs(np.asarray(F.gelu(x, approximate=False)) - ref).max() < 1e-12


def test_multi_head_attention_matches_torch(F):
    rng = np.random.default_rng(1)
    B, S, D, H = 2, 9, 64, 4
    hd = D // H
    x = rng.standard_normal((B, S, D))
    p = {n: {"kernel": rng.standard_normal((D, H, hd)) * 0.2, "bias": rng.standard_normal((H, hd)) * 0.1} for n in ("query", "key", "value")}
    p["out"] = {"kernel": rng.standard_normal((H, hd, D)) * 0.2, "bias": rng.standard_normal(D) * 0.1}
    mask = rng.uniform(size=(B, 1, S, S)) > 0.3
    mask[..., 0] = True                                          # no fully masked row
    y = F.MultiHeadDotProductAttention(num_heads=H).apply({"params": p}, x, x, mask=mask)
    mha = torch.nn.MultiheadAttention(D, H, batch_first=True, dtype=torch.float64)
    with torch.no_grad():
        w = np.concatenate([p[n]["kernel"].reshape(D, D).T for n in ("query", "key", "value")], 0)
        mha.in_proj_weight.copy_(torch.from_numpy(w))
        mha.in_proj_bias.copy_(torch.from_numpy(np.concatenate([p[n]["bias"].reshape(D) for n in ("query", "key", "value")])))
        mha.out_proj.weight.copy_(torch.from_numpy(p["out"]["kernel"].reshape(D, D).T.copy()))
        mha.out_proj.bias.copy_(torch.from_numpy(p["out"]["bias"]))
        am = torch.from_numpy(~np.repeat(mask, H, axis=1).reshape(B * H, S, S))       # torch: True = masked out
        ref, _ = mha(torch.from_numpy(x), torch.from_numpy(x), torch.from_numpy(x), attn_mask=am, need_weights=False)
    assert np.abs(np.asarray(y) - ref.numpy()).max() < 1e-10


def test_module_scoping_and_auto_names(F):
    """setup() children are named by attribute (dicts: attr_key), compact children Class_N, explicit names win; params are
    looked up by that path -- the rules behind 'output_head_<leaf>' / 'encoderblock_i' / 'Dense_0' in the reference pytree."""
    class Inner(F.Module):
        feats: int = 3

        @F.compact
        def __call__(self, x):
            x = F.Dense(self.feats)(x)
            return F.Dense(self.feats, name="last")(F.Dense(self.feats)(x))

    class Outer(F.Module):
        def setup(self):
            self.proj = F.Dense(4)
            self.named = F.Dense(4, name="explicit_name")
            self.heads = {"a_b": F.Dense(2), "c": F.Dense(2)}
            self.inner = Inner()

        def __call__(self, x):
            return self.heads["a_b"](self.named(self.proj(x))), self.inner(x)

    variables = Outer().init(np.zeros(2, np.uint32), np.ones((1, 5)))
    names = variables["params"]
    assert set(names) == {"proj", "explicit_name", "heads_a_b", "inner"}          # heads_c is never called: no parameters
    assert set(names["inner"]) == {"Dense_0", "Dense_1", "last"}
    assert names["proj"]["kernel"].shape == (5, 4) and names["heads_a_b"]["kernel"].shape == (4, 2)
    out1 = Outer().apply(variables, np.ones((1, 5)))
    out2 = Outer().apply(variables, np.ones((1, 5)))
    assert np.array_equal(out1[0], out2[0]) and np.array_equal(out1[1], out2[1])
    with pytest.raises(KeyError):
        Outer().apply({"params": {}}, np.ones((1, 5)))
