"""Generates tests/golden/ref_disc*.npz by EXECUTING THE REFERENCE'S OWN CODE for the DiscreteActionHead configuration
(base_net_kwargs.action_head_type = "discrete"; hypervla/components/base_network.py:22-33, 92-99, action_heads.py:252-396,
octo/model/components/tokenizers.py:235-275) in float64 through oracle/refshim, like make_ref_golden.py does for the README mix head.

    python tests/golden/make_ref_discrete_golden.py     # writes ref_disc4_b2_t2.npz (4 readout tokens), ref_disc28_b2_t2.npz (28)

Per case: generated rows (sampled), context embedding, per-env readout-token embeddings, vocab logits, argmax tokens and decoded actions
from model.hypernet.apply + base_net.apply(method=BaseNetwork.predict_action) per env (the body of scripts/train.py:559-577).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from hvla import config as C, metadata as M, params as P, synthetic as S   # noqa: E402
from make_ref_golden import example_batch, f64                             # noqa: E402

CASES = {"ref_disc4_b2_t2": ("action_horizon", 11, 2, 2), "ref_disc28_b2_t2": ("action_dim_and_action_horizon", 12, 2, 2)}
WEIGHT_STRIDE = 97


def discrete_config(token_type: str) -> dict:
    cfg = C.default_config()
    cfg["base_net_kwargs"]["action_head_type"] = "discrete"
    cfg["base_net_kwargs"]["action_head_kwargs"] = {"discrete_token_type": token_type}
    return cfg


def main():
    from oracle import refshim
    refshim.install()
    import jax
    import hypervla.model as RM
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (token_type, ci, B, T) in CASES.items():
        cfg = discrete_config(token_type)
        spec = M.HeadSpec.from_config(cfg)
        params = P.init_params(2025, "P1", spec)
        model = RM.HyperVLA.from_config(cfg, example_batch(), rng=jax.random.PRNGKey(0))
        ref_names = sorted(k for k in model.params if k.startswith("output_head_"))
        assert ref_names == sorted(k for k in params if k.startswith("output_head_")), "generated-leaf set differs from the reference's init_base_net"
        model = model.replace(params=f64(params))
        inp = S.make_inputs(ci, B, T)
        idict, istate = f64(inp["instruction_dict"]), f64(inp["initial_state"])
        tasks = {"pad_mask_dict": {"language_instruction": np.ones(T, bool)}, "language_instruction": idict["language_instruction"]}
        gen_T, ctx = model.hypernet.apply({"params": model.params}, tasks, train=False, initial_states=istate)
        rows = np.zeros((T, M.n_generated(spec)), np.float64)
        for path, (off, shape) in M.packed_offsets(spec).items():
            leaf = gen_T
            for k in path:
                leaf = leaf[k]
            rows[:, off:off + int(np.prod(shape))] = np.asarray(leaf).reshape(T, -1)

        def head_outputs(mdl, images, tok):
            emb, _ = mdl.encode(images, tok, train=False)
            return emb["readout_action"].tokens, mdl.action_head(emb, train=False)
        actions, tokens, logits, hs = [], [], [], []
        for i in range(B):
            bp_i = jax.tree_util.tree_map_with_path(
                lambda path, x: np.asarray(x)[0] if any("image_encoder" in str(p.key) for p in path) else np.asarray(x)[inp["task_index"][i]], gen_T)
            act = model.base_net.apply({"params": bp_i}, inp["images"][i], None, np.ones((1, 1), bool), train=False,
                                       rng=jax.random.PRNGKey(1), rngs={"dropout": jax.random.PRNGKey(1)}, method=RM.BaseNetwork.predict_action)
            h, lg = model.base_net.apply({"params": bp_i}, inp["images"][i], None, method=head_outputs)
            lg = np.asarray(lg, np.float64)[0, 0]                       # (horizon, action_dim, vocab)
            actions.append(np.asarray(act, np.float64)[0])
            logits.append(lg)
            tokens.append(lg.argmax(-1))
            hs.append(np.asarray(h, np.float64)[0, 0])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), config_index=ci, B=B, T=T, n_action_tokens=spec.n_action_tokens,
                            task_index=inp["task_index"], action=np.stack(actions), tokens=np.stack(tokens).astype(np.int32), logits=np.stack(logits),
                            h=np.stack(hs), ctx=np.asarray(ctx)[:, 0], rows_sample=rows[:, ::WEIGHT_STRIDE], rows_abs=np.abs(rows).sum(1),
                            source="reference code executed through oracle/refshim (float64)")
        print(name, "action[0,0]=", actions[0][0], "tokens[0,0]=", tokens[0][0])


if __name__ == "__main__":
    main()
