"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN CODE (/root/reference/hypervla/model.py and
hypervla/components/*.py, unmodified) in float64 through oracle/refshim (NumPy stand-ins for the jax / flax
primitives, HF torch DINOv2 for the un-vendored FlaxDinov2Model).  Needs /root/reference, so it runs in the build
container only; the fixtures it writes are committed and travel to the GPU box.

    python tests/golden/make_ref_golden.py [case ...]   # writes ref_c1_b1_t1.npz, ref_c2_b3_t3.npz, ref_c5_b6_t2.npz, ref_c2_b64_t64.npz

Call shapes used (all reference code):
  * model = HyperVLA.from_config(config, example_batch)                    hypervla/model.py:286-368
    (its params are then replaced by the seeded P1 tree hvla.params.init_params(2025, "P1"), same pytree)
  * B = T = 1: model.create_tasks(...) ; model.sample_actions(...)          hypervla/model.py:35-137
  * batched:   model.hypernet.apply({'params': p}, tasks, train=False, initial_states=...)
               jax.vmap(per_sample_predict_action)(...) with
               model.base_net.apply({'params': base_params_i}, image_i[None], ..., method=BaseNetwork.predict_action)
               — the body of scripts/train.py:546-579 (validation_action_loss), the reference's only batched
               generate-then-act.  With a task_index (config 5) env i takes the generated tree of task task_index[i],
               which is what the reference computes when the same task is repeated in the batch.
  * gripper logits / action-token embedding: base_net.apply(..., method=<lambda calling the reference's
    BaseNetwork.encode and MixActionHead.__call__>)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))

from hvla import config as C, metadata as M, params as P, synthetic as S   # noqa: E402

CASES = {"ref_c1_b1_t1": (1, 1, 1), "ref_c2_b3_t3": (2, 3, 3), "ref_c5_b6_t2": (5, 6, 2),
         # BASELINE configs[1] at its full size: 64 envs, one task each -- the batch the headline bench line runs (LayerNorm-free flow, GEMM chain)
         "ref_c2_b64_t64": (2, 64, 64)}
WEIGHT_STRIDE = 97


def f64(tree):
    import jax
    return jax.tree_map(lambda x: np.asarray(x, np.float64) if np.asarray(x).dtype == np.float32 else np.asarray(x), tree)


def example_batch():
    inp = S.make_inputs(1, 1, 1)
    return {"observation": {"image_primary": inp["images"], "timestep_pad_mask": np.ones((1, 1), bool)},
            "task": {"language_instruction": inp["instruction_dict"]["language_instruction"],
                     "pad_mask_dict": {"language_instruction": np.ones(1, bool)}},
            "initial_state": inp["initial_state"]}


def build_reference_model(params=None):
    """HyperVLA.from_config through the shim; returns (reference model with `params` installed, reference module)."""
    from oracle import refshim
    refshim.install()
    import jax
    import hypervla.model as RM
    model = RM.HyperVLA.from_config(C.default_config(), example_batch(), rng=jax.random.PRNGKey(0))
    if params is not None:
        model = model.replace(params=f64(params))
    return model, RM


def pack_rows(tree, T):
    rows = np.zeros((T, M.N_GENERATED), np.float64)
    for path, (off, shape) in M.packed_offsets().items():
        leaf = tree
        for k in path:
            leaf = leaf[k]
        n = int(np.prod(shape))
        rows[:, off:off + n] = np.asarray(leaf).reshape(T, n)
    return rows


def run_reference(model, RM, ci, B, T):
    import jax
    inp = S.make_inputs(ci, B, T)
    idict, istate = f64(inp["instruction_dict"]), f64(inp["initial_state"])
    if B == 1 and T == 1:
        base_params, tasks, _ = model.create_tasks(instruction_dict=idict, initial_state=istate)
        ctx = None
        action, _ = model.sample_actions(inp["images"], idict, tasks, inp["timestep_pad_mask"], base_params,
                                         rng=jax.random.PRNGKey(1))
        per_env = jax.tree_map(lambda x: np.asarray(x)[None], base_params)
        gen_T = per_env
    else:
        tasks = {"pad_mask_dict": {"language_instruction": np.ones(T, bool)},
                 "language_instruction": idict["language_instruction"]}
        gen_T, ctx = model.hypernet.apply({"params": model.params}, tasks, train=False, initial_states=istate)
        ti = inp["task_index"]
        # env i <- generated tree of task ti[i]; the shared DINOv2 leaves are identical for every task (broadcast views)
        per_env = jax.tree_util.tree_map_with_path(
            lambda path, x: np.broadcast_to(np.asarray(x)[0], (B,) + np.asarray(x).shape[1:])
            if any("image_encoder" in str(p.key) for p in path) else np.asarray(x)[ti], gen_T)
        batch = {"observation": {"image_primary": inp["images"], "timestep_pad_mask": np.ones((B, 1), bool)},
                 "task": {"language_instruction": jax.tree_map(lambda x: np.asarray(x)[ti], idict["language_instruction"])}}

        def per_sample_predict_action(base_params, sample_data, dropout_rng):          # scripts/train.py:559-577
            sample_data = jax.tree_map(lambda x: np.expand_dims(x, 0), sample_data)
            return model.base_net.apply(
                {"params": base_params}, sample_data["observation"]["image_primary"],
                sample_data["task"]["language_instruction"]["token_embedding"],
                sample_data["observation"]["timestep_pad_mask"], train=False, rng=dropout_rng,
                rngs={"dropout": dropout_rng}, method=RM.BaseNetwork.predict_action)
        action = jax.vmap(per_sample_predict_action, in_axes=(0, 0, 0))(per_env, batch, jax.random.split(jax.random.PRNGKey(1), B))
        action = np.asarray(action)[:, 0]

    # logits + action-token embedding through the reference's encode / MixActionHead.__call__
    def head_outputs(mdl, images, tok):
        emb, _ = mdl.encode(images, tok, train=False)
        cont, logits = mdl.action_head(emb, train=False)
        return emb["readout_action"].tokens, cont, logits
    hs, logits = [], []
    for i in range(B):
        bp_i = jax.tree_map(lambda x: np.asarray(x)[i], per_env)
        tokens, _, lg = model.base_net.apply({"params": bp_i}, inp["images"][i], None, method=head_outputs)
        hs.append(np.asarray(tokens).reshape(-1))
        logits.append(np.asarray(lg).reshape(-1))
    rows = pack_rows({k: v for k, v in gen_T.items()}, T)
    if ctx is None:
        _, ctx = model.hypernet.apply({"params": model.params}, tasks, train=False, initial_states=istate)
    return dict(rows=rows, ctx=np.asarray(ctx)[:, 0], action=np.asarray(action, np.float64), logit=np.stack(logits),
                h=np.stack(hs), task_index=inp["task_index"])


def save_case(path, ci, B, T, r):
    np.savez_compressed(
        path, config_index=ci, B=B, T=T, task_index=r["task_index"], action=r["action"], logit=r["logit"], ctx=r["ctx"],
        h=r["h"], rows_sample=r["rows"][:, ::WEIGHT_STRIDE], rows_sum=r["rows"].sum(1), rows_abs=np.abs(r["rows"]).sum(1),
        source="reference code executed through oracle/refshim (float64)")


def main():
    params = P.init_params(2025, "P1")
    model, RM = build_reference_model(params)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (ci, B, T) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        r = run_reference(model, RM, ci, B, T)
        save_case(os.path.join(out_dir, name + ".npz"), ci, B, T, r)
        print(name, "action[0,0]=", r["action"][0, 0], "logit[0]=", r["logit"][0])


if __name__ == "__main__":
    main()
