"""Generates the golden fixtures under tests/golden/ from the fp64 oracle.

The reference ships no golden vectors and cannot be imported here (no jax/flax),
so these vectors come from the NumPy restatement in oracle/hypervla_oracle.py evaluated in
float64 on seeded synthetic parameters (hvla.params.init_params(2025, "P1")) and inputs
(hvla.synthetic.make_inputs).  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))

from hvla import metadata as M, params as P, synthetic as S   # noqa: E402
from oracle import hypervla_oracle as O                         # noqa: E402

# name -> (config_index, B, T)
CASES = {"c1_b1_t1": (1, 1, 1), "c2_b3_t3": (2, 3, 3), "c5_b6_t2": (5, 6, 2)}
WEIGHT_STRIDE = 97


def pack_rows(gen_flat, T):
    rows = np.zeros((T, M.N_GENERATED), np.float64)
    for path, (off, shape) in M.packed_offsets().items():
        n = int(np.prod(shape))
        rows[:, off:off + n] = gen_flat[path].reshape(T, n)
    return rows


def run_case(params, dino, ci, B, T, dtype):
    inp = S.make_inputs(ci, B, T)
    lang = inp["instruction_dict"]["language_instruction"]
    gen, emb = O.generate(params, lang["token_embedding"], lang["attention_mask"],
                          inp["initial_state"]["patch_embeddings"][:, 0], dtype=dtype,
                          generated_paths=M.generated_leaves_canonical())
    per_env = O.to_tree(O.take_tasks(gen, inp["task_index"]))
    act, logit, hidden, h = O.sample_actions(dino, per_env, inp["images"][:, 0], dtype=dtype, return_all=True)
    return dict(rows=pack_rows(gen, T), ctx=emb[:, 0], action=act, logit=logit, hidden=hidden, h=h,
                task_index=inp["task_index"])


def main():
    params = P.init_params(2025, "P1")
    dino = P.dino_tree_from_params(params)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (ci, B, T) in CASES.items():
        r = run_case(params, dino, ci, B, T, np.float64)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            config_index=ci, B=B, T=T, task_index=r["task_index"],
            action=r["action"].astype(np.float32), logit=r["logit"].astype(np.float64), ctx=r["ctx"].astype(np.float64),
            h=r["h"].astype(np.float64),
            rows_sample=r["rows"][:, ::WEIGHT_STRIDE].astype(np.float64), rows_sum=r["rows"].sum(1), rows_abs=np.abs(r["rows"]).sum(1),
            hidden_sample=r["hidden"][:, ::16, ::48].astype(np.float64), hidden_abs_mean=np.abs(r["hidden"]).mean((1, 2)))
        print(name, "action[0,0]=", r["action"][0, 0], "logit[0]=", r["logit"][0])


if __name__ == "__main__":
    main()
