#!/usr/bin/env python
"""GEMM chain (gemm_chain.cuh) against one launch per GEMM: the DINOv2 embeddings of the same images must be BIT-IDENTICAL
(the arithmetic per element is the same, only the schedule differs) and the chained forward is timed.
  python tools/chain_check.py [B ...]      (runs each setting in its own process: the switches are read once per process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(B, out):
    sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))
    import numpy as np
    import torch
    from hvla import params as P
    from hvla.runtime import Runtime
    rt = Runtime(P.init_params(2025, "P1"), precision="bf16", device="cuda:0")
    rng = np.random.default_rng(7)
    img = torch.from_numpy(rng.integers(0, 256, size=(B, 224, 224, 3), dtype=np.uint8)).cuda()
    emb = rt.dino_forward(img)
    torch.cuda.synchronize()
    ts = []
    for _ in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        emb = rt.dino_forward(img)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    e2 = rt.dino_forward(img)
    torch.cuda.synchronize()
    rep = bool((e2.view(torch.int16) == emb.view(torch.int16)).all())
    np.save(out, emb.view(torch.int16).cpu().numpy())
    if hasattr(rt.lib, "hvla_debug_chain_stats"):
        import ctypes
        buf = (ctypes.c_ulonglong * 8)()
        rt.lib.hvla_debug_chain_stats(buf)
        n = 14
        print(f"   chain stats per forward: units with a dependency {buf[0] / n:.0f}, had to wait {buf[1] / n:.0f}, spins {buf[2] / n:.0f}, "
              f"cycles waited {buf[3] / n:.0f} (sum over pairs; {buf[3] / n / 74 / 1.9e3:.1f} us per pair at 1.9 GHz)")
    print(f"B={B} chain={os.environ.get('HVLA_CHAIN', 'default')} lag={os.environ.get('HVLA_CHAIN_LAG', 'default')}: dino forward median {ts[len(ts) // 2]:.3f} ms, min {ts[0]:.3f} ms, "
          f"repeatable {rep}, finite {bool(torch.isfinite(emb.float()).all())}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    import numpy as np
    ok = True
    for B in [int(a) for a in sys.argv[1:]] or [64]:
        outs = []
        sweep = os.environ.get("CHAIN_SWEEP")
        cfgs = [("0", None), ("1", None)] + ([("1", c) for c in sweep.split(";")] if sweep else [("1", "0"), ("1", "3"), ("1", "1000")])
        for chain, lag in cfgs:
            env = dict(os.environ, HVLA_CHAIN=chain)
            if lag is not None:
                env["HVLA_CHAIN_LAG"] = lag
            out = f"/tmp/chain_{B}_{chain}_{lag}.npy"
            r = subprocess.run([sys.executable, __file__, "--child", str(B), out], env=env, timeout=300)
            if r.returncode != 0:
                print(f"B={B} chain={chain} lag={lag}: FAILED rc={r.returncode}")
                ok = False
                continue
            outs.append((chain, lag, np.load(out)))
        for chain, lag, o in outs[1:]:
            same = np.array_equal(outs[0][2], o)
            ok &= same
            print(f"B={B} chain={chain} lag={lag} vs unchained: {'bit-identical' if same else 'DIFFERENT: %d of %d' % ((outs[0][2] != o).sum(), o.size)}")
    sys.exit(0 if ok else 1)
