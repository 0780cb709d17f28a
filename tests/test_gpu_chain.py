"""GEMM chain (csrc/gemm_chain.cuh: proj -> fc1 -> fc2 -> q|k|v of the next layer in one persistent launch, dynamic unit scheduling,
dependency counters between row blocks) against one launch per GEMM.  The arithmetic per element is identical, so the DINOv2 hidden
states must be BIT-IDENTICAL whatever the schedule: sequential (default), fully interleaved (lag 0: every unit waits for the one
before it -- the stress test of the release / acquire protocol) and a short wavefront (lag 3)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [25, 64])
def test_gemm_chain_is_bit_identical_to_one_launch_per_gemm(batch):
    env = dict(os.environ, CHAIN_SWEEP="0;3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "chain_check.py"), str(batch)], env=env, capture_output=True, text=True,
                       timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert r.stdout.count("bit-identical") == 3 and "DIFFERENT" not in r.stdout and "FAILED" not in r.stdout
    assert "repeatable False" not in r.stdout
