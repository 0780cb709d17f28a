"""Batch-1 act step for an ncu launch list (eager launches so every kernel is visible):
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b1_launches.csv python tools/b1_profile.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, 'hyper-vla_b200')
from hvla import config as C, synthetic as S
from hvla.model import HyperVLA
B = int(os.environ.get("B", "1"))
m = HyperVLA.from_config(C.default_config(), precision='bf16', params_variant='P1')
rt = m.runtime
inp = S.make_inputs(2, B, B)
bp, tasks, _ = m.create_tasks(instruction_dict=inp['instruction_dict'], initial_state=inp['initial_state'])
img = torch.from_numpy(inp['images'][:, 0]).cuda()
for _ in range(int(os.environ.get("N", "4"))):
    rt.act_device(img, bp.weights, None)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lat = []
for _ in range(50):
    a.record(); rt.act_device(img, bp.weights, None); b.record(); torch.cuda.synchronize(); lat.append(a.elapsed_time(b))
print("B", B, "p50 ms", float(np.median(lat)))
