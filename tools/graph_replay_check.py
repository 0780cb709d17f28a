import os, sys, time, numpy as np, torch
sys.path.insert(0, 'hyper-vla_b200')
from hvla import config as C, synthetic as S, _native as N
from hvla.model import HyperVLA
B = 64
m = HyperVLA.from_config(C.default_config(), precision='bf16', params_variant='P1')
rt = m.runtime
inp = S.make_inputs(2, B, B)
bp, tasks, _ = m.create_tasks(instruction_dict=inp['instruction_dict'], initial_state=inp['initial_state'])
img = torch.from_numpy(inp['images'][:, 0]).cuda()
W = bp.weights
for _ in range(5): rt.act_device(img, W, None)
torch.cuda.synchronize()
def timeit(fn, n=30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print('eager ms/step', timeit(lambda: rt.act_device(img, W, None)))
g = torch.cuda.CUDAGraph()
act = torch.empty((B, 4, 7), device='cuda'); logit = torch.empty((B, 4), device='cuda')
ws, wsb = rt.workspace(B, 0)
def launch():
    N.check(rt.lib.hvla_act(rt.stream(), rt.dino_vec.data_ptr(), rt.dino_mat.data_ptr(), img.data_ptr(), W.data_ptr(), None, B, B, act.data_ptr(), logit.data_ptr(), ws, wsb, rt.dtype), 'act')
with torch.cuda.graph(g):
    launch()
for _ in range(3): g.replay()
print('graph ms/step', timeit(lambda: g.replay()))
