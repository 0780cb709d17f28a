import sys
sys.path.insert(0, 'hyper-vla_b200')
import torch
from hvla import config as C, synthetic as S
from hvla.model import HyperVLA
m = HyperVLA.from_config(C.default_config(), precision='bf16', params_variant='P1')
inp = S.make_inputs(2, 64, 64)
for _ in range(3):
    bp, tasks, _ = m.create_tasks(instruction_dict=inp['instruction_dict'], initial_state=inp['initial_state'])
torch.cuda.synchronize()
