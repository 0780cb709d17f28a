"""Batched, GPU-resident action post-processing.

Mirrors, for B environments at once, what ``InferenceWrapper.step`` does on the host for one environment after the
model call (data/utils/hypervla_interface.py:219-300): un-normalise with the dataset statistics, temporally ensemble
the last ``pred_action_horizon`` predictions (data/utils/action_ensemble.py:6-27), convert the euler rotation to
axis-angle and post-process the gripper per ``policy_setup``.  State (prediction history, sticky gripper) lives on the
device; the work is one kernel launch (``hvla_postprocess``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

POLICIES = {"google_robot": 0, "widowx_bridge": 1, "libero": 2}
NORMS = {"normal": 0, "bounds": 1}


class BatchedActionPostprocessor:
    def __init__(self, num_envs: int, policy_setup: str, normalization_type: str, statistics: dict,
                 action_ensemble: bool = True, action_ensemble_temp: float = 0.0, device=None):
        import torch
        if policy_setup not in POLICIES:
            raise ValueError(f"Unknown policy setup: {policy_setup}")             # hypervla_interface.py:53-54
        if normalization_type not in NORMS:
            raise ValueError(f"Unknown normalization type: {normalization_type}")  # :241-242
        if not torch.cuda.is_available():
            raise N.HvlaError("no CUDA device: hvla post-processing is CUDA-only")
        self.lib = N.lib()
        self.B = int(num_envs)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.policy, self.norm = POLICIES[policy_setup], NORMS[normalization_type]
        self.sticky_repeat = 15 if policy_setup == "google_robot" else 1           # :46-51
        if self.norm == 0:
            a, b = statistics["std"], statistics["mean"]
        else:
            a, b = statistics["p01"], statistics["p99"]
        self._a = np.ascontiguousarray(np.asarray(a, np.float32)[:7])
        self._b = np.ascontiguousarray(np.asarray(b, np.float32)[:7])
        self._mask = np.ascontiguousarray(np.asarray(statistics.get("mask", np.ones(7, bool)), np.uint8)[:7])
        self.ensemble, self.temp = int(bool(action_ensemble)), float(action_ensemble_temp)
        self.state = torch.zeros((self.B, int(self.lib.hvla_postprocess_state_floats())), dtype=torch.float32, device=self.device)
        self._pending_reset = torch.ones((self.B,), dtype=torch.uint8, device=self.device)

    def reset(self, env_mask=None) -> None:
        """Start new episodes (InferenceWrapper.reset, :141-162) for all envs or the envs selected by ``env_mask``."""
        import torch
        if env_mask is None:
            self._pending_reset.fill_(1)
        else:
            self._pending_reset |= torch.as_tensor(np.asarray(env_mask)).to(self.device).to(torch.uint8)

    def step(self, raw_actions):
        """raw_actions (B,4,7) float32 (CUDA tensor or numpy) -> (raw_action (B,7), action (B,7)) CUDA tensors."""
        import torch
        ra = raw_actions if torch.is_tensor(raw_actions) else torch.from_numpy(np.ascontiguousarray(raw_actions, np.float32))
        ra = ra.to(self.device, torch.float32).contiguous()
        if tuple(ra.shape) != (self.B, 4, 7):
            raise ValueError(f"raw_actions must be ({self.B}, 4, 7)")              # :249
        out_raw = torch.empty((self.B, 7), dtype=torch.float32, device=self.device)
        out_act = torch.empty((self.B, 7), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):       # launches go to the current device's context
            st = self.lib.hvla_postprocess(int(torch.cuda.current_stream(self.device).cuda_stream), ra.data_ptr(), self.state.data_ptr(),
                                           self._pending_reset.data_ptr(), self.B, self.norm,
                                           self._a.ctypes.data_as(C.c_void_p), self._b.ctypes.data_as(C.c_void_p),
                                           self._mask.ctypes.data_as(C.c_void_p), self.ensemble, C.c_float(self.temp), self.policy,
                                           self.sticky_repeat, out_raw.data_ptr(), out_act.data_ptr())
        N.check(st, "hvla_postprocess")
        self._pending_reset.zero_()
        return out_raw, out_act
