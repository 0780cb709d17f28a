"""CPU oracle of the action post-processing that follows the model call (TEST INFRASTRUCTURE ONLY).

NumPy restatement of, per environment,
  * un-normalisation          data/utils/hypervla_interface.py:219-242
  * temporal action ensemble  data/utils/action_ensemble.py:6-27 (BatchActionEnsembler; same rule as the
                              simpler_env ActionEnsembler the wrapper uses at :250-253)
  * euler -> axis-angle and gripper handling per policy setup   hypervla_interface.py:261-300
`transforms3d` (euler2axangle) is not installed here; its published algorithm (euler2quat 'sxyz' -> quat2axangle)
is restated in `euler2axangle`.  The reference ships no fixtures for this step: parity unpinned.
"""
from collections import deque

import numpy as np


def euler2axangle(ai, aj, ak):
    """transforms3d.euler.euler2axangle(ai, aj, ak, axes='sxyz') -> (axis(3), angle), float64."""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si, cj, sj, ck, sk = np.cos(ai), np.sin(ai), np.cos(aj), np.sin(aj), np.cos(ak), np.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    w, x, y, z = cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc
    eps = np.finfo(np.float64).eps
    Nq = w * w + x * x + y * y + z * z
    if Nq < eps:
        return np.array([1.0, 0, 0]), 0.0
    if Nq != 1:
        s = np.sqrt(Nq)
        w, x, y, z = w / s, x / s, y / s, z / s
    len2 = x * x + y * y + z * z
    if len2 < eps ** 2:
        return np.array([1.0, 0, 0]), 0.0
    theta = 2 * np.arccos(max(min(w, 1), -1))
    return np.array([x, y, z]) / np.sqrt(len2), theta


class EnvPostprocessor:
    """One environment's InferenceWrapper post-processing state (hypervla_interface.py:19-87, 141-162)."""

    def __init__(self, policy_setup, norm_type, stats, action_ensemble=True, temp=0.0, horizon=4):
        self.policy_setup, self.norm_type, self.stats = policy_setup, norm_type, stats
        self.action_ensemble, self.temp, self.horizon = action_ensemble, temp, horizon
        self.sticky_gripper_num_repeat = 15 if policy_setup == "google_robot" else 1
        self.reset()

    def reset(self):
        self.history = deque(maxlen=self.horizon)
        self.sticky_action_is_on, self.gripper_action_repeat = False, 0
        self.sticky_gripper_action, self.previous_gripper_action = 0.0, None

    def step(self, raw_actions):
        raw_actions = np.asarray(raw_actions, np.float32)                      # (4, 7)
        st = self.stats
        mask = np.asarray(st.get("mask", np.ones(7, bool)))
        if self.norm_type == "normal":
            raw_actions = np.where(mask, raw_actions * np.float32(1) * st["std"].astype(np.float32) + st["mean"].astype(np.float32), raw_actions)
        else:
            p01, p99 = st["p01"].astype(np.float32), st["p99"].astype(np.float32)
            raw_actions = np.where(mask, (raw_actions + 1) * (p99 - p01 + np.float32(1e-8)) / 2 + p01, raw_actions)
        raw_actions = raw_actions.astype(np.float32)
        if self.action_ensemble:
            self.history.append(raw_actions)
            num = len(self.history)
            preds = np.stack([pa[i] for i, pa in zip(range(num - 1, -1, -1), self.history)])
            w = np.exp(-self.temp * np.arange(num))
            w = w / w.sum()
            raw_action = np.sum(w[:, None] * preds, axis=0)
        else:
            raw_action = np.array(raw_actions[0])
        roll, pitch, yaw = np.asarray(raw_action[3:6], dtype=np.float64)
        ax, ang = euler2axangle(roll, pitch, yaw)
        rot = ax * ang
        if self.policy_setup == "google_robot":
            cur = float(raw_action[-1])
            rel = 0 if self.previous_gripper_action is None else self.previous_gripper_action - cur
            self.previous_gripper_action = cur
            if np.abs(rel) > 0.5 and self.sticky_action_is_on is False:
                self.sticky_action_is_on, self.sticky_gripper_action = True, rel
            if self.sticky_action_is_on:
                self.gripper_action_repeat += 1
                rel = self.sticky_gripper_action
            if self.gripper_action_repeat == self.sticky_gripper_num_repeat:
                self.sticky_action_is_on, self.gripper_action_repeat, self.sticky_gripper_action = False, 0, 0.0
            grip = rel
        elif self.policy_setup == "widowx_bridge":
            grip = 2.0 * (raw_action[-1] > 0.5) - 1.0
        else:
            grip = 2 * raw_action[-1] - 1
        action = np.concatenate([np.asarray(raw_action[:3], np.float32), rot.astype(np.float32), np.array([grip]).astype(np.float32)])
        return np.asarray(raw_action, np.float64), action
