"""Generates tests/golden/refwrap_*.npz by EXECUTING THE REFERENCE'S OWN `InferenceWrapper.step`
(/root/reference/data/utils/hypervla_interface.py:164-304, unmodified) through oracle/refshim, with a stand-in model
that returns canned raw actions.  Pins the code either side of the model call:
un-normalisation, temporal ensembling, euler->axis-angle, the gripper rules of the three policy setups incl. the sticky
google_robot gripper, and the order resize -> crop -> round/clip.

Third-party pieces the wrapper imports that are absent here get stand-ins (everything else is the reference's code):
  * simpler_env ActionEnsembler      -> the reference's own data/utils/action_ensemble.py:BatchActionEnsembler on a batch of 1
  * transforms3d.euler.euler2axangle -> oracle/postprocess_oracle.py:euler2axangle (pinned against scipy in tests)
  * tensorflow image ops             -> oracle/preprocess_oracle.py (the TF kernels restated; still unpinned)

    python tests/golden/make_ref_wrapper_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-vla_b200"))

CASES = [("google_robot", "normal", False), ("widowx_bridge", "bounds", True), ("libero", "normal", False)]
STEPS, RESET_AT, TEMP = 24, 11, 0.0


def stats_and_actions(seed=42):
    rng = np.random.default_rng(seed)
    stats = {"mean": rng.normal(0, 0.1, 7), "std": rng.uniform(0.05, 0.5, 7), "p01": rng.uniform(-1, -0.2, 7),
             "p99": rng.uniform(0.2, 1, 7), "mask": np.array([1, 1, 1, 1, 1, 1, 0], bool)}
    raw = rng.uniform(-3, 3, (STEPS, 4, 7)).astype(np.float32)
    raw[..., 6] = (rng.uniform(size=(STEPS, 4)) > 0.5)
    frames = rng.integers(0, 256, (2, 120, 160, 3), dtype=np.uint8)
    return stats, raw, frames


class _TFArr(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def _install_wrapper_standins():
    from oracle import postprocess_oracle as PP, preprocess_oracle as PO, refshim
    refshim.install()
    from data.utils.action_ensemble import BatchActionEnsembler         # the reference's own ensembler

    class ActionEnsembler:                                               # simpler_env's single-env interface
        def __init__(self, pred_action_horizon, action_ensemble_temp=0.0):
            self._b = BatchActionEnsembler(pred_action_horizon, action_ensemble_temp)

        def reset(self):
            self._b.reset()

        def ensemble_action(self, cur_action):
            return self._b.ensemble_action(np.asarray(cur_action)[None])[0]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    mod("transforms3d"); mod("transforms3d.euler", euler2axangle=PP.euler2axangle)
    for n in ("simpler_env", "simpler_env.utils", "simpler_env.utils.action"):
        mod(n)
    mod("simpler_env.utils.action.action_ensemble", ActionEnsembler=ActionEnsembler)

    import tensorflow as tf                                              # the refshim stub module
    arr = lambda x: np.asarray(x, np.float32).view(_TFArr)

    def resize(image, size, method, antialias):
        assert method == "lanczos3" and antialias and size[0] == size[1]
        return arr(PO.resize_lanczos3(np.asarray(image), size[0]))

    def crop_and_resize(images, boxes, box_indices, crop_size):
        b = np.asarray(boxes, np.float64)[0]
        assert images.shape[0] == 1 and abs(b[2] - b[0] - np.sqrt(0.9)) < 1e-12
        return arr(PO.crop_and_resize_center(np.asarray(images)[0], crop_size[0])[None])
    tf.image = types.SimpleNamespace(resize=resize, crop_and_resize=crop_and_resize)
    tf.range = lambda n: np.arange(n)
    tf.round = lambda x: arr(np.rint(np.asarray(x)))
    tf.clip_by_value = lambda x, lo, hi: arr(np.clip(np.asarray(x), lo, hi))
    tf.uint8 = np.uint8
    tf.cast = lambda x, dt: np.asarray(x).astype(dt).view(_TFArr)


class FakeModel:
    """Stands in for HyperVLA: hands back canned (1,4,7) raw actions and records the frames it was given."""

    def __init__(self, stats, norm, raw):
        self.dataset_statistics = {"action": stats}
        self.config = {"dataset_kwargs": {"dataset_kwargs": {"action_proprio_normalization_type": norm}}}
        self.raw, self.t, self.seen = raw, 0, []

    def create_tasks(self, instruction_dict=None, initial_state=None):
        return None, None, None

    def sample_actions(self, images, instruction_dict, task, pad_mask, base_params, rng=None, image_embeddings=None):
        self.seen.append(np.asarray(images)[0, 0].copy())
        out = self.raw[self.t][None]
        self.t += 1
        return out, {}


def run_reference_wrapper(policy, norm, crop):
    _install_wrapper_standins()
    from data.utils.hypervla_interface import InferenceWrapper
    stats, raw, frames = stats_and_actions()
    model = FakeModel(stats, norm, raw)
    w = InferenceWrapper(model=model, policy_setup=policy, horizon=1, pred_action_horizon=4, exec_horizon=1, image_size=224,
                         action_ensemble=True, crop=crop)
    w.action_ensemble_temp = TEMP
    w.reset("task a", instruction_dict={})
    raws, acts = [], []
    for t in range(STEPS):
        if t == RESET_AT:
            w.reset("task b", instruction_dict={})
        raw_action, action, image, _, _ = w.step(frames[t % 2])
        raws.append(np.asarray(raw_action, np.float64))
        acts.append(np.asarray(action, np.float64))
    return dict(raw_action=np.stack(raws), action=np.stack(acts), frames_out=np.stack(model.seen[:2]))


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for policy, norm, crop in CASES:
        r = run_reference_wrapper(policy, norm, crop)
        np.savez_compressed(os.path.join(out_dir, f"refwrap_{policy}_{norm}.npz"), policy=policy, norm=norm, crop=crop,
                            steps=STEPS, reset_at=RESET_AT, temp=TEMP, **r,
                            source="reference InferenceWrapper.step executed through oracle/refshim")
        print(policy, norm, "crop" if crop else "", r["action"][0], r["action"][-1])


if __name__ == "__main__":
    main()
