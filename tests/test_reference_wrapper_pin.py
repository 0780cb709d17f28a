"""Pins the code either side of the model call to the REFERENCE'S OWN
`InferenceWrapper.step` (data/utils/hypervla_interface.py:164-304), executed unmodified through oracle/refshim by
tests/golden/make_ref_wrapper_golden.py with canned model outputs -> committed tests/golden/refwrap_*.npz."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("google_robot", "normal", False), ("widowx_bridge", "bounds", True), ("libero", "normal", False)]
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/data/utils"), reason="reference checkout not present")


def _gen():
    spec = importlib.util.spec_from_file_location("make_ref_wrapper_golden", os.path.join(HERE, "golden", "make_ref_wrapper_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _fixture(policy, norm):
    return np.load(os.path.join(HERE, "golden", f"refwrap_{policy}_{norm}.npz"))


@pytest.mark.parametrize("policy,norm,crop", CASES)
def test_postprocess_oracle_matches_reference_wrapper(policy, norm, crop):
    from oracle.postprocess_oracle import EnvPostprocessor
    gen, g = _gen(), _fixture(policy, norm)
    stats, raw, _ = gen.stats_and_actions()
    env = EnvPostprocessor(policy, norm, stats, True, float(g["temp"]))
    for t in range(int(g["steps"])):
        if t == int(g["reset_at"]):
            env.reset()
        r_raw, r_act = env.step(raw[t])
        assert np.abs(r_raw - g["raw_action"][t]).max() < 1e-6, t
        assert np.abs(r_act - g["action"][t]).max() < 1e-6, t
        if policy != "libero":
            assert r_act[6] == g["action"][t][6], t                      # sticky / binarised gripper: exact


@pytest.mark.parametrize("policy,norm,crop", CASES)
def test_preprocess_oracle_matches_reference_wrapper_frames(policy, norm, crop):
    from oracle import preprocess_oracle as PO
    gen, g = _gen(), _fixture(policy, norm)
    _, _, frames = gen.stats_and_actions()
    assert bool(g["crop"]) == crop
    for i in range(2):
        assert np.array_equal(PO.resize_image(frames[i], 224, crop=crop), g["frames_out"][i])


@needs_reference
def test_reference_wrapper_rerun_reproduces_fixture():
    gen = _gen()
    policy, norm, crop = CASES[0]
    r, g = gen.run_reference_wrapper(policy, norm, crop), _fixture(policy, norm)
    assert np.array_equal(r["action"], g["action"]) and np.array_equal(r["raw_action"], g["raw_action"])
    assert np.array_equal(r["frames_out"], g["frames_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("policy,norm,crop", CASES)
def test_gpu_pre_and_postprocessing_match_reference_wrapper(policy, norm, crop):
    import torch
    assert torch.cuda.is_available()
    from hvla.postprocess import BatchedActionPostprocessor
    from hvla.preprocess import BatchedImagePreprocessor
    gen, g = _gen(), _fixture(policy, norm)
    stats, raw, frames = gen.stats_and_actions()
    out = BatchedImagePreprocessor(224, crop=crop)(frames).cpu().numpy()
    assert np.array_equal(out, g["frames_out"])
    B = 3                                                                    # three identical envs, one launch per step
    pp = BatchedActionPostprocessor(B, policy, norm, stats, action_ensemble=True, action_ensemble_temp=float(g["temp"]))
    for t in range(int(g["steps"])):
        if t == int(g["reset_at"]):
            pp.reset()
        g_raw, g_act = pp.step(np.repeat(raw[t][None], B, 0))
        g_raw, g_act = g_raw.cpu().numpy(), g_act.cpu().numpy()
        for e in range(B):
            assert np.abs(g_raw[e] - g["raw_action"][t]).max() < 2e-6, (t, e)
            assert np.abs(g_act[e] - g["action"][t]).max() < 2e-6, (t, e)
