"""CPU-only checks of bench.py's host logic: both arms describe the same workload with the same config dict (the driver
compares them), the whole-step algorithmic-byte formula is SURVEY section 8d's, and the 3-camera rig is three views."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import helpers as Hh  # noqa: E402,F401  (path set-up)


def test_config_is_identical_in_both_arms_and_names_the_workload():
    wl = bench.WORKLOADS["kitti-375x1242-1M"]
    a = bench.make_config("kitti-375x1242-1M", wl, 1, 1)
    b = bench.make_config("kitti-375x1242-1M", wl, 1, 1)
    assert a == b and a["workload"] == "kitti-375x1242-1M" and a["image"] == [375, 1242] and a["gaussians"] == 1_000_000
    assert "model" not in a
    assert bench.make_config("kitti-375x1242-1M", wl, 8, 1)["views_per_step"] == 8
    rig = bench.WORKLOADS["waymo-3cam-1066x1600-3M"]
    assert bench.make_config("waymo-3cam-1066x1600-3M", rig, 1, 3)["views_per_step"] == 3


def test_algorithmic_bytes_is_the_survey_formula():
    # B_alg = N_s*1220 + N_o*2790 + R*404 + Px*84 per view; the verdict's recomputation for configs[1]
    total, stages = bench.algorithmic_bytes(750_000, 250_000, 2_413_910, 375 * 1242)
    assert total == 750_000 * 1220 + 250_000 * 2790 + 2_413_910 * 404 + 465_750 * 84 == 2_626_842_640
    assert stages["blend_backward"] == 2_413_910 * 172 + 465_750 * 44


def test_rig_workload_is_three_cameras_sharing_one_timestep():
    wl = bench.WORKLOADS["waymo-3cam-1066x1600-3M"]
    views = bench.views_of_step(wl, 0, "cpu")
    assert len(views) == 3
    times = {t for _, t, _ in views}
    assert len(times) == 1
    # the three optical axes are 45 degrees apart (scene/dataset_readers.py:261-357: front-left / front / front-right)
    axes = [np.asarray(cam.world_view_transform)[:3, 2] for cam, _, _ in views]
    for a, b in ((axes[0], axes[1]), (axes[1], axes[2])):
        ang = math.degrees(math.acos(float(np.clip(np.dot(a, b), -1, 1))))
        assert abs(ang - 45.0) < 1e-3
    assert len(bench.views_of_step(bench.WORKLOADS["kitti-375x1242-1M"], 3, "cpu")) == 1


def test_elementwise_metric_catches_what_the_max_norm_hides():
    import torch
    b = torch.tensor([1.0, 1e-3, 1e-3])
    a = torch.tensor([1.0, 1e-3, 2e-3])             # a 100 % error on a small entry
    assert abs(Hh.rel_err(a, b) - 1e-3) < 1e-9      # "1e-3 relative" in the max norm ...
    frac, worst = Hh.elementwise_err(a, b, rtol=1e-4, atol_frac=1e-6)
    assert abs(frac - 1 / 3) < 1e-12 and worst > 500   # ... but one element in three is off by ~900x its bound
    assert Hh.elementwise_err(b, b)[0] == 0.0
