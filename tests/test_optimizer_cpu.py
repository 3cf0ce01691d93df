"""CPU-only checks of the optimizer host logic (adgs_b200/optimizer.py): group names and order of
scene/gaussian_model.py:346-370, the learning-rate schedule of utils/general_utils.py:29-62, and
argument validation of adgs_adam_step without a GPU."""
import ctypes as C
import math

import numpy as np

from adgs_b200 import _lib as L
from adgs_b200.optimizer import GROUP_NAMES, get_expon_lr_func


def test_group_names_follow_the_reference_order():
    assert GROUP_NAMES[:6] == ("scene_xyz", "scene_shs_dc", "scene_shs_rest", "scene_opacity", "scene_scaling",
                               "scene_rotation")
    assert GROUP_NAMES[6:12] == ("obj_xyz", "obj_shs_dc", "obj_shs_rest", "obj_opacity", "obj_scaling", "obj_rotation")
    assert GROUP_NAMES[12:] == ("deform_rotation", "deform_shs_scene", "deform_shs_obj", "deform_xyz",
                                "deform_background", "time_sigma")
    assert len(GROUP_NAMES) == 18


def test_expon_lr_schedule():
    f = get_expon_lr_func(lr_init=1.6e-4, lr_final=1.6e-6, lr_delay_mult=0.01, max_steps=30_000)
    assert math.isclose(f(0), 1.6e-4, rel_tol=1e-12)
    assert math.isclose(f(30_000), 1.6e-6, rel_tol=1e-12)
    assert math.isclose(f(15_000), 1.6e-5, rel_tol=1e-9)           # log-linear midpoint
    assert math.isclose(f(10**9), 1.6e-6, rel_tol=1e-12)           # clipped
    assert f(-1) == 0.0
    assert get_expon_lr_func(0.0, 0.0)(10) == 0.0                  # disabled parameter
    g = get_expon_lr_func(1e-2, 1e-4, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=1000)
    assert math.isclose(g(0), 0.1 * 1e-2, rel_tol=1e-12)           # warm-up starts at lr_init * mult
    exp = (0.1 + 0.9 * np.sin(0.5 * np.pi * 0.5)) * np.exp(np.log(1e-2) * 0.95 + np.log(1e-4) * 0.05)
    assert math.isclose(g(50), exp, rel_tol=1e-12)


def test_adam_step_rejects_bad_arguments_without_a_gpu():
    lib = L.load()
    seg = (L.AdamSegment * 1)()
    assert C.sizeof(L.AdamSegment) == 4 * 8 + 3 * 8 + 16 + 16 + 8
    assert lib.adgs_adam_step(seg, L.ADAM_MAX_SEGMENTS + 1, 0.9, 0.999, 1e-15, 1, None) == -1
    assert lib.adgs_adam_step(seg, 1, 0.9, 0.999, 1e-15, 0, None) == -1       # step counts from 1
    assert lib.adgs_adam_step(seg, 1, 1.0, 0.999, 1e-15, 1, None) == -1       # beta1 must be < 1
    assert lib.adgs_adam_step(None, 1, 0.9, 0.999, 1e-15, 1, None) == -1
    seg[0].n = 16                                                             # pointers missing
    assert lib.adgs_adam_step(seg, 1, 0.9, 0.999, 1e-15, 1, None) == -1
    seg[0].n = 0                                                              # empty segments are skipped: no launch
    assert lib.adgs_adam_step(seg, 1, 0.9, 0.999, 1e-15, 1, None) == 0


def _cpu_model(ns=5, no=3):
    import torch
    from adgs_b200.gaussian_model import GaussianModel
    oa = {"xyz": [6, 3, 0, 2, 0, 0], "rotation": [0, 0, 0, 0, 5, 2], "shs": [0, 0, 0, 2, 0, 0], "background": [6, 3, 0, 2, 0, 0]}
    g = torch.Generator().manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g)
    ref = dict(scene_xyz=R(ns, 3), obj_xyz=R(no, 3), scene_shs_dc=R(ns, 1, 3), obj_shs_dc=R(no, 1, 3),
               scene_shs_rest=R(ns, 15, 3), obj_shs_rest=R(no, 15, 3), scene_opacity=R(ns, 1), obj_opacity=R(no, 1),
               scene_scaling=R(ns, 3), obj_scaling=R(no, 3), scene_rotation=R(ns, 4), obj_rotation=R(no, 4),
               xyz_deform_param=R(no, 3, 10), rotation_deform_param=R(no, 4, 5), shs_deform_param_scene=R(ns, 3, 4),
               shs_deform_param_obj=R(no, 3, 4), background_deform_param=R(1, 3, 10), gs_time=torch.rand(no, 1),
               gs_time_sigma=R(no, 2))
    return GaussianModel.from_reference(ref, oa, device="cpu")


def test_segments_skip_arrays_without_gradient_like_torch_adam():
    """torch.optim.Adam skips parameters whose .grad is None (the iteration of a densification: every
    per-Gaussian tensor is new and only deform_background still has its gradient) and counts steps per parameter;
    FusedAdam builds segments only for arrays that have a gradient and keeps one step count per array."""
    import torch
    from adgs_b200.gaussian_model import PARAM_NAMES
    from adgs_b200.optimizer import FusedAdam
    m = _cpu_model()
    opt = FusedAdam(m, {n: 1e-3 for n in GROUP_NAMES})
    assert opt._segments() == [] and opt.step_count == 0
    m.background_deform.grad = torch.zeros_like(m.background_deform)
    segs = opt._segments()
    assert [k for k, _ in segs] == ["background_deform"] and segs[0][1].n == m.background_deform.numel()
    for k in PARAM_NAMES:
        getattr(m, k).grad = torch.zeros_like(getattr(m, k))
    assert [k for k, _ in opt._segments()] == list(PARAM_NAMES)
    # the xyz segment carries both learning rates and the scene / object split
    xyz = dict(opt._segments())["xyz"]
    assert xyz.lr_rule == L.ADAM_LR_SPLIT and xyz.split == 3 * m.n_scene
    opt.step_counts["background_deform"] = 7
    assert opt.step_count == 7
    opt.step_count = 3
    assert set(opt.step_counts.values()) == {3}


def test_shared_arrays_need_equal_learning_rates():
    import pytest
    import torch
    from adgs_b200.gaussian_model import PARAM_NAMES
    from adgs_b200.optimizer import FusedAdam
    m = _cpu_model()
    opt = FusedAdam(m, {n: 1e-3 for n in GROUP_NAMES})
    for k in PARAM_NAMES:
        getattr(m, k).grad = torch.zeros_like(getattr(m, k))
    for g in opt.param_groups:
        if g["name"] == "obj_scaling":
            g["lr"] = 2e-3
    with pytest.raises(ValueError):
        opt._segments()
