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
