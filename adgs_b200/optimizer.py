"""Optimizer step of the hot path's parameters.

Mirror of what `GaussianModel.training_setup` / `update_learning_rate` build and `train.py:163-167`
steps in the reference (scene/gaussian_model.py:337-411): a `torch.optim.Adam(l, lr=0.0, eps=1e-15)`
over 18 named parameter groups with three exponentially decayed position learning rates
(utils/general_utils.py:29-62). Here the 18 groups live in the 10 planar arrays of
`adgs_b200.gaussian_model.GaussianModel`, the moments have the same planar layout, and ONE kernel
launch (`adgs_adam_step`, adgs_b200/csrc/adam.cu) performs the whole step.

`param_groups` keeps the reference's names and order, so code that walks
`optimizer.param_groups` and sets `group['lr']` by `group['name']` (update_learning_rate) works
unchanged. Moments can be exported / imported in the reference's tensor layout
(`state_in_reference_layout`, `load_reference_state`).

No fallback: the step raises if the CUDA library is missing.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .gaussian_model import PARAM_NAMES, get_param_num

# reference group order (scene/gaussian_model.py:346-370)
GROUP_NAMES = (
    "scene_xyz", "scene_shs_dc", "scene_shs_rest", "scene_opacity", "scene_scaling", "scene_rotation",
    "obj_xyz", "obj_shs_dc", "obj_shs_rest", "obj_opacity", "obj_scaling", "obj_rotation",
    "deform_rotation", "deform_shs_scene", "deform_shs_obj", "deform_xyz", "deform_background", "time_sigma",
)


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear learning-rate decay with an optional sine warm-up, utils/general_utils.py:29-62."""

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return delay_rate * log_lerp

    return helper


MAX_WINDOW_COLUMNS = 128   # bits of adgs_adam_segment.active


class FusedAdam:
    """Adam over a planar GaussianModel with the reference's 18 parameter groups.

    lrs: {group name: learning rate}; missing names start at 0.0 like the reference's position groups.
    window_aware=True: the control-point arrays' gradients are only read for the columns the
    backward of this step wrote (`model.active_columns()`), all others count as zero -- identical
    results to dense Adam on zero-filled gradients, without the fill and its read.
    """

    def __init__(self, model, lrs=None, betas=(0.9, 0.999), eps=1e-15, window_aware=False):
        self.model = model
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.window_aware = bool(window_aware)
        # torch.optim.Adam keeps one step count per parameter and skips parameters without a gradient (the
        # iteration after a densification: every per-Gaussian tensor is new, only deform_background steps)
        self.step_counts = {k: 0 for k in PARAM_NAMES}
        lrs = dict(lrs or {})
        unknown = set(lrs) - set(GROUP_NAMES)
        if unknown:
            raise ValueError(f"unknown parameter group(s): {sorted(unknown)}")
        self.param_groups = [{"name": n, "lr": float(lrs.get(n, 0.0)), "params": self._group_views(n),
                              "betas": self.betas, "eps": self.eps} for n in GROUP_NAMES]
        self.state = {k: {"exp_avg": torch.zeros_like(getattr(model, k)),
                          "exp_avg_sq": torch.zeros_like(getattr(model, k))} for k in PARAM_NAMES}

    # ---- reference-shaped views of the planar arrays (introspection only; the kernel uses the arrays) ----
    def _group_views(self, name):
        m = self.model
        ns = m.n_scene
        rows = (lambda t: t[:ns]) if name.startswith("scene_") or name.endswith("_scene") else (lambda t: t[ns:])
        if name in ("scene_xyz", "obj_xyz"):
            return [rows(m.xyz)]
        if name in ("scene_opacity", "obj_opacity"):
            return [rows(m.opacity)]
        if name in ("scene_scaling", "obj_scaling"):
            return [rows(m.scaling)]
        if name in ("scene_rotation", "obj_rotation"):
            return [rows(m.rotation)]
        if name in ("scene_shs_dc", "obj_shs_dc"):
            return [rows(m.sh4[0])[:, :3]]
        if name in ("scene_shs_rest", "obj_shs_rest"):
            return [rows(m.sh4[0])[:, 3:], m.sh4[1:, :ns] if name.startswith("scene_") else m.sh4[1:, ns:]]
        if name in ("deform_shs_scene", "deform_shs_obj"):
            return [m.shs_deform4[:, :ns] if name.endswith("_scene") else m.shs_deform4[:, ns:]]
        return [{"deform_rotation": m.rot_deform, "deform_xyz": m.xyz_deform, "deform_background": m.background_deform,
                 "time_sigma": m.gs_time_sigma}[name]]

    def _lr(self, name):
        for g in self.param_groups:
            if g["name"] == name:
                return float(g["lr"])
        raise KeyError(name)

    def _check_shared(self, a, b):
        if self._lr(a) != self._lr(b):
            raise ValueError(f"groups {a} and {b} share one planar array and must keep the same learning rate "
                             f"(the reference gives them the same value, scene/gaussian_model.py:346-370)")
        return self._lr(a)

    def window_eligible(self):
        """Can the step take 'zero without reading' for inactive control-point planes? The column mask of
        adgs_adam_segment has MAX_WINDOW_COLUMNS bits; any plane size works (the kernel decides per element where a
        float4 straddles two planes). The ONE place this is decided: GaussianModel.sparse_deform_grads asks here, so
        the render backward zero-fills exactly when the step is going to read everything."""
        m = self.model
        return bool(self.window_aware and m.xyz_deform.shape[0] <= MAX_WINDOW_COLUMNS and
                    m.rot_deform.shape[0] <= MAX_WINDOW_COLUMNS)

    @property
    def step_count(self):
        return max(self.step_counts.values())

    @step_count.setter
    def step_count(self, value):
        self.step_counts = {k: int(value) for k in PARAM_NAMES}

    def _segments(self):
        """[(array name, adgs_adam_segment)] of the arrays that have a gradient."""
        m = self.model
        n, ns, no = m.n_scene + m.n_obj, m.n_scene, m.n_obj
        need_active = self.window_eligible() and (m.xyz_deform.grad is not None or m.rot_deform.grad is not None)
        active = m.active_columns() if need_active else None
        segs = []

        def add(key, lr_a, lr_b=0.0, rule=L.ADAM_LR_UNIFORM, split=0, plane=0, cols=None):
            p = getattr(m, key)
            if p.numel() == 0:
                return
            if p.grad is None:      # torch.optim.Adam skips it
                return
            st = self.state[key]
            if not (p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError(f"FusedAdam.step(): {key} and its gradient must be contiguous")
            s = L.AdamSegment(param=p.data_ptr(), grad=p.grad.data_ptr(), exp_avg=st["exp_avg"].data_ptr(),
                              exp_avg_sq=st["exp_avg_sq"].data_ptr(), n=p.numel(), split=split, plane=0,
                              lr_a=lr_a, lr_b=lr_b, lr_rule=rule)
            if cols is not None and plane > 0:
                # eligibility was decided by window_eligible() when the backward ran (sparse_deform_grads): here the
                # planes outside `cols` are UNWRITTEN memory, so there is no dense fallback to take
                assert p.numel() // plane <= MAX_WINDOW_COLUMNS
                s.plane = plane
                bits = [0, 0]
                for c in cols:
                    bits[c >> 6] |= 1 << (c & 63)
                s.active[0], s.active[1] = bits
            segs.append((key, s))

        add("xyz", self._lr("scene_xyz"), self._lr("obj_xyz"), L.ADAM_LR_SPLIT, split=3 * ns)
        add("scaling", self._check_shared("scene_scaling", "obj_scaling"))
        add("rotation", self._check_shared("scene_rotation", "obj_rotation"))
        add("opacity", self._check_shared("scene_opacity", "obj_opacity"))
        add("sh4", self._check_shared("scene_shs_dc", "obj_shs_dc"), self._check_shared("scene_shs_rest", "obj_shs_rest"),
            L.ADAM_LR_SH4, split=4 * n)
        add("shs_deform4", self._check_shared("deform_shs_scene", "deform_shs_obj"))
        add("xyz_deform", self._lr("deform_xyz"), plane=3 * no, cols=None if active is None else active["xyz"])
        add("rot_deform", self._lr("deform_rotation"), plane=4 * no, cols=None if active is None else active["rotation"])
        add("background_deform", self._lr("deform_background"))
        add("gs_time_sigma", self._lr("time_sigma"))
        return segs

    @torch.no_grad()
    def step(self):
        lib = L.load()
        segs = self._segments()
        if not segs:
            return
        dev = self.model.xyz.device
        by_step = {}
        for key, s in segs:
            self.step_counts[key] += 1
            by_step.setdefault(self.step_counts[key], []).append(s)
        # one launch; a second one only while deform_background is ahead of the per-Gaussian arrays
        for step, group in by_step.items():
            arr = (L.AdamSegment * len(group))(*group)
            with torch.cuda.device(dev):
                st = lib.adgs_adam_step(arr, len(group), self.betas[0], self.betas[1], self.eps, step,
                                        torch.cuda.current_stream(dev).cuda_stream)
            L.check(st, "adam_step")
        if self.window_aware:
            self.model.reset_active_columns()   # (also counts the backwards of the step when not eligible)

    def zero_grad(self, set_to_none=True):
        for k in PARAM_NAMES:
            p = getattr(self.model, k)
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    # ---- moments in the reference's tensor layout (checkpoint interop, parity tests) ------------------
    def state_in_reference_layout(self):
        """{reference tensor name: {"exp_avg", "exp_avg_sq"}} for the 18 groups' tensors."""
        out = {}
        for which in ("exp_avg", "exp_avg_sq"):
            ref = self.model.reference_layout({k: self.state[k][which] for k in PARAM_NAMES})
            for name, t in ref.items():
                out.setdefault(name, {})[which] = t
        return out

    def load_reference_state(self, ref_state, step):
        """Inverse of state_in_reference_layout; `step` = the reference optimizer's step count."""
        for which in ("exp_avg", "exp_avg_sq"):
            planar = self.model.planar_layout({name: d[which] for name, d in ref_state.items()})
            for k in PARAM_NAMES:
                self.state[k][which].copy_(planar[k])
        self.step_count = int(step)   # all arrays


def training_setup(model, training_args, window_aware=False):
    """GaussianModel.training_setup (scene/gaussian_model.py:337-400): builds the optimizer with the
    reference's learning rates and the three position schedulers; returns the FusedAdam."""
    model.percent_dense = training_args.percent_dense
    model.object_extent = (training_args.object_extent if getattr(training_args, "object_extent", None) is not None
                           else getattr(model, "object_extent", 10.0))
    model.cameras_extent = max(getattr(model, "cameras_extent", 0.0), training_args.min_camera_extent)
    model.scene_extent = getattr(model, "scene_extent", 0.0)
    a = training_args
    lrs = {
        "scene_xyz": 0.0, "obj_xyz": 0.0, "deform_xyz": 0.0, "deform_background": 0.0,
        "scene_shs_dc": a.feature_lr, "obj_shs_dc": a.feature_lr,
        "scene_shs_rest": a.feature_lr / 20.0, "obj_shs_rest": a.feature_lr / 20.0,
        "scene_opacity": a.opacity_lr, "obj_opacity": a.opacity_lr,
        "scene_scaling": a.scaling_lr, "obj_scaling": a.scaling_lr,
        "scene_rotation": a.rotation_lr, "obj_rotation": a.rotation_lr,
        "deform_rotation": a.rotation_deform_lr, "deform_shs_scene": a.shs_deform_lr, "deform_shs_obj": a.shs_deform_lr,
        "time_sigma": a.gs_time_sigma_lr,
    }
    model.optimizer = FusedAdam(model, lrs, eps=1e-15, window_aware=window_aware)
    mk = lambda extent, scale: get_expon_lr_func(
        lr_init=a.position_lr_init * extent * scale, lr_final=a.position_lr_final * extent * scale,
        lr_delay_mult=a.position_lr_delay_mult, max_steps=a.position_lr_max_steps)
    model.obj_xyz_scheduler_args = mk(model.object_extent, a.obj_position_lr_scale)
    model.scene_xyz_scheduler_args = mk(model.cameras_extent, a.scene_position_lr_scale)
    model.deform_scheduler_args = mk(model.scene_extent, a.position_deform_lr_scale)
    # densification statistics and the near-index table (gaussian_model.py:340-341, 395-397)
    from .densify import set_obj_near_idx, setup_statistics
    setup_statistics(model)
    lam = lambda k: float(getattr(a, k, 0.0) or 0.0)
    if model.use_time_mask is None:     # scene/gaussian_model.py:394
        model.use_time_mask = lam("lambda_sigma") > 0.0
        model.__dict__.pop("_basis_cache", None)
    model.use_near_idx = lam("lambda_reg") > 0.0 or (lam("lambda_sigma") > 0.0 and lam("lambda_sigma_reg") > 0.0)
    model.near_num = int(getattr(a, "near_num", 0) or 0)
    if model.use_near_idx and model.xyz.is_cuda:
        set_obj_near_idx(model)
    return model.optimizer


def update_learning_rate(model, iteration):
    """GaussianModel.update_learning_rate (scene/gaussian_model.py:402-413)."""
    for group in model.optimizer.param_groups:
        if group["name"] in ("scene_xyz", "deform_background"):
            group["lr"] = model.scene_xyz_scheduler_args(iteration)
        elif group["name"] == "obj_xyz":
            group["lr"] = model.obj_xyz_scheduler_args(iteration)
        elif group["name"] == "deform_xyz":
            group["lr"] = model.deform_scheduler_args(iteration)
