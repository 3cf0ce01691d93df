"""Generates tests/golden/densify.npz by importing the REFERENCE's own scene/gaussian_model.py from
/root/reference (so it only runs in the build container) and calling, on CPU,
GaussianModel.training_setup / add_densification_stats / densify_and_prune / reset_opacity on a seeded
synthetic model whose Adam moments were populated by two real optimizer steps.

Non-invasive shims (the reference file itself is unmodified):
  * modules that are not installed and not used on this path -- roma, plyfile, simple_knn._C,
    pytorch3d.ops -- are replaced by empty stand-ins;
  * the reference hard-codes device="cuda" in torch.zeros / torch.ones (...) calls: the factory functions
    are wrapped to map "cuda" to "cpu" while the generator runs;
  * torch.normal(mean=0.0, std=stds) is wrapped to draw z = randn and return z * stds (which is how ATen
    implements it) so that the unit normals z can be stored in the fixture: the CUDA path and the oracle
    are fed the same z.
Usage: python tests/golden/make_densify_golden.py
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _shim(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_shim("roma", unitquat_slerp=None, unitquat_to_rotvec=None, rotvec_to_unitquat=None, quat_conjugation=None,
      quat_product=None)
_shim("plyfile", PlyData=None, PlyElement=None)
_shim("simple_knn")
_shim("simple_knn._C", distCUDA2=None)
_shim("pytorch3d")
_shim("pytorch3d.ops", knn_points=None)



def _on_cpu(fn):
    def wrapped(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


for _name in ("zeros", "ones", "full", "tensor", "rand", "randperm", "empty"):
    setattr(torch, _name, _on_cpu(getattr(torch, _name)))
recorded_z = []
_gen = torch.Generator().manual_seed(1234)


def normal_recorded(mean=0.0, std=None, **k):
    assert mean == 0.0 and torch.is_tensor(std)
    z = torch.randn(std.shape, generator=_gen)
    recorded_z.append(z.clone())
    return z * std


torch.normal = normal_recorded
sys.path.insert(0, REF)
import importlib.util  # noqa: E402

# scene/__init__.py pulls in the dataset readers (open3d, ...): load scene/gaussian_model.py as a file instead
_spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)          # the reference's file, unmodified
GaussianModel = _mod.GaussianModel
from torch import nn  # noqa: E402

ORDER_ARGS = {"xyz": [6, 3, 0, 2, 0, 0], "rotation": [0, 0, 0, 0, 5, 2], "shs": [0, 0, 0, 2, 0, 0],
              "background": [6, 3, 0, 2, 0, 0]}
PARAMS = ("_scene_xyz", "_scene_shs_dc", "_scene_shs_rest", "_scene_opacity", "_scene_scaling", "_scene_rotation",
          "_obj_xyz", "_obj_shs_dc", "_obj_shs_rest", "_obj_opacity", "_obj_scaling", "_obj_rotation",
          "xyz_deform_param", "rotation_deform_param", "shs_deform_param_scene", "shs_deform_param_obj",
          "background_deform_param", "gs_time_sigma")
GROUP_OF = {"_scene_xyz": "scene_xyz", "_scene_shs_dc": "scene_shs_dc", "_scene_shs_rest": "scene_shs_rest",
            "_scene_opacity": "scene_opacity", "_scene_scaling": "scene_scaling", "_scene_rotation": "scene_rotation",
            "_obj_xyz": "obj_xyz", "_obj_shs_dc": "obj_shs_dc", "_obj_shs_rest": "obj_shs_rest",
            "_obj_opacity": "obj_opacity", "_obj_scaling": "obj_scaling", "_obj_rotation": "obj_rotation",
            "xyz_deform_param": "deform_xyz", "rotation_deform_param": "deform_rotation",
            "shs_deform_param_scene": "deform_shs_scene", "shs_deform_param_obj": "deform_shs_obj",
            "background_deform_param": "deform_background", "gs_time_sigma": "time_sigma"}


def training_args():
    return SimpleNamespace(
        percent_dense=0.01, object_extent=5.0, min_camera_extent=5.0, feature_lr=0.0025, opacity_lr=0.05,
        scaling_lr=0.005, rotation_lr=0.001, rotation_deform_lr=0.001, shs_deform_lr=0.0025, gs_time_sigma_lr=1e-2,
        position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=60000,
        obj_position_lr_scale=0.8, scene_position_lr_scale=1.0, position_deform_lr_scale=0.2,
        lambda_reg=0.0, lambda_sigma=0.0, lambda_sigma_reg=0.0, near_num=8)


def make_model(ns, no, seed):
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.rand(*s, generator=g)
    Nn = lambda *s: torch.randn(*s, generator=g)
    m = GaussianModel(3, ORDER_ARGS)
    m.scene_extent, m.cameras_extent, m.object_extent, m.frame_gap, m.use_time_mask = 20.0, 8.0, 10.0, 0.02, True
    P = lambda t: nn.Parameter(t.contiguous().requires_grad_(True))
    cx, cr, cs = 6 + 4, 5, 4

    def part(n, split_size):
        # log-scales straddling the clone / split size and the "big point" sizes
        sc = torch.log(split_size * torch.exp(1.6 * Nn(n, 3)))
        return (P(Nn(n, 3) * 5), P(Nn(n, 1, 3)), P(Nn(n, 15, 3) * 0.1), P(Nn(n, 1) * 3.0 - 2.0), P(sc), P(Nn(n, 4)))

    (m._scene_xyz, m._scene_shs_dc, m._scene_shs_rest, m._scene_opacity, m._scene_scaling,
     m._scene_rotation) = part(ns, 20.0 * 0.01)
    (m._obj_xyz, m._obj_shs_dc, m._obj_shs_rest, m._obj_opacity, m._obj_scaling, m._obj_rotation) = part(no, 5.0 * 0.01)
    m.xyz_deform_param = P(Nn(no, 3, cx) * 1e-2)
    m.rotation_deform_param = P(Nn(no, 4, cr) * 1e-2)
    m.shs_deform_param_scene = P(Nn(ns, 3, cs) * 1e-2)
    m.shs_deform_param_obj = P(Nn(no, 3, cs) * 1e-2)
    m.background_deform_param = P(Nn(1, 3, cx) * 1e-2)
    m.gs_time = R(no, 1)
    m.gs_time_sigma = P(torch.full((no, 2), float(np.log(0.02))) + 0.1 * Nn(no, 2))
    m.max_radii2D = torch.zeros((ns + no,))
    m.training_setup(training_args())
    for grp in m.optimizer.param_groups:      # non-zero learning rates for the position groups too
        if grp["lr"] == 0.0:
            grp["lr"] = 1e-3
    for _ in range(2):                        # real Adam steps so that every moment is populated
        for name in PARAMS:
            p = getattr(m, name)
            p.grad = Nn(*p.shape) * 1e-2
        m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    return m, g


def snapshot(m, prefix, out):
    for name in PARAMS:
        p = getattr(m, name)
        out[f"{prefix}{name.lstrip('_')}"] = p.detach().numpy().copy()
        st = m.optimizer.state.get(p)
        if st is not None:
            out[f"{prefix}{name.lstrip('_')}.exp_avg"] = st["exp_avg"].numpy().copy()
            out[f"{prefix}{name.lstrip('_')}.exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}gs_time"] = m.gs_time.numpy().copy()
    out[f"{prefix}xyz_gradient_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}denom"] = m.denom.numpy().copy()
    out[f"{prefix}max_radii2D"] = m.max_radii2D.numpy().copy()


def case(tag, ns, no, seed, prune_big, out):
    m, g = make_model(ns, no, seed)
    n = ns + no
    # ---- add_densification_stats + the max_radii2D update of train.py:151 -----------------------------
    for it in range(3):
        radii = (torch.rand(n, generator=g) * 40 - 8).to(torch.int32).clamp(min=0)
        vsp = SimpleNamespace(grad=torch.randn(n, 3, generator=g) * 4e-4)
        vis = radii > 0
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis])
        m.add_densification_stats({"viewspace_points": vsp, "visibility_filter": vis})
        out[f"{tag}.stats{it}.radii"] = radii.numpy().copy()
        out[f"{tag}.stats{it}.grad"] = vsp.grad.numpy().copy()
    snapshot(m, f"{tag}.before.", out)
    recorded_z.clear()
    m.densify_and_prune(0.0002, 0.0002, 0.005, prune_big)
    out[f"{tag}.z_scene"], out[f"{tag}.z_obj"] = recorded_z[0].numpy().copy(), recorded_z[1].numpy().copy()
    snapshot(m, f"{tag}.after.", out)
    # ---- reset_opacity on the densified model ----------------------------------------------------------
    m.reset_opacity()
    snapshot(m, f"{tag}.reset.", out)
    out[f"{tag}.prune_big"] = np.array(int(prune_big))
    out[f"{tag}.extents"] = np.array([m.scene_extent, m.object_extent, m.percent_dense], dtype=np.float64)


if __name__ == "__main__":
    data = {}
    case("a", 300, 120, 1, False, data)
    case("b", 257, 255, 2, True, data)
    case("c", 40, 0, 3, True, data)
    np.savez_compressed(os.path.join(HERE, "densify.npz"), **data)
    for t in "abc":
        print(t, data[f"{t}.before.scene_xyz"].shape, data[f"{t}.before.obj_xyz"].shape, "->",
              data[f"{t}.after.scene_xyz"].shape, data[f"{t}.after.obj_xyz"].shape,
              data[f"{t}.z_scene"].shape, data[f"{t}.z_obj"].shape)
