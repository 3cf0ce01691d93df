"""GPU parity of the fused optimizer step (adgs_b200/optimizer.py -> adgs_adam_step) against
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` over the reference's 18 parameter groups in the
REFERENCE's tensor layout (scene/gaussian_model.py:346-372, stepped by train.py:163-167).
torch.optim.Adam is what the reference calls, so it is the oracle here; tolerance 2e-6 relative
(fp32: a few ulp of rounding-order difference per step), and bit-exact between the dense and the
window-aware mode."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from adgs_b200 import scenes
from adgs_b200.gaussian_model import GaussianModel, PARAM_NAMES
from adgs_b200.gaussian_renderer import render
from adgs_b200.optimizer import FusedAdam, GROUP_NAMES

pytestmark = pytest.mark.gpu

# reference tensor name of every group (scene/gaussian_model.py:346-370)
GROUP_TENSOR = {
    "scene_xyz": "scene_xyz", "scene_shs_dc": "scene_shs_dc", "scene_shs_rest": "scene_shs_rest",
    "scene_opacity": "scene_opacity", "scene_scaling": "scene_scaling", "scene_rotation": "scene_rotation",
    "obj_xyz": "obj_xyz", "obj_shs_dc": "obj_shs_dc", "obj_shs_rest": "obj_shs_rest", "obj_opacity": "obj_opacity",
    "obj_scaling": "obj_scaling", "obj_rotation": "obj_rotation", "deform_rotation": "rotation_deform_param",
    "deform_shs_scene": "shs_deform_param_scene", "deform_shs_obj": "shs_deform_param_obj",
    "deform_xyz": "xyz_deform_param", "deform_background": "background_deform_param", "time_sigma": "gs_time_sigma",
}
LRS = {  # all different where the layouts allow it, so that every per-element rule is exercised
    "scene_xyz": 1.6e-4, "obj_xyz": 1.28e-3, "scene_shs_dc": 2.5e-3, "obj_shs_dc": 2.5e-3,
    "scene_shs_rest": 1.25e-4, "obj_shs_rest": 1.25e-4, "scene_opacity": 5e-2, "obj_opacity": 5e-2,
    "scene_scaling": 5e-3, "obj_scaling": 5e-3, "scene_rotation": 1e-3, "obj_rotation": 1e-3,
    "deform_rotation": 1.1e-3, "deform_shs_scene": 2.4e-3, "deform_shs_obj": 2.4e-3, "deform_xyz": 3.2e-4,
    "deform_background": 0.0, "time_sigma": 1e-2,
}


def _model(n_scene, n_obj, seed=0):
    cam = scenes.make_camera(160, 96, 90.0, device="cuda")
    cloud = scenes.random_cloud(n_scene + n_obj, cam, seed=seed, median_radius_px=4.0)
    tensors = scenes.random_model_tensors(n_scene, n_obj, scenes.BENCH_ORDER_ARGS, cloud, seed=seed + 1, device="cuda")
    return GaussianModel.from_reference(tensors, scenes.BENCH_ORDER_ARGS), tensors, cam


@pytest.mark.parametrize("n_scene,n_obj", [(1501, 777), (64, 0), (0, 130)])
def test_fused_adam_matches_torch_adam_on_reference_groups(n_scene, n_obj):
    model, tensors, _ = _model(n_scene, n_obj)
    ref = {k: torch.nn.Parameter(v.detach().clone()) for k, v in tensors.items() if k != "gs_time"}
    groups = [{"params": [ref[GROUP_TENSOR[n]]], "lr": LRS[n], "name": n} for n in GROUP_NAMES]
    ref_opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    opt = FusedAdam(model, LRS, eps=1e-15)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for it in range(4):
        if it == 2:   # learning rates change between steps (update_learning_rate)
            for g in ref_opt.param_groups + opt.param_groups:
                if g["name"] in ("scene_xyz", "obj_xyz", "deform_xyz"):
                    g["lr"] = g["lr"] * 0.5
        gref = {k: torch.randn(p.shape, generator=gen, device="cuda") * (10.0 ** float(it - 2)) for k, p in ref.items()}
        if it == 3:   # sparse gradients: untouched elements still move with their momentum
            gref = {k: g * (torch.rand(g.shape, generator=gen, device="cuda") < 0.3) for k, g in gref.items()}
        for k, p in ref.items():
            p.grad = gref[k]
        planar = model.planar_layout(gref)
        for k in PARAM_NAMES:
            getattr(model, k).grad = planar[k]
        ref_opt.step()
        opt.step()
    torch.cuda.synchronize()
    ours = model.to_reference()
    for k, p in ref.items():
        if p.numel() == 0:
            continue
        err = (ours[k] - p.detach()).abs().max().item() / max(p.detach().abs().max().item(), 1e-12)
        assert err <= 2e-6, f"{k}: parameters differ from torch.optim.Adam by {err:.3e}"
    # moments, in the reference layout
    st = opt.state_in_reference_layout()
    for g in ref_opt.param_groups:
        p = g["params"][0]
        if p.numel() == 0:
            continue
        name = GROUP_TENSOR[g["name"]]
        for which in ("exp_avg", "exp_avg_sq"):
            a, b = st[name][which], ref_opt.state[p][which]
            err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
            assert err <= 2e-6, f"{name}.{which}: {err:.3e}"
    assert opt.step_count == 4


def test_reference_state_round_trip():
    model, tensors, _ = _model(300, 200)
    opt = FusedAdam(model, LRS)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for k in PARAM_NAMES:
        opt.state[k]["exp_avg"].copy_(torch.randn(getattr(model, k).shape, generator=gen, device="cuda"))
        opt.state[k]["exp_avg_sq"].copy_(torch.rand(getattr(model, k).shape, generator=gen, device="cuda"))
    # padding lanes of the float4 SH-deform planes do not exist in the reference layout
    before = {k: {w: t.clone() for w, t in opt.state[k].items()} for k in PARAM_NAMES}
    opt2 = FusedAdam(model, LRS)
    opt2.load_reference_state(opt.state_in_reference_layout(), step=7)
    assert opt2.step_count == 7
    for k in PARAM_NAMES:
        if k == "shs_deform4":
            continue
        for w in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(opt2.state[k][w], before[k][w]), (k, w)


def _train_steps(window_aware, steps=3, n_obj=2000):
    model, _, cam = _model(4000, n_obj, seed=3)
    args = SimpleNamespace(percent_dense=0.01, object_extent=10.0, min_camera_extent=10.0, feature_lr=0.0025,
                           opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, rotation_deform_lr=0.001,
                           shs_deform_lr=0.0025, gs_time_sigma_lr=1e-2, position_lr_init=0.00016,
                           position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=60_000,
                           position_deform_lr_scale=0.2, obj_position_lr_scale=0.8, scene_position_lr_scale=1.0)
    model.scene_extent = 20.0
    opt = model.training_setup(args, window_aware=window_aware)
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=False)
    gen = torch.Generator(device="cuda").manual_seed(11)
    cots = [torch.randn(c, 96, 160, generator=gen, device="cuda") for c in (3, 1, 1, 3, 1)]
    for it in range(1, steps + 1):
        model.update_learning_rate(it)
        # poison the caching allocator's free blocks: torch.empty() then hands out NaN-filled memory, so a read of a
        # gradient plane that the sparse backward left unwritten turns the parameters NaN instead of passing by luck
        poison = torch.full((96 * 1024 * 1024,), float("nan"), device="cuda")
        del poison
        t = 0.2 + 0.25 * it            # a different B-spline window every step
        vcam = SimpleNamespace(image_height=96, image_width=160, FoVx=cam.FoVx, FoVy=cam.FoVy,
                               world_view_transform=cam.world_view_transform,
                               full_proj_transform=cam.full_proj_transform, camera_center=cam.camera_center, time=t)
        res = render(vcam, model, None, pipe, flow_pkg=[t + 0.04, None, None, None, None, None], render_objmask=True)
        outs = (res["render"], res["depth"], res["img_opacity"], res["img_flow"], res["img_semantic"])
        torch.autograd.backward(outs, (cots[0], cots[1][0], cots[2][0], cots[3], cots[4]))
        opt.step()
        opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    return model, opt


@pytest.mark.parametrize("n_obj", [600, 601, 602, 603])   # plane = 3 * n_obj: a float4 may straddle two planes
def test_window_aware_step_ignores_inactive_planes_bit_exactly(n_obj):
    """Same gradients, two optimizers: dense reads zero-filled planes; window-aware is given NaN in every
    inactive control-point plane and must never read them. Results must be bit-identical -- for ANY number of
    object Gaussians (after densification / pruning it is never a multiple of 4)."""
    t, flow_t = 0.43, 0.47
    a, _, _ = _model(900, n_obj, seed=4)
    b, _, _ = _model(900, n_obj, seed=4)
    tb = a.time_basis(t, flow_t)
    b._note_active_columns(b.time_basis(t, flow_t))
    act = b.active_columns()
    assert 0 < len(act["xyz"]) < a.xyz_deform.shape[0] and 0 < len(act["rotation"]) < a.rot_deform.shape[0]
    oa, ob = FusedAdam(a, LRS), FusedAdam(b, LRS, window_aware=True)
    gen = torch.Generator(device="cuda").manual_seed(9)
    for it in range(3):
        for k in PARAM_NAMES:
            g = torch.randn(getattr(a, k).shape, generator=gen, device="cuda")
            ga, gb = g.clone(), g.clone()
            if k in ("xyz_deform", "rot_deform"):
                cols = act["xyz" if k == "xyz_deform" else "rotation"]
                inactive = torch.ones(g.shape[0], dtype=torch.bool, device="cuda")
                inactive[torch.tensor(cols, device="cuda")] = False
                ga[inactive] = 0.0
                gb[inactive] = float("nan")
            getattr(a, k).grad, getattr(b, k).grad = ga, gb
        oa.step()
        b._note_active_columns(b.time_basis(t, flow_t))
        ob.step()
    torch.cuda.synchronize()
    for k in PARAM_NAMES:
        assert torch.equal(getattr(a, k).detach(), getattr(b, k).detach()), k
        for w in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(oa.state[k][w], ob.state[k][w]), (k, w)
        assert torch.isfinite(getattr(b, k).detach()).all()


@pytest.mark.parametrize("n_obj", [2000, 2001])
def test_window_aware_training_matches_dense_training(n_obj):
    """End to end: render -> backward (inactive planes left unwritten, the allocator's free blocks poisoned with
    NaN) -> window-aware step, three iterations with a different B-spline window each, against the dense path. Not
    bit-exact: the blend backward's floating-point REDs are unordered, so two runs differ in the last bits of every
    gradient."""
    dense, od = _train_steps(False, n_obj=n_obj)
    sparse, os_ = _train_steps(True, n_obj=n_obj)
    for k in PARAM_NAMES:
        pd, ps = getattr(dense, k).detach(), getattr(sparse, k).detach()
        assert torch.isfinite(ps).all(), k
        scale = max(pd.abs().max().item(), 1e-12)
        rel = (pd - ps).abs() / scale
        # Adam with eps = 1e-15 turns ANY non-zero gradient into a full-size step, so an element whose gradient is
        # pure RED-ordering noise around zero may move by +-lr in one run and not in the other: a handful of such
        # elements is tolerated (and bounded by the three steps taken); a wrong or unwritten plane would show up as
        # thousands of elements
        bad = int((rel > 1e-4).sum())
        assert bad <= max(16, pd.numel() // 20_000), f"{k}: {bad} of {pd.numel()} elements differ, max {rel.max().item():.3e}"
        assert rel.max().item() <= 0.7, f"{k}: {rel.max().item():.3e}"
    # and training moved the parameters
    fresh, _, _ = _model(4000, n_obj, seed=3)
    assert not torch.equal(fresh.xyz_deform.detach(), dense.xyz_deform.detach())
    assert not torch.equal(fresh.sh4.detach(), dense.sh4.detach())


def test_window_aware_rejects_two_backwards_per_step():
    model, _, cam = _model(500, 300)
    model.optimizer = FusedAdam(model, LRS, window_aware=True)
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=False)
    vcam = SimpleNamespace(image_height=96, image_width=160, FoVx=cam.FoVx, FoVy=cam.FoVy,
                           world_view_transform=cam.world_view_transform, full_proj_transform=cam.full_proj_transform,
                           camera_center=cam.camera_center, time=0.3)
    render(vcam, model, None, pipe)["render"].sum().backward()
    with pytest.raises(RuntimeError, match="one render backward per optimizer step"):
        render(vcam, model, None, pipe)["render"].sum().backward()
