"""GPU integration: render -> fused losses -> backward -> FusedAdam (adgs_b200/train_step.py, the inner part
of train.py:74-167) drives the loss down on a small synthetic scene, in the dense and in the window-aware
optimizer mode."""
from types import SimpleNamespace

import pytest
import torch

from adgs_b200 import scenes
from adgs_b200.gaussian_model import GaussianModel
from adgs_b200.gaussian_renderer import render
from adgs_b200.train_step import training_iteration

pytestmark = pytest.mark.gpu

ARGS = SimpleNamespace(percent_dense=0.01, object_extent=10.0, min_camera_extent=10.0, feature_lr=0.0025, opacity_lr=0.05,
                       scaling_lr=0.005, rotation_lr=0.001, rotation_deform_lr=0.001, shs_deform_lr=0.0025,
                       gs_time_sigma_lr=1e-2, position_lr_init=0.00016, position_lr_final=0.0000016,
                       position_lr_delay_mult=0.01, position_lr_max_steps=60_000, position_deform_lr_scale=0.2,
                       obj_position_lr_scale=0.8, scene_position_lr_scale=1.0)
OPT = SimpleNamespace(lambda_dssim=0.2, lambda_l1=1.0, lambda_depth=0.1, lambda_flow=0.0, lambda_obj=0.1, lambda_sky=0.05,
                      lambda_sigma=0.01)


def _build(seed):
    W, H = 160, 96
    cam = scenes.make_camera(W, H, 90.0, device="cuda")
    cloud = scenes.random_cloud(6000, cam, seed=seed, median_radius_px=4.0)
    tensors = scenes.random_model_tensors(4000, 2000, scenes.BENCH_ORDER_ARGS, cloud, seed=seed + 1, device="cuda")
    return GaussianModel.from_reference(tensors, scenes.BENCH_ORDER_ARGS), cam, W, H


def _view(cam, t, **targets):
    return SimpleNamespace(image_height=cam.image_height, image_width=cam.image_width, FoVx=cam.FoVx, FoVy=cam.FoVy,
                           world_view_transform=cam.world_view_transform, full_proj_transform=cam.full_proj_transform,
                           camera_center=cam.camera_center, time=t, **targets)


@pytest.mark.parametrize("window_aware", [False, True])
def test_training_iterations_reduce_the_loss(window_aware):
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
    # targets: a render of the SAME scene with different colours / opacities (so the optimum is reachable)
    target_model, cam, W, H = _build(seed=1)
    with torch.no_grad():
        target_model.sh4.mul_(0.3).add_(0.2)
        target_model.opacity.add_(1.0)
        tgt = render(_view(cam, 0.4), target_model, None, pipe, render_objmask=True)
    targets = dict(original_image=tgt["render"].clamp(0, 1).clone(), depth=tgt["depth"].clone(),
                   semantic=(tgt["img_semantic"][0] > 0.5).float(), sky=(tgt["img_opacity"] < 0.05).float())
    model, _, _, _ = _build(seed=1)
    model.scene_extent = 20.0
    model.training_setup(ARGS, window_aware=window_aware)
    view = _view(cam, 0.4, **targets)
    losses = []
    for it in range(1, 41):
        logs, _ = training_iteration(model, view, OPT, pipe, it, frame_gap=1.0 / 96)
        losses.append(logs["total_loss"])
    losses = torch.stack(losses).cpu()
    assert torch.isfinite(losses).all()
    assert losses[-5:].mean() < 0.8 * losses[:5].mean(), losses.tolist()
    assert model.optimizer.step_count == 40


@pytest.mark.parametrize("window_aware", [False, True])
def test_training_with_densification_and_near_regularisers(window_aware):
    """train.py:74-167 including the densification block (:149-160) and, in the dense optimizer mode, the
    near-index regularisers (:101-110): the Gaussian set changes every 10 iterations, the optimizer state follows
    it, deform_background runs one Adam step ahead of the per-Gaussian arrays after each densification (torch's
    per-parameter step counts), and the loss stays finite."""
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
    target_model, cam, W, H = _build(seed=2)
    with torch.no_grad():
        target_model.sh4.mul_(0.3).add_(0.2)
        tgt = render(_view(cam, 0.4), target_model, None, pipe, render_objmask=True)
    targets = dict(original_image=tgt["render"].clamp(0, 1).clone(), depth=tgt["depth"].clone(),
                   semantic=(tgt["img_semantic"][0] > 0.5).float(), sky=(tgt["img_opacity"] < 0.05).float())
    model, _, _, _ = _build(seed=2)
    model.scene_extent = 20.0
    args = SimpleNamespace(**vars(ARGS), lambda_reg=0.0 if window_aware else 0.5, lambda_sigma=0.01,
                           lambda_sigma_reg=0.0 if window_aware else 0.5, near_num=8)
    opt = SimpleNamespace(**vars(OPT), lambda_reg=args.lambda_reg, lambda_sigma_reg=args.lambda_sigma_reg,
                          densify_until_iter=26, densify_from_iter=0, densification_interval=10,
                          opacity_reset_interval=15, densify_scene_grad_threshold=1e-7, densify_obj_grad_threshold=1e-7,
                          near_idx_reset_interval=5)
    model.training_setup(args, window_aware=window_aware)
    assert model.use_near_idx == (not window_aware)
    if not window_aware:
        assert model.obj_near_idx.shape == (2000 // 8, 8)
    view = _view(cam, 0.4, **targets)
    counts, losses = [], []
    for it in range(1, 31):
        logs, pkg = training_iteration(model, view, opt, pipe, it, frame_gap=1.0 / 96, densify=opt)
        counts.append(model.get_pts_num)
        losses.append(logs["total_loss"])
        if not window_aware:
            assert logs["reg_loss"] is not None and logs["reg_sigma_loss"] is not None
    assert torch.isfinite(torch.stack(losses)).all()
    assert counts[9] != counts[8] and counts[19] != counts[18]      # densified at iterations 10 and 20
    assert counts[29] == counts[19]                                 # densify_until_iter
    sc = model.optimizer.step_counts
    assert sc["background_deform"] == 30 and sc["xyz"] == 28 and sc["rot_deform"] == 28
    assert model.xyz_gradient_accum.shape[0] == model.get_pts_num
    assert model.optimizer.state["sh4"]["exp_avg"].shape == model.sh4.shape
    if not window_aware:
        assert int(model.obj_near_idx.max()) < model.n_obj
    assert float(torch.sigmoid(model.opacity).max()) <= 1.0
