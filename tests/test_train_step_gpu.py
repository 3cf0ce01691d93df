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
