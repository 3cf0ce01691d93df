"""GPU parity of the fused environment map (adgs_b200/env.py -> adgs_env_*) against the golden vectors of the
reference's own scene/env.py, against the oracle + torch.optim.Adam over several training steps (the
touched-tile step must equal the dense optimizer), and through render()."""
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from adgs_b200.env import EnvironmentMap

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "env.npz"))
T = lambda k: torch.tensor(G[k], device="cuda")


def _rel(a, b):
    a, b = a.detach().double().cpu().numpy(), np.asarray(b, dtype=np.float64)
    return max(np.abs(a - b).max() - 5e-7, 0.0) / max(np.abs(b).max(), 1e-12)


def _env_from(grid):
    env = EnvironmentMap(grid.shape[-1], num_channel=grid.shape[1])
    with torch.no_grad():
        env.grid_map.copy_(grid)
    return env


@pytest.mark.parametrize("c", ["a", "b", "c"])
def test_env_matches_reference_golden(c):
    env = _env_from(T(f"{c}_grid_map"))
    fg, op = T(f"{c}_fg").requires_grad_(True), T(f"{c}_op").requires_grad_(True)
    H, W = fg.shape[1:]
    cam = SimpleNamespace(FoVx=float(G[f"{c}_fovx"]), image_width=W, image_height=H, world_view_transform=T(f"{c}_wvt"))
    rendered, bg = env.composite(fg, op, cam)
    assert _rel(bg, G[f"{c}_background"]) <= 1e-4 and _rel(rendered, G[f"{c}_rendered"]) <= 1e-4
    (rendered * T(f"{c}_cot")).sum().backward()
    assert _rel(fg.grad, G[f"{c}_d_fg"]) <= 1e-6 and _rel(op.grad, G[f"{c}_d_op"]) <= 1e-4
    assert _rel(env.grad_buffer, G[f"{c}_d_grid"]) <= 1e-4
    assert _rel(env.get_image_background(cam), G[f"{c}_background"]) <= 1e-4


def _cam(H, W, yaw_deg):
    yaw = math.radians(yaw_deg)
    wvt = torch.eye(4, device="cuda")
    wvt[:3, :3] = torch.tensor([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
    return SimpleNamespace(FoVx=math.radians(90.0), image_width=W, image_height=H, world_view_transform=wvt)


def test_touched_tile_adam_equals_dense_torch_adam():
    from oracle import env_oracle as EO
    R, H, W = 192, 40, 64
    g = torch.Generator(device="cuda").manual_seed(2)
    grid0 = torch.randn(1, 3, R, R, generator=g, device="cuda") * 0.5
    env = _env_from(grid0)
    env.training_setup(SimpleNamespace(env_lr=1e-2))
    ref = torch.nn.Parameter(grid0.clone())
    ref_opt = torch.optim.Adam([{"params": [ref], "lr": 1e-2, "name": "env"}], lr=0.0, eps=1e-15)
    for it in range(6):
        cam = _cam(H, W, yaw_deg=25.0 * it)          # the camera sweeps: new tiles are touched, old ones keep decaying
        fg = torch.rand(3, H, W, generator=g, device="cuda")
        op = torch.rand(1, H, W, generator=g, device="cuda")
        cot = torch.randn(3, H, W, generator=g, device="cuda")
        rendered, _ = env.composite(fg, op, cam)
        (rendered * cot).sum().backward()
        env.optimizer.step()
        env.optimizer.zero_grad(set_to_none=True)
        bg = EO.get_image_background(ref, cam.FoVx, H, W, cam.world_view_transform)
        (EO.composite(fg, op, bg) * cot).sum().backward()
        ref_opt.step()
        ref_opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    assert _rel(env.grid_map, ref.detach().cpu().numpy()) <= 2e-5
    assert float(env.grad_buffer.abs().max()) == 0.0            # every visited tile was cleared by the step
    touched = env._state["touched"].float().mean().item()
    assert 0.0 < touched < 1.0                                   # only part of the map is ever stepped
    assert env.grid_map.grad is None


def test_render_uses_the_fused_composite():
    from adgs_b200 import scenes
    from adgs_b200.gaussian_model import GaussianModel
    from adgs_b200.gaussian_renderer import render
    from oracle import env_oracle as EO
    W, H = 160, 96
    cam = scenes.make_camera(W, H, 90.0, device="cuda")
    cloud = scenes.random_cloud(3000, cam, seed=5, median_radius_px=4.0)
    tensors = scenes.random_model_tensors(2000, 1000, scenes.BENCH_ORDER_ARGS, cloud, seed=6, device="cuda")
    model = GaussianModel.from_reference(tensors, scenes.BENCH_ORDER_ARGS)
    env = EnvironmentMap(128)
    with torch.no_grad():
        env.grid_map.normal_(0.0, 1.0)
    view = SimpleNamespace(image_height=H, image_width=W, FoVx=cam.FoVx, FoVy=cam.FoVy,
                           world_view_transform=cam.world_view_transform, full_proj_transform=cam.full_proj_transform,
                           camera_center=cam.camera_center, time=0.4)
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=False)
    res = render(view, model, env, pipe)
    bg = EO.get_image_background(env.grid_map.detach(), cam.FoVx, H, W, cam.world_view_transform)
    expect = res["foreground"].detach() + (1.0 - res["img_opacity"].detach()) * bg
    assert _rel(res["render"], expect.cpu().numpy()) <= 1e-4 and _rel(res["background"], bg.cpu().numpy()) <= 1e-4
    res["render"].sum().backward()
    assert float(env.grad_buffer.abs().max()) > 0.0 and model.opacity.grad is not None


def test_env_map_iteration_cost_at_reference_resolution():
    """8192^2 x 3 map (arguments/__init__.py:66-67), KITTI frame: composite + backward + optimizer step, fused
    vs the reference's formulation (oracle = scene/env.py restated: grid_sample autograd + dense torch Adam).
    Numbers go to gpurun_out/env_timing.json for profiles/."""
    import json
    from oracle import env_oracle as EO
    R, H, W = 8192, 375, 1242
    cam = _cam(H, W, 10.0)
    fg = torch.rand(3, H, W, device="cuda", requires_grad=True)
    op = torch.rand(1, H, W, device="cuda", requires_grad=True)
    cot = torch.randn(3, H, W, device="cuda")

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    env = EnvironmentMap(R)
    env.training_setup(SimpleNamespace(env_lr=1e-3))

    def ours():
        fg.grad = op.grad = None
        rendered, _ = env.composite(fg, op, cam)
        (rendered * cot).sum().backward()
        env.optimizer.step()
        env.optimizer.zero_grad(set_to_none=True)

    ms_ours = timed(ours)
    grid0 = env.grid_map.detach().clone()
    del env
    torch.cuda.empty_cache()
    ref = torch.nn.Parameter(grid0)
    ref_opt = torch.optim.Adam([{"params": [ref], "lr": 1e-3}], lr=0.0, eps=1e-15)

    def reference():
        fg.grad = op.grad = None
        bg = EO.get_image_background(ref, cam.FoVx, H, W, cam.world_view_transform)
        (EO.composite(fg, op, bg) * cot).sum().backward()
        ref_opt.step()
        ref_opt.zero_grad(set_to_none=True)

    ms_ref = timed(reference)
    out = {"resolution": R, "H": H, "W": W, "fused_ms_per_iteration": round(ms_ours, 4),
           "torch_ms_per_iteration": round(ms_ref, 4), "speedup": round(ms_ref / ms_ours, 1)}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.isdir(os.path.join(root, "gpurun_out")):
        with open(os.path.join(root, "gpurun_out", "env_timing.json"), "w") as f:
            json.dump(out, f)
    assert ms_ours < ms_ref, out
