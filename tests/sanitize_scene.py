"""Small fwd+bwd of the three entry paths for `compute-sanitizer` (tools/sanitize.sh):
strict drop-in rasterizer, fused render(), splat exchange at world size 1 (multi-view kernels). Not a pytest test."""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import helpers as Hh  # noqa: E402
from adgs_b200 import scenes  # noqa: E402
from adgs_b200.gaussian_model import GaussianModel  # noqa: E402
from adgs_b200.gaussian_renderer import render  # noqa: E402
from adgs_b200.parallel import SplatExchangeStep  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
W, H = 64, 48
if which in ("all", "strict"):
    c = Hh.make_case(n=10_000, W=W, H=H, seed=1)
    out = Hh.OURS.rasterize_gaussians(*Hh.fwd_args(c))
    cot = Hh.cotangents(c)
    Hh.OURS.rasterize_gaussians_backward(*Hh.bwd_args(c, out, cot), opacities=c["opacity"])
    torch.cuda.synchronize()
    print("strict ok, num_rendered", out[0])
if which in ("all", "fused", "exchange"):
    cam = scenes.make_camera(W, H, 90.0, time=0.37, device="cuda")
    cloud = scenes.random_cloud(10_000, cam, seed=2, median_radius_px=4.0)
    tensors = scenes.random_model_tensors(7000, 3000, scenes.BENCH_ORDER_ARGS, cloud, seed=3, device="cuda")
    model = GaussianModel.from_reference(tensors, scenes.BENCH_ORDER_ARGS, device="cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    cot = {k: torch.randn(ch, H, W, generator=g).cuda() for k, ch in (("color", 3), ("depth", 1), ("opacity", 1),
                                                                      ("flow", 3), ("semantic", 1))}
if which in ("all", "fused"):
    for sync_free in (False, True):
        pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=sync_free)
        res = render(cam, model, None, pipe, flow_pkg=[0.41, None, None, None, None, None], render_objmask=True)
        torch.autograd.backward((res["render"], res["depth"], res["img_opacity"], res["img_flow"], res["img_semantic"]),
                                (cot["color"], cot["depth"][0], cot["opacity"][0], cot["flow"], cot["semantic"]))
    torch.cuda.synchronize()
    print("fused ok")
if which in ("all", "exchange"):
    pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
    ex = SplatExchangeStep(model)
    views = [(cam._replace(time=0.2), 0.25), (cam._replace(time=0.5), 0.55), (cam._replace(time=0.8), 0.75)]
    ex.run(views, lambda v, r: cot, pipe)
    ex.run(views, lambda v, r: cot, pipe, views_per_rank=3)
    torch.cuda.synchronize()
    print("exchange (world 1) ok")
