"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from adgs_b200 import _lib as L  # noqa: E402
from adgs_b200 import scenes  # noqa: E402
from adgs_b200.rasterizer import GaussianRasterizationSettings, _C as OURS  # noqa: E402


def has_cuda():
    return torch.cuda.is_available()


def make_case(n=5000, W=160, H=96, seed=0, sh_degree=3, inv_depth=True, flow=True, D_S=1, bg=(0.0, 0.0, 0.0),
              median_radius_px=4.0, device="cuda", colors_precomp=False, cov3D_precomp=False, scale_modifier=1.0,
              yaw_deg=0.0, cluster=None):
    cam = scenes.make_camera(W, H, 90.0, yaw_deg=yaw_deg, device=device)
    cloud = scenes.random_cloud(n, cam, seed=seed, median_radius_px=median_radius_px, cluster=cluster)
    inp = scenes.activated_inputs(cloud, device=device)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    e = torch.Tensor([])
    case = dict(
        cam=cam, n=n, W=W, H=H,
        background=torch.tensor(bg, dtype=torch.float32, device=device),
        means3D=inp["means3D"], opacity=inp["opacities"], scales=inp["scales"], rotations=inp["rotations"],
        sh=inp["shs"], colors=e, cov3D_precomp=e,
        flow_points=(inp["means3D"] + 0.05 * torch.randn(n, 3, generator=g).to(device)) if flow else e,
        semantic=(torch.rand(n, D_S, generator=g).to(device) if D_S else e),
        scale_modifier=scale_modifier, tan_fovx=math.tan(cam.FoVx * 0.5), tan_fovy=math.tan(cam.FoVy * 0.5),
        degree=sh_degree, inv_depth=inv_depth,
    )
    if colors_precomp:
        case["colors"] = torch.rand(n, 3, generator=g).to(device)
        case["sh"] = e
    if cov3D_precomp:
        from oracle import raster_oracle as O
        c3 = O.cov3d(inp["scales"].cpu().numpy(), scale_modifier, inp["rotations"].cpu().numpy())
        case["cov3D_precomp"] = torch.tensor(c3, device=device)
        case["scales"] = e
        case["rotations"] = e
    return case


def fwd_args(c, debug=False):
    cam = c["cam"]
    return (c["background"], c["means3D"], c["colors"], c["opacity"], c["scales"], c["rotations"],
            c["scale_modifier"], c["cov3D_precomp"], cam.world_view_transform, cam.full_proj_transform,
            c["tan_fovx"], c["tan_fovy"], c["H"], c["W"], c["sh"], c["flow_points"], c["semantic"], c["degree"],
            cam.camera_center, False, c["inv_depth"], debug)


def cotangents(c, seed=7, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, W = c["H"], c["W"]
    D_S = c["semantic"].shape[1] if c["semantic"].numel() else 0
    mk = lambda ch: torch.randn(ch, H, W, generator=g).to(device)
    return dict(color=mk(3), depth=mk(1), flow=mk(3), semantic=mk(D_S), opacity=mk(1))


def bwd_args(c, fwd_out, cot, debug=False):
    cam = c["cam"]
    (rendered, color, depth, img_opacity, radii, geom, binning, img, img_flow, img_sem) = fwd_out
    return (c["background"], c["means3D"], radii, c["colors"], c["scales"], c["rotations"], c["scale_modifier"],
            c["cov3D_precomp"], cam.world_view_transform, cam.full_proj_transform, c["tan_fovx"], c["tan_fovy"],
            cot["color"], cot["depth"], cot["flow"], cot["semantic"], c["semantic"], c["flow_points"], c["sh"],
            c["degree"], cam.camera_center, geom, rendered, binning, img, img_opacity, cot["opacity"],
            c["inv_depth"], debug)


def _view(buf, off, count, dtype):
    esz = torch.empty((), dtype=dtype).element_size()
    return buf[off:off + count * esz].view(dtype)


def inspect_ours(geom, binning, img, P, R, W, H):
    """Named views into our three arenas (adgs_geometry_offsets & co.)."""
    lib = L.load()
    res = {}
    gl = L.GeometryLayout()
    lib.adgs_geometry_offsets(P, C.byref(gl))
    base = geom.data_ptr()
    pad = (-base) % 128
    g = geom[pad:]
    res["counters"] = _view(g, gl.counters, 4, torch.int32)
    res["sorted_depth_bits"] = _view(g, gl.depths, P, torch.int32)
    res["tiles_touched"] = _view(g, gl.tiles_touched, P, torch.int32)
    res["record"] = _view(g, gl.record, 16 * P, torch.float32).view(P, 16)
    res["cov3D"] = _view(g, gl.cov3D, 6 * P, torch.float32).view(P, 6)
    res["clamped"] = _view(g, gl.clamped, P, torch.uint8)
    res["depth_order"] = _view(g, gl.depth_order, P, torch.int32)
    res["point_offsets"] = _view(g, gl.point_offsets, P, torch.int32)
    if R > 0:
        bl = L.BinningLayout()
        lib.adgs_binning_offsets(R, C.byref(bl))
        b = binning[(-binning.data_ptr()) % 128:]
        alt = lib.adgs_binning_result_in_alt(W, H)
        res["point_list"] = _view(b, bl.point_list_alt if alt else bl.point_list, R, torch.int32)
        res["point_list_tile"] = _view(b, bl.point_list_tile_alt if alt else bl.point_list_tile, R, torch.int32)
    il = L.ImageLayout()
    lib.adgs_image_offsets(W, H, C.byref(il))
    im = img[(-img.data_ptr()) % 128:]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    res["ranges"] = _view(im, il.ranges, 2 * tiles, torch.int32).view(tiles, 2)
    res["n_contrib"] = _view(im, il.n_contrib, W * H, torch.int32).view(H, W)
    return res


def rel_err(a, b):
    """max |a-b| / max(|b|max, tiny): the '1e-4 relative' of BASELINE.json on whole tensors."""
    a, b = a.double(), b.double()
    if a.numel() == 0:
        return 0.0
    denom = max(b.abs().max().item(), 1e-12)
    return (a - b).abs().max().item() / denom


def elementwise_err(a, b, rtol=1e-4, atol_frac=1e-6):
    """Element-wise companion of rel_err: an element FAILS if |a-b| > rtol*|b| + atol with
    atol = atol_frac * max|b| (so that entries down to a millionth of the largest one are held to the relative
    bound instead of hiding behind the max norm). Returns (failing fraction, worst |a-b| / (rtol*|b| + atol))."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    if a.numel() == 0:
        return 0.0, 0.0
    atol = atol_frac * max(b.abs().max().item(), 1e-30)
    bound = rtol * b.abs() + atol
    excess = (a - b).abs() / bound
    return (excess > 1.0).double().mean().item(), excess.max().item()


def to_np(t):
    return None if (t is None or t.numel() == 0) else t.detach().cpu().numpy()


def oracle_settings(c):
    from oracle import raster_oracle as O
    cam = c["cam"]
    return O.Settings(c["H"], c["W"], c["tan_fovx"], c["tan_fovy"], to_np(c["background"]), c["scale_modifier"],
                      to_np(cam.world_view_transform.contiguous()), to_np(cam.full_proj_transform.contiguous()),
                      c["degree"], to_np(cam.camera_center), False, c["inv_depth"], False)
