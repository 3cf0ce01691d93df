"""Generates tests/golden/env.npz by running the REFERENCE's own scene/env.py (EnvironmentMap.get_image_background
through get_image_cam_rays, vector_to_theta and torch grid_sample) and the composite of
gaussian_renderer/__init__.py:92-94 on the CPU, with torch autograd for the gradients.

Non-invasive shims (the file itself is loaded unmodified from /root/reference):
  * `open3d`, `scene.cameras`, `utils.system_utils` (imports of scene/env.py that the evaluated functions do not
    use) are replaced by empty modules; the `scene` package is not imported (its __init__ pulls in the dataset
    readers);
  * the hard-coded device='cuda' (scene/env.py:16-21,31-35) and `.cuda()` (:60) are redirected to the CPU (torch.tensor / arange / rand / ones wrappers).
Usage: python tests/golden/make_env_golden.py
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
for name in ("open3d", "scene", "scene.cameras", "utils.system_utils"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["scene.cameras"].Camera = object
sys.modules["utils.system_utils"].searchForMaxIteration = None


def _cpu(fn):
    def wrapped(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


for fname in ("tensor", "arange", "rand", "ones"):
    setattr(torch, fname, _cpu(getattr(torch, fname)))
torch.Tensor.cuda = lambda self, *a, **k: self

spec = importlib.util.spec_from_file_location("ref_env", os.path.join(REF, "scene", "env.py"))
E = importlib.util.module_from_spec(spec)
spec.loader.exec_module(E)  # the reference's file, unmodified


def case(R, H, W, fovx_deg, yaw_deg, seed):
    g = torch.Generator().manual_seed(seed)
    env = E.EnvironmentMap(R, num_channel=3, use_cache=False)
    with torch.no_grad():
        env.grid_map.copy_(torch.randn(1, 3, R, R, generator=g))
    yaw, pitch = math.radians(yaw_deg), math.radians(7.0)
    Ry = torch.tensor([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]], dtype=torch.float32)
    Rx = torch.tensor([[1, 0, 0], [0, math.cos(pitch), -math.sin(pitch)], [0, math.sin(pitch), math.cos(pitch)]], dtype=torch.float32)
    wvt = torch.eye(4)
    wvt[:3, :3] = Ry @ Rx
    wvt[3, :3] = torch.tensor([0.3, -0.2, 1.0])
    cam = types.SimpleNamespace(FoVx=math.radians(fovx_deg), image_width=W, image_height=H, world_view_transform=wvt, cam_id=0)
    fg = torch.rand(3, H, W, generator=g).requires_grad_(True)
    op = torch.rand(1, H, W, generator=g).requires_grad_(True)
    bg = env.get_image_background(cam)
    rendered = fg + (1.0 - op) * bg                       # gaussian_renderer/__init__.py:92-94
    cot = torch.randn(3, H, W, generator=g)
    (rendered * cot).sum().backward()
    return dict(grid_map=env.grid_map.detach(), wvt=wvt, fovx=np.float32(cam.FoVx), fg=fg.detach(), op=op.detach(),
                background=bg.detach(), rendered=rendered.detach(), cot=cot, d_grid=env.grid_map.grad, d_fg=fg.grad,
                d_op=op.grad)


if __name__ == "__main__":
    data = {}
    for name, args in {"a": (64, 24, 40, 90.0, 20.0, 1), "b": (33, 17, 9, 120.0, -140.0, 2), "c": (128, 30, 50, 60.0, 175.0, 3)}.items():
        for k, v in case(*args).items():
            data[f"{name}_{k}"] = np.asarray(v.numpy() if torch.is_tensor(v) else v)
    np.savez_compressed(os.path.join(HERE, "env.npz"), **data)
    print("wrote env.npz", len(data), "arrays")
