"""TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's environment map (SURVEY.md section 8f
rank 3). Only tests/ may import this module.

Follows, device-agnostic: get_image_cam_rays (scene/env.py:11-27), EnvironmentMap.get_image_background (:44-64),
get_env_color (:66-76), utils/graphics_utils.py:fov2focal (:82-83) and vector_to_theta (:95-100), and the
composite of gaussian_renderer/__init__.py:92-94. Pinned by tests/golden/env.npz = outputs and autograd
gradients of the reference's OWN scene/env.py run on the CPU (tests/golden/make_env_golden.py).
"""
import math

import torch
from torch.nn.functional import grid_sample, normalize


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def get_image_cam_rays(focal, height, width, device):
    K = torch.tensor([[focal, 0.0, width / 2], [0.0, focal, height / 2], [0.0, 0.0, 1.0]], dtype=torch.float32, device=device)
    K_inv = torch.inverse(K)
    grid = torch.stack(torch.meshgrid(torch.arange(0, width, dtype=torch.float32, device=device),
                                      torch.arange(0, height, dtype=torch.float32, device=device), indexing='xy'), dim=-1)
    pts = torch.cat([grid, torch.ones((grid.shape[0], grid.shape[1], 1), dtype=torch.float32, device=device)], dim=-1)[..., None]
    rays = (K_inv @ pts)
    return normalize(rays[..., 0], p=2, dim=-1)


def vector_to_theta(x):
    x, y, z = x[..., 0:1], x[..., 1:2], x[..., 2:3]
    hxy = torch.hypot(x, y)
    return torch.cat([torch.arctan2(y, x), torch.arctan2(z, hxy)], dim=-1)


def get_image_background(grid_map, fovx, height, width, world_view_transform):
    dev = grid_map.device
    rays = get_image_cam_rays(fov2focal(fovx, width), height, width, dev)
    rays = (world_view_transform[:3, :3].to(dev) @ rays[..., None]).squeeze(-1)
    angle = vector_to_theta(normalize(rays, p=2, dim=-1))
    scale = torch.tensor([1.0 / torch.pi, 2.0 / torch.pi], dtype=torch.float32, device=dev)
    rgb = grid_sample(grid_map, grid=(angle * scale)[None, ...], align_corners=True)
    return torch.sigmoid(rgb).squeeze(0)


def composite(foreground, img_opacity, background):
    return foreground + (1.0 - img_opacity) * background
