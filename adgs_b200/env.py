"""Environment map of the scene background: mirror of
`scene/env.py:EnvironmentMap` -- same constructor, `grid_map` parameter (1,C,R,R), `get_image_background`,
`get_env_color`, `training_setup`, `save_weights` / `load_weights` -- on the fused kernels of
adgs_b200/csrc/env.cu.

What changes underneath:
  * `composite(foreground, img_opacity, cam)` evaluates the background AND the blend
    `foreground + (1 - img_opacity) * background` (gaussian_renderer/__init__.py:92-94) in one kernel;
    `adgs_b200.gaussian_renderer.render` calls it when it is handed this class;
  * the texel gradients never pass through autograd: the backward adds them into a persistent dense buffer
    (`grad_buffer`) and marks the 32x32-texel tiles it touched; `optimizer.step()` runs Adam over the tiles ever
    touched and clears their gradient in the same pass. This is exactly the dense
    `torch.optim.Adam([grid_map], lr=env_lr, eps=1e-15)` of scene/env.py:78-83 (an untouched texel has zero
    moments and a zero update), without the 805 MB zero-filled gradient and the 201 M-element optimizer pass that
    the reference pays per iteration for its 8192^2 map. `grid_map.grad` therefore stays None.
No fallback: the methods raise if the CUDA library is missing.
"""
import ctypes as C
import math

import torch
from torch import nn

from . import _lib as L


def fov2focal(fov, pixels):
    """utils/graphics_utils.py:82-83"""
    return pixels / (2 * math.tan(fov / 2))


class _EnvComposite(torch.autograd.Function):
    """(foreground (C,H,W) | None, img_opacity (1,H,W)|(H,W) | None, grid_map) -> (rendered, background).
    grid_map is an input only so that the outputs require grad whenever the map does; its gradient goes to the
    environment's persistent buffer, not through autograd (backward returns None for it)."""

    @staticmethod
    def forward(ctx, foreground, img_opacity, grid_map, env, H, W, focal, view):
        lib = L.load()
        dev = env.grid_map.device
        Cc = env.num_channel
        fg = None if foreground is None else foreground.detach().to(torch.float32).contiguous()
        op = None if img_opacity is None else img_opacity.detach().to(torch.float32).contiguous()
        background = torch.empty((Cc, H, W), dtype=torch.float32, device=dev)
        rendered = torch.empty((Cc, H, W), dtype=torch.float32, device=dev)
        view_c = view.detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            st = lib.adgs_env_forward(C.byref(env._c_env(False)), H, W, float(focal), view_c.data_ptr(), L.ptr(fg), L.ptr(op),
                                      background.data_ptr(), rendered.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "env_forward")
        ctx.env, ctx.geom = env, (H, W, float(focal))
        ctx.save_for_backward(view_c, op if op is not None else torch.empty(0, device=dev))
        ctx.has = (foreground is not None, img_opacity is not None,
                   None if img_opacity is None else tuple(img_opacity.shape))
        return rendered, background

    @staticmethod
    def backward(ctx, g_rendered, g_background):
        lib = L.load()
        env = ctx.env
        H, W, focal = ctx.geom
        view_c, op = ctx.saved_tensors
        dev = view_c.device
        has_fg, has_op, op_shape = ctx.has
        gr = None if g_rendered is None else g_rendered.to(torch.float32).contiguous()
        gb = None if g_background is None else g_background.to(torch.float32).contiguous()
        d_op = torch.empty((H, W), dtype=torch.float32, device=dev) if (has_op and ctx.needs_input_grad[1]) else None
        if gr is not None or gb is not None:
            with torch.cuda.device(dev):
                st = lib.adgs_env_backward(C.byref(env._c_env(True)), H, W, focal, view_c.data_ptr(),
                                           op.data_ptr() if has_op else None, L.ptr(gr), L.ptr(gb), L.ptr(d_op),
                                           torch.cuda.current_stream(dev).cuda_stream)
            L.check(st, "env_backward")
            env._backwards += 1
        elif d_op is not None:
            d_op.zero_()
        d_fg = gr if (has_fg and ctx.needs_input_grad[0]) else None
        if d_fg is None and has_fg and ctx.needs_input_grad[0]:
            d_fg = torch.zeros_like(view_c.new_empty((env.num_channel, H, W)))
        return d_fg, (None if d_op is None else d_op.reshape(op_shape)), None, None, None, None, None, None


class EnvAdam:
    """`torch.optim.Adam([grid_map], lr, eps=1e-15)` of scene/env.py:78-83 over the touched tiles."""

    def __init__(self, env, lr, betas=(0.9, 0.999), eps=1e-15):
        self.env = env
        self.param_groups = [{"params": [env.grid_map], "lr": float(lr), "name": "env", "betas": betas, "eps": eps}]
        self.step_count = 0

    @torch.no_grad()
    def step(self):
        lib = L.load()
        env = self.env
        if env._backwards == 0:
            # torch.optim.Adam skips parameters whose .grad is None: nothing to do before the first backward
            return
        self.step_count += 1
        g = self.param_groups[0]
        dev = env.grid_map.device
        with torch.cuda.device(dev):
            st = lib.adgs_env_adam_step(C.byref(env._c_env(True)), float(g["lr"]), float(g["betas"][0]),
                                        float(g["betas"][1]), float(g["eps"]), self.step_count,
                                        torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "env_adam_step")

    def zero_grad(self, set_to_none=True):
        """The step clears the gradient of every tile it visits; nothing else holds a gradient."""
        return None


class EnvironmentMap:
    def __init__(self, resolution, num_channel=3, use_cache=True, device="cuda"):
        self.resolution = int(resolution)
        self.num_channel = int(num_channel)
        grid_map = (torch.rand((1, num_channel, resolution, resolution), dtype=torch.float32, device=device) * 2.0 - 1.0) * 1e-4
        self.grid_map = nn.Parameter(grid_map.requires_grad_(True))
        self.scale = torch.tensor([1.0 / torch.pi, 2.0 / torch.pi], dtype=torch.float32, device=device)
        self.optimizer = None
        self.use_cache = use_cache      # kept for signature compatibility: the rays are recomputed in registers
        self._state = None
        self._backwards = 0

    # ---- native views --------------------------------------------------------------------------------
    def _ensure_state(self):
        if self._state is None or self._state["grad"].shape != self.grid_map.shape:
            lib = L.load()
            z = lambda: torch.zeros_like(self.grid_map.detach())
            self._state = {"grad": z(), "exp_avg": z(), "exp_avg_sq": z(),
                           "touched": torch.zeros((lib.adgs_env_touched_bytes(self.resolution),), dtype=torch.uint8,
                                                  device=self.grid_map.device),
                           "tile_list": torch.zeros((lib.adgs_env_tile_list_bytes(self.resolution) // 4,),
                                                    dtype=torch.int32, device=self.grid_map.device)}
        return self._state

    def _c_env(self, with_state):
        if not self.grid_map.is_contiguous():
            raise RuntimeError("EnvironmentMap.grid_map must be contiguous")
        e = L.EnvMap(R=self.resolution, C=self.num_channel, grid=self.grid_map.data_ptr())
        if with_state:
            s = self._ensure_state()
            e.grad, e.exp_avg, e.exp_avg_sq = s["grad"].data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr()
            e.touched = s["touched"].data_ptr()
            e.tile_list = s["tile_list"].data_ptr()
        return e

    @property
    def grad_buffer(self):
        """Persistent dense gradient of grid_map (zero outside the tiles touched since their last step)."""
        return self._ensure_state()["grad"]

    # ---- reference interface ---------------------------------------------------------------------------
    def composite(self, foreground, img_opacity, cam):
        """(rendered, background) with rendered = foreground + (1 - img_opacity) * background."""
        H, W = int(cam.image_height), int(cam.image_width)
        focal = fov2focal(cam.FoVx, W)
        return _EnvComposite.apply(foreground, img_opacity, self.grid_map, self, H, W, focal, cam.world_view_transform)

    def get_image_background(self, cam, use_cache=True, return_grid=False):
        """scene/env.py:44-64 -> (C,H,W). Differentiable w.r.t. grid_map through the side-channel gradient buffer."""
        if return_grid:
            raise NotImplementedError("return_grid (the pixel grid used for flow visualisation) is not part of the hot path")
        return self.composite(None, None, cam)[1]

    def get_env_color(self, view, input_angle=False):
        """scene/env.py:66-76 for arbitrary directions: evaluation helper (extract_env_map), plain torch, no gradients kept."""
        from torch.nn.functional import grid_sample, normalize
        with torch.no_grad():
            if not input_angle:
                v = normalize(view, p=2, dim=-1)
                x, y, z = v[..., 0:1], v[..., 1:2], v[..., 2:3]
                angle = torch.cat([torch.arctan2(y, x), torch.arctan2(z, torch.hypot(x, y))], dim=-1)
            else:
                angle = view
            rgb = grid_sample(self.grid_map, grid=(angle * self.scale)[None, ...], align_corners=True)
            return torch.sigmoid(rgb).squeeze(0)

    def training_setup(self, training_args):
        """scene/env.py:78-83"""
        self._ensure_state()
        self.optimizer = EnvAdam(self, lr=training_args.env_lr, eps=1e-15)

    def save_weights(self, weights_path):
        torch.save(self.grid_map, weights_path)          # scene/env.py:85-86: same file content

    def load_weights(self, weights_path):
        grid_map = torch.load(weights_path, map_location=self.grid_map.device)
        self.grid_map = nn.Parameter(grid_map.requires_grad_(True))
        self.resolution, self.num_channel = int(grid_map.shape[-1]), int(grid_map.shape[1])
        self._state = None
        self._backwards = 0
