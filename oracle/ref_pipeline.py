"""TEST INFRASTRUCTURE ONLY -- the reference's per-iteration pipeline, assembled from
  * oracle/trajectory_oracle.py  (torch restatement of utils/func_utils.py + GaussianModel accessors,
    run on the GPU the way the reference runs it: many small ATen kernels + autograd), and
  * oracle/ref_module.py         (the UNMODIFIED reference rasterizer built in oracle/_ref),
glued exactly like gaussian_renderer/__init__.py:18-115 + _RasterizeGaussians
(diff_gaussian_rasterization/__init__.py:48-174). Used by tests/ as the end-to-end oracle and by
`bench.py --impl reference` as the reference arm. Never imported by adgs_b200/.
"""
import torch


class RefRasterize(torch.autograd.Function):
    """autograd shell around a rasterizer backend exposing the reference's `_C` signatures
    (backend = oracle.ref_module, or adgs_b200.rasterizer._C for A/B tests)."""

    @staticmethod
    def forward(ctx, backend, c, means3D, opacity, scales, rotations, sh, flow_points, semantic):
        cam = c["cam"]
        e = torch.Tensor([])
        args = (c["background"], means3D, e, opacity, scales, rotations, 1.0, e, cam.world_view_transform,
                cam.full_proj_transform, c["tan_fovx"], c["tan_fovy"], c["H"], c["W"], sh, flow_points, semantic,
                c["degree"], cam.camera_center, False, c["inv_depth"], False)
        out = backend.rasterize_gaussians(*args)
        ctx.backend, ctx.c, ctx.out = backend, c, out
        ctx.save_for_backward(means3D, opacity, scales, rotations, sh, flow_points, semantic)
        return out[1], out[4], out[2], out[3], out[8], out[9]

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_opacity, g_flow, g_sem):
        means3D, opacity, scales, rotations, sh, flow_points, semantic = ctx.saved_tensors
        c, out, cam = ctx.c, ctx.out, ctx.c["cam"]
        e = torch.Tensor([])
        args = (c["background"], means3D, out[4], e, scales, rotations, 1.0, e, cam.world_view_transform,
                cam.full_proj_transform, c["tan_fovx"], c["tan_fovy"], g_color, g_depth, g_flow, g_sem, semantic,
                flow_points, sh, c["degree"], cam.camera_center, out[5], out[0], out[6], out[7], out[3], g_opacity,
                c["inv_depth"], False)
        if getattr(ctx.backend, "__name__", "") == "_C":     # our drop-in takes the opacities too
            g = ctx.backend.rasterize_gaussians_backward(*args, opacities=opacity)
        else:
            g = ctx.backend.rasterize_gaussians_backward(*args)
        return None, None, g[3], g[2], g[6], g[7], g[5], g[8], None


def reference_render(ref_model, c, t, flow_t, backend, render_objmask=True):
    """gaussian_renderer.render() of the reference on a trajectory_oracle.ReferenceModel.
    Returns ((color, radii, depth, img_opacity, img_flow, img_semantic), deform_pkg)."""
    flow = ref_model.get_deformed_xyz(flow_t) if flow_t is not None else torch.Tensor([])
    pkg = ref_model.get_deformed_pkg(t)
    sem = ref_model.get_obj_mask().float()[..., None] if render_objmask else torch.Tensor([])
    out = RefRasterize.apply(backend, c, pkg['xyz'], pkg['opacity'], ref_model.get_scaling(), pkg['rotation'],
                             pkg['shs'], flow, sem)
    return out, pkg
