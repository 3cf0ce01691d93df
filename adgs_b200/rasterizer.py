"""Drop-in replacement for the `diff_gaussian_rasterization` package of AD-GS.

Mirrors, name for name, the reference's plugin surface
(submodules/depth-diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py):
`GaussianRasterizationSettings` (:176-189), `GaussianRasterizer` (:191-251),
`rasterize_gaussians` (:21-46), `_RasterizeGaussians` (:48-174), and the three native entry
points of `_C` (ext.cpp:15-19) as the `_C` namespace below. Same argument order, same 6-tuple
result `(color, radii, depth, img_opacity, img_flow, img_semantic)`, same gradient order, same
errors. The compute goes through the C ABI of libadgs_b200.so (include/adgs_b200.h) on
torch's current CUDA stream; there is no CPU or PyTorch fallback.
"""
import ctypes as C
import threading
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib as L


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    inv_depth: bool
    debug: bool


# ---------------------------------------------------------------------------------------------
# arena allocation through callbacks (the reference's resizeFunctional lambdas,
# rasterize_points.cu:27-33): the native side asks for N bytes, we hand out a torch uint8 tensor.
# ---------------------------------------------------------------------------------------------
_tls = threading.local()


def _make_alloc(slot):
    def _alloc(nbytes, _user):
        ctx = _tls.ctx
        try:
            buf = torch.empty((int(nbytes),), dtype=torch.uint8, device=ctx["device"])
        except Exception as ex:  # out of memory -> null -> ADGS_ERR_ALLOC
            ctx["error"] = ex
            return None
        ctx[slot] = buf
        return buf.data_ptr()

    return L.ALLOC_FN(_alloc)


_GEOM_CB = _make_alloc("geom")
_BINNING_CB = _make_alloc("binning")
_IMAGE_CB = _make_alloc("image")


def _f32(t, name):
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t.contiguous()


def _camera(s: GaussianRasterizationSettings, keep):
    bg = _f32(s.bg, "bg")
    view = _f32(s.viewmatrix, "viewmatrix")
    proj = _f32(s.projmatrix, "projmatrix")
    campos = _f32(s.campos, "campos")
    keep.extend([bg, view, proj, campos])
    return L.Camera(
        image_height=int(s.image_height), image_width=int(s.image_width),
        tanfovx=float(s.tanfovx), tanfovy=float(s.tanfovy), scale_modifier=float(s.scale_modifier),
        sh_degree=int(s.sh_degree), prefiltered=int(bool(s.prefiltered)), inv_depth=int(bool(s.inv_depth)),
        debug=int(bool(s.debug)), _pad=0,
        bg=L.ptr(bg), viewmatrix=L.ptr(view), projmatrix=L.ptr(proj), campos=L.ptr(campos))


def _gaussians(means3D, sh, colors, flow_points, semantic, opacity, scales, rotations, cov3D, keep):
    means3D = _f32(means3D, "means3D")
    sh = _f32(sh, "sh")
    colors = _f32(colors, "colors_precomp")
    flow_points = _f32(flow_points, "flow_points")
    semantic = _f32(semantic, "semantic")
    opacity = _f32(opacity, "opacities")
    scales = _f32(scales, "scales")
    rotations = _f32(rotations, "rotations")
    cov3D = _f32(cov3D, "cov3D_precomp")
    keep.extend([means3D, sh, colors, flow_points, semantic, opacity, scales, rotations, cov3D])
    P = 0 if means3D is None else means3D.shape[0]
    M = 0 if sh is None else sh.shape[1]
    D_S = 0 if semantic is None else semantic.shape[1]
    if D_S > L.MAX_SEMANTIC:
        raise RuntimeError(f"semantic has {D_S} channels; at most {L.MAX_SEMANTIC} are supported")
    if flow_points is not None and flow_points.shape[1] != 3:
        raise RuntimeError("flow_points must have dimensions (num_points, 3)")
    return L.Gaussians(P=P, M=M, D_S=D_S, _pad=0, means3D=L.ptr(means3D), shs=L.ptr(sh),
                       colors_precomp=L.ptr(colors), flow_points=L.ptr(flow_points), semantic=L.ptr(semantic),
                       opacities=L.ptr(opacity), scales=L.ptr(scales), rotations=L.ptr(rotations),
                       cov3D_precomp=L.ptr(cov3D))


class _C:
    """The native entry points of the reference's pybind module (RZ/ext.cpp:15-19), same
    positional signatures and result tuples (RZ/rasterize_points.h:18-78)."""

    @staticmethod
    def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                            viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, flow_points,
                            semantic, degree, campos, prefiltered, inv_depth, debug):
        if means3D.ndimension() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        lib = L.load()
        dev = means3D.device
        P, H, W = means3D.shape[0], int(image_height), int(image_width)
        D_S = semantic.shape[1] if (semantic is not None and semantic.numel() != 0) else 0
        opts = dict(dtype=torch.float32, device=dev)
        has_color = (colors is not None and colors.numel() != 0) or (sh is not None and sh.numel() != 0)
        out_color = torch.empty((3, H, W), **opts) if has_color else torch.zeros((3, H, W), **opts)
        out_depth = torch.empty((1, H, W), **opts)
        img_opacity = torch.empty((1, H, W), **opts)
        img_flow = torch.empty((3, H, W), **opts)
        img_semantic = torch.empty((D_S, H, W), **opts)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        settings = GaussianRasterizationSettings(H, W, tan_fovx, tan_fovy, background, scale_modifier, viewmatrix,
                                                 projmatrix, degree, campos, prefiltered, inv_depth, debug)
        keep = []
        with torch.cuda.device(dev):
            cam = _camera(settings, keep)
            g = _gaussians(means3D, sh, colors, flow_points, semantic, opacity, scales, rotations, cov3D_precomp,
                           keep)
            out = L.Images(color=L.ptr(out_color) if has_color else None, depth=L.ptr(out_depth),
                           opacity=L.ptr(img_opacity), flow=L.ptr(img_flow), semantic=L.ptr(img_semantic),
                           radii=L.ptr(radii))
            ctx = {"device": dev, "geom": None, "binning": None, "image": None, "error": None}
            _tls.ctx = ctx
            stream = torch.cuda.current_stream(dev).cuda_stream
            try:
                rendered = lib.adgs_rasterize_forward(C.byref(cam), C.byref(g), C.byref(out), _GEOM_CB, _BINNING_CB,
                                                      _IMAGE_CB, None, stream)
            finally:
                _tls.ctx = None
            if rendered < 0 and ctx["error"] is not None:
                raise ctx["error"]
            L.check(rendered, "rasterize_gaussians")
        empty = torch.empty((0,), dtype=torch.uint8, device=dev)
        geom = ctx["geom"] if ctx["geom"] is not None else empty
        binning = ctx["binning"] if ctx["binning"] is not None else empty
        img = ctx["image"] if ctx["image"] is not None else empty
        return rendered, out_color, out_depth, img_opacity, radii, geom, binning, img, img_flow, img_semantic

    @staticmethod
    def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                     cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                     dL_dout_depth, dL_dout_flow, dL_dout_semantic, semantic, flow_points, sh, degree,
                                     campos, geomBuffer, R, binningBuffer, imageBuffer, img_opacity,
                                     grad_img_opacity, inv_depth, debug, opacities=None, needs=None):
        """`needs` (optional, 10 bools in result order) lets the autograd wrapper skip gradient
        tensors nobody consumes; omitted => all ten are produced, like the reference."""
        lib = L.load()
        dev = means3D.device
        P = means3D.shape[0]
        H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
        M = sh.shape[1] if (sh is not None and sh.numel() != 0) else 0
        D_S = semantic.shape[1] if (semantic is not None and semantic.numel() != 0) else 0
        if needs is None:
            needs = [True] * 10
        opts = dict(dtype=torch.float32, device=dev)

        def mk(i, shape):
            return torch.empty(shape, **opts) if needs[i] else None

        dL_dmeans2D = mk(0, (P, 3))
        dL_dcolors = mk(1, (P, 3))
        dL_dopacity = mk(2, (P, 1))
        dL_dmeans3D = mk(3, (P, 3))
        dL_dcov3D = mk(4, (P, 6))
        dL_dsh = mk(5, (P, M, 3))
        dL_dscales = mk(6, (P, 3))
        dL_drotations = mk(7, (P, 4))
        dL_dflow = mk(8, (P, 3))
        dL_dsem = mk(9, (P, D_S))
        if D_S > 1 and dL_dsem is None:
            dL_dsem = torch.empty((P, D_S), **opts)
        if P != 0:
            settings = GaussianRasterizationSettings(H, W, tan_fovx, tan_fovy, background, scale_modifier,
                                                     viewmatrix, projmatrix, degree, campos, False, inv_depth, debug)
            keep = []
            with torch.cuda.device(dev):
                cam = _camera(settings, keep)
                g = _gaussians(means3D, sh, colors, flow_points, semantic, opacities, scales, rotations,
                               cov3D_precomp, keep)
                dpix = [_f32(t, n) for t, n in ((dL_dout_color, "dL_dout_color"), (dL_dout_depth, "dL_dout_depth"),
                                                (dL_dout_flow, "dL_dout_flow"),
                                                (dL_dout_semantic, "dL_dout_semantic"),
                                                (grad_img_opacity, "grad_img_opacity"))]
                ig = L.ImageGrads(*[L.ptr(t) for t in dpix])
                gg = L.GaussianGrads(*[L.ptr(t) for t in (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D,
                                                          dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dflow,
                                                          dL_dsem)])
                scratch = torch.empty((lib.adgs_backward_scratch_bytes(P),), dtype=torch.uint8, device=dev)
                img_opacity = _f32(img_opacity, "img_opacity")
                radii_c = radii.contiguous()
                stream = torch.cuda.current_stream(dev).cuda_stream
                st = lib.adgs_rasterize_backward(C.byref(cam), C.byref(g), L.ptr(radii_c), L.ptr(geomBuffer), int(R),
                                                 L.ptr(binningBuffer), L.ptr(imageBuffer), L.ptr(img_opacity),
                                                 C.byref(ig), C.byref(gg), L.ptr(scratch), stream)
                L.check(st, "rasterize_gaussians_backward")
        else:
            for t in (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                      dL_drotations, dL_dflow, dL_dsem):
                if t is not None:
                    t.zero_()
        return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
                dL_dflow, dL_dsem)

    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        lib = L.load()
        P = means3D.shape[0]
        present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
        if P != 0:
            with torch.cuda.device(means3D.device):
                m, v, p = _f32(means3D, "means3D"), _f32(viewmatrix, "viewmatrix"), _f32(projmatrix, "projmatrix")
                st = lib.adgs_mark_visible(P, L.ptr(m), L.ptr(v), L.ptr(p), present.data_ptr(),
                                           torch.cuda.current_stream(means3D.device).cuda_stream)
                L.check(st, "mark_visible")
        return present


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        flow_points, semantic, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, flow_points, semantic, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, flow_points,
                semantic, raster_settings):
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, flow_points, semantic,
                raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered,
                raster_settings.inv_depth, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (num_rendered, color, depth, img_opacity, radii, geomBuffer, binningBuffer, imgBuffer, img_flow,
                 img_semantic) = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            (num_rendered, color, depth, img_opacity, radii, geomBuffer, binningBuffer, imgBuffer, img_flow,
             img_semantic) = _C.rasterize_gaussians(*args)

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, img_opacity, flow_points, semantic, opacities)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, img_opacity, img_flow, img_semantic

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_img_opacity, grad_img_flow, grad_img_semantic):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
         img_opacity, flow_points, semantic, opacities) = ctx.saved_tensors

        # inputs: means3D, means2D, sh, colors, opacities, scales, rotations, cov3D, flow, semantic, settings
        n = ctx.needs_input_grad
        # result order of the native call: means2D, colors, opacity, means3D, cov3D, sh, scales, rots, flow, sem
        needs = [n[1], n[3], n[4], n[0], n[7], n[2], n[5], n[6], n[8], n[9]]

        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, grad_out_color,
                grad_depth, grad_img_flow, grad_img_semantic, semantic, flow_points, sh, raster_settings.sh_degree,
                raster_settings.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, img_opacity,
                grad_img_opacity, raster_settings.inv_depth, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                res = _C.rasterize_gaussians_backward(*args, opacities=opacities, needs=needs)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            res = _C.rasterize_gaussians_backward(*args, opacities=opacities, needs=needs)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations, grad_flow_points, grad_semantic) = res
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, grad_flow_points, grad_semantic, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, flow_points=None, semantic=None):
        raster_settings = self.raster_settings

        if shs is not None and colors_precomp is not None:
            raise Exception('Cannot provice both shs and colors_precomp')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if flow_points is None:
            flow_points = torch.Tensor([])
        if semantic is None:
            semantic = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, flow_points, semantic, raster_settings)
