"""Drop-in for `gaussian_renderer.render` (gaussian_renderer/__init__.py:18-115) on the fused
B200 path: one native call evaluates the trajectory at `camera.time` (and at `flow_time`),
projects, bins, sorts and blends; one native call does the whole backward down to the
parameter tensors of adgs_b200.gaussian_model.GaussianModel.

Same signature, same result-dict keys. Differences that are deliberate and documented in
DESIGN.md: the deformed tensors (`xyz`, `rotation`, `shs`) are only materialised on request
(`pipe.materialize_deformed=True`; no shipped caller reads them); `opacity` is always returned.
"""
import ctypes as C
import math
import threading
import torch

from . import _lib as L
from .gaussian_model import GaussianModel, PARAM_NAMES
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, _camera

_tls = threading.local()


def _binning_alloc(nbytes, _user):
    ctx = _tls.ctx
    ctx["binning"] = torch.empty((int(nbytes),), dtype=torch.uint8, device=ctx["device"])
    return ctx["binning"].data_ptr()


_BINNING_CB = L.ALLOC_FN(_binning_alloc)
_NULL_CB = L.ALLOC_FN()


class _FusedRender(torch.autograd.Function):
    """(screenspace_points, *model parameters) -> (color, radii, depth, img_opacity, img_flow, img_semantic, opacity)."""

    @staticmethod
    def forward(ctx, screenspace_points, xyz, scaling, rotation, opacity, sh4, shs_deform4, xyz_deform, rot_deform,
                background_deform, gs_time_sigma, model, settings, t, flow_t, render_objmask, sync_free,
                materialize):
        lib = L.load()
        dev = xyz.device
        # arena size for this forward from the counters of earlier steps; the newest `sync_free_outstanding` forwards
        # may still be in flight (the host then queues the next step while the device finishes this one)
        model._resolve_counter_checks(block=True, outstanding=int(getattr(model, "sync_free_outstanding", 0)))
        N, H, W = model.get_pts_num, int(settings.image_height), int(settings.image_width)
        o = dict(dtype=torch.float32, device=dev)
        color = torch.empty((3, H, W), **o)
        depth = torch.empty((1, H, W), **o)
        img_opacity = torch.empty((1, H, W), **o)
        img_flow = torch.empty((3, H, W), **o)
        D_S = 1 if render_objmask else 0
        img_sem = torch.empty((D_S, H, W), **o)
        radii = torch.empty((N,), dtype=torch.int32, device=dev)
        opacity_act = torch.empty((N, 1), **o)
        deformed = {"opacity": opacity_act}
        if materialize:
            deformed.update(xyz=torch.empty((N, 3), **o), rotation=torch.empty((N, 4), **o),
                            shs=torch.empty((N, 16, 3), **o), scaling=torch.empty((N, 3), **o))
            if flow_t is not None:
                deformed["flow_xyz"] = torch.empty((N, 3), **o)
        tensors = dict(zip(PARAM_NAMES, (xyz, scaling, rotation, opacity, sh4, shs_deform4, xyz_deform, rot_deform,
                                         background_deform, gs_time_sigma)))
        tensors = {k: v.contiguous() for k, v in tensors.items()}
        cm = model.c_model_from(tensors)
        tb = model.time_basis(t, flow_t)
        keep = []
        with torch.cuda.device(dev):
            cam = _camera(settings, keep)
            images = L.Images(color=L.ptr(color), depth=L.ptr(depth), opacity=L.ptr(img_opacity), flow=L.ptr(img_flow),
                              semantic=L.ptr(img_sem), radii=L.ptr(radii))
            dfm = L.Deformed(xyz=L.ptr(deformed.get("xyz")), rotation=L.ptr(deformed.get("rotation")),
                             shs=L.ptr(deformed.get("shs")), opacity=L.ptr(opacity_act),
                             scaling=L.ptr(deformed.get("scaling")), flow_xyz=L.ptr(deformed.get("flow_xyz")))
            geom = torch.empty((lib.adgs_geometry_bytes(N),), dtype=torch.uint8, device=dev)
            img = torch.empty((lib.adgs_image_bytes(W, H),), dtype=torch.uint8, device=dev)
            saved = torch.empty((lib.adgs_render_saved_bytes(N),), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            pending = None
            if sync_free and model._binning_capacity > 0:
                capacity = int(model._binning_capacity)
                binning = torch.empty((lib.adgs_binning_bytes(capacity),), dtype=torch.uint8, device=dev)
                st = lib.adgs_render_forward(C.byref(cam), C.byref(cm), C.byref(tb), int(render_objmask),
                                             C.byref(images), C.byref(dfm), L.ptr(geom), L.ptr(binning), capacity,
                                             _NULL_CB, None, L.ptr(img), L.ptr(saved), stream)
                L.check(st, "render_forward")
                counters = model._pinned_counters()
                L.check(lib.adgs_read_counters(L.ptr(geom), N, counters.data_ptr(), stream), "read_counters")
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                pending = (counters, ev)
            else:
                c = {"device": dev, "binning": None}
                _tls.ctx = c
                try:
                    R = lib.adgs_render_forward(C.byref(cam), C.byref(cm), C.byref(tb), int(render_objmask),
                                                C.byref(images), C.byref(dfm), L.ptr(geom), None, 0, _BINNING_CB,
                                                None, L.ptr(img), L.ptr(saved), stream)
                finally:
                    _tls.ctx = None
                L.check(R, "render_forward")
                binning, capacity = c["binning"], int(R)
                model._note_num_rendered(int(R))
        ctx.model, ctx.settings, ctx.tb, ctx.render_objmask = model, settings, tb, render_objmask
        ctx.capacity, ctx.pending = capacity, pending
        ctx.save_for_backward(*[tensors[k] for k in PARAM_NAMES], radii, geom, binning, img, saved, img_opacity)
        ctx.mark_non_differentiable(radii)
        ctx.deformed = deformed
        extra = tuple(deformed[k] for k in ("xyz", "rotation", "shs", "scaling", "flow_xyz") if k in deformed) \
            if materialize else ()
        for e in extra:
            ctx.mark_non_differentiable(e)
        ctx.mark_non_differentiable(opacity_act)
        return (color, radii, depth, img_opacity, img_flow, img_sem, opacity_act) + extra

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_opacity, g_flow, g_sem, g_opact, *g_extra):
        lib = L.load()
        model, settings, tb = ctx.model, ctx.settings, ctx.tb
        saved_t = ctx.saved_tensors
        tensors = dict(zip(PARAM_NAMES, saved_t[:len(PARAM_NAMES)]))
        radii, geom, binning, img, saved, img_opacity = saved_t[len(PARAM_NAMES):]
        dev = tensors["xyz"].device
        N = model.get_pts_num
        # Gradient storage: fresh tensors, or -- when a multi-GPU step installed a sink -- fresh views of
        # its flat all-reduce bucket, so that autograd adopts them without a copy (parallel.py).
        sink = getattr(model, "_grad_sink", None)
        grads = sink() if sink is not None else {k: torch.empty_like(v) for k, v in tensors.items()}
        d_means2D = torch.empty((N, 3), dtype=torch.float32, device=dev)
        # Sync-free forward: {num_rendered, overflow} is on its way to pinned memory. The backward does not
        # wait for it -- the blend backward reads the overflow flag on the device and leaves every gradient
        # zero if the arena was too small -- so the host never blocks inside a step; the counters are looked
        # at when they have arrived (here, if the forward is already done, else before the next forward).
        if ctx.pending is not None:
            model._defer_counter_check(ctx.pending[0], ctx.pending[1], ctx.capacity)
            model._resolve_counter_checks(block=False)
        keep = []
        with torch.cuda.device(dev):
            cam = _camera(settings, keep)
            cm = model.c_model_from(tensors)
            gm = model.c_model_from(grads, with_time=False)
            cot = [None if g is None else g.contiguous() for g in (g_color, g_depth, g_flow, g_sem, g_opacity)]
            ig = L.ImageGrads(dL_dcolor=L.ptr(cot[0]), dL_ddepth=L.ptr(cot[1]), dL_dflow=L.ptr(cot[2]),
                              dL_dsemantic=L.ptr(cot[3]) if ctx.render_objmask else None,
                              dL_dopacity=L.ptr(cot[4]))
            scratch = torch.empty((lib.adgs_render_scratch_bytes(N, model.n_obj),), dtype=torch.uint8, device=dev)
            # window-aware optimizer: inactive control-point planes stay unwritten (no zero-fill); the
            # optimizer reads only the columns recorded here
            sparse = model.sparse_deform_grads and sink is None
            if sparse and model.__dict__.get("_active_cols", {}).get("backwards", 0) > 0:
                raise RuntimeError("window-aware FusedAdam supports one render backward per optimizer step")
            model._note_active_columns(tb)
            tb.sparse_grads = int(sparse)
            st = lib.adgs_render_backward(C.byref(cam), C.byref(cm), C.byref(tb), int(ctx.render_objmask),
                                          L.ptr(radii), L.ptr(geom), L.ptr(binning), int(ctx.capacity), L.ptr(img),
                                          L.ptr(saved), L.ptr(img_opacity), C.byref(ig), C.byref(gm),
                                          L.ptr(d_means2D), L.ptr(scratch),
                                          torch.cuda.current_stream(dev).cuda_stream)
            L.check(st, "render_backward")
        return (d_means2D,) + tuple(grads[k] for k in PARAM_NAMES) + (None,) * 7


def render(viewpoint_camera, pc: GaussianModel, env_map, pipe, scaling_modifier=1.0, override_color=None,
           flow_pkg=None, render_objmask=False):
    """Render the scene (same contract as gaussian_renderer/__init__.py:18-115)."""
    # zero tensor whose .grad receives the screen-space mean gradients (gaussian_renderer/__init__.py:26-30);
    # a leaf, so no retain_grad() and no extra "+ 0" kernel is needed
    screenspace_points = torch.zeros((pc.get_pts_num, 3), dtype=torch.float32, device=pc.xyz.device,
                                     requires_grad=True)

    dev = pc.xyz.device
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=_zeros3(dev), scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform.to(dev),
        projmatrix=viewpoint_camera.full_proj_transform.to(dev), sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center.to(dev), prefiltered=False,
        inv_depth=getattr(pipe, "inv_depth", False), debug=getattr(pipe, "debug", False))

    flow_time = None
    if flow_pkg is not None:
        flow_time = flow_pkg[0]
    t = viewpoint_camera.time

    if override_color is not None:
        # Rare visualisation path (render.py): trajectory values + the strict drop-in rasterizer.
        pkg = pc.get_deformed_pkg(t)
        flow_points = pc.get_deformed_xyz(flow_time) if flow_time is not None else None
        semantic = pc.get_obj_mask.float()[..., None] if render_objmask else None
        foreground, radii, depth, img_opacity, img_flow, img_semantic = GaussianRasterizer(raster_settings)(
            means3D=pkg['xyz'], means2D=screenspace_points, shs=None, colors_precomp=override_color,
            opacities=pkg['opacity'], scales=pc.get_scaling, rotations=pkg['rotation'], flow_points=flow_points,
            semantic=semantic)
        deform_pkg = pkg
        opacity = pkg['opacity']
    else:
        materialize = bool(getattr(pipe, "materialize_deformed", False))
        sync_free = bool(getattr(pipe, "sync_free", True)) and torch.is_grad_enabled()
        out = _FusedRender.apply(screenspace_points, *pc.hot_parameters(), pc, raster_settings, float(t),
                                 None if flow_time is None else float(flow_time), bool(render_objmask), sync_free,
                                 materialize)
        foreground, radii, depth, img_opacity, img_flow, img_semantic, opacity = out[:7]
        deform_pkg = {'opacity': opacity}
        if materialize:
            # pipe.materialize_deformed: the tensors the reference's get_deformed_pkg / get_deformed_xyz(flow_time) /
            # get_scaling would have produced, exactly as the fused kernel used them (tests/test_fused_gpu.py)
            deform_pkg.update(xyz=out[7], rotation=out[8], shs=out[9], scaling=out[10])
            if flow_time is not None:
                deform_pkg["flow_xyz"] = out[11]

    if env_map is not None and hasattr(env_map, "composite"):
        # adgs_b200.env.EnvironmentMap: background + blend in one kernel (gaussian_renderer/__init__.py:92-94)
        rendered_image, background = env_map.composite(foreground, img_opacity, viewpoint_camera)
    elif env_map is not None:
        background = env_map.get_image_background(viewpoint_camera)
        rendered_image = foreground + (1.0 - img_opacity) * background
    else:
        background = torch.zeros_like(foreground)
        rendered_image = foreground

    res = {
        "render": rendered_image,
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
        "depth": depth.squeeze(0),
        "opacity": opacity,
        'img_opacity': img_opacity.squeeze(0),
        'foreground': foreground,
        'background': background,
        'img_flow': img_flow if flow_time is not None else None,
        'img_semantic': img_semantic if render_objmask else None,
    }
    res.update(deform_pkg)
    return res


_ZEROS3 = {}


def _zeros3(dev):
    key = str(dev)
    if key not in _ZEROS3:
        _ZEROS3[key] = torch.zeros(3, dtype=torch.float32, device=dev)
    return _ZEROS3[key]
