"""Loss front-end of the training iteration: drop-in for the image
term of `train.py:79-80,113` -- `utils/loss_utils.py:l1_loss` (:20-21) and `ssim` (:35-58) -- on two
fused CUDA kernels (adgs_b200/csrc/loss.cu) instead of five grouped 11x11 convolutions plus a dozen
element-wise kernels and their autograd mirror images.

    from adgs_b200.losses import l1_loss, ssim, image_loss
    Ll1, s = l1_loss(image, gt), ssim(image, gt)            # same call sites as the reference
    loss = image_loss(image, gt, opt.lambda_dssim, opt.lambda_l1)   # both terms: one forward + one backward launch

Gradients flow to the first argument only (the rendered image), like in training where the ground
truth is a constant. No fallback: the functions raise if the CUDA library is missing.
"""
import ctypes as C

import torch

from . import _lib as L


def _check(img, gt):
    if img.dim() != 3 or img.shape != gt.shape:
        raise RuntimeError("image loss expects two (C,H,W) tensors of the same shape")
    if not img.is_cuda or img.dtype != torch.float32 or gt.dtype != torch.float32:
        raise RuntimeError("image loss expects float32 CUDA tensors")


def _forward(ctx, img, gt, w_l1, w_dssim):
    lib = L.load()
    _check(img, gt)
    img_c, gt_c = img.contiguous(), gt.detach().contiguous()
    C_, H, W = img_c.shape
    dev = img_c.device
    need = ctx.needs_input_grad[0]
    planes = torch.empty((3, C_, H, W), dtype=torch.float32, device=dev) if need else None
    partial = torch.empty((lib.adgs_image_loss_partial_floats(C_, H, W),), dtype=torch.float32, device=dev)
    out = torch.empty((3,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = lib.adgs_image_loss_forward(C_, H, W, img_c.data_ptr(), gt_c.data_ptr(),
                                         planes[0].data_ptr() if need else None,
                                         planes[1].data_ptr() if need else None,
                                         planes[2].data_ptr() if need else None, partial.data_ptr(),
                                         float(w_l1), float(w_dssim), out.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream)
    L.check(st, "image_loss_forward")
    if need:
        ctx.save_for_backward(img_c, gt_c, planes)
    return out


def _backward(ctx, g_l1, w_l1, g_ssim, w_ssim):
    lib = L.load()
    img_c, gt_c, planes = ctx.saved_tensors
    C_, H, W = img_c.shape
    dev = img_c.device
    g_l1 = g_l1.to(torch.float32).contiguous()
    g_ssim = g_ssim.to(torch.float32).contiguous()
    d_img = torch.empty_like(img_c)
    with torch.cuda.device(dev):
        st = lib.adgs_image_loss_backward(C_, H, W, img_c.data_ptr(), gt_c.data_ptr(), planes[0].data_ptr(),
                                          planes[1].data_ptr(), planes[2].data_ptr(), g_l1.data_ptr(), float(w_l1),
                                          g_ssim.data_ptr(), float(w_ssim), d_img.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
    L.check(st, "image_loss_backward")
    return d_img


class _L1AndSsim(torch.autograd.Function):
    """(img (C,H,W), gt (C,H,W)) -> tensor (2,) = (mean |img - gt|, mean ssim(img, gt))."""

    @staticmethod
    def forward(ctx, img, gt):
        return _forward(ctx, img, gt, 0.0, 0.0)[:2]

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return _backward(ctx, g[0:1], 1.0, g[1:2], 1.0), None


class _WeightedImageLoss(torch.autograd.Function):
    """w_l1 * mean |img - gt| + w_dssim * (1 - mean ssim) as one 0-d tensor."""

    @staticmethod
    def forward(ctx, img, gt, w_l1, w_dssim):
        ctx.w = (float(w_l1), float(w_dssim))
        return _forward(ctx, img, gt, w_l1, w_dssim)[2]

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(1)
        return _backward(ctx, g, ctx.w[0], g, -ctx.w[1]), None, None, None


def l1_and_ssim(img, gt):
    """Both scalars from one launch pair."""
    out = _L1AndSsim.apply(img, gt)
    return out[0], out[1]


def l1_loss(network_output, gt):
    """utils/loss_utils.py:20-21 for (C,H,W) images."""
    return _L1AndSsim.apply(network_output, gt)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:35-58 (window 11, size_average=True: the only form the reference calls, train.py:80)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("adgs_b200.losses.ssim implements the reference's training call: window_size=11, "
                                  "size_average=True")
    return _L1AndSsim.apply(img1, img2)[1]


def image_loss(image, gt_image, lambda_dssim=0.2, lambda_l1=1.0):
    """(1 - lambda_dssim) * lambda_l1 * L1 + lambda_dssim * (1 - ssim), train.py:79-80,113."""
    return _WeightedImageLoss.apply(image, gt_image, (1.0 - lambda_dssim) * lambda_l1, lambda_dssim)


# ---- per-pixel terms: depth, object mask, sky, flow (train.py:82-100) ---------------------------------------

class _PixelLosses(torch.autograd.Function):
    """(depth (H,W), img_semantic (1,H,W), img_opacity (H,W), img_flow (3,H,W)) ->
    tensor (6,) = (depth, obj, sky, flow losses, lambda-weighted sum, selected flow pixels).
    Gradients flow from element 4 (the weighted sum) only."""

    @staticmethod
    def forward(ctx, depth, img_semantic, img_opacity, img_flow, targets, lambdas):
        lib = L.load()
        ref = next(t for t in (depth, img_semantic, img_opacity, img_flow) if t is not None)
        dev = ref.device
        H, W = int(ref.shape[-2]), int(ref.shape[-1])
        f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep = dict(depth=f32(depth), sem=f32(img_semantic), opac=f32(img_opacity), flow_pts=f32(img_flow),
                    gt_depth=f32(targets.get("gt_depth")), gt_sem=f32(targets.get("gt_semantic")),
                    gt_sky=f32(targets.get("gt_sky")))
        for k in ("depth", "sem", "opac", "flow_pts", "gt_depth", "gt_sem", "gt_sky"):
            t = keep[k]
            if t is not None and (t.shape[-2] != H or t.shape[-1] != W):
                raise RuntimeError(f"pixel losses: {k} has shape {tuple(t.shape)}, expected (..., {H}, {W})")
        inp = L.PixelLossInputs(H=H, W=W)
        use_depth = keep["gt_depth"] is not None and keep["depth"] is not None and lambdas.get("depth", 0.0) > 0.0
        use_obj = keep["gt_sem"] is not None and keep["sem"] is not None and lambdas.get("obj", 0.0) > 0.0
        use_sky = keep["gt_sky"] is not None and keep["opac"] is not None and lambdas.get("sky", 0.0) > 0.0
        flow_pkg = targets.get("flow_pkg")
        use_flow = flow_pkg is not None and keep["flow_pts"] is not None and lambdas.get("flow", 0.0) > 0.0
        if use_depth:
            inp.depth, inp.gt_depth = keep["depth"].data_ptr(), keep["gt_depth"].data_ptr()
        if use_obj:
            inp.img_semantic, inp.gt_semantic = keep["sem"].data_ptr(), keep["gt_sem"].data_ptr()
        if use_sky:
            inp.img_opacity, inp.gt_sky = keep["opac"].data_ptr(), keep["gt_sky"].data_ptr()
        if use_flow:
            _, K, R, T, flow, flow_vis = flow_pkg
            keep["flow"], keep["flow_vis"] = f32(flow), f32(flow_vis)
            inp.img_flow, inp.flow, inp.flow_vis = keep["flow_pts"].data_ptr(), keep["flow"].data_ptr(), keep["flow_vis"].data_ptr()
            if keep["opac"] is not None:
                inp.flow_opacity = keep["opac"].data_ptr()
            # small host matrices (the reference keeps K, R, T as 3x3 / 3 tensors in flow_pkg)
            for name, src, n in (("K", K, 9), ("R", R, 9), ("T", T, 3)):
                vals = [float(x) for x in torch.as_tensor(src).detach().reshape(-1).cpu().tolist()]
                if len(vals) != n:
                    raise RuntimeError(f"pixel losses: flow_pkg {name} must have {n} elements")
                arr = getattr(inp, name)
                for i, x in enumerate(vals):
                    arr[i] = x
            inp.flow_dist = float(targets.get("flow_dist", 1e-3))
        inp.lambda_depth = float(lambdas.get("depth", 0.0)) if use_depth else 0.0
        inp.lambda_obj = float(lambdas.get("obj", 0.0)) if use_obj else 0.0
        inp.lambda_sky = float(lambdas.get("sky", 0.0)) if use_sky else 0.0
        inp.lambda_flow = float(lambdas.get("flow", 0.0)) if use_flow else 0.0
        scratch = torch.empty((lib.adgs_pixel_loss_scratch_bytes(H, W) // 8 + 1,), dtype=torch.float64, device=dev)
        out = torch.empty((6,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = lib.adgs_pixel_loss(C.byref(inp), 1, scratch.data_ptr(), None, None, None, None, None, out.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "pixel_loss")
        ctx.inp, ctx.keep, ctx.scratch, ctx.out = inp, keep, scratch, out
        ctx.shapes = tuple(None if t is None else tuple(t.shape) for t in (depth, img_semantic, img_opacity, img_flow))
        ctx.used = (use_depth, use_obj, use_sky or (use_flow and keep["opac"] is not None), use_flow)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        inp, dev = ctx.inp, ctx.out.device
        need = ctx.needs_input_grad[:4]
        g_total = g[4:5].to(torch.float32).contiguous()
        planes = []
        for shape, used, nd in zip(ctx.shapes, ctx.used, need):
            planes.append(torch.empty(shape, dtype=torch.float32, device=dev) if (shape is not None and used and nd)
                          else None)
        with torch.cuda.device(dev):
            st = lib.adgs_pixel_loss(C.byref(inp), 2, ctx.scratch.data_ptr(), g_total.data_ptr(), L.ptr(planes[0]),
                                     L.ptr(planes[1]), L.ptr(planes[2]), L.ptr(planes[3]), ctx.out.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "pixel_loss backward")
        grads = []
        for shape, p, nd in zip(ctx.shapes, planes, need):
            if shape is None or not nd:
                grads.append(None)
            else:
                grads.append(p if p is not None else torch.zeros(shape, dtype=torch.float32, device=dev))
        return tuple(grads) + (None, None)


def pixel_losses(depth=None, img_semantic=None, img_opacity=None, img_flow=None, gt_depth=None, gt_semantic=None,
                 gt_sky=None, flow_pkg=None, flow_dist=1e-3, lambda_depth=0.0, lambda_obj=0.0, lambda_sky=0.0,
                 lambda_flow=0.0):
    """The per-pixel terms of train.py:82-100 in one call; a term is active when its lambda is > 0 and its
    prediction and target are given (the reference's `if opt.lambda_* > 0.0` guards). Returns a dict with
    'depth_loss', 'obj_loss', 'sky_loss', 'flow_loss' (detached values, for logging) and 'weighted' =
    lambda_depth * depth + lambda_obj * obj + lambda_sky * sky + lambda_flow * flow (differentiable)."""
    out = _PixelLosses.apply(depth, img_semantic, img_opacity, img_flow,
                             dict(gt_depth=gt_depth, gt_semantic=gt_semantic, gt_sky=gt_sky, flow_pkg=flow_pkg,
                                  flow_dist=flow_dist),
                             dict(depth=lambda_depth, obj=lambda_obj, sky=lambda_sky, flow=lambda_flow))
    d = out.detach()
    return {"depth_loss": d[0], "obj_loss": d[1], "sky_loss": d[2], "flow_loss": d[3], "weighted": out[4],
            "flow_pixels": d[5]}


def get_depth_loss(pred, gt, mask=None):
    """utils/loss_utils.py:60-65 (the training call passes no mask, train.py:86)."""
    if mask is not None:
        raise NotImplementedError("adgs_b200.losses.get_depth_loss implements the training call: mask=None")
    return _PixelLosses.apply(pred, None, None, None, dict(gt_depth=gt), dict(depth=1.0))[4]


def get_flow_loss(img_flow, flow_pkg, img_opacity=None, dist=1e-3):
    """utils/loss_utils.py:88-108. Returns a 0-d tensor (0 when no pixel is selected, where the reference
    returns the python float 0.0) -- and never synchronises with the host."""
    return _PixelLosses.apply(None, None, img_opacity, img_flow, dict(flow_pkg=flow_pkg, flow_dist=dist),
                              dict(flow=1.0))[4]


def near_reg_loss(model):
    """train.py:101-103: mean over anchors and axes of the variance (over the K near neighbours, unbiased) of the
    position control points, summed over the control-point axis -- on the planar (Cx, 3, N_obj) array:
    `xyz_deform_param[obj_near_idx]` (P,K,3,C) -> var(dim=1).sum(-1).mean()  ==  xyz_deform[:, :, idx] (C,3,P,K)
    -> var(dim=-1).sum(0).mean(). Plain torch on the device (a gather + reduction, differentiable);
    `obj_near_idx` comes from adgs_b200.densify.set_obj_near_idx."""
    idx = model.obj_near_idx
    return torch.mean(torch.sum(torch.var(model.xyz_deform[:, :, idx], dim=-1), dim=0))


def near_sigma_reg_loss(model):
    """train.py:108-110: `gs_time_sigma[obj_near_idx]` (P,K,2) -> var(dim=1).sum(-1).mean()."""
    return torch.mean(torch.sum(torch.var(model.gs_time_sigma[model.obj_near_idx], dim=1), dim=-1))
