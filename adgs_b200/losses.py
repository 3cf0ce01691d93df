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
