"""TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's training losses that produce the
five dL/dpixel planes. Only tests/ may import this module; the product
path (adgs_b200/losses.py) never does.

Follows, line by line but device-agnostic and without the visualisation imports:
  l1_loss, gaussian, create_window, ssim, _ssim, get_depth_loss, get_flow_loss   utils/loss_utils.py:20-73,88-108
  normalized_depth_scale_and_shift, get_scaled_shifted_depth                    utils/depth_utils.py:3-45
  flow_points_project                                                            utils/flow_utils.py:5-10
  object-mask / sky binary cross entropy and the weighted total                  train.py:79-115
Pinned by tests/golden/loss.npz = outputs and autograd gradients of the reference's OWN
utils/loss_utils.py run on the CPU (tests/golden/make_loss_golden.py).
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def gaussian(window_size, sigma):
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def create_window(window_size, channel):
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11, size_average=True):
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean() if size_average else ssim_map.mean(1).mean(1).mean(1)


def normalized_depth_scale_and_shift(prediction, target, mask=None):
    if mask is None:
        mask = torch.ones_like(prediction)
    a_00 = torch.sum(mask * prediction * prediction)
    a_01 = torch.sum(mask * prediction)
    a_11 = torch.sum(mask)
    b_0 = torch.sum(mask * prediction * target)
    b_1 = torch.sum(mask * target)
    det = a_00 * a_11 - a_01 * a_01
    if det == 0:
        return 0.0, 0.0
    return (a_11 * b_0 - a_01 * b_1) / det, (-a_01 * b_0 + a_00 * b_1) / det


def get_depth_loss(pred, gt, mask=None):
    scale, shift = normalized_depth_scale_and_shift(pred, gt, mask)
    pred = scale * pred + shift
    if mask is None:
        mask = torch.ones_like(pred)
    return torch.sum(torch.abs(pred - gt) * mask) / torch.sum(mask)


def flow_points_project(flow_pts, K, R, T, dist=1e-3):
    proj = (K @ (R @ flow_pts[..., None] + T[..., None]))[..., 0]
    mask = proj[..., 2] > dist
    return proj[..., :2] / torch.clamp_min(proj[..., 2:], dist), mask


def get_flow_loss(img_flow, flow_pkg, img_opacity=None, dist=1e-3):
    _, K, R, T, flow, flow_vis = flow_pkg
    H, W = flow.shape[1:]
    vis = (flow_vis > 0.5) & (flow[0] <= W - 1.0) & (flow[0] >= 0.0) & (flow[1] <= H - 1.0) & (flow[1] >= 0.0)
    sel = torch.nonzero(vis, as_tuple=True)
    if sel[0].numel() == 0:
        return 0.0
    vis = vis.float()
    if img_opacity is not None:
        vis = vis * img_opacity
    pts = torch.permute(img_flow[:, sel[0], sel[1]], (1, 0))
    tgt = torch.permute(flow[:, sel[0], sel[1]], (1, 0))
    vis = vis[sel[0], sel[1]]
    proj, mask = flow_points_project(pts, K, R, T, dist=dist)
    vis = vis * mask.float()
    loss = torch.abs(proj - tgt) * vis[..., None]
    loss = torch.cat([loss[..., :1] / W, loss[..., 1:] / H], dim=-1)
    return torch.mean(torch.sum(loss, dim=-1))


def obj_loss(img_semantic, gt_semantic):
    """train.py:91-94"""
    pred = torch.clip(img_semantic, 1e-3, 1.0 - 1e-3)
    return F.binary_cross_entropy(pred[0], (gt_semantic > 0).float())


def sky_loss(img_opacity, gt_sky):
    """train.py:96-99"""
    pred = torch.clip(img_opacity, 1e-3, 1.0 - 1e-3)
    return F.binary_cross_entropy(1.0 - pred, gt_sky)


def image_loss(image, gt_image, lambda_dssim=0.2, lambda_l1=1.0):
    """train.py:79-80,113"""
    return (1.0 - lambda_dssim) * lambda_l1 * l1_loss(image, gt_image) + lambda_dssim * (1.0 - ssim(image, gt_image))
