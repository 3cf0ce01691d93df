"""Generates tests/golden/loss.npz by importing the REFERENCE's own utils/loss_utils.py and
utils/depth_utils.py from /root/reference (so it only runs in the build container) and calling
l1_loss / ssim / get_depth_loss / get_flow_loss and the two binary-cross-entropy terms of
train.py:94-102 on seeded CPU inputs, with torch autograd for the gradients.

One non-invasive shim: `flow_vis` (a visualisation dependency of utils/flow_utils.py:3, not
installed, not used by the losses) is replaced by an empty module.
Usage: python tests/golden/make_loss_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
fv = types.ModuleType("flow_vis")
fv.flow_to_color = None
sys.modules["flow_vis"] = fv
sys.path.insert(0, REF)
import utils.loss_utils as LU  # noqa: E402  (the reference's file, unmodified)


def case(H, W, seed, lambda_dssim=0.2, lambda_l1=1.0):
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.rand(*s, generator=g)
    out = {}
    # ---- image: (1 - l_dssim) * l_l1 * L1 + l_dssim * (1 - ssim), train.py:79-80,113 -------------------
    gt = R(3, H, W)
    img = (gt + 0.25 * torch.randn(3, H, W, generator=g)).clamp(0, 1.2).requires_grad_(True)
    l1 = LU.l1_loss(img, gt)
    ss = LU.ssim(img, gt)
    loss = (1.0 - lambda_dssim) * lambda_l1 * l1 + lambda_dssim * (1.0 - ss)
    loss.backward()
    out.update(img=img.detach(), gt=gt, l1=l1.detach(), ssim=ss.detach(), image_loss=loss.detach(), d_img=img.grad)
    # ---- depth: scale/shift-invariant L1 on inverse depth, train.py:83-86 ------------------------------
    gt_depth = R(H, W) * 2.0 + 0.1
    pred = (0.6 * gt_depth + 0.3 + 0.2 * torch.randn(H, W, generator=g)).requires_grad_(True)
    dl = LU.get_depth_loss(pred, gt_depth)
    dl.backward()
    out.update(depth_pred=pred.detach(), depth_gt=gt_depth, depth_loss=dl.detach(), d_depth=pred.grad)
    # ---- object mask / sky: clipped binary cross entropy, train.py:91-100 ------------------------------
    sem = (R(1, H, W) * 1.2 - 0.1).requires_grad_(True)          # leaves [0,1] on purpose: the clip has zero gradient
    gt_sem = (R(H, W) > 0.6).float() * 3.0                         # class ids; the loss uses (> 0)
    obj = torch.nn.functional.binary_cross_entropy(torch.clip(sem, 1e-3, 1.0 - 1e-3)[0], (gt_sem > 0).float())
    obj.backward()
    out.update(sem=sem.detach(), gt_sem=gt_sem, obj_loss=obj.detach(), d_sem=sem.grad)
    opac = (R(H, W) * 1.1 - 0.05).requires_grad_(True)
    gt_sky = (R(H, W) > 0.7).float()
    sky = torch.nn.functional.binary_cross_entropy(1.0 - torch.clip(opac, 1e-3, 1.0 - 1e-3), gt_sky)
    sky.backward()
    out.update(opac=opac.detach(), gt_sky=gt_sky, sky_loss=sky.detach(), d_opac_sky=opac.grad)
    # ---- flow: utils/loss_utils.py:get_flow_loss ---------------------------------------------------------
    focal = 0.9 * W
    K = torch.tensor([[focal, 0.0, W / 2], [0.0, focal, H / 2], [0.0, 0.0, 1.0]])
    ang = 0.05
    Rm = torch.tensor([[np.cos(ang), 0.0, np.sin(ang)], [0.0, 1.0, 0.0], [-np.sin(ang), 0.0, np.cos(ang)]],
                      dtype=torch.float32)
    T = torch.tensor([0.1, -0.05, 0.2])
    pts = torch.stack([(R(H, W) - 0.5) * 8.0, (R(H, W) - 0.5) * 3.0, R(H, W) * 20.0 - 1.0], 0)   # some behind the camera
    img_flow = pts.clone().requires_grad_(True)
    flow = torch.stack([R(H, W) * (W + 6.0) - 3.0, R(H, W) * (H + 6.0) - 3.0], 0)              # (2,H,W) target pixels, some outside
    flow_vis = R(H, W)
    opac2 = R(H, W).requires_grad_(True)
    fl = LU.get_flow_loss(img_flow, [0.4, K, Rm, T, flow, flow_vis], opac2, dist=0.02)
    fl.backward()
    out.update(flow_K=K, flow_R=Rm, flow_T=T, img_flow=pts, flow=flow, flow_vis=flow_vis, flow_opac=opac2.detach(),
               flow_loss=fl.detach(), d_img_flow=img_flow.grad, d_opac_flow=opac2.grad, flow_dist=torch.tensor(0.02))
    return {k: np.asarray(v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()}


if __name__ == "__main__":
    torch.manual_seed(0)
    data = {}
    for name, (H, W, seed) in {"a": (37, 53, 1), "b": (16, 16, 2), "c": (9, 70, 3)}.items():
        for k, v in case(H, W, seed).items():
            data[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **data)
    print("wrote", os.path.join(HERE, "loss.npz"), len(data), "arrays")
