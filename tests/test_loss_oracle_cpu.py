"""CPU: the loss oracle (oracle/loss_oracle.py) against the golden vectors produced by the reference's
own utils/loss_utils.py (tests/golden/loss.npz, tests/golden/make_loss_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss.npz"))
T = lambda k: torch.tensor(G[k])


def _close(a, b, tol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("c", ["a", "b", "c"])
def test_image_depth_bce_flow_losses_match_the_reference(c):
    img = T(f"{c}_img").requires_grad_(True)
    loss = LO.image_loss(img, T(f"{c}_gt"))
    loss.backward()
    assert _close(LO.l1_loss(img, T(f"{c}_gt")).item(), G[f"{c}_l1"])
    assert _close(LO.ssim(img, T(f"{c}_gt")).item(), G[f"{c}_ssim"])
    assert _close(loss.item(), G[f"{c}_image_loss"]) and _close(img.grad, G[f"{c}_d_img"])

    pred = T(f"{c}_depth_pred").requires_grad_(True)
    dl = LO.get_depth_loss(pred, T(f"{c}_depth_gt"))
    dl.backward()
    assert _close(dl.item(), G[f"{c}_depth_loss"]) and _close(pred.grad, G[f"{c}_d_depth"])

    sem = T(f"{c}_sem").requires_grad_(True)
    ol = LO.obj_loss(sem, T(f"{c}_gt_sem"))
    ol.backward()
    assert _close(ol.item(), G[f"{c}_obj_loss"]) and _close(sem.grad, G[f"{c}_d_sem"])
    op = T(f"{c}_opac").requires_grad_(True)
    sl = LO.sky_loss(op, T(f"{c}_gt_sky"))
    sl.backward()
    assert _close(sl.item(), G[f"{c}_sky_loss"]) and _close(op.grad, G[f"{c}_d_opac_sky"])

    pts = T(f"{c}_img_flow").requires_grad_(True)
    op2 = T(f"{c}_flow_opac").requires_grad_(True)
    fl = LO.get_flow_loss(pts, [0.4, T(f"{c}_flow_K"), T(f"{c}_flow_R"), T(f"{c}_flow_T"), T(f"{c}_flow"),
                                T(f"{c}_flow_vis")], op2, dist=float(G[f"{c}_flow_dist"]))
    fl.backward()
    assert _close(fl.item(), G[f"{c}_flow_loss"])
    assert _close(pts.grad, G[f"{c}_d_img_flow"]) and _close(op2.grad, G[f"{c}_d_opac_flow"])
