"""GPU parity of the fused loss front-end (adgs_b200/losses.py -> adgs_image_loss_*) against (1) the golden
vectors produced by the reference's own utils/loss_utils.py (tests/golden/loss.npz) and (2) the torch
restatement (oracle/loss_oracle.py) at the KITTI frame size. Tolerance 1e-5 relative on the scalars, 1e-4 of
the gradient plane's maximum (fp32; the separable window sums in a different order than a 2-D convolution)."""
import os

import numpy as np
import pytest
import torch

from adgs_b200 import losses as LS

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss.npz"))
T = lambda k: torch.tensor(G[k], device="cuda")


def _rel(a, b):
    """max |a - b| relative to max |b|, after an absolute allowance of 5e-7 (fp32 noise around exact zeros)."""
    a, b = a.detach().double().cpu().numpy(), np.asarray(b, dtype=np.float64)
    return max(np.abs(a - b).max() - 5e-7, 0.0) / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("c", ["a", "b", "c"])
def test_image_loss_matches_reference_golden(c):
    img = T(f"{c}_img").requires_grad_(True)
    gt = T(f"{c}_gt")
    l1, ss = LS.l1_and_ssim(img, gt)
    assert _rel(l1, G[f"{c}_l1"]) <= 1e-5 and _rel(ss, G[f"{c}_ssim"]) <= 1e-5
    loss = LS.image_loss(img, gt, 0.2, 1.0)
    loss.backward()
    assert _rel(loss, G[f"{c}_image_loss"]) <= 1e-5
    assert _rel(img.grad, G[f"{c}_d_img"]) <= 1e-4
    # the reference's separate call sites
    assert _rel(LS.l1_loss(img, gt), G[f"{c}_l1"]) <= 1e-5 and _rel(LS.ssim(img, gt), G[f"{c}_ssim"]) <= 1e-5


@pytest.mark.parametrize("H,W", [(375, 1242), (1, 1), (11, 5), (33, 16)])
def test_image_loss_matches_oracle_at_size(H, W):
    from oracle import loss_oracle as LO
    g = torch.Generator(device="cuda").manual_seed(H * 7 + W)
    gt = torch.rand(3, H, W, generator=g, device="cuda")
    base = (gt + 0.2 * torch.randn(3, H, W, generator=g, device="cuda")).clamp(0, 1)
    base[:, ::7, ::5] = gt[:, ::7, ::5]                      # exact ties: sign(0) = 0 like torch.abs' backward
    a = base.clone().requires_grad_(True)
    b = base.clone().requires_grad_(True)
    w = (torch.rand(2, generator=g, device="cuda") + 0.5)
    l1, ss = LS.l1_and_ssim(a, gt)
    (w[0] * l1 - w[1] * ss).backward()
    (w[0] * LO.l1_loss(b, gt) - w[1] * LO.ssim(b, gt)).backward()
    assert _rel(l1, LO.l1_loss(b, gt).item()) <= 1e-5 and _rel(ss, LO.ssim(b, gt).item()) <= 1e-5
    assert _rel(a.grad, b.grad.cpu().numpy()) <= 1e-4


def test_image_loss_without_grad_and_bad_shapes():
    gt = torch.rand(3, 20, 30, device="cuda")
    with torch.no_grad():
        l1, ss = LS.l1_and_ssim(gt * 0.5, gt)
    assert torch.isfinite(l1) and torch.isfinite(ss)
    same = LS.l1_and_ssim(gt, gt)
    assert float(same[0]) == 0.0 and abs(float(same[1]) - 1.0) <= 1e-6
    with pytest.raises(RuntimeError):
        LS.l1_and_ssim(gt[0], gt)
    with pytest.raises(NotImplementedError):
        LS.ssim(gt, gt, window_size=7)


def test_fused_image_loss_is_faster_than_the_torch_composition():
    """Timing beside the reference's own formulation (the oracle = utils/loss_utils.py restated) at the KITTI
    frame size; the numbers are written to gpurun_out/loss_timing.json for profiles/."""
    import json
    from oracle import loss_oracle as LO
    H, W = 375, 1242
    gt = torch.rand(3, H, W, device="cuda")
    x = (gt + 0.1 * torch.randn_like(gt)).requires_grad_(True)

    def run(fn, n=30):
        for _ in range(5):
            x.grad = None
            fn().backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            x.grad = None
            fn().backward()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_fused = run(lambda: LS.image_loss(x, gt, 0.2, 1.0))
    ms_torch = run(lambda: LO.image_loss(x, gt, 0.2, 1.0))
    out = {"H": H, "W": W, "fused_ms_fwd_bwd": round(ms_fused, 4), "torch_ms_fwd_bwd": round(ms_torch, 4),
           "speedup": round(ms_torch / ms_fused, 2)}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.isdir(os.path.join(root, "gpurun_out")):
        with open(os.path.join(root, "gpurun_out", "loss_timing.json"), "w") as f:
            json.dump(out, f)
    assert ms_fused < ms_torch, out
