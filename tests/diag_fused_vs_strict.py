"""Diagnostic (not a test): which fields of the per-Gaussian state differ between the fused and the strict front end."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import helpers as Hh
import test_fused_gpu as TF
from adgs_b200.gaussian_model import PARAM_NAMES

model, c, _ = TF._bench_scene(TF.FULL["small"])
P, W, H = c["n"], c["W"], c["H"]
res = TF._fused(model, c, 0.37, 0.41)
node = res["foreground"].grad_fn
saved = node.saved_tensors
geom, binning, img = saved[len(PARAM_NAMES) + 1: len(PARAM_NAMES) + 4]
sc = TF._strict_case(c, res, model)
out = Hh.OURS.rasterize_gaussians(*Hh.fwd_args(sc))
R = out[0]
io = Hh.inspect_ours(geom, binning, img, P, R, W, H)
ist = Hh.inspect_ours(out[5], out[6], out[7], P, R, W, H)
vis = res["radii"] > 0
names = ["x", "y", "A", "B", "C", "op", "depth", "ext", "r", "g", "b", "dfeat", "f0", "f1", "f2", "sem"]
for i, nm in enumerate(names):
    a, b = io["record"][vis, i], ist["record"][vis, i]
    nd = (a.view(torch.int32) != b.view(torch.int32)).sum().item()
    print(f"{nm:6s} bit mismatches {nd:6d} of {int(vis.sum())}   max abs diff {(a - b).abs().max().item():.3e}")
print("cov3D mismatches", (io["cov3D"][vis].view(torch.int32) != ist["cov3D"][vis].view(torch.int32)).sum().item())
print("n_contrib mismatches", (io["n_contrib"] != ist["n_contrib"]).sum().item())
