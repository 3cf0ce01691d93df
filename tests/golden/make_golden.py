"""Generates tests/golden/raster_*.npz by running the UNMODIFIED reference rasterizer
(oracle/_ref/libadgs_ref.so, built from /root/reference by oracle/Makefile) on seeded inputs.
Must run on a GPU box:  python tests/golden/make_golden.py [outdir]
The fixtures hold inputs AND reference outputs (forward images, radii, internal binning state,
all ten gradients) so CPU-only tests can pin oracle/raster_oracle.py against the real reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as Hh  # noqa: E402
from oracle import ref_module as REF  # noqa: E402

CASES = {
    "raster_a": dict(n=400, W=64, H=48, seed=101, sh_degree=3, inv_depth=True, flow=True, D_S=1,
                     median_radius_px=5.0),
    "raster_b": dict(n=300, W=53, H=37, seed=102, sh_degree=1, inv_depth=False, flow=False, D_S=0,
                     bg=(0.3, 0.5, 0.7), median_radius_px=8.0),
    "raster_c": dict(n=300, W=48, H=32, seed=103, sh_degree=2, inv_depth=True, flow=True, D_S=3,
                     median_radius_px=4.0, yaw_deg=10.0),
    "raster_d": dict(n=250, W=48, H=32, seed=104, colors_precomp=True, cov3D_precomp=True, D_S=1,
                     median_radius_px=6.0),
}

GRADS = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
         "dL_drotations", "dL_dflow_points", "dL_dsemantic"]


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else HERE
    os.makedirs(outdir, exist_ok=True)
    for name, kw in CASES.items():
        c = Hh.make_case(**kw)
        fwd = REF.rasterize_gaussians(*Hh.fwd_args(c))
        R = fwd[0]
        ins = REF.inspect(fwd[5], fwd[6], fwd[7], c["n"], R, c["W"], c["H"])
        cot = Hh.cotangents(c)
        grads = REF.rasterize_gaussians_backward(*Hh.bwd_args(c, fwd, cot))
        d = {"kw": np.array(repr(kw))}
        cam = c["cam"]
        for k in ("background", "means3D", "opacity", "scales", "rotations", "sh", "colors", "cov3D_precomp",
                  "flow_points", "semantic"):
            d["in_" + k] = c[k].detach().cpu().numpy()
        d["in_viewmatrix"] = cam.world_view_transform.contiguous().cpu().numpy()
        d["in_projmatrix"] = cam.full_proj_transform.contiguous().cpu().numpy()
        d["in_campos"] = cam.camera_center.cpu().numpy()
        d["in_scalars"] = np.array([c["tan_fovx"], c["tan_fovy"], c["scale_modifier"], c["H"], c["W"], c["degree"],
                                    int(c["inv_depth"])], dtype=np.float64)
        for k, v in cot.items():
            d["cot_" + k] = v.cpu().numpy()
        d["num_rendered"] = np.array(R)
        for k, i in (("color", 1), ("depth", 2), ("opacity", 3), ("radii", 4), ("flow", 8), ("semantic", 9)):
            d["out_" + k] = fwd[i].cpu().numpy()
        for k in ("tiles_touched", "point_offsets", "point_list", "point_list_keys", "n_contrib", "ranges", "means2D",
                  "depths", "conic_opacity"):
            if k in ins:
                d["state_" + k] = ins[k].cpu().numpy()
        for k, v in zip(GRADS, grads):
            d["grad_" + k] = v.cpu().numpy()
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "R", R, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
