"""GPU diagnostic: libadgs_b200 vs the reference built in oracle/_ref vs the numpy oracle.
Prints mismatch statistics instead of asserting (the asserting versions live in tests/).
Lives under tests/ (it imports oracle/). Usage (on the GPU box): python tests/gpu_diag.py [--big]
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as Hh  # noqa: E402
from oracle import ref_module as REF  # noqa: E402
from oracle import raster_oracle as O  # noqa: E402

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations", "dL_dflow_points", "dL_dsemantic"]


def compare_case(name, c, check_oracle=False, backward=True):
    print(f"\n=== {name}: n={c['n']} {c['W']}x{c['H']} deg={c['degree']} inv_depth={c['inv_depth']} "
          f"flow={c['flow_points'].numel() > 0} D_S={c['semantic'].shape[1] if c['semantic'].numel() else 0}")
    P, W, H = c["n"], c["W"], c["H"]
    ours = Hh.OURS.rasterize_gaussians(*Hh.fwd_args(c))
    torch.cuda.synchronize()
    ref = REF.rasterize_gaussians(*Hh.fwd_args(c))
    R_o, R_r = ours[0], ref[0]
    print(f"num_rendered ours={R_o} ref={R_r}")
    io = Hh.inspect_ours(ours[5], ours[6], ours[7], P, R_o, W, H)
    ir = REF.inspect(ref[5], ref[6], ref[7], P, R_r, W, H)
    radii_o, radii_r = ours[4], ref[4]
    print("radii mismatches:", (radii_o != radii_r).sum().item(), "of", P, " visible:", (radii_r > 0).sum().item())
    print("tiles_touched mismatches:", (io["tiles_touched"] != ir["tiles_touched"]).sum().item())
    vis = radii_r > 0
    rec = io["record"]
    both = vis & (radii_o > 0)
    if both.any():
        print("means2D max abs diff:", (rec[both, 0:2] - ir["means2D"][both]).abs().max().item())
        print("depth bit mismatches:", (rec[both, 6].view(torch.int32) != ir["depths"][both].view(torch.int32)).sum().item())
        con_o = torch.stack([rec[both, 2], rec[both, 3], rec[both, 4], rec[both, 5]], 1)
        print("conic_opacity bit mismatches:", (con_o.view(torch.int32) != ir["conic_opacity"][both].view(torch.int32)).sum().item(),
              " rel err:", Hh.rel_err(con_o, ir["conic_opacity"][both]))
        if c["sh"].numel():
            print("rgb rel err:", Hh.rel_err(rec[both, 8:11], ir["rgb"][both]),
                  " rgb bit mismatches:", (rec[both, 8:11].contiguous().view(torch.int32) != ir["rgb"][both].contiguous().view(torch.int32)).sum().item())
            cl_o = io["clamped"][both]
            cl_r = ir["clamped"][both]
            cl_r_bits = (cl_r[:, 0].int() | (cl_r[:, 1].int() << 1) | (cl_r[:, 2].int() << 2))
            print("clamped mismatches:", (cl_o.int() != cl_r_bits).sum().item())
        if ir.get("cov3D") is not None and c["scales"].numel():
            print("cov3D bit mismatches:", (io["cov3D"][both].contiguous().view(torch.int32) != ir["cov3D"][both].contiguous().view(torch.int32)).sum().item())
    if R_o == R_r and R_o > 0:
        keys_o = (io["point_list_tile"].long() << 32) | (rec[io["point_list"].long(), 6].view(torch.int32).long() & 0xFFFFFFFF)
        print("sorted key mismatches:", (keys_o != ir["point_list_keys"]).sum().item(), "of", R_o)
        print("point_list mismatches:", (io["point_list"] != ir["point_list"]).sum().item())
    print("ranges mismatches:", (io["ranges"] != ir["ranges"]).sum().item())
    print("n_contrib mismatches:", (io["n_contrib"] != ir["n_contrib"]).sum().item(), "of", W * H)
    for nm, i in (("color", 1), ("depth", 2), ("img_opacity", 3), ("img_flow", 8), ("img_semantic", 9)):
        a, b = ours[i], ref[i]
        if a.numel():
            print(f"{nm}: rel err {Hh.rel_err(a, b):.3e}  max abs {(a - b).abs().max().item():.3e} bit-equal {(a == b).all().item()}")
    if backward:
        cot = Hh.cotangents(c)
        go = Hh.OURS.rasterize_gaussians_backward(*Hh.bwd_args(c, ours, cot), opacities=c["opacity"])
        torch.cuda.synchronize()
        gr = REF.rasterize_gaussians_backward(*Hh.bwd_args(c, ref, cot))
        for nm, a, b in zip(GRAD_NAMES, go, gr):
            if a is not None and a.numel():
                print(f"{nm}: rel err {Hh.rel_err(a, b):.3e}  max|ref| {b.abs().max().item():.3e}")
    if check_oracle:
        s = Hh.oracle_settings(c)
        n = Hh.to_np
        t0 = time.time()
        out, st = O.rasterize_forward(s, n(c["means3D"]), n(c["opacity"]), n(c["scales"]), n(c["rotations"]),
                                      n(c["cov3D_precomp"]), n(c["sh"]), n(c["colors"]), n(c["flow_points"]),
                                      n(c["semantic"]))
        print(f"[oracle fwd {time.time() - t0:.1f}s] R={st['num_rendered']}")
        print("oracle radii mismatches vs ref:", (torch.tensor(out["radii"]).cuda() != radii_r).sum().item())
        print("oracle n_contrib mismatches vs ref:", (torch.tensor(out["n_contrib"].astype(np.int32)).cuda() != ir["n_contrib"]).sum().item())
        for nm, i in (("color", 1), ("depth", 2), ("opacity", 3), ("flow", 8), ("semantic", 9)):
            if ref[i].numel():
                print(f"oracle {nm} rel err vs ref: {Hh.rel_err(torch.tensor(out[nm]).cuda(), ref[i]):.3e}")
        if backward:
            t0 = time.time()
            go_ = O.rasterize_backward(s, st, out, n(c["means3D"]), n(cot["color"]), n(cot["depth"]), n(cot["flow"]),
                                       n(cot["semantic"]), n(cot["opacity"]), n(c["scales"]), n(c["rotations"]),
                                       n(c["cov3D_precomp"]), n(c["sh"]), n(c["flow_points"]), n(c["semantic"]))
            print(f"[oracle bwd {time.time() - t0:.1f}s]")
            for nm, b in zip(GRAD_NAMES, gr):
                if b.numel():
                    print(f"oracle {nm} rel err vs ref: {Hh.rel_err(torch.tensor(go_[nm]).cuda().reshape(b.shape), b):.3e}")


def time_case(name, c, iters=5):
    cot = Hh.cotangents(c)
    for impl, F, B in (("ours", Hh.OURS.rasterize_gaussians, Hh.OURS.rasterize_gaussians_backward),
                       ("ref", REF.rasterize_gaussians, REF.rasterize_gaussians_backward)):
        tf, tb = [], []
        for _ in range(iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = F(*Hh.fwd_args(c))
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if impl == "ours":
                B(*Hh.bwd_args(c, out, cot), opacities=c["opacity"])
            else:
                B(*Hh.bwd_args(c, out, cot))
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tf.append(t1 - t0)
            tb.append(t2 - t1)
        print(f"[time {name}] {impl}: fwd {min(tf) * 1e3:.3f} ms  bwd {min(tb) * 1e3:.3f} ms  R={out[0]}")


def sort_check():
    import ctypes as C
    lib = Hh.L.load()
    for n, bits in ((1, 32), (1000, 32), (4096, 11), (4097, 13), (1_000_003, 32), (3_000_000, 11)):
        g = torch.Generator(device="cpu").manual_seed(n)
        keys = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64).to(torch.int32).cuda()
        if bits < 32:
            keys = keys & ((1 << bits) - 1)
        vals = torch.arange(n, dtype=torch.int32).cuda()
        k_in, v_in = keys.clone(), vals.clone()
        k_out, v_out = torch.empty_like(keys), torch.empty_like(vals)
        ws = torch.empty(lib.adgs_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
        sel = lib.adgs_sort_pairs(k_in.data_ptr(), v_in.data_ptr(), k_out.data_ptr(), v_out.data_ptr(), n, 0, bits,
                                  ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        rk, rv = (k_out, v_out) if sel == 0 else (k_in, v_in)
        ek, ei = torch.sort(keys.long(), stable=True)
        ok_k = (rk.long() == ek).all().item()
        ok_v = (rv.long() == ei).all().item()
        print(f"sort n={n} bits={bits} sel={sel}: keys ok={ok_k} stable values ok={ok_v}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), "cpus", os.cpu_count())
    for fn in (sort_check,):
        try:
            fn()
        except Exception:
            traceback.print_exc()
    cases = [
        ("small", dict(n=2000, W=96, H=64, seed=1), True),
        ("ragged", dict(n=5000, W=171, H=99, seed=2, D_S=0, flow=False, inv_depth=False, bg=(0.3, 0.5, 0.7)), False),
        ("deg1-sem4", dict(n=5000, W=160, H=96, seed=3, sh_degree=1, D_S=4), False),
        ("precomp", dict(n=3000, W=128, H=80, seed=4, colors_precomp=True, cov3D_precomp=True, D_S=1), False),
        ("medium", dict(n=200_000, W=640, H=360, seed=5), False),
    ]
    if args.big:
        cases.append(("kitti-1M", dict(n=1_000_000, W=1242, H=375, seed=6, median_radius_px=3.0), False))
    for name, kw, chk in cases:
        try:
            c = Hh.make_case(**kw)
            compare_case(name, c, check_oracle=chk)
            if c["n"] >= 200_000:
                time_case(name, c)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
