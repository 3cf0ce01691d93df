"""CPU simulation (numpy oracle; not a test, not collected by pytest): how many warp iterations does the blend forward
need on a KITTI-density scene when survivors are culled per 8x4 sub-tile (the shipped kernel), per 4x4 half-warp
block or per 4x2 quarter-warp block? Result (seed 6, 100 k Gaussians at 393x118 = the bench scene's density):
1.20x fewer iterations for half-warps, 1.33x for quarter-warps, 1.46x more (splat, block) items. The half-warp kernels
built from this estimate were parity-green but 3-6 % slower on the B200 (profiles/r1_z_ab_half_warp_blend.txt).
Usage: python tests/sim_cull_granularity.py"""
import sys, math, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hh
from oracle import raster_oracle as O
n, W, H = 100_000, 393, 118
c = Hh.make_case(n=n, W=W, H=H, seed=6, median_radius_px=3.0, device="cpu")
s = Hh.oracle_settings(c)
N = Hh.to_np
g = O.preprocess(s, N(c["means3D"]), N(c["opacity"]), N(c["scales"]), N(c["rotations"]), None, N(c["sh"]), None)
keys, pl, ranges, _ = O.binning(s, g)
print("R/N", len(pl)/n, "tiles", s.grid_x*s.grid_y, "avg list", len(pl)/(s.grid_x*s.grid_y))
mx, my = g["means2D"][:,0].astype(np.float64), g["means2D"][:,1].astype(np.float64)
A, B, C = [g["conic"][:,i].astype(np.float64) for i in range(3)]
op = g["opacity"].astype(np.float64)
thresh = -np.log(255.0*np.maximum(op,1e-12))
def may_touch(idx, X0, Y0, X1, Y1):
    m_x, m_y, a, b, cc, th = mx[idx], my[idx], A[idx], B[idx], C[idx], thresh[idx]
    dx0 = m_x - np.clip(m_x, X0, X1); dy0 = m_y - np.clip(m_y, Y0, Y1)
    inside = (dx0 == 0) & (dy0 == 0)
    dx_lo, dx_hi, dy_lo, dy_hi = m_x - X1, m_x - X0, m_y - Y1, m_y - Y0
    pw = lambda dx, dy: -0.5*(a*dx*dx + cc*dy*dy) - b*dx*dy
    best = np.full(len(idx), -3e38)
    v = np.clip(-b*dx0/cc, dy_lo, dy_hi); best = np.where(dx0 != 0, pw(dx0, v), best)
    v = np.clip(-b*dy0/a, dx_lo, dx_hi); best = np.where(dy0 != 0, np.maximum(best, pw(v, dy0)), best)
    return (th <= 0.01) & (inside | ~(best < th - 0.01))
rng = np.random.default_rng(0)
tiles = rng.choice(s.grid_x*s.grid_y, 60, replace=False)
it_cur = it_half = it_q = ev_pairs = 0; surv_cur = surv_half = 0; act_lanes = 0
for t in tiles:
    r0, r1 = ranges[t]
    idx = pl[r0:r1].astype(np.int64)
    if len(idx) == 0: continue
    tx, ty = (t % s.grid_x)*16, (t // s.grid_x)*16
    for w in range(8):
        X0, Y0 = tx + (w & 1)*8, ty + (w >> 1)*4
        full = may_touch(idx, X0, Y0, X0+7, Y0+3)
        hA = may_touch(idx, X0, Y0, X0+3, Y0+3); hB = may_touch(idx, X0+4, Y0, X0+7, Y0+3)
        # quarter: 4x2 blocks
        qs = [may_touch(idx, X0+4*(k&1), Y0+2*(k>>1), X0+4*(k&1)+3, Y0+2*(k>>1)+1) for k in range(4)]
        nchunk = (len(idx)+31)//32
        for ch in range(nchunk):
            sl = slice(ch*32, ch*32+32)
            it_cur += full[sl].sum(); it_half += max(hA[sl].sum(), hB[sl].sum()); it_q += max(q[sl].sum() for q in qs)
        surv_cur += full.sum(); surv_half += hA.sum() + hB.sum()
print("iterations per warp: current", it_cur, "half-warp", it_half, "ratio", it_cur/it_half, "quarter", it_q, it_cur/it_q)
print("survivor items: 8x4", surv_cur, "4x4 halves", surv_half, surv_half/surv_cur)
