"""GPU timing of the densification path at the 1 M-Gaussian workload (BASELINE configs[1] size):

  * adgs_b200 densify_and_prune (classify + scan + plan + one gather + split; adgs_b200/csrc/densify.cu)
  * the reference's formulation -- clone, split, prune as three rounds of boolean-mask indexing and torch.cat over
    the 17 per-Gaussian tensors and their 34 Adam moments (scene/gaussian_model.py:560-861) -- restated with torch
    ops on the same device in the reference's layout (measurement baseline only; the parity oracle is
    oracle/densify_oracle.py)
  * knn_points (adgs_knn_points) vs a chunked torch cdist + topk
  * add_densification_stats (one launch) vs the reference's four indexed torch ops

Lives under tests/ because it reuses the parity tests' model builders (which import oracle/); it is measurement
infrastructure, not collected by pytest. Writes gpurun_out/densify_timing.json. Usage (GPU box): python tests/densify_timing.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from adgs_b200 import densify as D  # noqa: E402
from adgs_b200 import scenes  # noqa: E402
import test_densify_gpu as TD  # noqa: E402  (model builders)

SCENE = ("scene_xyz", "scene_shs_dc", "scene_shs_rest", "scene_opacity", "scene_scaling", "scene_rotation",
         "shs_deform_param_scene")
OBJ = ("obj_xyz", "obj_shs_dc", "obj_shs_rest", "obj_opacity", "obj_scaling", "obj_rotation", "shs_deform_param_obj",
       "xyz_deform_param", "rotation_deform_param", "gs_time_sigma")


def build_rotation(r):
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)


def torch_reference_densify(st, max_grad, min_opacity, prune_big, scene_extent, object_extent, percent_dense, N=2):
    """The reference's sequence with torch ops (masks, cats, and their host synchronisations)."""
    def select(names, keep):
        for n in names:
            for suf in ("", ".exp_avg", ".exp_avg_sq"):
                st[n + suf] = st[n + suf][keep]

    def prune(sm, om):
        select(SCENE, ~sm)
        select(OBJ, ~om)
        st["gs_time"] = st["gs_time"][~om]
        valid = torch.cat([~sm, ~om])
        for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
            st[k] = st[k][valid]

    def postfix(new):
        for n, ext in new.items():
            if n == "gs_time":
                continue
            st[n] = torch.cat([st[n], ext], 0)
            for suf in (".exp_avg", ".exp_avg_sq"):
                st[n + suf] = torch.cat([st[n + suf], torch.zeros_like(ext)], 0)
        st["gs_time"] = torch.cat([st["gs_time"], new["gs_time"]], 0)
        n = st["scene_xyz"].shape[0] + st["obj_xyz"].shape[0]
        st["xyz_gradient_accum"] = torch.zeros((n, 1), device="cuda")
        st["denom"] = torch.zeros((n, 1), device="cuda")
        st["max_radii2D"] = torch.zeros((n,), device="cuda")

    grads = st["xyz_gradient_accum"] / st["denom"]
    grads[grads.isnan()] = 0.0
    grads = torch.norm(grads, dim=-1)
    ns = st["scene_xyz"].shape[0]
    ssel, osel = grads[:ns] >= max_grad, grads[ns:] >= max_grad
    ssz, osz = scene_extent * percent_dense, object_extent * percent_dense
    # clone
    s1 = ssel & (torch.exp(st["scene_scaling"]).max(dim=1).values <= ssz)
    o1 = osel & (torch.exp(st["obj_scaling"]).max(dim=1).values <= osz)
    new = {n: st[n][s1] for n in SCENE}
    new.update({n: st[n][o1] for n in OBJ})
    new["gs_time"] = st["gs_time"][o1]
    postfix(new)
    # split
    pad = lambda m, n: torch.cat([m, torch.zeros(n - m.shape[0], dtype=torch.bool, device="cuda")])
    ssel, osel = pad(ssel, st["scene_xyz"].shape[0]), pad(osel, st["obj_xyz"].shape[0])
    s2 = ssel & (torch.exp(st["scene_scaling"]).max(dim=1).values > ssz)
    o2 = osel & (torch.exp(st["obj_scaling"]).max(dim=1).values > osz)
    new = {}
    for part, sel in (("scene", s2), ("obj", o2)):
        stds = torch.exp(st[f"{part}_scaling"][sel]).repeat(N, 1)
        samples = torch.normal(mean=0.0, std=stds)
        rots = build_rotation(st[f"{part}_rotation"][sel]).repeat(N, 1, 1)
        new[f"{part}_xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + st[f"{part}_xyz"][sel].repeat(N, 1)
        new[f"{part}_scaling"] = torch.log(stds / (0.8 * N))
        for k in ("rotation", "opacity"):
            new[f"{part}_{k}"] = st[f"{part}_{k}"][sel].repeat(N, 1)
        for k in ("shs_dc", "shs_rest"):
            new[f"{part}_{k}"] = st[f"{part}_{k}"][sel].repeat(N, 1, 1)
    new["shs_deform_param_scene"] = st["shs_deform_param_scene"][s2].repeat(N, 1, 1)
    for k in ("shs_deform_param_obj", "xyz_deform_param", "rotation_deform_param"):
        new[k] = st[k][o2].repeat(N, 1, 1)
    new["gs_time_sigma"] = st["gs_time_sigma"][o2].repeat(N, 1)
    new["gs_time"] = st["gs_time"][o2].repeat(N, 1)
    postfix(new)
    prune(torch.cat([s2, torch.zeros(N * int(s2.sum()), dtype=torch.bool, device="cuda")]),
          torch.cat([o2, torch.zeros(N * int(o2.sum()), dtype=torch.bool, device="cuda")]))
    # final prune
    sp = (torch.sigmoid(st["scene_opacity"]) < min_opacity).squeeze()
    op = (torch.sigmoid(st["obj_opacity"]) < min_opacity).squeeze()
    if prune_big:
        sp = sp | (torch.exp(st["scene_scaling"]).max(dim=1).values > scene_extent * 0.05)
        op = op | (torch.exp(st["obj_scaling"]).max(dim=1).values > object_extent * 0.1)
    prune(sp, op)
    return st["scene_xyz"].shape[0], st["obj_xyz"].shape[0]


def wall_ms(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3, out


def main():
    out = {"device": torch.cuda.get_device_name(0)}
    n_scene, n_obj = 750_000, 250_000
    st = TD.random_state(n_scene, n_obj, seed=3)
    # ---- ours (second run timed: allocator warm, like a training run after its first densification) ----------
    ms_ours = []
    for rep in range(3):
        model = TD.model_from_state(st, scenes.BENCH_ORDER_ARGS, 20.0, 5.0, 0.01)
        ms, _ = wall_ms(lambda: model.densify_and_prune(0.0002, 0.0002, 0.005, True))
        ms_ours.append(ms)
        rows = (model.n_scene, model.n_obj)
        del model
    out["adgs_densify_and_prune_ms"] = ms_ours
    out["rows_after"] = rows
    # ---- reference formulation with torch ops --------------------------------------------------------------
    ms_ref = []
    for rep in range(3):
        tst = {k: torch.from_numpy(v).cuda() for k, v in st.items()}
        ms, rows_ref = wall_ms(lambda: torch_reference_densify(tst, 0.0002, 0.005, True, 20.0, 5.0, 0.01))
        ms_ref.append(ms)
        del tst
    out["torch_reference_formulation_ms"] = ms_ref
    out["rows_after_reference_formulation"] = rows_ref
    # ---- statistics ---------------------------------------------------------------------------------------
    model = TD.model_from_state(st, scenes.BENCH_ORDER_ARGS, 20.0, 5.0, 0.01)
    n = n_scene + n_obj
    grad = torch.randn((n, 3), device="cuda") * 1e-4
    radii = torch.randint(-5, 30, (n,), device="cuda", dtype=torch.int32).clamp(min=0)
    pkg = {"viewspace_points": type("V", (), {"grad": grad})(), "radii": radii}
    model.add_densification_stats(pkg)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        model.add_densification_stats(pkg)
    e1.record()
    torch.cuda.synchronize()
    out["adgs_stats_ms"] = e0.elapsed_time(e1) / 20
    acc, den, mr = model.xyz_gradient_accum.clone(), model.denom.clone(), model.max_radii2D.clone()
    e0.record()
    for _ in range(20):
        vis = radii > 0
        mr[vis] = torch.max(mr[vis], radii[vis])
        acc[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
        den[vis] += 1
    e1.record()
    torch.cuda.synchronize()
    out["torch_reference_stats_ms"] = e0.elapsed_time(e1) / 20
    # ---- knn ----------------------------------------------------------------------------------------------
    pts = torch.cat([model.xyz.detach()[n_scene:], model.gs_time.reshape(-1, 1) * 20.0], dim=-1).contiguous()
    anchors = pts[torch.randperm(n_obj, device="cuda")[:n_obj // 8]].contiguous()
    D.knn_points(anchors, pts, 8)
    e0.record()
    for _ in range(5):
        idx = D.knn_points(anchors, pts, 8)
    e1.record()
    torch.cuda.synchronize()
    out["adgs_knn_points_ms"] = e0.elapsed_time(e1) / 5
    e0.record()
    ref_idx = []
    for i in range(0, anchors.shape[0], 4096):
        d = torch.cdist(anchors[i:i + 4096], pts)
        ref_idx.append(torch.topk(d, 8, dim=1, largest=False).indices)
    e1.record()
    torch.cuda.synchronize()
    out["torch_cdist_topk_ms"] = e0.elapsed_time(e1)
    ref_idx = torch.cat(ref_idx)
    out["knn_index_agreement_with_torch"] = float((torch.sort(idx, 1).values == torch.sort(ref_idx, 1).values).float().mean())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "densify_timing.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
