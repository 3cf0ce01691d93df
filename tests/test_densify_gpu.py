"""GPU parity of densification / pruning / opacity reset / near-index K-NN (adgs_b200/densify.py ->
adgs_b200/csrc/densify.cu, knn_points.cu) against

  * tests/golden/densify.npz -- outputs of the reference's OWN scene/gaussian_model.py (CPU), and
  * oracle/densify_oracle.py -- the sequential clone -> split -> prune restatement, on seeded random models,

through the public mirror methods of GaussianModel and the C ABI. Which rows survive and in which order, and every
copied array (parameters and Adam moments), must be BIT-EXACT; the split children's positions (a 3x3 product
whose summation order differs between bmm and the kernel) and scales within 2e-6.
At full size (1 M Gaussians) the checks are size-independent properties: row conservation, order preservation,
zero moments on new rows, prune idempotence."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from adgs_b200 import densify as D
from adgs_b200 import scenes
from adgs_b200.gaussian_model import GaussianModel, PARAM_NAMES
from adgs_b200.optimizer import FusedAdam
from oracle import densify_oracle as O

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "densify.npz"))
GOLD_ORDER_ARGS = {"xyz": [6, 3, 0, 2, 0, 0], "rotation": [0, 0, 0, 0, 5, 2], "shs": [0, 0, 0, 2, 0, 0],
                   "background": [6, 3, 0, 2, 0, 0]}
REF_NAMES = O.SCENE_ROWS + O.OBJ_ROWS + ("background_deform_param",)


def gold_state(tag, stage):
    pre = f"{tag}.{stage}."
    return {k[len(pre):]: GOLD[k].copy() for k in GOLD.files if k.startswith(pre)}


def model_from_state(st, order_args, scene_extent=20.0, object_extent=5.0, percent_dense=0.01):
    """GaussianModel + FusedAdam (moments loaded) + statistics from a reference-layout numpy state."""
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ref = {k: T(st[k]) for k in REF_NAMES}
    ref["gs_time"] = T(st["gs_time"])
    model = GaussianModel.from_reference(ref, order_args)
    model.scene_extent, model.object_extent, model.percent_dense = scene_extent, object_extent, percent_dense
    opt = FusedAdam(model, {}, eps=1e-15)
    model.optimizer = opt
    if "scene_xyz.exp_avg" in st:
        opt.load_reference_state({k: {"exp_avg": T(st[k + ".exp_avg"]), "exp_avg_sq": T(st[k + ".exp_avg_sq"])}
                                  for k in REF_NAMES}, step=2)
    D.setup_statistics(model)
    if "xyz_gradient_accum" in st:
        model.xyz_gradient_accum.copy_(T(st["xyz_gradient_accum"]))
        model.denom.copy_(T(st["denom"]))
        model.max_radii2D.copy_(T(st["max_radii2D"]))
    return model


def state_of(model):
    """Reference-layout numpy state of a model (parameters, moments, statistics)."""
    out = {k: v.detach().cpu().numpy() for k, v in model.to_reference().items()}
    for name, d in model.optimizer.state_in_reference_layout().items():
        out[name + ".exp_avg"] = d["exp_avg"].detach().cpu().numpy()
        out[name + ".exp_avg_sq"] = d["exp_avg_sq"].detach().cpu().numpy()
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        out[k] = getattr(model, k).cpu().numpy()
    return out


def compare(got, want, what):
    exact, close = 0, 0
    for k, w in want.items():
        if k.startswith("background_deform_param"):
            continue
        g = got[k]
        assert g.shape == w.shape, f"{what}: {k} shape {g.shape} != {w.shape}"
        if k in ("scene_xyz", "obj_xyz", "scene_scaling", "obj_scaling"):
            np.testing.assert_allclose(g, w, rtol=2e-6, atol=2e-6, err_msg=f"{what}: {k}")
            close += 1
        else:
            assert np.array_equal(g, w), f"{what}: {k} differs, max abs {np.abs(g - w).max() if g.size else 0}"
            exact += 1
    assert exact >= 30 and close == 4


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_statistics_match_reference_golden(tag):
    st = gold_state(tag, "before")
    model = model_from_state({k: v for k, v in st.items() if k not in ("xyz_gradient_accum", "denom", "max_radii2D")},
                             GOLD_ORDER_ARGS)
    for it in range(3):
        vsp = SimpleNamespace(grad=torch.from_numpy(GOLD[f"{tag}.stats{it}.grad"]).cuda())
        radii = torch.from_numpy(GOLD[f"{tag}.stats{it}.radii"]).cuda()
        model.add_densification_stats({"viewspace_points": vsp, "radii": radii, "visibility_filter": radii > 0})
    np.testing.assert_allclose(model.xyz_gradient_accum.cpu().numpy(), st["xyz_gradient_accum"], rtol=1e-6, atol=1e-12)
    assert np.array_equal(model.denom.cpu().numpy(), st["denom"])
    assert np.array_equal(model.max_radii2D.cpu().numpy(), st["max_radii2D"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_densify_and_prune_matches_reference_golden(tag):
    st = gold_state(tag, "before")
    scene_extent, object_extent, percent_dense = [float(v) for v in GOLD[f"{tag}.extents"]]
    model = model_from_state(st, GOLD_ORDER_ARGS, scene_extent, object_extent, percent_dense)
    zs = [torch.from_numpy(GOLD[f"{tag}.z_scene"]), torch.from_numpy(GOLD[f"{tag}.z_obj"])]
    calls = []

    def sample_fn(rows):
        z = zs[len(calls)]
        calls.append(rows)
        assert z.shape[0] == rows, "the library asked for a different number of normals than the reference drew"
        return z.cuda()

    model.densify_and_prune(0.0002, 0.0002, 0.005, bool(GOLD[f"{tag}.prune_big"]), sample_fn=sample_fn)
    want = gold_state(tag, "after")
    assert (model.n_scene, model.n_obj) == (want["scene_xyz"].shape[0], want["obj_xyz"].shape[0])
    compare(state_of(model), want, f"golden {tag}")
    # reset_opacity on the densified model
    model.reset_opacity()
    got, want = state_of(model), gold_state(tag, "reset")
    for part in ("scene", "obj"):
        np.testing.assert_allclose(got[f"{part}_opacity"], want[f"{part}_opacity"], rtol=2e-6, atol=2e-6)
        assert not got[f"{part}_opacity.exp_avg"].any() and not got[f"{part}_opacity.exp_avg_sq"].any()
    assert np.array_equal(got["scene_xyz.exp_avg"], want["scene_xyz.exp_avg"])


def random_state(n_scene, n_obj, seed, order_args=scenes.BENCH_ORDER_ARGS):
    cam = scenes.make_camera(160, 96, 90.0, device="cuda")
    cloud = scenes.random_cloud(n_scene + n_obj, cam, seed=seed, median_radius_px=4.0)
    tensors = scenes.random_model_tensors(n_scene, n_obj, order_args, cloud, seed=seed + 1, device="cuda")
    st = {k: v.detach().cpu().numpy() for k, v in tensors.items()}
    rng = np.random.default_rng(seed)
    n = n_scene + n_obj
    # scales straddling the clone / split / big thresholds, opacities straddling min_opacity
    for part, rows, size in (("scene", n_scene, 0.2), ("obj", n_obj, 0.05)):
        st[f"{part}_scaling"] = np.log(size * np.exp(1.5 * rng.standard_normal((rows, 3)))).astype(np.float32)
        st[f"{part}_opacity"] = (3.0 * rng.standard_normal((rows, 1)) - 2.0).astype(np.float32)
    for k in REF_NAMES:
        st[k + ".exp_avg"] = (1e-3 * rng.standard_normal(st[k].shape)).astype(np.float32)
        st[k + ".exp_avg_sq"] = (1e-6 * rng.random(st[k].shape)).astype(np.float32)
    st["denom"] = rng.integers(0, 4, (n, 1)).astype(np.float32)
    st["xyz_gradient_accum"] = (st["denom"] * 0.0002 * np.exp(rng.standard_normal((n, 1)))).astype(np.float32)
    st["max_radii2D"] = rng.integers(0, 50, (n,)).astype(np.float32)
    return st


@pytest.mark.parametrize("n_scene,n_obj,prune_big", [(3000, 1300, True), (513, 0, False), (0, 700, True),
                                                     (256, 256, False)])
def test_densify_and_prune_matches_oracle(n_scene, n_obj, prune_big):
    st = random_state(n_scene, n_obj, seed=n_scene + n_obj)
    model = model_from_state(st, scenes.BENCH_ORDER_ARGS, 20.0, 5.0, 0.01)
    drawn = []

    def sample_fn(rows):
        g = torch.Generator().manual_seed(100 + len(drawn))
        drawn.append(torch.randn((rows, 3), generator=g))
        return drawn[-1].cuda()

    src, tag = model.densify_and_prune(0.0002, 0.0002, 0.005, prune_big, sample_fn=sample_fn)
    want = {k: v.copy() for k, v in st.items()}
    O.densify_and_prune(want, 0.0002, 0.0002, 0.005, prune_big, 20.0, 5.0, 0.01, drawn[0].numpy(), drawn[1].numpy(),
                        N=2, gpu_division=True)
    assert (model.n_scene, model.n_obj) == (want["scene_xyz"].shape[0], want["obj_xyz"].shape[0])
    compare(state_of(model), want, "oracle")
    kinds = (tag[:model.n_scene + model.n_obj] & 3).cpu().numpy()
    assert set(np.unique(kinds)) <= {0, 1, 2}
    # the model still renders-compatible: planar shapes are consistent
    assert model.xyz_deform.shape[2] == model.n_obj and model.sh4.shape[1] == model.n_scene + model.n_obj
    assert model.gs_time.shape[0] == model.n_obj


def test_prune_points_matches_oracle_and_is_idempotent():
    st = random_state(1000, 600, seed=5)
    model = model_from_state(st, scenes.BENCH_ORDER_ARGS)
    rng = np.random.default_rng(9)
    sm, om = rng.random(1000) < 0.3, rng.random(600) < 0.6
    model.prune_points(torch.from_numpy(sm).cuda(), torch.from_numpy(om).cuda())
    want = {k: v.copy() for k, v in st.items()}
    O.prune_points(want, sm, om)
    got = state_of(model)
    for k, w in want.items():
        if not k.startswith("background"):
            assert np.array_equal(got[k], w), k
    before = state_of(model)
    model.prune_points(torch.zeros(model.n_scene, dtype=torch.bool), torch.zeros(model.n_obj, dtype=torch.bool))
    after = state_of(model)
    for k in before:
        assert np.array_equal(before[k], after[k]), k
    # empty result
    model.prune_points(torch.ones(model.n_scene, dtype=torch.bool), torch.ones(model.n_obj, dtype=torch.bool))
    assert model.get_pts_num == 0 and model.xyz.shape == (0, 3)


def test_densify_full_size_properties():
    """1 M Gaussians (BASELINE configs[1] size): conservation and ordering properties, then a render-free timing."""
    n_scene, n_obj = 750_000, 250_000
    st = random_state(n_scene, n_obj, seed=3)
    model = model_from_state(st, scenes.BENCH_ORDER_ARGS, 20.0, 5.0, 0.01)
    old_xyz, old_m = model.xyz.detach().clone(), model.optimizer.state["sh4"]["exp_avg"].clone()
    old_rot_deform = model.rot_deform.detach().clone()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    src, tag = model.densify_and_prune(0.0002, 0.0002, 0.005, True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    n2, ns2 = model.get_pts_num, model.n_scene
    src, kind = src[:n2].long(), (tag[:n2] & 3)
    assert n2 > 0 and src.shape[0] == n2
    # partitions stay partitions; within a partition the kinds appear as [keep | clone | child copy 0 | child copy 1]
    assert bool((src[:ns2] < n_scene).all()) and bool((src[ns2:] >= n_scene).all())
    for lo, hi in ((0, ns2), (ns2, n2)):
        k = kind[lo:hi]
        assert bool((k[1:] >= k[:-1]).all())
        s_keep = src[lo:hi][k == 0]
        assert bool((s_keep[1:] > s_keep[:-1]).all())             # source order preserved, no duplicates
        s_clone = src[lo:hi][k == 1]
        assert bool((s_clone[1:] > s_clone[:-1]).all())
        s_child = src[lo:hi][k == 2]
        half = s_child.shape[0] // 2
        assert s_child.shape[0] % 2 == 0 and bool((s_child[:half] == s_child[half:]).all())
    # kept / cloned rows are bit-copies; new rows have zero moments; kept rows carry their moments
    same = kind != 2
    assert torch.equal(model.xyz.detach()[same], old_xyz[src[same]])
    new_m = model.optimizer.state["sh4"]["exp_avg"]
    assert torch.equal(new_m[:, kind == 0], old_m[:, src[kind == 0]])
    assert not bool(new_m[:, kind != 0].any())
    assert torch.equal(model.rot_deform.detach(), old_rot_deform[:, src[ns2:] - n_scene])
    assert not bool(model.xyz_gradient_accum.any()) and model.max_radii2D.shape[0] == n2
    print(f"densify_and_prune 1M -> {n2} rows: {ms:.2f} ms")
    assert ms < 200.0


def test_densified_model_renders_and_steps():
    """After densification the model, its optimizer and the renderer still work together."""
    from adgs_b200.gaussian_renderer import render
    st = random_state(4000, 1500, seed=8)
    model = model_from_state(st, scenes.BENCH_ORDER_ARGS, 20.0, 5.0, 0.01)
    for g in model.optimizer.param_groups:
        g["lr"] = 1e-3
    cam = scenes.make_camera(160, 96, 90.0, time=0.37, device="cuda")
    pipe = SimpleNamespace(debug=False, inv_depth=True, sync_free=False)
    model.densify_and_prune(0.0002, 0.0002, 0.005, False)
    pkg = render(cam, model, None, pipe, flow_pkg=None, render_objmask=True)
    (pkg["render"].sum() + pkg["depth"].sum()).backward()
    model.add_densification_stats(pkg)
    assert float(model.denom.sum()) == float((pkg["radii"] > 0).sum())
    model.optimizer.step()
    assert all(torch.isfinite(getattr(model, k)).all() for k in PARAM_NAMES)


@pytest.mark.parametrize("D,K,P,A", [(3, 8, 5000, 625), (4, 8, 20000, 2500), (4, 5, 300, 60), (3, 16, 4097, 130),
                                     (4, 32, 1000, 31), (4, 1, 64, 64)])
def test_knn_points_matches_oracle(D, K, P, A):
    g = torch.Generator().manual_seed(P + K)
    pts = torch.randn((P, D), generator=g)
    anchors = pts[torch.randperm(P, generator=g)[:A]].clone()
    idx, d = D_knn(anchors, pts, K)
    widx, wd = O.knn_points(anchors.numpy(), pts.numpy(), K)
    np.testing.assert_allclose(d, wd, rtol=1e-5, atol=1e-6)
    assert (np.diff(d, axis=1) >= 0).all()
    # indices agree except where two candidates are closer than float rounding (FMA contraction in the kernel)
    diff = idx != widx
    if diff.any():
        gap_ok = np.abs(d - wd)[diff] <= 1e-5 * np.maximum(wd[diff], 1e-6)
        assert gap_ok.all() and diff.mean() < 1e-3
    assert (idx[:, 0] == np.asarray([np.where((pts.numpy() == a).all(1))[0][0] for a in anchors.numpy()])).all()


def D_knn(anchors, pts, K):
    idx, d = D.knn_points(anchors.cuda(), pts.cuda(), K, return_dists=True)
    return idx.cpu().numpy(), d.cpu().numpy()


def test_set_obj_near_idx_mirrors_reference():
    st = random_state(500, 2400, seed=4)
    model = model_from_state(st, scenes.BENCH_ORDER_ARGS)
    model.use_near_idx, model.near_num = True, 8
    torch.manual_seed(0)
    model.set_obj_near_idx()
    idx = model.obj_near_idx
    assert idx.shape == (2400 // 8, 8) and idx.dtype == torch.int64
    # same anchors as the reference draws with this seed: randperm on the device
    torch.manual_seed(0)
    perm = torch.randperm(2400, device="cuda")[:300]
    assert torch.equal(idx[:, 0], perm)                      # every anchor finds itself first
    xyz4 = torch.cat([model.xyz.detach()[500:], model.gs_time.reshape(-1, 1) * 20.0], dim=-1)
    widx, _ = O.knn_points(xyz4[perm].cpu().numpy(), xyz4.cpu().numpy(), 8)
    assert (idx.cpu().numpy() == widx).mean() > 0.999


def test_knn_points_large_timing():
    """near-index table at the 1 M-Gaussian workload's size: 250 k object Gaussians, 31 k anchors, D = 4, K = 8."""
    g = torch.Generator(device="cuda").manual_seed(1)
    pts = torch.randn((250_000, 4), generator=g, device="cuda")
    anchors = pts[torch.randperm(250_000, device="cuda")[:31_250]].contiguous()
    D.knn_points(anchors, pts, 8)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    idx, d = D.knn_points(anchors, pts, 8, return_dists=True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    # spot-check 64 anchors against torch
    sel = torch.arange(0, 31_250, 500, device="cuda")
    full = ((anchors[sel, None, :] - pts[None]) ** 2).sum(-1)
    wd, _ = torch.topk(full, 8, dim=1, largest=False)
    torch.testing.assert_close(d[sel], wd, rtol=1e-5, atol=1e-6)
    print(f"knn_points 31k x 250k, K=8: {ms:.2f} ms ({31_250 * 250_000 / ms / 1e6:.1f} G pairs/s)")
    assert ms < 100.0


def test_create_from_pcd_mirrors_reference():
    """scene/gaussian_model.py:255-333 on the planar storage: scales from distCUDA2, DC colour = RGB2SH, opacity 0.1,
    identity rotations, scene / object split by obj_id, deformation parameters U(-1,1)*1e-5, sigma = log(frame_gap)."""
    from adgs_b200.simple_knn import distCUDA2
    rng = np.random.default_rng(0)
    P = 5000
    pcd = SimpleNamespace(points=rng.standard_normal((P, 3)) * 10, colors=rng.random((P, 3)),
                          time=rng.random((P, 1)), obj_id=(rng.random((P, 1)) < 0.3).astype(np.float64))
    torch.manual_seed(0)
    m = GaussianModel.create_from_pcd(pcd, scene_extent=30.0, cameras_extent=8.0, frame_gap=1.0 / 96,
                                      default_order_downsample_ratio=3)
    obj = pcd.obj_id[:, 0] > 0.5
    assert (m.n_scene, m.n_obj) == (int((~obj).sum()), int(obj.sum()))
    assert m.order_args["xyz"] == [32, 5, 0, 6, 0, 0] and m.order_args["rotation"] == [0, 0, 0, 0, 32, 1]
    ref = m.to_reference()
    pts = torch.tensor(pcd.points).float().cuda()
    d2 = torch.clamp_min(distCUDA2(pts), 1e-7)
    want_scale = torch.log(torch.sqrt(d2))[:, None].repeat(1, 3)
    assert torch.equal(ref["scene_scaling"], want_scale[torch.from_numpy(~obj).cuda()])
    assert torch.equal(ref["obj_xyz"], pts[torch.from_numpy(obj).cuda()])
    want_dc = (torch.tensor(pcd.colors).float().cuda() - 0.5) / 0.28209479177387814
    torch.testing.assert_close(ref["scene_shs_dc"][:, 0], want_dc[torch.from_numpy(~obj).cuda()])
    assert not bool(ref["obj_shs_rest"].any())
    torch.testing.assert_close(torch.sigmoid(ref["obj_opacity"]), torch.full_like(ref["obj_opacity"], 0.1))
    assert bool((ref["scene_rotation"] == torch.tensor([1.0, 0, 0, 0]).cuda()).all())
    assert ref["xyz_deform_param"].shape == (m.n_obj, 3, 44) and float(ref["xyz_deform_param"].abs().max()) <= 1e-5
    torch.testing.assert_close(ref["gs_time"][:, 0], torch.tensor(pcd.time[obj, 0]).float().cuda())
    assert torch.allclose(ref["gs_time_sigma"], torch.tensor(float(np.log(1.0 / 96))).cuda())
    assert m.active_sh_degree == 0 and m.max_radii2D.shape == (P,)
