"""Densification, pruning, opacity reset and the near-index K-NN of `scene/gaussian_model.py:GaussianModel`
over the planar B200 storage.

Same names, arguments and results as the reference:

    add_densification_stats(render_pkg)                              gaussian_model.py:863-867 (+ train.py:151)
    densify_and_prune(max_scene_grad, max_obj_grad, min_opacity, prune_big_points)   :835-861
    prune_points(scene_mask, obj_mask)                               :585-614
    reset_opacity()                                                  :463-467
    set_obj_near_idx(K=None)                                         :825-833

but clone -> split -> prune is ONE classification pass, one scan, and one gather launch that writes every new
parameter and Adam-moment array exactly once (adgs_b200/csrc/densify.cu) instead of three rounds of boolean-mask
indexing and torch.cat over 17 tensors and their 34 moments. The only host synchronisation is the read-back of
the eight row counts (the new array sizes); the reference synchronises on every mask.

The split children's random offsets are `torch.randn` drawn on the model's device in the reference's order
(scene rows first, then object rows) -- `torch.normal(0, stds)` is `randn * stds` in ATen -- so a run seeded like
the reference consumes the same random stream.

No fallback: every function raises if the CUDA library is missing. Nothing here imports `oracle/`.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .gaussian_model import PARAM_NAMES

# arrays with one row per Gaussian of [scene ; object] / per object Gaussian: (name, rows are object-only)
_PER_GAUSSIAN = (("xyz", False), ("scaling", False), ("rotation", False), ("opacity", False), ("sh4", False),
                 ("shs_deform4", False), ("xyz_deform", True), ("rot_deform", True), ("gs_time_sigma", True))


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def setup_statistics(model):
    """The three running statistics of training_setup / create_from_pcd (gaussian_model.py:285,340-341)."""
    n, dev = model.get_pts_num, model.xyz.device
    model.xyz_gradient_accum = torch.zeros((n, 1), dtype=torch.float32, device=dev)
    model.denom = torch.zeros((n, 1), dtype=torch.float32, device=dev)
    model.max_radii2D = torch.zeros((n,), dtype=torch.float32, device=dev)


@torch.no_grad()
def add_densification_stats(model, render_pkg, update_max_radii=True):
    """`gaussians.max_radii2D[vis] = max(max_radii2D[vis], radii[vis])` (train.py:151) and
    `add_densification_stats(render_pkg)` (gaussian_model.py:863-867) in one launch.
    update_max_radii=False leaves max_radii2D alone (the reference method on its own)."""
    grad = render_pkg["viewspace_points"].grad
    if grad is None:
        raise RuntimeError("add_densification_stats: viewspace_points has no gradient (call backward first)")
    radii = render_pkg["radii"]
    if radii.dtype != torch.int32:
        radii = radii.to(torch.int32)
    n, dev = model.get_pts_num, model.xyz.device
    if grad.shape[0] != n or not hasattr(model, "xyz_gradient_accum") or model.xyz_gradient_accum.shape[0] != n:
        raise RuntimeError("add_densification_stats: statistics do not match the model (call training_setup)")
    grad, radii = grad.contiguous(), radii.contiguous()
    with torch.cuda.device(dev):
        st = L.load().adgs_densify_stats(n, L.ptr(grad), L.ptr(radii), L.ptr(model.xyz_gradient_accum),
                                         L.ptr(model.denom), L.ptr(model.max_radii2D) if update_max_radii else None,
                                         _stream(dev))
    L.check(st, "densify_stats")


def _planes_width(model, name, t):
    """(planes, width, rows) of a planar array as the gather sees it."""
    if name in ("xyz", "scaling"):
        return 1, 3, t.shape[0]
    if name == "rotation":
        return 1, 4, t.shape[0]
    if name == "opacity":
        return 1, 1, t.shape[0]
    if name in ("sh4", "shs_deform4"):
        return t.shape[0], 4, t.shape[1]
    if name == "xyz_deform":
        return t.shape[0] * 3, 1, t.shape[2]
    if name == "rot_deform":
        return t.shape[0], 4, t.shape[1]
    if name == "gs_time_sigma":
        return 1, 2, t.shape[0]
    raise KeyError(name)


def _new_shape(name, t, rows):
    if name in ("sh4", "shs_deform4", "rot_deform"):
        return (t.shape[0], rows) + tuple(t.shape[2:])
    if name == "xyz_deform":
        return (t.shape[0], 3, rows)
    return (rows,) + tuple(t.shape[1:])


def _apply_plan(model, params, totals, z_scene=None, z_obj=None, keep_statistics=False):
    """plan + gather (+ split) for the row counts `totals`; installs the new arrays in model and optimizer."""
    lib = L.load()
    dev = model.xyz.device
    ns, no = model.n_scene, model.n_obj
    n_split = params.n_split if params.mode == L.DENSIFY_AND_PRUNE else 1
    ns2 = totals[0] + totals[1] + n_split * totals[2]
    no2 = totals[4] + totals[5] + n_split * totals[6]
    n2 = ns2 + no2
    src = torch.empty((max(n2, 1),), dtype=torch.int32, device=dev)
    tag = torch.empty((max(n2, 1),), dtype=torch.int32, device=dev)
    opt = model.__dict__.get("optimizer")
    with torch.cuda.device(dev):
        stream = _stream(dev)
        L.check(lib.adgs_densify_plan(C.byref(params), L.ptr(model._densify_ws), totals, L.ptr(src), L.ptr(tag), stream),
                "densify_plan")
        segs, new, new_state = [], {}, {}

        def add(old, name, obj_only, zero_new):
            planes, width, rows = _planes_width(model, name, old)
            out = torch.empty(_new_shape(name, old, no2 if obj_only else n2), dtype=torch.float32, device=dev)
            segs.append(L.GatherSegment(src=L.ptr(old), dst=L.ptr(out), planes=planes, width=width, src_rows=rows,
                                        dst_rows=no2 if obj_only else n2, src_row0=ns if obj_only else 0,
                                        dst_row0=ns2 if obj_only else 0, zero_new=int(zero_new)))
            return out

        for name, obj_only in _PER_GAUSSIAN:
            p = getattr(model, name)
            new[name] = add(p.detach(), name, obj_only, False)
            if opt is not None:
                new_state[name] = {w: add(opt.state[name][w], name, obj_only, True) for w in ("exp_avg", "exp_avg_sq")}
        gs_time_new = torch.empty((no2,), dtype=torch.float32, device=dev)
        segs.append(L.GatherSegment(src=L.ptr(model.gs_time), dst=L.ptr(gs_time_new), planes=1, width=1, src_rows=no,
                                    dst_rows=no2, src_row0=ns, dst_row0=ns2, zero_new=0))
        stats_new = {}
        if keep_statistics and hasattr(model, "xyz_gradient_accum"):
            for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
                old = getattr(model, k)
                out = torch.empty((n2,) + tuple(old.shape[1:]), dtype=torch.float32, device=dev)
                segs.append(L.GatherSegment(src=L.ptr(old), dst=L.ptr(out), planes=1, width=1, src_rows=ns + no,
                                            dst_rows=n2, src_row0=0, dst_row0=0, zero_new=0))
                stats_new[k] = out
        segs = [s for s in segs if s.dst_rows > 0 and s.planes > 0]
        if len(segs) > L.GATHER_MAX_SEGMENTS:
            raise RuntimeError("densify: too many arrays for one gather launch")
        if segs:
            arr = (L.GatherSegment * len(segs))(*segs)
            L.check(lib.adgs_densify_gather(arr, len(segs), L.ptr(src), L.ptr(tag), stream), "densify_gather")
        if params.mode == L.DENSIFY_AND_PRUNE and (totals[2] or totals[6]):
            L.check(lib.adgs_densify_split(n2, ns2, L.ptr(src), L.ptr(tag), L.ptr(model.xyz), L.ptr(model.scaling),
                                           L.ptr(model.rotation), L.ptr(z_scene), L.ptr(z_obj), n_split,
                                           L.ptr(new["xyz"]), L.ptr(new["scaling"]), stream), "densify_split")
    # ---- install (replace_tensor / cat_tensors_to_optimizer: new Parameters, moments carried over) ---------
    model.n_scene, model.n_obj = int(ns2), int(no2)
    for name, _ in _PER_GAUSSIAN:
        setattr(model, name, nn.Parameter(new[name]))
    model.gs_time = gs_time_new
    if opt is not None:
        for name, _ in _PER_GAUSSIAN:
            opt.state[name] = new_state[name]
        for g in opt.param_groups:
            g["params"] = opt._group_views(g["name"])
    if keep_statistics:
        for k, v in stats_new.items():
            setattr(model, k, v)
    elif hasattr(model, "xyz_gradient_accum"):
        setup_statistics(model)             # densification_postfix resets the statistics (gaussian_model.py:708-711)
    model.__dict__.pop("_active_cols", None)
    return src, tag


def _classify(model, params, prune_mask=None):
    lib = L.load()
    dev = model.xyz.device
    ws_bytes = lib.adgs_densify_workspace_bytes(model.n_scene, model.n_obj)
    ws = model.__dict__.get("_densify_ws")
    if ws is None or ws.numel() < ws_bytes or ws.device != dev:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        model.__dict__["_densify_ws"] = ws
    totals = (C.c_int32 * 8)()
    with torch.cuda.device(dev):
        st = lib.adgs_densify_classify(
            C.byref(params), L.ptr(getattr(model, "xyz_gradient_accum", None)), L.ptr(getattr(model, "denom", None)),
            L.ptr(model.scaling), L.ptr(model.opacity), L.ptr(prune_mask), L.ptr(ws), totals, _stream(dev))
    L.check(st, "densify_classify")
    return totals


@torch.no_grad()
def densify_and_prune(model, max_scene_grad, max_obj_grad, min_opacity, prune_big_points, N=2, sample_fn=None):
    """GaussianModel.densify_and_prune (gaussian_model.py:835-861). `sample_fn(rows) -> (rows, 3)` unit normals,
    default torch.randn on the model's device; it is called for the scene rows first, then the object rows,
    like the two torch.normal calls of densify_and_split. Returns (src, tag): the source row and kind of every
    new row (tag & 3: 0 kept, 1 clone, 2 split child)."""
    if not hasattr(model, "xyz_gradient_accum"):
        raise RuntimeError("densify_and_prune: no densification statistics (call training_setup first)")
    dev = model.xyz.device
    scene_extent, object_extent = float(model.scene_extent), float(model.object_extent)
    pd = float(model.percent_dense)
    params = L.DensifyParams(
        N_scene=model.n_scene, N_obj=model.n_obj, mode=L.DENSIFY_AND_PRUNE, n_split=int(N),
        max_scene_grad=max_scene_grad, max_obj_grad=max_obj_grad, scene_split_size=scene_extent * pd,
        obj_split_size=object_extent * pd, min_opacity=min_opacity, prune_big=int(bool(prune_big_points)),
        scene_big_size=scene_extent * 0.05, obj_big_size=object_extent * 0.1)
    totals = _classify(model, params)
    sample_fn = sample_fn or (lambda rows: torch.randn((rows, 3), dtype=torch.float32, device=dev))
    z_scene = sample_fn(int(N) * totals[3]).to(device=dev, dtype=torch.float32).contiguous()
    z_obj = sample_fn(int(N) * totals[7]).to(device=dev, dtype=torch.float32).contiguous()
    if z_scene.shape != (int(N) * totals[3], 3) or z_obj.shape != (int(N) * totals[7], 3):
        raise ValueError("densify_and_prune: sample_fn returned the wrong shape")
    plan = _apply_plan(model, params, totals, z_scene, z_obj)
    set_obj_near_idx(model)
    return plan


@torch.no_grad()
def prune_points(model, scene_mask, obj_mask):
    """GaussianModel.prune_points: boolean masks (scene rows, object rows) of the Gaussians to REMOVE."""
    dev = model.xyz.device
    if scene_mask.shape[0] != model.n_scene or obj_mask.shape[0] != model.n_obj:
        raise ValueError("prune_points: mask sizes do not match the model")
    mask = torch.cat([scene_mask.to(dev).reshape(-1), obj_mask.to(dev).reshape(-1)]).to(torch.uint8).contiguous()
    params = L.DensifyParams(N_scene=model.n_scene, N_obj=model.n_obj, mode=L.DENSIFY_PRUNE_ONLY, n_split=1)
    totals = _classify(model, params, prune_mask=mask)
    return _apply_plan(model, params, totals, keep_statistics=True)


@torch.no_grad()
def reset_opacity(model, cap=0.01):
    """GaussianModel.reset_opacity: opacity = inverse_sigmoid(min(sigmoid(opacity), 0.01)) and zeroed Adam
    moments for both opacity groups (replace_tensor_to_optimizer)."""
    dev = model.xyz.device
    opt = model.__dict__.get("optimizer")
    m = opt.state["opacity"]["exp_avg"] if opt is not None else None
    v = opt.state["opacity"]["exp_avg_sq"] if opt is not None else None
    with torch.cuda.device(dev):
        st = L.load().adgs_reset_opacity(model.get_pts_num, cap, L.ptr(model.opacity.data), L.ptr(m), L.ptr(v),
                                         _stream(dev))
    L.check(st, "reset_opacity")
    # the reference installs a NEW Parameter (replace_tensor_to_optimizer, scene/gaussian_model.py:447-461): it has
    # no gradient, so an optimizer.step() later in the same iteration skips the freshly reset opacities
    model.opacity.grad = None


@torch.no_grad()
def knn_points(anchors, points, K, return_dists=False):
    """idx (A,K) int64 of the K nearest `points` of every anchor, ascending squared distance
    (pytorch3d.ops.knn_points(anchor[None], xyz[None], K=K).idx.squeeze(0))."""
    lib = L.load()
    if anchors.dim() != 2 or points.dim() != 2 or anchors.shape[1] != points.shape[1]:
        raise ValueError("knn_points: anchors (A,D) and points (P,D) expected")
    dev = points.device
    a = anchors.to(device=dev, dtype=torch.float32).contiguous()
    p = points.to(dtype=torch.float32).contiguous()
    A, P, D = a.shape[0], p.shape[0], p.shape[1]
    idx = torch.empty((A, K), dtype=torch.int64, device=dev)
    dists = torch.empty((A, K), dtype=torch.float32, device=dev) if return_dists else None
    ws = torch.empty((lib.adgs_knn_points_workspace_bytes(A, P, K),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = lib.adgs_knn_points(A, P, D, K, L.ptr(a), L.ptr(p), L.ptr(idx), L.ptr(dists), L.ptr(ws), _stream(dev))
    L.check(st, "knn_points")
    return (idx, dists) if return_dists else idx


@torch.no_grad()
def set_obj_near_idx(model, K=None):
    """GaussianModel.set_obj_near_idx (gaussian_model.py:825-833): P // K random anchors among the object
    Gaussians (torch.randperm on the device, like the reference) and their K nearest object Gaussians in
    (x, y, z[, gs_time * scene_extent])."""
    if not getattr(model, "use_near_idx", False):
        return
    K = int(model.near_num if K is None else K)
    no, ns = model.n_obj, model.n_scene
    xyz = model.xyz.detach()[ns:]
    if model.use_time_mask:
        xyz = torch.cat([xyz, model.gs_time.reshape(no, 1) * float(model.scene_extent)], dim=-1)
    anchor = xyz[torch.randperm(no, device=xyz.device)[:no // K]]
    if anchor.shape[0] == 0:
        model.obj_near_idx = torch.empty((0, K), dtype=torch.int64, device=xyz.device)
        return
    model.obj_near_idx = knn_points(anchor, xyz, K)
