"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product path) for densification / pruning:
a numpy restatement of the reference's sequential algorithm on the reference's tensor layout,

    add_densification_stats      scene/gaussian_model.py:863-867  (+ max_radii2D update, train.py:151)
    densify_and_prune            scene/gaussian_model.py:835-861
    densify_and_clone            scene/gaussian_model.py:775-823
    densify_and_split            scene/gaussian_model.py:714-773
    prune_points/_prune_optimizer scene/gaussian_model.py:560-614
    cat_tensors_to_optimizer / densification_postfix   scene/gaussian_model.py:616-712
    reset_opacity / replace_tensor_to_optimizer        scene/gaussian_model.py:463-467, 547-558
    build_rotation               utils/general_utils.py:77-94

It deliberately keeps the reference's three-stage structure (clone, then split + prune of the split
sources, then the final prune, each by boolean masks and concatenation) so that it checks the
single-pass composition of adgs_b200/csrc/densify.cu rather than restating it.

Pinned by tests/golden/densify.npz = outputs of the reference's OWN scene/gaussian_model.py run on CPU
(tests/golden/make_densify_golden.py). `gpu_division=True` evaluates `scaling / (0.8 N)` as the torch CUDA
kernel does (multiplication by the float32 reciprocal, ATen BinaryDivTrueKernel.cu) instead of the CPU
kernel's true division; the two differ by at most one ulp.

State = dict of numpy arrays keyed by the reference's attribute names without the leading underscore;
Adam moments under "<name>.exp_avg" / "<name>.exp_avg_sq".
"""
import numpy as np

SCENE_ROWS = ("scene_xyz", "scene_shs_dc", "scene_shs_rest", "scene_opacity", "scene_scaling", "scene_rotation",
              "shs_deform_param_scene")
OBJ_ROWS = ("obj_xyz", "obj_shs_dc", "obj_shs_rest", "obj_opacity", "obj_scaling", "obj_rotation",
            "shs_deform_param_obj", "xyz_deform_param", "rotation_deform_param", "gs_time_sigma")
MOMENTS = (".exp_avg", ".exp_avg_sq")
f32 = np.float32


def sigmoid(x):
    x = x.astype(f32)
    return (f32(1) / (f32(1) + np.exp(-x))).astype(f32)


def inverse_sigmoid(x):
    """utils/general_utils.py:20-21"""
    x = x.astype(f32)
    return np.log(x / (f32(1) - x)).astype(f32)


def build_rotation(r):
    """utils/general_utils.py:77-94"""
    r = r.astype(f32)
    norm = np.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=-1)
    return R.reshape(-1, 3, 3).astype(f32)


def add_densification_stats(state, grad_means2D, radii):
    """train.py:151-152 then gaussian_model.py:863-867 (visibility_filter = radii > 0)."""
    vis = radii > 0
    state["max_radii2D"][vis] = np.maximum(state["max_radii2D"][vis], radii[vis].astype(f32))
    g = grad_means2D[vis, :2].astype(f32)
    state["xyz_gradient_accum"][vis] += np.sqrt(g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1])[:, None]
    state["denom"][vis] += f32(1)


def _select(state, names, mask):
    """_prune_optimizer: parameters and moments by a keep mask."""
    for n in names:
        state[n] = state[n][mask]
        for m in MOMENTS:
            if n + m in state:
                state[n + m] = state[n + m][mask]


def prune_points(state, scene_mask, obj_mask):
    """gaussian_model.py:585-614: masks of the rows to REMOVE."""
    _select(state, SCENE_ROWS, ~scene_mask)
    _select(state, OBJ_ROWS, ~obj_mask)
    state["gs_time"] = state["gs_time"][~obj_mask]
    valid = np.concatenate([~scene_mask, ~obj_mask])
    for k in ("xyz_gradient_accum", "denom", "max_radii2D"):
        state[k] = state[k][valid]


def _postfix(state, new):
    """cat_tensors_to_optimizer + densification_postfix: append rows, zero moments, reset the statistics."""
    for n, ext in new.items():
        if n == "gs_time":
            continue
        state[n] = np.concatenate([state[n], ext], axis=0)
        for m in MOMENTS:
            if n + m in state:
                state[n + m] = np.concatenate([state[n + m], np.zeros_like(ext)], axis=0)
    state["gs_time"] = np.concatenate([state["gs_time"], new["gs_time"]], axis=0)
    n_pts = state["scene_xyz"].shape[0] + state["obj_xyz"].shape[0]
    state["xyz_gradient_accum"] = np.zeros((n_pts, 1), f32)
    state["denom"] = np.zeros((n_pts, 1), f32)
    state["max_radii2D"] = np.zeros((n_pts,), f32)


def _max_scale(state, key):
    return np.exp(state[key].astype(f32)).max(axis=1) if state[key].shape[0] else np.zeros((0,), f32)


def densify_and_clone(state, scene_sel, obj_sel, scene_size, obj_size):
    scene_sel = scene_sel & (_max_scale(state, "scene_scaling") <= f32(scene_size))
    obj_sel = obj_sel & (_max_scale(state, "obj_scaling") <= f32(obj_size))
    new = {n: state[n][scene_sel] for n in SCENE_ROWS}
    new.update({n: state[n][obj_sel] for n in OBJ_ROWS})
    new["gs_time"] = state["gs_time"][obj_sel]
    _postfix(state, new)


def densify_and_split(state, scene_sel, obj_sel, scene_size, obj_size, z_scene, z_obj, N=2, gpu_division=True):
    scene_sel = scene_sel & (_max_scale(state, "scene_scaling") > f32(scene_size))
    obj_sel = obj_sel & (_max_scale(state, "obj_scaling") > f32(obj_size))
    new = {}
    rep = lambda a: np.concatenate([a] * N, axis=0)
    for part, sel, z in (("scene", scene_sel, z_scene), ("obj", obj_sel, z_obj)):
        get_scaling = np.exp(state[f"{part}_scaling"][sel].astype(f32))
        stds = rep(get_scaling)
        samples = (z.astype(f32).reshape(stds.shape) * stds).astype(f32)   # torch.normal(0, stds) = randn * stds
        rots = rep(build_rotation(state[f"{part}_rotation"][sel]))
        new[f"{part}_xyz"] = (np.einsum("nij,nj->ni", rots, samples).astype(f32) + rep(state[f"{part}_xyz"][sel])).astype(f32)
        if gpu_division:
            scaled = stds * (f32(1.0) / f32(0.8 * N))
        else:
            scaled = stds / f32(0.8 * N)
        new[f"{part}_scaling"] = np.log(scaled.astype(f32)).astype(f32)
        for k in ("rotation", "shs_dc", "shs_rest", "opacity"):
            new[f"{part}_{k}"] = rep(state[f"{part}_{k}"][sel])
    new["shs_deform_param_scene"] = rep(state["shs_deform_param_scene"][scene_sel])
    for k in ("shs_deform_param_obj", "xyz_deform_param", "rotation_deform_param", "gs_time_sigma", "gs_time"):
        new[k] = rep(state[k][obj_sel])
    _postfix(state, new)
    n_new_s, n_new_o = N * int(scene_sel.sum()), N * int(obj_sel.sum())
    prune_points(state, np.concatenate([scene_sel, np.zeros(n_new_s, bool)]),
                 np.concatenate([obj_sel, np.zeros(n_new_o, bool)]))


def densify_and_prune(state, max_scene_grad, max_obj_grad, min_opacity, prune_big_points, scene_extent, object_extent,
                      percent_dense, z_scene, z_obj, N=2, gpu_division=True):
    """gaussian_model.py:835-861 (without set_obj_near_idx). z_scene / z_obj: unit normals, (N * selected, 3)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        grads = state["xyz_gradient_accum"].astype(f32) / state["denom"].astype(f32)
    grads[np.isnan(grads)] = 0.0
    grads = np.abs(grads[:, 0])                                  # torch.norm over the size-1 last dim
    ns = state["scene_xyz"].shape[0]
    scene_sel = grads[:ns] >= f32(max_scene_grad)
    obj_sel = grads[ns:] >= f32(max_obj_grad)
    scene_size, obj_size = scene_extent * percent_dense, object_extent * percent_dense
    densify_and_clone(state, scene_sel, obj_sel, scene_size, obj_size)
    pad = lambda m, n: np.concatenate([m, np.zeros(n - m.shape[0], bool)])
    scene_sel = pad(scene_sel, state["scene_xyz"].shape[0])
    obj_sel = pad(obj_sel, state["obj_xyz"].shape[0])
    densify_and_split(state, scene_sel, obj_sel, scene_size, obj_size, z_scene, z_obj, N, gpu_division)
    scene_prune = sigmoid(state["scene_opacity"])[:, 0] < f32(min_opacity)
    obj_prune = sigmoid(state["obj_opacity"])[:, 0] < f32(min_opacity)
    if prune_big_points:
        scene_prune |= _max_scale(state, "scene_scaling") > f32(scene_extent * 0.05)
        obj_prune |= _max_scale(state, "obj_scaling") > f32(object_extent * 0.1)
    prune_points(state, scene_prune, obj_prune)


def reset_opacity(state):
    """gaussian_model.py:463-467: opacities capped at 0.01, moments zeroed."""
    for part in ("scene", "obj"):
        k = f"{part}_opacity"
        state[k] = inverse_sigmoid(np.minimum(sigmoid(state[k]), f32(0.01)))
        for m in MOMENTS:
            if k + m in state:
                state[k + m] = np.zeros_like(state[k])


def knn_points(anchors, points, K):
    """pytorch3d.ops.knn_points(anchor[None], xyz[None], K).idx / .dists (brute force, squared distances
    accumulated over the coordinates in order, ascending, ties by smaller index). pytorch3d is not installed
    and not vendored by the reference (parity unpinned; restated from its published behaviour)."""
    a, p = anchors.astype(f32), points.astype(f32)
    d = np.zeros((a.shape[0], p.shape[0]), f32)
    for c in range(a.shape[1]):
        diff = a[:, c:c + 1] - p[None, :, c]
        d = (d + diff * diff).astype(f32)
    idx = np.argsort(d, axis=1, kind="stable")[:, :K]
    return idx.astype(np.int64), np.take_along_axis(d, idx, axis=1)
