"""Host-side mirror of the hot-path part of `scene/gaussian_model.py:GaussianModel`.

Same names and meaning for everything `gaussian_renderer.render()` touches
(`get_xyz`, `get_scaling`, `get_obj_mask`, `get_deformed_xyz`, `get_deformed_pkg`,
`active_sh_degree`, `order_args`, ... scene/gaussian_model.py:88-231), but the parameters live in
the B200 layout of include/adgs_b200.h:adgs_model -- Gaussians ordered [scene ; object], wide
per-Gaussian blocks planar so that warps read whole 128-byte lines:

    xyz (N,3)  scaling (N,3)  rotation (N,4)  opacity (N,1)             raw, pre-activation
    sh4 (12,N,4)                 the (16,3) SH block, flattened, in float4 chunks
    shs_deform4 (ceil(3*Cs/4),N,4)
    xyz_deform (Cx,3,N_obj)      rot_deform (Cr,N_obj,4)      background_deform (3,Cb)
    gs_time (N_obj,)             gs_time_sigma (N_obj,2)

`from_reference()` / `to_reference()` convert from / to the reference's tensors
(`_scene_xyz`, `_obj_xyz`, `xyz_deform_param (N_obj,3,Cx)`, ... gaussian_model.py:46-84,285-328),
so checkpoints and parity tests speak the reference layout.

The optimizer (adgs_b200/optimizer.py), densification / pruning / near-index K-NN (adgs_b200/densify.py) and the
PLY + deform.pth checkpoint formats (adgs_b200/checkpoint.py) hang off the same class under the reference's
method names.
"""
import ctypes as C
import math
import time

import numpy as np
import torch
from torch import nn

from . import _lib as L

DEFAULT_ORDER_ARGS = {  # arguments/__init__.py:71-77
    'xyz': [None, None, 0, None, 0, 0],
    'rotation': [0, 0, 0, 0, None, None],
    'shs': [0, 0, 0, None, 0, 0],
    'background': [0, 0, 0, 0, 0, 0],
}


def get_param_num(args):
    """utils/func_utils.py:79-80"""
    return args[0] + args[2] + 2 * args[3] + args[4]


def set_default_param_order(order_args: dict, frame_num: int, downsample_ratio: int = 3):
    """utils/func_utils.py:82-119: fill the `None`s of the per-attribute 6-tuples
    [bspline_ctrl, bspline_order, poly, fft, quat_ctrl, quat_order]."""
    res = dict()
    for k, v in order_args.items():
        a = v if v is not None else [None] * 6
        assert a[0] is None or a[0] >= 0, f'The B-Spline ctrl pts num cannot be negative in {k}, but find {a[0]}.'
        n_b = a[0] if a[0] is not None else int(frame_num // downsample_ratio)
        k_b = 0
        if n_b > 0:
            assert a[1] is None or a[1] >= 0, f'The B-Spline order cannot be negative in {k}, but find {a[1]}.'
            if a[1] is not None and a[1] + 1 > n_b:
                print('[WARNING] The B-Spline order should be lower than the ctrl pts num. Set order to', n_b - 1)
            k_b = min(a[1] if a[1] is not None else 5, n_b - 1)
        assert a[2] is None or a[2] >= 0, f'The poly order cannot be negative in {k}, but find {a[2]}.'
        n_poly = a[2] if a[2] is not None else int(frame_num // downsample_ratio)
        assert a[3] is None or a[3] >= 0, f'The fft order cannot be negative in {k}, but find {a[3]}.'
        n_fft = a[3] if a[3] is not None else 6
        assert a[4] is None or a[4] >= 0, \
            f'The quaternion spline ctrl pts num cannot be negative in {k}, but find {a[4]}.'
        n_q = a[4] if a[4] is not None else int(frame_num // downsample_ratio)
        k_q = 0
        if n_q > 0:
            assert a[5] is None or a[5] >= 0, f'The quaternion spline order cannot be negative in {k}, but find {a[5]}.'
            if a[5] is not None and a[5] + 1 > n_q:
                print('[WARNING] The quaternion spline order should be lower than the ctrl pts num. Set order to',
                      n_q - 1)
            k_q = min(a[5] if a[5] is not None else 1, n_q - 1)
        res[k] = [n_b, k_b, n_poly, n_fft, n_q, k_q]
    return res


_M = {}


def deboor_cox_matrix(order: int) -> np.ndarray:
    """Uniform B-spline basis matrix M_k, B(u) = [1,u,...,u^k] M_k, via the de Boor-Cox recursion
    (same float32 arithmetic as utils/func_utils.py:33-50 so M_1..M_5 are identical)."""
    if order in _M:
        return _M[order]
    if order == 0:
        m = np.array([[1.0]], dtype=np.float32)
    else:
        prev = deboor_cox_matrix(order - 1)
        z = np.zeros((1, prev.shape[1]), dtype=np.float32)
        up, lo = np.concatenate([prev, z], 0), np.concatenate([z, prev], 0)
        a = np.zeros((order, order + 1), dtype=np.float32)
        b = np.zeros((order, order + 1), dtype=np.float32)
        i = np.arange(order)
        a[i, i] = i + 1
        a[i, i + 1] = order - i - 1
        b[i, i] = -1
        b[i, i + 1] = 1
        m = (up @ a + lo @ b) / order
    _M[order] = m
    return m


def _bspline_window(v: float, n: int, k: int):
    """(first control index, basis values) of func_utils.py:127-132."""
    interval = n - k
    start = min(int(v * interval), interval - 1)
    u = v * interval - start
    powers = np.array([u ** i for i in range(k + 1)], dtype=np.float64)
    return start, powers @ deboor_cox_matrix(k).astype(np.float64)


def linear_terms(v: float, args):
    """{column: weight} such that get_func_result's linear part == sum_c param[..., c] * weight[c]
    (B-spline window + polynomial + Fourier, func_utils.py:127-153)."""
    n_b, k_b, n_poly, n_fft, n_q, k_q = args
    terms, off = {}, 0
    if n_b != 0:
        s, basis = _bspline_window(v, n_b, k_b)
        for j in range(k_b + 1):
            terms[off + s + j] = float(basis[j])
        off += n_b
    if n_poly != 0:
        for i in range(1, n_poly + 1):
            terms[off + i - 1] = float(v ** i)
        off += n_poly
    if n_fft != 0:
        for i in range(1, n_fft + 1):
            terms[off + i - 1] = math.sin(v * i * math.pi)
            terms[off + n_fft + i - 1] = math.cos(v * i * math.pi)
        off += 2 * n_fft
    return terms, off


def _fill_lin(dst: L.LinBasis, args, t, t2=None):
    terms0, _ = linear_terms(t, args)
    terms1 = linear_terms(t2, args)[0] if t2 is not None else {}
    cols = sorted(set(terms0) | set(terms1))
    if len(cols) > L.MAX_TERMS:
        raise ValueError(f"{len(cols)} active trajectory columns exceed ADGS_MAX_TERMS={L.MAX_TERMS}")
    dst.n = len(cols)
    dst.n_cols = get_param_num(args)
    for i, c in enumerate(cols):
        dst.col[i] = c
        dst.w0[i] = terms0.get(c, 0.0)
        dst.w1[i] = terms1.get(c, 0.0)


def make_time_basis(order_args, t: float, flow_t=None, use_time_mask=True) -> L.TimeBasis:
    """Everything time-dependent the kernels need for one render, computed on the host in float64
    (the reference also derives segment index and u on the host, func_utils.py:128-132)."""
    tb = L.TimeBasis()
    _fill_lin(tb.xyz, order_args['xyz'], t, flow_t)
    _fill_lin(tb.background, order_args['background'], t, flow_t)
    _fill_lin(tb.shs, order_args['shs'], t)
    ra = order_args['rotation']
    _fill_lin(tb.rotation, ra, t)
    tb.quat.n_ctrl, tb.quat.k = ra[4], ra[5]
    if ra[4] != 0:
        if ra[5] > L.MAX_QUAT_ORDER:
            raise ValueError("quaternion spline order > 7 is not supported")
        s, basis = _bspline_window(t, ra[4], ra[5])
        tb.quat.start = ra[0] + ra[2] + 2 * ra[3] + s
        for i in range(1, ra[5] + 1):
            tb.quat.cum[i] = float(basis[i:].sum())
    tb.t = float(t)
    tb.use_time_mask = int(bool(use_time_mask))
    tb.has_flow = int(flow_t is not None)
    return tb


PARAM_NAMES = ("xyz", "scaling", "rotation", "opacity", "sh4", "shs_deform4", "xyz_deform", "rot_deform",
               "background_deform", "gs_time_sigma")


class GaussianModel(nn.Module):
    """Hot-path subset of scene/gaussian_model.py:GaussianModel over planar B200 storage."""

    def __init__(self, sh_degree: int, order_args: dict):
        super().__init__()
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.order_args = order_args
        self.use_time_mask = None    # resolved by training_setup (lambda_sigma > 0) unless set, like gaussian_model.py:78,394
        self.n_scene = 0
        self.n_obj = 0
        self._binning_capacity = 0   # running bound on num_rendered for the sync-free path
        # sync-free forwards whose counters the next forward does not wait for. 0: the host trails the device by less
        # than one step; 1-2 made no difference to the end-to-end step time at 1 M Gaussians (the host needs 0.87 ms
        # to queue a 1.41 ms step: profiles/r2_sf_sync_free_depth.txt), so the conservative value stays
        self.sync_free_outstanding = 0
        self._pending = None         # (pinned counters, event) of the last sync-free forward

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_reference(cls, ref: dict, order_args: dict, sh_degree=3, use_time_mask=True, device="cuda"):
        """ref: the reference's tensors by attribute name without the leading underscore
        (scene_xyz, obj_xyz, scene_shs_dc, ..., xyz_deform_param, ..., gs_time, gs_time_sigma)."""
        m = cls(sh_degree, order_args)
        m.use_time_mask = use_time_mask
        g = lambda k: ref[k].detach().to(device=device, dtype=torch.float32)
        cat = lambda a, b: torch.cat([g(a), g(b)], dim=0)
        m.n_scene, m.n_obj = ref["scene_xyz"].shape[0], ref["obj_xyz"].shape[0]
        n, no = m.n_scene + m.n_obj, m.n_obj
        shs = torch.cat([cat("scene_shs_dc", "obj_shs_dc"), cat("scene_shs_rest", "obj_shs_rest")], dim=1)  # (N,16,3)
        shsd = cat("shs_deform_param_scene", "shs_deform_param_obj").reshape(n, -1)                        # (N,3*Cs)
        pad = (-shsd.shape[1]) % 4
        if pad:
            shsd = torch.cat([shsd, shsd.new_zeros(n, pad)], dim=1)
        params = dict(
            xyz=cat("scene_xyz", "obj_xyz"), scaling=cat("scene_scaling", "obj_scaling"),
            rotation=cat("scene_rotation", "obj_rotation"), opacity=cat("scene_opacity", "obj_opacity"),
            sh4=shs.reshape(n, 12, 4).permute(1, 0, 2), shs_deform4=shsd.reshape(n, -1, 4).permute(1, 0, 2),
            xyz_deform=g("xyz_deform_param").permute(2, 1, 0), rot_deform=g("rotation_deform_param").permute(2, 0, 1),
            background_deform=g("background_deform_param").reshape(3, -1), gs_time_sigma=g("gs_time_sigma"))
        for k, v in params.items():
            setattr(m, k, nn.Parameter(v.contiguous().clone()))
        m.register_buffer("gs_time", g("gs_time").reshape(no).contiguous().clone())
        m.active_sh_degree = sh_degree
        return m

    @classmethod
    def create_from_pcd(cls, pcd, scene_extent, cameras_extent, frame_gap, default_order_downsample_ratio,
                        sh_degree=3, order_args=None, use_time_mask=None, device="cuda"):
        """GaussianModel.create_from_pcd (scene/gaussian_model.py:255-333): initial Gaussians from a point cloud
        with `.points (P,3)`, `.colors (P,3)` in [0,1], `.time (P,1)`, `.obj_id (P,1)` (utils/graphics_utils.py:17-22).
        Scales from the mean squared distance to the 3 nearest neighbours (distCUDA2 -> adgs_dist_cuda2), identity
        rotations, opacity 0.1, DC colour = RGB2SH, deformation parameters U(-1,1) * 1e-5 drawn with torch.rand in
        the reference's order (xyz, rotation, shs, background), gs_time_sigma = log(frame_gap)."""
        from .simple_knn import distCUDA2
        order_args = set_default_param_order(order_args if order_args is not None else DEFAULT_ORDER_ARGS,
                                             int(1.0 / frame_gap), default_order_downsample_ratio)
        pts = torch.tensor(np.asarray(pcd.points)).float().to(device)
        n = pts.shape[0]
        fused_color = (torch.tensor(np.asarray(pcd.colors)).float().to(device) - 0.5) / 0.28209479177387814  # RGB2SH
        shs = torch.zeros((n, 3, (sh_degree + 1) ** 2), dtype=torch.float32, device=device)
        shs[:, :3, 0] = fused_color
        shs = shs.transpose(1, 2)                                                     # (P,16,3)
        dist2 = torch.clamp_min(distCUDA2(pts), 0.0000001)
        scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
        rots = torch.zeros((n, 4), device=device)
        rots[:, 0] = 1.0
        opac = torch.full((n, 1), math.log(0.1 / 0.9), dtype=torch.float32, device=device)  # inverse_sigmoid(0.1)
        scene_mask = torch.tensor(np.asarray(pcd.obj_id)[..., 0] <= 0.5, dtype=torch.bool, device=device)
        obj_mask = torch.logical_not(scene_mask)
        no = int(obj_mask.sum())
        U = lambda *shape: (torch.rand(shape, device=device, dtype=torch.float32) * 2.0 - 1.0) * 1e-5
        xyz_deform = U(no, 3, get_param_num(order_args['xyz']))
        rot_deform = U(no, 4, get_param_num(order_args['rotation']))
        shs_deform = U(n, 3, get_param_num(order_args['shs']))
        bg_deform = U(1, 3, get_param_num(order_args['background']))
        ref = dict(
            scene_xyz=pts[scene_mask], obj_xyz=pts[obj_mask],
            scene_shs_dc=shs[scene_mask, 0:1], obj_shs_dc=shs[obj_mask, 0:1],
            scene_shs_rest=shs[scene_mask, 1:], obj_shs_rest=shs[obj_mask, 1:],
            scene_scaling=scales[scene_mask], obj_scaling=scales[obj_mask],
            scene_rotation=rots[scene_mask], obj_rotation=rots[obj_mask],
            scene_opacity=opac[scene_mask], obj_opacity=opac[obj_mask],
            xyz_deform_param=xyz_deform, rotation_deform_param=rot_deform,
            shs_deform_param_scene=shs_deform[scene_mask], shs_deform_param_obj=shs_deform[obj_mask],
            background_deform_param=bg_deform,
            gs_time=torch.tensor(np.asarray(pcd.time), device=device, dtype=torch.float32)[obj_mask],
            gs_time_sigma=torch.full((no, 2), float(np.log(frame_gap)), dtype=torch.float32, device=device))
        m = cls.from_reference(ref, order_args, sh_degree=sh_degree, use_time_mask=use_time_mask, device=device)
        m.active_sh_degree = 0
        m.scene_extent, m.cameras_extent, m.object_extent, m.frame_gap = scene_extent, cameras_extent, 10.0, frame_gap
        m.max_radii2D = torch.zeros((n,), device=device)
        return m

    def reference_layout(self, t: dict) -> dict:
        """Planar arrays (keys = PARAM_NAMES, e.g. parameters, gradients or optimizer moments) -> the
        reference's tensors by attribute name (inverse of planar_layout)."""
        ns, n = self.n_scene, self.n_scene + self.n_obj
        xyz, sc, rot, op = t["xyz"], t["scaling"], t["rotation"], t["opacity"]
        shs = t["sh4"].permute(1, 0, 2).reshape(n, 16, 3)
        cs = get_param_num(self.order_args['shs'])
        shsd = t["shs_deform4"].permute(1, 0, 2).reshape(n, -1)[:, :3 * cs].reshape(n, 3, cs)
        return dict(
            scene_xyz=xyz[:ns], obj_xyz=xyz[ns:], scene_scaling=sc[:ns], obj_scaling=sc[ns:],
            scene_rotation=rot[:ns], obj_rotation=rot[ns:], scene_opacity=op[:ns], obj_opacity=op[ns:],
            scene_shs_dc=shs[:ns, 0:1], obj_shs_dc=shs[ns:, 0:1], scene_shs_rest=shs[:ns, 1:], obj_shs_rest=shs[ns:, 1:],
            shs_deform_param_scene=shsd[:ns], shs_deform_param_obj=shsd[ns:],
            xyz_deform_param=t["xyz_deform"].permute(2, 1, 0),
            rotation_deform_param=t["rot_deform"].permute(1, 2, 0),
            background_deform_param=t["background_deform"].reshape(1, 3, -1),
            gs_time_sigma=t["gs_time_sigma"])

    def planar_layout(self, ref: dict) -> dict:
        """The reference's tensors by attribute name -> planar arrays (keys = PARAM_NAMES)."""
        dev = self.xyz.device
        g = lambda k: ref[k].detach().to(device=dev, dtype=torch.float32)
        cat = lambda a, b: torch.cat([g(a), g(b)], dim=0)
        n = self.n_scene + self.n_obj
        shs = torch.cat([cat("scene_shs_dc", "obj_shs_dc"), cat("scene_shs_rest", "obj_shs_rest")], dim=1)
        shsd = cat("shs_deform_param_scene", "shs_deform_param_obj").reshape(n, -1)
        pad = (-shsd.shape[1]) % 4
        if pad:
            shsd = torch.cat([shsd, shsd.new_zeros(n, pad)], dim=1)
        out = dict(
            xyz=cat("scene_xyz", "obj_xyz"), scaling=cat("scene_scaling", "obj_scaling"),
            rotation=cat("scene_rotation", "obj_rotation"), opacity=cat("scene_opacity", "obj_opacity"),
            sh4=shs.reshape(n, 12, 4).permute(1, 0, 2), shs_deform4=shsd.reshape(n, -1, 4).permute(1, 0, 2),
            xyz_deform=g("xyz_deform_param").permute(2, 1, 0), rot_deform=g("rotation_deform_param").permute(2, 0, 1),
            background_deform=g("background_deform_param").reshape(3, -1), gs_time_sigma=g("gs_time_sigma"))
        return {k: v.contiguous() for k, v in out.items()}

    def to_reference(self, grads=False) -> dict:
        """Inverse of from_reference (parameters, or their .grad when grads=True)."""
        pick = (lambda p: p.grad) if grads else (lambda p: p.detach())
        out = self.reference_layout({k: pick(getattr(self, k)) for k in PARAM_NAMES})
        if not grads:
            out["gs_time"] = self.gs_time.reshape(-1, 1)
        return out

    # ---- optimizer (scene/gaussian_model.py:337-413), see adgs_b200/optimizer.py ------------------------
    def training_setup(self, training_args, window_aware=False):
        from .optimizer import training_setup
        return training_setup(self, training_args, window_aware=window_aware)

    def update_learning_rate(self, iteration):
        from .optimizer import update_learning_rate
        update_learning_rate(self, iteration)

    # ---- densification (scene/gaussian_model.py:463-467, 560-867), see adgs_b200/densify.py ----------------
    def add_densification_stats(self, render_pkg, update_max_radii=True):
        from .densify import add_densification_stats
        add_densification_stats(self, render_pkg, update_max_radii=update_max_radii)

    def densify_and_prune(self, max_scene_grad, max_obj_grad, min_opacity, prune_big_points, **kw):
        from .densify import densify_and_prune
        return densify_and_prune(self, max_scene_grad, max_obj_grad, min_opacity, prune_big_points, **kw)

    def prune_points(self, scene_mask, obj_mask):
        from .densify import prune_points
        return prune_points(self, scene_mask, obj_mask)

    def reset_opacity(self):
        from .densify import reset_opacity
        reset_opacity(self)

    def set_obj_near_idx(self, K=None):
        from .densify import set_obj_near_idx
        set_obj_near_idx(self, K)

    # ---- checkpoints (scene/gaussian_model.py:415-543), see adgs_b200/checkpoint.py -------------------------
    def save_ply(self, path):
        from .checkpoint import save_ply
        save_ply(self, path)

    def load_ply(self, path, device="cuda"):
        from .checkpoint import load_ply
        load_ply(self, path, device=device)

    def _note_active_columns(self, tb):
        """Control-point columns the backward of this render writes (window-aware optimizer step)."""
        act = self.__dict__.setdefault("_active_cols", {"xyz": set(), "rotation": set(), "backwards": 0})
        act["xyz"].update(tb.xyz.col[i] for i in range(tb.xyz.n))
        act["rotation"].update(tb.rotation.col[i] for i in range(tb.rotation.n))
        if tb.quat.n_ctrl:
            act["rotation"].update(range(tb.quat.start, tb.quat.start + tb.quat.k + 1))
        act["backwards"] += 1

    def active_columns(self):
        act = self.__dict__.get("_active_cols")
        if act is None or act["backwards"] == 0:
            raise RuntimeError("window-aware optimizer step without a render backward since the last step")
        return {"xyz": sorted(act["xyz"]), "rotation": sorted(act["rotation"])}

    def reset_active_columns(self):
        self.__dict__["_active_cols"] = {"xyz": set(), "rotation": set(), "backwards": 0}

    @property
    def sparse_deform_grads(self):
        """True while a window-aware optimizer owns the gradients: render backward then leaves inactive
        control-point planes unwritten (one backward per optimizer step only)."""
        opt = self.__dict__.get("optimizer")
        return bool(opt is not None and getattr(opt, "window_aware", False) and opt.window_eligible())

    def shard(self, rank: int, world: int):
        """Rank `rank`'s slice of the model for the splat-exchange multi-GPU path: every world-th scene Gaussian
        and every world-th object Gaussian starting at `rank` (row i of a block belongs to rank i % world -- a strided
        cut balances culled / visible and near / far Gaussians across the ranks, whatever order the arrays are in),
        padded to equal sizes across ranks with fully transparent Gaussians (opacity logit -1e30 => alpha = 0: they
        contribute nothing anywhere)."""
        ref = self.to_reference()
        ns, no = -(-self.n_scene // world), -(-self.n_obj // world)

        def cut(t, n_per, pad_value=0.0):
            part = t[rank::world]
            if part.shape[0] < n_per:
                pad = torch.full((n_per - part.shape[0],) + tuple(t.shape[1:]), pad_value, dtype=t.dtype, device=t.device)
                part = torch.cat([part, pad], dim=0)
            return part.contiguous()

        out = {}
        for k, v in ref.items():
            if k == "background_deform_param":
                out[k] = v.clone()
                continue
            per = ns if (k.startswith("scene_") or k == "shs_deform_param_scene") else no
            out[k] = cut(v, per, -1e30 if k.endswith("_opacity") else 0.0)
        m = GaussianModel.from_reference(out, self.order_args, sh_degree=self.max_sh_degree,
                                         use_time_mask=self.use_time_mask, device=self.xyz.device)
        m.active_sh_degree = self.active_sh_degree
        return m

    def hot_parameters(self):
        return [getattr(self, k) for k in PARAM_NAMES]

    # ---- reference-named accessors (scene/gaussian_model.py:88-171) ---------------------------
    @property
    def get_pts_num(self):
        return self.n_scene + self.n_obj

    @property
    def get_scene_pts_num(self):
        return self.n_scene

    @property
    def get_obj_pts_num(self):
        return self.n_obj

    @property
    def get_xyz(self):
        return self.xyz

    @property
    def get_scaling(self):
        return torch.exp(self.scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self.rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self.opacity)

    @property
    def get_shs(self):
        n = self.get_pts_num
        return self.sh4.permute(1, 0, 2).reshape(n, 16, 3)

    @property
    def get_obj_mask(self):
        mask = torch.zeros((self.get_pts_num,), dtype=torch.bool, device=self.xyz.device)
        mask[self.n_scene:] = True
        return mask

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    # ---- C-ABI views ----------------------------------------------------------------------------
    def c_model_from(self, t: dict, with_time=True) -> L.Model:
        """adgs_model over the given tensors (parameters, or gradient buffers with with_time=False)."""
        return L.Model(N_scene=self.n_scene, N_obj=self.n_obj, xyz=L.ptr(t["xyz"]), scaling=L.ptr(t["scaling"]),
                       rotation=L.ptr(t["rotation"]), opacity=L.ptr(t["opacity"]), sh4=L.ptr(t["sh4"]),
                       shs_deform4=L.ptr(t["shs_deform4"]), xyz_deform=L.ptr(t["xyz_deform"]),
                       rot_deform=L.ptr(t["rot_deform"]), background_deform=L.ptr(t["background_deform"]),
                       gs_time=L.ptr(self.gs_time) if with_time else None, gs_time_sigma=L.ptr(t["gs_time_sigma"]))

    def c_model(self) -> L.Model:
        return self.c_model_from({k: getattr(self, k) for k in PARAM_NAMES})

    def _note_num_rendered(self, R: int):
        """Sync-free binning: keep the arena 30 % above the largest num_rendered seen so far."""
        self._last_num_rendered = int(R)
        self._binning_capacity = max(self._binning_capacity, int(1.3 * R) + 65536)

    def _defer_counter_check(self, counters, event, capacity):
        self.__dict__.setdefault("_counter_checks", []).append((counters, event, int(capacity)))

    def _resolve_counter_checks(self, block: bool, outstanding: int = 0):
        """Look at the {num_rendered, overflow} words of earlier sync-free forwards: grow the binning arena,
        and report a dropped iteration (the device already zeroed its gradients). block=False only takes the
        ones whose copy has landed; block=True waits for all but the newest `outstanding` ones (the host may then
        run that many forwards ahead of the device, and an overflow is acted on that many iterations later)."""
        checks = self.__dict__.get("_counter_checks")
        if not checks:
            return
        rest = []
        for i, (counters, event, capacity) in enumerate(checks):
            must = block and i < len(checks) - outstanding
            if not must and not event.query():
                rest.append((counters, event, capacity))
                continue
            if not event.query():
                w0 = time.perf_counter()
                event.synchronize()
                self.__dict__["_host_wait_s"] = self.__dict__.get("_host_wait_s", 0.0) + time.perf_counter() - w0
            R, ovf = int(counters[0]), bool(counters[1])
            self._note_num_rendered(R)
            if ovf or R > capacity:
                import warnings
                warnings.warn("adgs_b200: the binning arena overflowed in a sync-free forward; that iteration's "
                              "images are invalid and its gradients were zeroed on the device; the arena has been "
                              "enlarged (see DESIGN.md, 'sync-free binning')")
        self.__dict__["_counter_checks"] = rest

    def time_basis(self, t, flow_t=None) -> L.TimeBasis:
        """Host-side basis for (t, flow_t); cached -- a training run revisits the same frame times."""
        key = (float(t), None if flow_t is None else float(flow_t), bool(self.use_time_mask))
        cache = self.__dict__.setdefault("_basis_cache", {})
        tb = cache.get(key)
        if tb is None:
            if len(cache) > 4096:
                cache.clear()
            tb = make_time_basis(self.order_args, t, flow_t, self.use_time_mask)
            cache[key] = tb
        return tb

    def _pinned_counters(self):
        """Small ring of pinned int32[2] buffers for the asynchronous {num_rendered, overflow} read-back."""
        ring = self.__dict__.setdefault("_counter_ring", [])
        if len(ring) < 8:
            ring.append(torch.empty((2,), dtype=torch.int32).pin_memory())
            return ring[-1]
        self.__dict__["_counter_idx"] = (self.__dict__.get("_counter_idx", -1) + 1) % len(ring)
        return ring[self.__dict__["_counter_idx"]]

    # ---- trajectory alone (drop-in for get_deformed_* ; values only, see gaussian_renderer.render for grads)
    @torch.no_grad()
    def _deform(self, t, flow_t=None, want=("xyz", "rotation", "shs", "opacity")):
        lib = L.load()
        n, dev = self.get_pts_num, self.xyz.device
        shapes = dict(xyz=(n, 3), rotation=(n, 4), shs=(n, 16, 3), opacity=(n, 1), scaling=(n, 3), flow_xyz=(n, 3))
        out = {k: torch.empty(shapes[k], dtype=torch.float32, device=dev) for k in want}
        d = L.Deformed(**{k: L.ptr(out.get(k)) for k in shapes})
        tb = self.time_basis(t, flow_t)
        with torch.cuda.device(dev):
            st = lib.adgs_trajectory_forward(C.byref(self.c_model()), C.byref(tb), C.byref(d),
                                             torch.cuda.current_stream(dev).cuda_stream)
        L.check(st, "trajectory_forward")
        return out

    def get_deformed_xyz(self, t):
        return self._deform(t, want=("xyz",))["xyz"]

    def get_deformed_rotation(self, t):
        return self._deform(t, want=("rotation",))["rotation"]

    def get_deformed_shs(self, t):
        return self._deform(t, want=("shs",))["shs"]

    def get_time_masked_opacity(self, t):
        return self._deform(t, want=("opacity",))["opacity"]

    def get_deformed_pkg(self, t):
        return self._deform(t)
