"""TEST INFRASTRUCTURE ONLY -- device-parametrised torch restatement of the reference trajectory.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.

Follows
  utils/func_utils.py:33-50   get_deboor_cox_mat          (de Boor-Cox basis matrix recursion)
  utils/func_utils.py:52-77   get_fft/poly/bspline_basic_func
  utils/func_utils.py:79-80   get_param_num
  utils/func_utils.py:121-173 get_func_result             (B-spline + poly + Fourier + cumulative quaternion B-spline)
  scene/gaussian_model.py:88-91,141-144,154-159,173-231   get_scaling / get_opacity / get_obj_mask / get_deformed_*
  utils/sh_utils.py:57-112    eval_sh                     (CPU restatement of computeColorFromSH)
The reference hard-codes device='cuda' (func_utils.py:54,61,72,75,159); here the device follows
the parameters so the path can run on the host cores (BASELINE config 1).

Third-party arithmetic: the quaternion maps come from roma==1.5.1 (environment.yaml:261; call
sites func_utils.py:164-169), which is NOT vendored in /root/reference and not installable here.
`quat_conjugation`, `quat_product`, `unitquat_to_rotvec` (shortest_arc=True), `rotvec_to_unitquat`
below restate roma's published algorithm (xyzw convention; small-angle Taylor branches at
|angle| <= 1e-3, mirroring scipy.spatial.transform.Rotation). PARITY UNPINNED at this boundary:
no reference test or golden vector pins roma's results; tests/test_oracle_cpu.py checks these
maps against scipy's Rotation (same convention) instead, and tests/golden/trajectory_*.npz pins
everything else against the reference's own func_utils.py run with these maps injected.
"""
import math

import numpy as np
import torch
from torch.nn.functional import normalize

_M_CACHE = {}


def get_deboor_cox_mat(order: int) -> np.ndarray:
    """(k+1)x(k+1) basis matrix M_k with B(u) = [1,u,..,u^k] @ M_k (func_utils.py:33-50)."""
    if order == 0:
        return np.array([[1.0]], dtype=np.float32)
    prev = get_deboor_cox_mat(order - 1)
    zero_row = np.zeros((1, prev.shape[1]), dtype=np.float32)
    upper = np.concatenate([prev, zero_row], axis=0)
    lower = np.concatenate([zero_row, prev], axis=0)
    left = np.zeros((order, order + 1), dtype=np.float32)
    right = np.zeros((order, order + 1), dtype=np.float32)
    i = np.arange(order, dtype=np.int32)
    left[i, i] = i + 1
    left[i, i + 1] = order - i - 1
    right[i, i] = -1
    right[i, i + 1] = 1
    return (upper @ left + lower @ right) / order


def bspline_basis(u, order, device, dtype=torch.float32):
    key = (order, str(device), dtype)
    if key not in _M_CACHE:
        _M_CACHE[key] = torch.tensor(get_deboor_cox_mat(order), dtype=dtype, device=device)
    powers = torch.arange(0.0, order + 1.0, 1.0, dtype=dtype, device=device)
    return (u ** powers) @ _M_CACHE[key]


def fft_basis(v, order, device, dtype=torch.float32):
    freq = torch.linspace(1.0, order, order, dtype=dtype, device=device) * np.pi
    return torch.cat([torch.sin(v * freq), torch.cos(v * freq)], dim=-1)


def poly_basis(v, order, device, dtype=torch.float32):
    freq = torch.linspace(1.0, order, order, dtype=dtype, device=device)
    return v ** freq


def get_param_num(args):
    return args[0] + args[2] + 2 * args[3] + args[4]


# ---- roma 1.5.1 restatement (xyzw) ------------------------------------------------------------
def quat_conjugation(q):
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def quat_product(p, q):
    vec = p[..., 3:] * q[..., :3] + q[..., 3:] * p[..., :3] + torch.cross(p[..., :3], q[..., :3], dim=-1)
    last = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat([vec, last[..., None]], dim=-1)


def unitquat_to_rotvec(quat, shortest_arc=True):
    shape = quat.shape[:-1]
    q = quat.reshape(-1, 4)
    if shortest_arc:
        q = torch.where((q[:, 3:] < 0), -q, q)
    half_angle = torch.atan2(torch.norm(q[:, :3], dim=1), q[:, 3])
    angle = 2 * half_angle
    small = torch.abs(angle) <= 1e-3
    safe = torch.where(small, torch.ones_like(angle), angle)
    scale = torch.where(small, 2 + angle ** 2 / 12 + 7 * angle ** 4 / 2880, safe / torch.sin(safe / 2))
    return (scale[:, None] * q[:, :3]).reshape(*shape, 3)


def rotvec_to_unitquat(rotvec):
    shape = rotvec.shape[:-1]
    r = rotvec.reshape(-1, 3)
    norms = torch.norm(r, dim=-1)
    small = norms <= 1e-3
    safe = torch.where(small, torch.ones_like(norms), norms)
    scale = torch.where(small, 0.5 - norms ** 2 / 48 + norms ** 4 / 3840, torch.sin(safe / 2) / safe)
    return torch.cat([scale[:, None] * r, torch.cos(norms / 2)[:, None]], dim=-1).reshape(*shape, 4)


# ---- func_utils.get_func_result ----------------------------------------------------------------
def get_func_result(v: float, param: torch.Tensor, order_args):
    device, dtype = param.device, param.dtype
    result = 0.0
    offset = 0
    n_b, k_b, n_poly, n_fft, n_q, k_q = order_args
    if n_b != 0:
        interval = n_b - k_b
        start = min(int(v * interval), interval - 1)
        ctrl = param[..., start + offset: start + k_b + offset + 1]
        u = v * interval - start
        result = result + torch.sum(ctrl * bspline_basis(u, k_b, device, dtype), dim=-1)
        offset += n_b
    if n_poly != 0:
        result = result + torch.sum(param[..., offset: offset + n_poly] * poly_basis(v, n_poly, device, dtype), dim=-1)
        offset += n_poly
    if n_fft != 0:
        result = result + torch.sum(param[..., offset: offset + 2 * n_fft] * fft_basis(v, n_fft, device, dtype), dim=-1)
        offset += 2 * n_fft
    if n_q != 0:
        interval = n_q - k_q
        start = min(int(v * interval), interval - 1)
        ident = torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=dtype, device=device).reshape(-1, 1)
        ctrl = param[..., start + offset: start + k_q + offset + 1] + ident          # N,4,k+1 (wxyz)
        ctrl = normalize(torch.permute(ctrl, (0, 2, 1)), dim=-1)[..., [1, 2, 3, 0]]    # N,k+1,4 (xyzw)
        u = v * interval - start
        func = bspline_basis(u, k_q, device, dtype)
        func_cum = torch.flip(torch.cumsum(torch.flip(func, dims=(-1,)), dim=-1), dims=(-1,))[..., 1:]
        rel = quat_product(quat_conjugation(ctrl[:, :-1, :]), ctrl[:, 1:, :])
        vec = unitquat_to_rotvec(rel)
        quat = rotvec_to_unitquat(vec * func_cum[None, :, None])
        out = ctrl[:, 0]
        for i in range(quat.shape[1]):
            out = quat_product(out, quat[:, i])
        result = result + out[..., [3, 0, 1, 2]]
        offset += n_q
    return result


# ---- scene/gaussian_model.py deform accessors ---------------------------------------------------
class ReferenceModel:
    """The parameter tensors of scene/gaussian_model.py:GaussianModel in REFERENCE layout and the
    accessors on the hot path (get_deformed_pkg & co.), device-agnostic."""

    FIELDS = ("scene_xyz", "obj_xyz", "scene_shs_dc", "obj_shs_dc", "scene_shs_rest", "obj_shs_rest",
              "scene_scaling", "obj_scaling", "scene_rotation", "obj_rotation", "scene_opacity", "obj_opacity",
              "xyz_deform_param", "rotation_deform_param", "shs_deform_param_scene", "shs_deform_param_obj",
              "background_deform_param", "gs_time", "gs_time_sigma")

    def __init__(self, order_args, use_time_mask=True, **tensors):
        self.order_args = order_args
        self.use_time_mask = use_time_mask
        for f in self.FIELDS:
            setattr(self, f, tensors[f])

    def trainable(self):
        return [f for f in self.FIELDS if f != "gs_time"]

    def get_deformed_xyz(self, t):
        obj_xyz = self.obj_xyz + get_func_result(t, self.xyz_deform_param, self.order_args['xyz'])
        xyz = torch.cat([self.scene_xyz, obj_xyz], dim=0)
        return xyz + get_func_result(t, self.background_deform_param, self.order_args['background'])

    def get_deformed_rotation(self, t):
        obj_rotation = get_func_result(t, self.rotation_deform_param, self.order_args['rotation'])
        if self.order_args['rotation'][4] == 0:
            obj_rotation = self.obj_rotation + obj_rotation
        return normalize(torch.cat([self.scene_rotation, obj_rotation], dim=0))

    def get_deformed_shs(self, t):
        deform = torch.cat([self.shs_deform_param_scene, self.shs_deform_param_obj], dim=0)
        dc = torch.cat([self.scene_shs_dc, self.obj_shs_dc], dim=0)
        dc = dc[:, 0] + get_func_result(t, deform, self.order_args['shs'])
        rest = torch.cat([self.scene_shs_rest, self.obj_shs_rest], dim=0)
        return torch.cat([dc[:, None], rest], dim=1)

    def get_time_masked_opacity(self, t):
        delta = t - self.gs_time
        sigma = torch.exp(self.gs_time_sigma)
        sigma = torch.where(delta < 0.0, sigma[:, :1], sigma[:, 1:])
        mask = torch.exp(-0.5 * (delta / sigma) ** 2)
        return torch.cat([torch.sigmoid(self.scene_opacity), torch.sigmoid(self.obj_opacity) * mask], dim=0)

    def get_opacity(self):
        return torch.sigmoid(torch.cat([self.scene_opacity, self.obj_opacity], dim=0))

    def get_scaling(self):
        return torch.exp(torch.cat([self.scene_scaling, self.obj_scaling], dim=0))

    def get_obj_mask(self):
        return torch.cat([torch.zeros(self.scene_xyz.shape[0], dtype=torch.bool, device=self.scene_xyz.device),
                          torch.ones(self.obj_xyz.shape[0], dtype=torch.bool, device=self.obj_xyz.device)])

    def get_deformed_pkg(self, t):
        return {'xyz': self.get_deformed_xyz(t), 'rotation': self.get_deformed_rotation(t),
                'shs': self.get_deformed_shs(t),
                'opacity': self.get_time_masked_opacity(t) if self.use_time_mask else self.get_opacity()}


# ---- utils/sh_utils.py:eval_sh -------------------------------------------------------------------
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, dirs):
    """sh: (..., C, (deg+1)^2); dirs unit (..., 3) -> (..., C) (sh_utils.py:57-112, deg <= 3)."""
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] +
                      C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + C2[3] * xz * sh[..., 7] +
                      C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
                          C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] +
                          C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                          C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] +
                          C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def sh_colors(model_shs, xyz, campos, deg):
    """RGB the rasterizer would compute from shs (N,16,3): clamp_min(eval_sh + 0.5, 0) (forward.cu:20-71)."""
    dirs = normalize(xyz - campos[None, :], dim=-1)
    return torch.clamp_min(eval_sh(deg, model_shs.transpose(1, 2), dirs) + 0.5, 0.0)


def random_reference_model(n_scene, n_obj, order_args, seed=0, device="cpu", dtype=torch.float32, deform_scale=1e-2,
                           cloud=None, requires_grad=False):
    """Seeded parameters with the shapes of create_from_pcd (gaussian_model.py:255-335)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = n_scene + n_obj

    def U(*shape, scale=1.0):
        return ((torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * scale).to(dtype).to(device)

    def N(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype).to(device)

    if cloud is None:
        xyz, scaling, rot, op = N(n, 3, scale=3.0), N(n, 3, scale=0.3) - 2.0, N(n, 4), N(n, 1)
        shs = torch.cat([N(n, 1, 3), N(n, 15, 3, scale=0.1)], dim=1)
    else:
        xyz, scaling, rot, op, shs = (torch.tensor(cloud[k]).to(dtype).to(device) for k in
                                      ("xyz", "scaling_raw", "rotation_raw", "opacity_raw", "shs"))
    Cx, Cr = get_param_num(order_args['xyz']), get_param_num(order_args['rotation'])
    Cs, Cb = get_param_num(order_args['shs']), get_param_num(order_args['background'])
    t = dict(
        scene_xyz=xyz[:n_scene], obj_xyz=xyz[n_scene:],
        scene_shs_dc=shs[:n_scene, 0:1].contiguous(), obj_shs_dc=shs[n_scene:, 0:1].contiguous(),
        scene_shs_rest=shs[:n_scene, 1:].contiguous(), obj_shs_rest=shs[n_scene:, 1:].contiguous(),
        scene_scaling=scaling[:n_scene], obj_scaling=scaling[n_scene:],
        scene_rotation=rot[:n_scene], obj_rotation=rot[n_scene:],
        scene_opacity=op[:n_scene], obj_opacity=op[n_scene:],
        xyz_deform_param=U(n_obj, 3, Cx, scale=deform_scale),
        rotation_deform_param=U(n_obj, 4, Cr, scale=deform_scale),
        shs_deform_param_scene=U(n_scene, 3, Cs, scale=deform_scale),
        shs_deform_param_obj=U(n_obj, 3, Cs, scale=deform_scale),
        background_deform_param=U(1, 3, Cb, scale=deform_scale),
        gs_time=(torch.rand(n_obj, 1, generator=g, dtype=torch.float64)).to(dtype).to(device),
        gs_time_sigma=U(n_obj, 2, scale=0.75) - 0.75,   # sigma in (0.22, 1): the time mask stays well above 0
    )
    t = {k: v.clone().contiguous() for k, v in t.items()}
    if requires_grad:
        for k, v in t.items():
            if k != "gs_time":
                v.requires_grad_(True)
    return ReferenceModel(order_args, True, **t)
