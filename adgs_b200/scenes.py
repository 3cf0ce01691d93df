"""Synthetic cameras and Gaussian clouds of the shapes BASELINE.json names.

Camera matrices are built the way the reference builds them (scene/cameras.py:77-80 with
utils/graphics_utils.py:getWorld2View2 :46-58 and getProjectionMatrix :60-80): the tensors handed
to the rasterizer are the TRANSPOSES of the math matrices, znear .01, zfar 100.
"""
import math
from typing import NamedTuple

import numpy as np
import torch


class Camera(NamedTuple):
    """The subset of scene/cameras.py:Camera that gaussian_renderer.render() reads."""
    image_height: int
    image_width: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # (4,4) transposed W2C
    full_proj_transform: torch.Tensor   # (4,4) transposed P @ W2C
    camera_center: torch.Tensor         # (3,)
    time: float


def make_camera(width, height, fovx_deg=90.0, yaw_deg=0.0, position=(0.0, 0.0, 0.0), time=0.37, znear=0.01,
                zfar=100.0, device="cpu") -> Camera:
    fovx = math.radians(fovx_deg)
    focal = width / (2.0 * math.tan(fovx / 2.0))
    fovy = 2.0 * math.atan(height / (2.0 * focal))
    yaw = math.radians(yaw_deg)
    # camera-to-world rotation (columns = camera axes in world): yaw about +y, camera looks along +z
    R = np.array([[math.cos(yaw), 0.0, math.sin(yaw)], [0.0, 1.0, 0.0], [-math.sin(yaw), 0.0, math.cos(yaw)]])
    C = np.asarray(position, dtype=np.float64)
    T = -R.T @ C                      # W2C translation
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T                  # getWorld2View2 receives R (C2W) and stores its transpose
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    w2c = torch.tensor(np.float32(Rt))
    world_view = w2c.transpose(0, 1).contiguous()
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1)
    full = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = world_view.inverse()[3, :3].contiguous()
    return Camera(int(height), int(width), fovx, fovy, world_view.to(device), full.to(device), center.to(device),
                  float(time))


def random_cloud(n, camera: Camera, seed=0, median_radius_px=3.0, behind_fraction=0.05, depth_range=(2.0, 80.0),
                 cluster=None, sh_coeffs=16, dtype=np.float32):
    """Gaussians uniform in the view frustum (log-uniform depth), log-normal scales chosen so the
    median projected radius is ~median_radius_px. Returns a dict of numpy arrays in reference layout:
    xyz (n,3), scaling_raw = log(s) (n,3), rotation_raw (n,4), opacity_raw = logit(o) (n,1),
    shs (n,16,3). `cluster` = (fraction, tile_fraction): that fraction of Gaussians is concentrated in
    a small screen region (stress config)."""
    rng = np.random.default_rng(seed)
    W, H = camera.image_width, camera.image_height
    tanx, tany = math.tan(camera.FoVx / 2), math.tan(camera.FoVy / 2)
    z = np.exp(rng.uniform(math.log(depth_range[0]), math.log(depth_range[1]), n))
    nb = int(n * behind_fraction)
    if nb:
        z[:nb] = -rng.uniform(0.5, 20.0, nb)
    u = rng.uniform(-1.05, 1.05, n)
    v = rng.uniform(-1.05, 1.05, n)
    if cluster is not None:
        frac, region = cluster
        k = int(n * frac)
        side = math.sqrt(region)
        u[nb:nb + k] = rng.uniform(-side, side, k) + 0.2
        v[nb:nb + k] = rng.uniform(-side, side, k) - 0.1
    xc = u * tanx * np.abs(z)
    yc = v * tany * np.abs(z)
    pts_cam = np.stack([xc, yc, z], 1)
    # camera -> world with the camera's own matrices (world_view_transform is W2C^T)
    w2c = camera.world_view_transform.cpu().numpy().astype(np.float64).T
    c2w = np.linalg.inv(w2c)
    xyz = pts_cam @ c2w[:3, :3].T + c2w[:3, 3]
    focal = W / (2.0 * tanx)
    # radius_px ~ 3 * sigma_px, sigma_px = s * focal / z
    s_med = (median_radius_px / 3.0) * np.abs(z) / focal
    s = s_med[:, None] * np.exp(rng.normal(0.0, 0.5, (n, 3)))
    rot = rng.normal(0.0, 1.0, (n, 4))
    op = rng.uniform(0.05, 0.95, (n, 1))
    shs = np.concatenate([rng.normal(0.0, 1.0, (n, 1, 3)), rng.normal(0.0, 0.1, (n, sh_coeffs - 1, 3))], 1)
    return dict(xyz=xyz.astype(dtype), scaling_raw=np.log(s).astype(dtype), rotation_raw=rot.astype(dtype),
                opacity_raw=np.log(op / (1 - op)).astype(dtype), shs=shs.astype(dtype))


def activated_inputs(cloud, device="cuda", requires_grad=False):
    """What the reference hands to the rasterizer when nothing is time-dependent: means3D,
    opacity = sigmoid, scales = exp, rotations = normalised quaternion, shs (n,16,3)."""
    t = {k: torch.tensor(v, device=device) for k, v in cloud.items()}
    out = dict(means3D=t["xyz"], opacities=torch.sigmoid(t["opacity_raw"]), scales=torch.exp(t["scaling_raw"]),
               rotations=torch.nn.functional.normalize(t["rotation_raw"]), shs=t["shs"])
    if requires_grad:
        for v in out.values():
            v.requires_grad_(True)
    return out


BENCH_ORDER_ARGS = {  # kitti-75 / waymo style with 32 control points: 96 frames // 3
    'xyz': [32, 5, 0, 6, 0, 0],
    'rotation': [0, 0, 0, 0, 32, 5],
    'shs': [0, 0, 0, 6, 0, 0],
    'background': [32, 5, 0, 6, 0, 0],
}


def random_model_tensors(n_scene, n_obj, order_args, cloud, seed=0, device="cuda", deform_scale=1e-2,
                         time_sigma=1.0 / 96.0):
    """Seeded parameters in the REFERENCE's tensor layout and naming (the shapes of
    scene/gaussian_model.py:create_from_pcd, :285-328): Gaussians [0, n_scene) of `cloud` are scene,
    the rest objects. Deformation parameters ~ U(-1,1)*deform_scale, gs_time ~ U(0,1),
    gs_time_sigma = log(time_sigma)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = n_scene + n_obj

    def num(a):
        return a[0] + a[2] + 2 * a[3] + a[4]

    def U(*shape):
        return ((torch.rand(*shape, generator=g) * 2 - 1) * deform_scale).to(device)

    t = {k: torch.tensor(cloud[k], dtype=torch.float32, device=device) for k in
         ("xyz", "scaling_raw", "rotation_raw", "opacity_raw", "shs")}
    out = dict(
        scene_xyz=t["xyz"][:n_scene], obj_xyz=t["xyz"][n_scene:],
        scene_shs_dc=t["shs"][:n_scene, 0:1], obj_shs_dc=t["shs"][n_scene:, 0:1],
        scene_shs_rest=t["shs"][:n_scene, 1:], obj_shs_rest=t["shs"][n_scene:, 1:],
        scene_scaling=t["scaling_raw"][:n_scene], obj_scaling=t["scaling_raw"][n_scene:],
        scene_rotation=t["rotation_raw"][:n_scene], obj_rotation=t["rotation_raw"][n_scene:],
        scene_opacity=t["opacity_raw"][:n_scene], obj_opacity=t["opacity_raw"][n_scene:],
        xyz_deform_param=U(n_obj, 3, num(order_args['xyz'])),
        rotation_deform_param=U(n_obj, 4, num(order_args['rotation'])),
        shs_deform_param_scene=U(n_scene, 3, num(order_args['shs'])),
        shs_deform_param_obj=U(n_obj, 3, num(order_args['shs'])),
        background_deform_param=U(1, 3, num(order_args['background'])),
        gs_time=torch.rand(n_obj, 1, generator=g).to(device),
        gs_time_sigma=torch.full((n_obj, 2), math.log(time_sigma), dtype=torch.float32, device=device),
    )
    return {k: v.contiguous() for k, v in out.items()}
