"""Checkpoint formats of the reference, for interop with scenes trained by AD-GS:

    point_cloud.ply   GaussianModel.save_ply / load_ply               scene/gaussian_model.py:413-446, 469-518
    deform.pth        the 10-tuple saved next to it                    scene/gaussian_model.py:448-459, 520-543

`point_cloud.ply` is what `plyfile.PlyData([PlyElement.describe(elements, 'vertex')]).write(path)` produces for
a structured float32 array: a `binary_little_endian 1.0` PLY with one `vertex` element whose properties are
x y z nx ny nz shs_dc_0..2 shs_rest_0..44 opacity scale_0..2 rot_0..3 obj (all `float`), rows = [scene ; object]
Gaussians. plyfile is not installed in this image, so the container format is written / parsed here with numpy
(the PLY header grammar is public; property order and dtypes are the reference's). `deform.pth` is a plain
torch.save of the reference's tensors in the reference's layout, so either side can load the other's files.

Host-side I/O only (no kernels): the planar <-> reference layout conversion is GaussianModel.to_reference /
from_reference.
"""
import os

import numpy as np
import torch
from torch import nn

from .gaussian_model import PARAM_NAMES, GaussianModel, get_param_num


def construct_list_of_attributes(sh_degree=3):
    """scene/gaussian_model.py:413-426"""
    n_rest = 3 * ((sh_degree + 1) ** 2 - 1)
    return (["x", "y", "z", "nx", "ny", "nz"] + [f"shs_dc_{i}" for i in range(3)] +
            [f"shs_rest_{i}" for i in range(n_rest)] + ["opacity"] + [f"scale_{i}" for i in range(3)] +
            [f"rot_{i}" for i in range(4)] + ["obj"])


def write_ply(path, names, table):
    """binary_little_endian PLY with one float32 `vertex` element (what plyfile writes for such an array)."""
    table = np.ascontiguousarray(table, dtype="<f4")
    assert table.ndim == 2 and table.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(table.tobytes())


_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4",
              "uint32": "u4", "float32": "f4", "float64": "f8"}


def read_ply(path):
    """-> structured numpy array of the first element (`plydata.elements[0]`); binary little/big endian or ascii."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, first = None, None, [], True
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is None:
                    count = int(tok[2])
                else:
                    first = False          # later elements are not read
            elif tok[0] == "property" and first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=[(n, "<" + t) for n, t in props])
            for i, (n, _) in enumerate(props):
                out[n] = rows[:, i]
            return out
        end = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, end + t) for n, t in props])
        data = f.read(dt.itemsize * count)
        if len(data) != dt.itemsize * count:
            raise ValueError(f"{path}: truncated PLY body")
        return np.frombuffer(data, dtype=dt, count=count)


def save_ply(model: GaussianModel, path):
    """GaussianModel.save_ply: point_cloud.ply + deform.pth next to it."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    ref = {k: v.detach().cpu() for k, v in model.to_reference().items()}
    cat = lambda a, b: torch.cat([ref[a], ref[b]], dim=0)
    xyz = cat("scene_xyz", "obj_xyz").numpy()
    shs_dc = cat("scene_shs_dc", "obj_shs_dc").transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    shs_rest = cat("scene_shs_rest", "obj_shs_rest").transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    obj = np.zeros((xyz.shape[0], 1), np.float32)
    obj[model.n_scene:] = 1.0
    table = np.concatenate([xyz, np.zeros_like(xyz), shs_dc, shs_rest, cat("scene_opacity", "obj_opacity").numpy(),
                            cat("scene_scaling", "obj_scaling").numpy(), cat("scene_rotation", "obj_rotation").numpy(),
                            obj], axis=1)
    write_ply(path, construct_list_of_attributes(model.max_sh_degree), table)
    torch.save((
        nn.Parameter(ref["xyz_deform_param"].contiguous()), nn.Parameter(ref["rotation_deform_param"].contiguous()),
        nn.Parameter(ref["shs_deform_param_scene"].contiguous()), nn.Parameter(ref["shs_deform_param_obj"].contiguous()),
        nn.Parameter(ref["background_deform_param"].contiguous()), ref["gs_time"].contiguous(),
        nn.Parameter(ref["gs_time_sigma"].contiguous()), model.use_time_mask, model.order_args,
        getattr(model, "scene_extent", 0.0),
    ), os.path.join(os.path.dirname(os.path.abspath(path)), "deform.pth"))


def load_ply(model: GaussianModel, path, device="cuda"):
    """GaussianModel.load_ply: fills `model` (planar storage) from point_cloud.ply + deform.pth."""
    el = read_ply(path)
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    xyz = np.stack([f32(el["x"]), f32(el["y"]), f32(el["z"])], axis=1)
    opacities = f32(el["opacity"])[:, None]
    obj_mask = f32(el["obj"]) > 0.5
    scene_mask = ~obj_mask
    names = el.dtype.names
    by_index = lambda prefix: sorted([n for n in names if n.startswith(prefix)], key=lambda x: int(x.split("_")[-1]))
    shs_dc = np.stack([f32(el[f"shs_dc_{i}"]) for i in range(3)], axis=1)[:, :, None]              # (P,3,1)
    rest_names = by_index("shs_rest_")
    n_coef = (model.max_sh_degree + 1) ** 2
    assert len(rest_names) == 3 * n_coef - 3
    shs_rest = np.stack([f32(el[n]) for n in rest_names], axis=1).reshape(-1, 3, n_coef - 1)         # (P,3,15)
    scales = np.stack([f32(el[n]) for n in by_index("scale_")], axis=1)
    rots = np.stack([f32(el[n]) for n in by_index("rot_")], axis=1)
    (xyz_deform, rot_deform, shs_deform_scene, shs_deform_obj, bg_deform, gs_time, gs_time_sigma, use_time_mask,
     order_args, scene_extent) = torch.load(os.path.join(os.path.dirname(os.path.abspath(path)), "deform.pth"),
                                            map_location="cpu", weights_only=True)   # tensors, Parameters, bool, dict, list, float only
    n_obj = int(obj_mask.sum())
    assert xyz_deform.shape[0] == n_obj
    assert xyz_deform.shape[-1] == get_param_num(order_args["xyz"])
    assert shs_deform_obj.shape[-1] == get_param_num(order_args["shs"])
    assert shs_deform_scene.shape[-1] == get_param_num(order_args["shs"])
    assert rot_deform.shape[-1] == get_param_num(order_args["rotation"])
    assert bg_deform.shape[-1] == get_param_num(order_args["background"])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    ref = {}
    for part, mask in (("scene", scene_mask), ("obj", obj_mask)):
        ref[f"{part}_xyz"] = T(xyz[mask])
        ref[f"{part}_shs_dc"] = T(shs_dc[mask]).transpose(1, 2).contiguous()
        ref[f"{part}_shs_rest"] = T(shs_rest[mask]).transpose(1, 2).contiguous()
        ref[f"{part}_opacity"] = T(opacities[mask])
        ref[f"{part}_scaling"] = T(scales[mask])
        ref[f"{part}_rotation"] = T(rots[mask])
    ref.update(xyz_deform_param=xyz_deform.detach(), rotation_deform_param=rot_deform.detach(),
               shs_deform_param_scene=shs_deform_scene.detach(), shs_deform_param_obj=shs_deform_obj.detach(),
               background_deform_param=bg_deform.detach(), gs_time=gs_time.detach(), gs_time_sigma=gs_time_sigma.detach())
    loaded = GaussianModel.from_reference(ref, order_args, sh_degree=model.max_sh_degree, use_time_mask=use_time_mask,
                                          device=device)
    model.order_args, model.use_time_mask, model.scene_extent = order_args, use_time_mask, scene_extent
    model.n_scene, model.n_obj = loaded.n_scene, loaded.n_obj
    for k in PARAM_NAMES:
        setattr(model, k, nn.Parameter(getattr(loaded, k).detach()))
    if "gs_time" in model._buffers:
        model.gs_time = loaded.gs_time
    else:
        model.register_buffer("gs_time", loaded.gs_time)
    model.active_sh_degree = model.max_sh_degree
    model.__dict__.pop("_basis_cache", None)
