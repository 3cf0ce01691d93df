"""CPU: the reference's checkpoint formats (adgs_b200/checkpoint.py): point_cloud.ply (property list and order of
scene/gaussian_model.py:413-446, binary_little_endian float32 vertex element as plyfile writes it) and deform.pth
(the 10-tuple of :448-459). Round trip through the planar model, header check, and a file laid out by hand the way
the reference's save_ply lays it out (SH coefficients channel-major: transpose(1, 2).flatten)."""
import os

import numpy as np
import torch

from adgs_b200 import checkpoint as CK
from adgs_b200.gaussian_model import GaussianModel, PARAM_NAMES

ORDER_ARGS = {"xyz": [6, 3, 0, 2, 0, 0], "rotation": [0, 0, 0, 0, 5, 2], "shs": [0, 0, 0, 2, 0, 0],
              "background": [6, 3, 0, 2, 0, 0]}


def ref_tensors(ns, no, seed=0):
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.randn(*s, generator=g)
    return dict(
        scene_xyz=R(ns, 3), obj_xyz=R(no, 3), scene_shs_dc=R(ns, 1, 3), obj_shs_dc=R(no, 1, 3),
        scene_shs_rest=R(ns, 15, 3), obj_shs_rest=R(no, 15, 3), scene_opacity=R(ns, 1), obj_opacity=R(no, 1),
        scene_scaling=R(ns, 3), obj_scaling=R(no, 3), scene_rotation=R(ns, 4), obj_rotation=R(no, 4),
        xyz_deform_param=R(no, 3, 10), rotation_deform_param=R(no, 4, 5), shs_deform_param_scene=R(ns, 3, 4),
        shs_deform_param_obj=R(no, 3, 4), background_deform_param=R(1, 3, 10), gs_time=torch.rand(no, 1, generator=g),
        gs_time_sigma=R(no, 2))


def test_attribute_list_matches_reference():
    names = CK.construct_list_of_attributes(3)
    assert names[:9] == ["x", "y", "z", "nx", "ny", "nz", "shs_dc_0", "shs_dc_1", "shs_dc_2"]
    assert names[9] == "shs_rest_0" and names[53] == "shs_rest_44" and names[54] == "opacity"
    assert names[55:] == ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "obj"]
    assert len(names) == 63


def test_save_load_round_trip(tmp_path):
    ref = ref_tensors(37, 21)
    m = GaussianModel.from_reference(ref, ORDER_ARGS, device="cpu")
    m.scene_extent = 12.5
    path = os.path.join(tmp_path, "point_cloud", "iteration_7", "point_cloud.ply")
    m.save_ply(path)
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n") + 11].decode()
    lines = head.split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex 58"
    assert lines[3] == "property float x" and lines[3 + 62] == "property float obj"
    assert len(raw) == len(head) + 58 * 63 * 4
    assert os.path.exists(os.path.join(os.path.dirname(path), "deform.pth"))
    m2 = GaussianModel(3, None)
    m2.load_ply(path, device="cpu")
    assert (m2.n_scene, m2.n_obj) == (37, 21) and m2.order_args == ORDER_ARGS and m2.scene_extent == 12.5
    assert m2.active_sh_degree == 3 and m2.use_time_mask is True
    for k in PARAM_NAMES:
        assert torch.equal(getattr(m, k).detach(), getattr(m2, k).detach()), k
    assert torch.equal(m.gs_time, m2.gs_time)
    back = m2.to_reference()
    for k, v in ref.items():
        assert torch.equal(back[k], v), k


def test_loads_a_file_in_the_reference_layout(tmp_path):
    """A PLY assembled column by column as save_ply does (gaussian_model.py:431-445): shs_dc / shs_rest are
    (P, K, 3) tensors written as transpose(1, 2).flatten(1), i.e. channel-major; rows are NOT sorted scene-first
    (load_ply splits them by the `obj` column)."""
    ref = ref_tensors(5, 4, seed=3)
    cat = lambda a, b: torch.cat([ref[a], ref[b]], 0)
    perm = torch.tensor([0, 5, 1, 6, 2, 7, 3, 8, 4])            # interleave scene and object rows
    obj = torch.cat([torch.zeros(5, 1), torch.ones(4, 1)], 0)
    table = torch.cat([cat("scene_xyz", "obj_xyz"), torch.zeros(9, 3),
                       cat("scene_shs_dc", "obj_shs_dc").transpose(1, 2).flatten(1),
                       cat("scene_shs_rest", "obj_shs_rest").transpose(1, 2).flatten(1),
                       cat("scene_opacity", "obj_opacity"), cat("scene_scaling", "obj_scaling"),
                       cat("scene_rotation", "obj_rotation"), obj], 1)[perm]
    path = os.path.join(tmp_path, "point_cloud.ply")
    CK.write_ply(path, CK.construct_list_of_attributes(3), table.numpy())
    torch.save((torch.nn.Parameter(ref["xyz_deform_param"]), torch.nn.Parameter(ref["rotation_deform_param"]),
                torch.nn.Parameter(ref["shs_deform_param_scene"]), torch.nn.Parameter(ref["shs_deform_param_obj"]),
                torch.nn.Parameter(ref["background_deform_param"]), ref["gs_time"],
                torch.nn.Parameter(ref["gs_time_sigma"]), True, ORDER_ARGS, 3.0), os.path.join(tmp_path, "deform.pth"))
    m = GaussianModel(3, None)
    m.load_ply(path, device="cpu")
    back = m.to_reference()
    for k, v in ref.items():
        assert torch.equal(back[k], v), k


def test_read_ply_ascii_and_errors(tmp_path):
    p = os.path.join(tmp_path, "a.ply")
    with open(p, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment hi\nelement vertex 2\nproperty float x\nproperty double y\n"
                "end_header\n1.5 2\n3 4.25\n")
    el = CK.read_ply(p)
    assert np.allclose(el["x"], [1.5, 3.0]) and np.allclose(el["y"], [2.0, 4.25])
    with open(p, "wb") as f:
        f.write(b"plx\n")
    try:
        CK.read_ply(p)
        assert False
    except ValueError:
        pass
