"""Generates tests/golden/trajectory.npz by importing the REFERENCE's own utils/func_utils.py
from /root/reference (so it only runs in the build container) and calling its get_func_result /
set_default_param_order / get_deboor_cox_mat on seeded inputs, on the CPU.

Two non-invasive shims make the unmodified file importable and runnable here:
  * `roma` (un-vendored, not installed) is replaced by a module exposing the restated maps of
    oracle/trajectory_oracle.py -- so the quaternion branch pins everything EXCEPT roma itself
    (parity at the roma boundary stays "unpinned");
  * the hard-coded device='cuda' (func_utils.py:54,61,72,75,159) is redirected to 'cpu' by
    wrapping torch.linspace / torch.arange / torch.tensor during the calls.
Usage: python tests/golden/make_trajectory_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import trajectory_oracle as TO  # noqa: E402

roma = types.ModuleType("roma")
roma.unitquat_slerp = None
roma.unitquat_to_rotvec = TO.unitquat_to_rotvec
roma.rotvec_to_unitquat = TO.rotvec_to_unitquat
roma.quat_conjugation = TO.quat_conjugation
roma.quat_product = TO.quat_product
sys.modules["roma"] = roma
sys.path.insert(0, REF)
import utils.func_utils as FU  # noqa: E402  (the reference's file, unmodified)


def _cpu(fn):
    def wrapped(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


ORDER_SETS = {
    "kitti75": dict(xyz=[None, 5, 0, 6, 0, 0], rotation=[0, 0, 0, 0, None, 5], shs=[0, 0, 0, 6, 0, 0],
                    background=[None, 5, 0, 6, 0, 0]),
    "kitti50": dict(xyz=[None, 2, 0, 6, 0, 0], rotation=[0, 0, 0, 0, None, 2], shs=[0, 0, 0, 6, 0, 0],
                    background=[None, 2, 0, 6, 0, 0]),
    "kitti25": dict(xyz=[None, 1, 0, 6, 0, 0], rotation=[0, 0, 0, 0, None, 1], shs=[0, 0, 0, 6, 0, 0],
                    background=[None, 1, 0, 6, 0, 0]),
    "generic": dict(xyz=[9, 3, 2, 4, 0, 0], rotation=[6, 2, 1, 2, 10, 3], shs=[4, 1, 1, 2, 0, 0],
                    background=[0, 0, 0, 0, 0, 0]),
}
TIMES = [0.0, 0.013, 0.25, 0.37, 0.5, 0.77, 0.999, 1.0]


def main():
    saved = (torch.linspace, torch.arange, torch.tensor)
    torch.linspace, torch.arange, torch.tensor = _cpu(torch.linspace), _cpu(torch.arange), _cpu(torch.tensor)
    try:
        d = {}
        for k in range(6):
            d[f"deboor_{k}"] = FU.get_deboor_cox_mat(k)
        g = torch.Generator().manual_seed(1234)
        for name, oa in ORDER_SETS.items():
            filled = FU.set_default_param_order(oa, 52, 3)
            d[f"{name}__order"] = np.array([filled[a] for a in ("xyz", "rotation", "shs", "background")])
            for attr, D, scale in (("xyz", 3, 0.5), ("rotation", 4, 0.3), ("shs", 3, 0.5), ("background", 3, 0.5)):
                args = filled[attr]
                C = FU.get_param_num(args)
                if C == 0:
                    continue
                param = (torch.rand(5, D, C, generator=g) * 2 - 1) * scale
                d[f"{name}__{attr}__param"] = param.numpy()
                for ti, t in enumerate(TIMES):
                    d[f"{name}__{attr}__t{ti}"] = FU.get_func_result(t, param, args).numpy()
        d["times"] = np.array(TIMES)
        # tiny-angle regime of the quaternion branch (initialisation scale 1e-5, gaussian_model.py:311-312)
        args = FU.set_default_param_order(ORDER_SETS["kitti75"], 52, 3)["rotation"]
        param = (torch.rand(5, 4, FU.get_param_num(args), generator=g) * 2 - 1) * 1e-5
        d["tiny__rotation__param"] = param.numpy()
        for ti, t in enumerate(TIMES):
            d[f"tiny__rotation__t{ti}"] = FU.get_func_result(t, param, args).numpy()
    finally:
        torch.linspace, torch.arange, torch.tensor = saved
    out = os.path.join(HERE, "trajectory.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes,", len(d), "arrays")


if __name__ == "__main__":
    main()
