"""Helper of test_fused_gpu.py::test_split_per_gaussian_backward_matches_one_kernel_form (not a test): renders the
small bench scene through the fused path, runs the backward with fixed cotangents and saves every parameter gradient.
The per-Gaussian backward variant is a process-wide choice (ADGS_TUNE_PGB, read once), hence one process per variant.
usage: ADGS_TUNE_PGB=<v> python tests/pgb_variant_grads.py OUT.pt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as Hh  # noqa: E402,F401  (puts the repo root on sys.path)
import test_fused_gpu as T  # noqa: E402


def main(out_path):
    model, c, _ = T._bench_scene(T.FULL["small"], seed=4)
    res = T._fused(model, c, 0.37, 0.41)
    cot = Hh.cotangents(c, seed=9)
    loss = ((res["render"] * cot["color"]).sum() + (res["depth"] * cot["depth"][0]).sum() +
            (res["img_opacity"] * cot["opacity"][0]).sum() + (res["img_flow"] * cot["flow"]).sum())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: v.detach().cpu() for k, v in model.to_reference(grads=True).items() if v is not None}
    grads["viewspace_points"] = res["viewspace_points"].grad.detach().cpu()
    torch.save(grads, out_path)


if __name__ == "__main__":
    main(sys.argv[1])
