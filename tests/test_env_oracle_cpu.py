"""CPU: the environment-map oracle (oracle/env_oracle.py) against the golden vectors produced by the reference's
own scene/env.py (tests/golden/env.npz, tests/golden/make_env_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import env_oracle as EO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "env.npz"))
T = lambda k: torch.tensor(G[k])


@pytest.mark.parametrize("c", ["a", "b", "c"])
def test_env_oracle_matches_reference(c):
    grid = T(f"{c}_grid_map").requires_grad_(True)
    fg, op = T(f"{c}_fg").requires_grad_(True), T(f"{c}_op").requires_grad_(True)
    H, W = fg.shape[1:]
    bg = EO.get_image_background(grid, float(G[f"{c}_fovx"]), H, W, T(f"{c}_wvt"))
    rendered = EO.composite(fg, op, bg)
    (rendered * T(f"{c}_cot")).sum().backward()
    close = lambda a, b: np.abs(a.detach().numpy() - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-12)
    assert close(bg, G[f"{c}_background"]) and close(rendered, G[f"{c}_rendered"])
    assert close(grid.grad, G[f"{c}_d_grid"]) and close(fg.grad, G[f"{c}_d_fg"]) and close(op.grad, G[f"{c}_d_op"])
