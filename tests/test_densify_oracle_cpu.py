"""CPU: oracle/densify_oracle.py against tests/golden/densify.npz, i.e. against the outputs of the reference's own
scene/gaussian_model.py (add_densification_stats, densify_and_prune, reset_opacity) run on CPU."""
import os

import numpy as np
import pytest

from oracle import densify_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "densify.npz"))
STATE_KEYS = ([n for n in O.SCENE_ROWS + O.OBJ_ROWS] + ["background_deform_param", "gs_time", "xyz_gradient_accum",
                                                        "denom", "max_radii2D"])


def load_state(tag, stage):
    pre = f"{tag}.{stage}."
    return {k[len(pre):]: GOLD[k].copy() for k in GOLD.files if k.startswith(pre)}


def assert_state(got, want, exact_rows=True, tol=0.0):
    for k, w in want.items():
        g = got[k]
        assert g.shape == w.shape, f"{k}: shape {g.shape} != {w.shape}"
        if tol == 0.0:
            assert np.array_equal(g, w), f"{k} differs (max abs {np.abs(g - w).max() if g.size else 0})"
        else:
            np.testing.assert_allclose(g, w, rtol=tol, atol=tol * 1e-2, err_msg=k)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stats_match_reference(tag):
    st = load_state(tag, "before")
    n = st["scene_xyz"].shape[0] + st["obj_xyz"].shape[0]
    st["xyz_gradient_accum"] = np.zeros((n, 1), np.float32)
    st["denom"] = np.zeros((n, 1), np.float32)
    st["max_radii2D"] = np.zeros((n,), np.float32)
    for it in range(3):
        O.add_densification_stats(st, GOLD[f"{tag}.stats{it}.grad"], GOLD[f"{tag}.stats{it}.radii"])
    want = load_state(tag, "before")
    np.testing.assert_allclose(st["xyz_gradient_accum"], want["xyz_gradient_accum"], rtol=1e-6, atol=1e-12)
    assert np.array_equal(st["denom"], want["denom"])
    assert np.array_equal(st["max_radii2D"], want["max_radii2D"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_densify_and_prune_matches_reference(tag):
    st = load_state(tag, "before")
    scene_extent, object_extent, percent_dense = GOLD[f"{tag}.extents"]
    O.densify_and_prune(st, 0.0002, 0.0002, 0.005, bool(GOLD[f"{tag}.prune_big"]), scene_extent, object_extent,
                        percent_dense, GOLD[f"{tag}.z_scene"], GOLD[f"{tag}.z_obj"], N=2, gpu_division=False)
    want = load_state(tag, "after")
    assert set(want) <= set(st)
    # which rows survive and in which order: exact (every copied array is bit-identical)
    copied = [k for k in want if not (k.startswith("scene_xyz") or k.startswith("obj_xyz") or
                                      k.startswith("scene_scaling") or k.startswith("obj_scaling"))]
    assert_state(st, {k: want[k] for k in copied})
    # split children: positions through a 3x3 product, scales through log(exp(s) / 1.6)
    for k in ("scene_xyz", "obj_xyz", "scene_scaling", "obj_scaling"):
        np.testing.assert_allclose(st[k], want[k], rtol=2e-6, atol=2e-6, err_msg=k)
        for m in O.MOMENTS:
            assert np.array_equal(st[k + m], want[k + m])
    # something happened in every class
    ns0, ns1 = GOLD[f"{tag}.before.scene_xyz"].shape[0], want["scene_xyz"].shape[0]
    assert ns1 != ns0


@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_division_differs_by_an_ulp_at_most(tag):
    a, b = load_state(tag, "before"), load_state(tag, "before")
    ext = GOLD[f"{tag}.extents"]
    args = (0.0002, 0.0002, 0.005, bool(GOLD[f"{tag}.prune_big"]), ext[0], ext[1], ext[2], GOLD[f"{tag}.z_scene"],
            GOLD[f"{tag}.z_obj"])
    O.densify_and_prune(a, *args, gpu_division=False)
    O.densify_and_prune(b, *args, gpu_division=True)
    for k in a:
        assert a[k].shape == b[k].shape
    np.testing.assert_allclose(a["scene_scaling"], b["scene_scaling"], rtol=2.5e-7, atol=2.5e-7)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_reset_opacity_matches_reference(tag):
    st = load_state(tag, "after")
    O.reset_opacity(st)
    want = load_state(tag, "reset")
    for part in ("scene", "obj"):
        k = f"{part}_opacity"
        np.testing.assert_allclose(st[k], want[k], rtol=1e-6, atol=1e-6)
        assert np.array_equal(st[k + ".exp_avg"], want[k + ".exp_avg"])
        assert np.array_equal(st[k + ".exp_avg_sq"], want[k + ".exp_avg_sq"])
        assert (O.sigmoid(st[k]) <= 0.01 + 1e-6).all()


def test_prune_points_keeps_order():
    st = load_state("a", "before")
    ns, no = st["scene_xyz"].shape[0], st["obj_xyz"].shape[0]
    rng = np.random.default_rng(0)
    sm, om = rng.random(ns) < 0.3, rng.random(no) < 0.5
    ref = {k: v.copy() for k, v in st.items()}
    O.prune_points(st, sm, om)
    assert np.array_equal(st["scene_xyz"], ref["scene_xyz"][~sm])
    assert np.array_equal(st["xyz_deform_param.exp_avg"], ref["xyz_deform_param.exp_avg"][~om])
    assert np.array_equal(st["max_radii2D"], ref["max_radii2D"][np.concatenate([~sm, ~om])])


def test_knn_points_oracle_small():
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((200, 4)).astype(np.float32)
    anchors = pts[rng.permutation(200)[:25]]
    idx, d = O.knn_points(anchors, pts, 8)
    assert idx.shape == (25, 8) and (np.diff(d, axis=1) >= 0).all()
    assert np.allclose(d[:, 0], 0.0)                       # an anchor finds itself first
    full = ((anchors[:, None, :].astype(np.float64) - pts[None].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(np.sort(idx, axis=1), np.sort(np.argsort(full, axis=1)[:, :8], axis=1))
