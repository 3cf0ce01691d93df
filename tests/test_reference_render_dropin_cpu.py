"""SURVEY Appendix C, 'API' row / INTEGRATION.md level 1: the reference's OWN `gaussian_renderer.render`
(/root/reference/gaussian_renderer/__init__.py:18-115) is executed UNMODIFIED with `adgs_b200/dropin` on sys.path,
so that its `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer` resolves to
the B200-native drop-in: every keyword it passes, the 6-tuple it unpacks and the autograd wiring it relies on
(screenspace_points.grad) must be accepted by our module surface.

This container has no GPU and the GPU box has no /root/reference, so the test runs here, on CPU tensors, with the one
thing that needs a device -- the native boundary `_C.rasterize_gaussians(_backward)` (RZ/ext.cpp:15-19) -- replaced by
a stand-in that checks the positional signature and shapes it receives. What the native side computes is the subject
of tests/test_parity_gpu.py; this test is about the Python surface above it. Skipped where the reference is absent."""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF_RENDER = "/root/reference/gaussian_renderer/__init__.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.exists(REF_RENDER), reason="reference checkout not present")


def _load_reference_render(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "adgs_b200", "dropin"))
    # `scene.*` is imported by the reference file for type annotations only; its real modules need uninstalled packages
    scene = types.ModuleType("scene")
    gm = types.ModuleType("scene.gaussian_model")
    env = types.ModuleType("scene.env")
    gm.GaussianModel = type("GaussianModel", (), {})
    env.EnvironmentMap = type("EnvironmentMap", (), {})
    for name, mod in (("scene", scene), ("scene.gaussian_model", gm), ("scene.env", env)):
        monkeypatch.setitem(sys.modules, name, mod)
    for name in ("diff_gaussian_rasterization",):
        monkeypatch.delitem(sys.modules, name, raising=False)
    spec = importlib.util.spec_from_file_location("reference_gaussian_renderer", REF_RENDER)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import diff_gaussian_rasterization as dgr
    import adgs_b200.rasterizer as ours
    assert dgr.GaussianRasterizer is ours.GaussianRasterizer, "the reference's import did not resolve to the drop-in"
    assert mod.GaussianRasterizer is ours.GaussianRasterizer
    return mod, ours


def _cpu_for_cuda(monkeypatch):
    """The reference hard-codes device="cuda" (gaussian_renderer/__init__.py:26,43-48); map it to CPU here."""
    if torch.cuda.is_available():
        return
    real_zeros_like, real_tensor = torch.zeros_like, torch.tensor

    def strip(kw):
        if str(kw.get("device", "")).startswith("cuda"):
            kw = dict(kw, device="cpu")
        return kw

    monkeypatch.setattr(torch, "zeros_like", lambda *a, **k: real_zeros_like(*a, **strip(k)))
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **strip(k)))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)


class _StubModel:
    """What render() reads of scene/gaussian_model.py:GaussianModel."""

    def __init__(self, n_scene, n_obj):
        g = torch.Generator().manual_seed(0)
        n = n_scene + n_obj
        self.n_scene, self.n = n_scene, n
        self.active_sh_degree = 3
        self._xyz = torch.randn(n, 3, generator=g, requires_grad=True)
        self._scale = torch.rand(n, 3, generator=g, requires_grad=True)
        self.times = []

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return self._scale

    @property
    def get_obj_mask(self):
        m = torch.zeros(self.n, dtype=torch.bool)
        m[self.n_scene:] = True
        return m

    def get_deformed_xyz(self, t):
        self.times.append(("flow", t))
        return self._xyz + 0.01 * t

    def get_deformed_pkg(self, t):
        self.times.append(("pkg", t))
        return {"xyz": self._xyz * 1.0, "rotation": torch.nn.functional.normalize(torch.ones(self.n, 4)),
                "shs": torch.zeros(self.n, 16, 3), "opacity": torch.full((self.n, 1), 0.5)}


def test_reference_render_runs_unmodified_on_the_dropin_surface(monkeypatch):
    ref, ours = _load_reference_render(monkeypatch)
    _cpu_for_cuda(monkeypatch)
    H, W, n_scene, n_obj = 24, 40, 30, 12
    calls = {}

    def fake_forward(*args):
        assert len(args) == 22, "positional signature of RasterizeGaussiansCUDA (RZ/rasterize_points.h:18-42)"
        (bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, view, proj, tanx, tany, h, w, sh, flow,
         semantic, degree, campos, prefiltered, inv_depth, debug) = args
        P = means3D.shape[0]
        assert (h, w) == (H, W) and bg.shape == (3,) and view.shape == (4, 4) and proj.shape == (4, 4)
        assert opacity.shape == (P, 1) and scales.shape == (P, 3) and rotations.shape == (P, 4)
        assert sh.shape == (P, 16, 3) and colors.numel() == 0 and cov3D.numel() == 0
        assert flow.shape == (P, 3) and semantic.shape == (P, 1) and degree == 3 and inv_depth is True
        calls["fwd"] = args
        z = lambda *s: torch.zeros(*s)
        e = torch.empty(0, dtype=torch.uint8)
        return 7, z(3, H, W), z(1, H, W), z(1, H, W), torch.ones(P, dtype=torch.int32), e, e, e, z(3, H, W), z(1, H, W)

    def fake_backward(*args, opacities=None, needs=None):
        assert len(args) == 29, "positional signature of RasterizeGaussiansBackwardCUDA (RZ/rasterize_points.h:44-75)"
        assert args[22] == 7, "num_rendered travels from the forward to the backward"
        calls["bwd"] = args
        P = args[1].shape[0]
        shapes = [(P, 3), (P, 3), (P, 1), (P, 3), (P, 6), (P, 16, 3), (P, 3), (P, 4), (P, 3), (P, 1)]
        return tuple(torch.zeros(s) if (needs is None or needs[i]) else None for i, s in enumerate(shapes))

    monkeypatch.setattr(ours._C, "rasterize_gaussians", staticmethod(fake_forward))
    monkeypatch.setattr(ours._C, "rasterize_gaussians_backward", staticmethod(fake_backward))

    cam = types.SimpleNamespace(FoVx=1.2, FoVy=0.8, image_height=H, image_width=W, time=0.37,
                                world_view_transform=torch.eye(4), full_proj_transform=torch.eye(4),
                                camera_center=torch.zeros(3))
    env_map = types.SimpleNamespace(get_image_background=lambda c: torch.full((3, H, W), 0.25))
    pipe = types.SimpleNamespace(inv_depth=True, debug=False)
    pc = _StubModel(n_scene, n_obj)
    res = ref.render(cam, pc, env_map, pipe, flow_pkg=[0.41, None, None, None, None, None], render_objmask=True)

    assert set(res) == {"render", "viewspace_points", "visibility_filter", "radii", "depth", "opacity", "img_opacity",
                        "foreground", "background", "img_flow", "img_semantic", "xyz", "rotation", "shs"}
    assert pc.times == [("flow", 0.41), ("pkg", 0.37)]
    assert res["render"].shape == (3, H, W) and torch.allclose(res["render"], torch.full((3, H, W), 0.25))
    assert res["depth"].shape == (H, W) and res["img_opacity"].shape == (H, W)
    assert res["visibility_filter"].all() and res["radii"].dtype == torch.int32
    assert res["img_flow"].shape == (3, H, W) and res["img_semantic"].shape == (1, H, W)
    # autograd wiring the training loop relies on (train.py:149-152: viewspace_point_tensor.grad)
    (res["render"].sum() + res["depth"].sum() + res["img_opacity"].sum()).backward()
    assert "bwd" in calls and res["viewspace_points"].grad is not None
    assert res["viewspace_points"].grad.shape == (n_scene + n_obj, 3)
    assert pc._xyz.grad is not None and pc._scale.grad is not None


def test_dropin_render_has_the_reference_signature():
    """`from gaussian_renderer import render` (train.py:22) -> same parameters, same defaults."""
    import inspect
    src = open(REF_RENDER).read()
    from adgs_b200.gaussian_renderer import render as ours_render
    params = list(inspect.signature(ours_render).parameters.values())
    assert [p.name for p in params] == ["viewpoint_camera", "pc", "env_map", "pipe", "scaling_modifier", "override_color",
                                        "flow_pkg", "render_objmask"]
    assert [p.default for p in params[4:]] == [1.0, None, None, False]
    for p in params:
        assert p.name in src
