"""adgs_b200 -- B200-native (sm_100a) implementation of the AD-GS per-iteration hot path.

Public surface (mirrors the reference's plugin API):
  adgs_b200.rasterizer   GaussianRasterizationSettings, GaussianRasterizer, _C   (diff_gaussian_rasterization)
  adgs_b200.simple_knn   distCUDA2                                               (simple_knn._C)
  adgs_b200.gaussian_model / gaussian_renderer   fused trajectory + render path  (scene.gaussian_model, gaussian_renderer)
All compute goes through libadgs_b200.so (include/adgs_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
