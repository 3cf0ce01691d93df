"""Put `adgs_b200/dropin` on sys.path and the reference's `import diff_gaussian_rasterization`
(gaussian_renderer/__init__.py:14) resolves to the B200-native implementation."""
from adgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,  # noqa: F401
                                  _RasterizeGaussians, _C, cpu_deep_copy_tuple)
