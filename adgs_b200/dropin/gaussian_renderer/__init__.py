"""`from gaussian_renderer import render` (train.py:22, render.py:19) -> fused B200-native path."""
from adgs_b200.gaussian_renderer import render  # noqa: F401
