"""`from simple_knn._C import distCUDA2` (scene/gaussian_model.py:20) -> B200-native implementation."""
