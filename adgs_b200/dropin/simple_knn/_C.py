from adgs_b200.simple_knn import distCUDA2  # noqa: F401
