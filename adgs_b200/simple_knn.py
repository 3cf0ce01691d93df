"""Drop-in for `simple_knn._C.distCUDA2` (submodules/simple-knn/ext.cpp:16, spatial.cu:16-26):
mean squared distance to the three nearest neighbours of every point, used once at
initialisation (scene/gaussian_model.py:277). Exact, bit-identical to the reference."""
import torch

from . import _lib as L


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor")
    lib = L.load()
    pts = points.contiguous().float()
    P = pts.shape[0]
    means = torch.zeros((P,), dtype=torch.float32, device=pts.device)
    if P == 0:
        return means
    with torch.cuda.device(pts.device):
        ws = torch.empty((lib.adgs_knn_workspace_bytes(P),), dtype=torch.uint8, device=pts.device)
        st = lib.adgs_dist_cuda2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(),
                                 torch.cuda.current_stream(pts.device).cuda_stream)
    L.check(st, "distCUDA2")
    return means


class _C:
    distCUDA2 = staticmethod(distCUDA2)
