"""CPU-only: the C-ABI library loads and exports every symbol include/adgs_b200.h declares
(no compute calls), and the host-visible sizing functions behave."""
import ctypes as C
import os
import re

from adgs_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "adgs_b200.h")).read()
    return sorted(set(re.findall(r"ADGS_API[^;(]*?\b(adgs_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/adgs_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in adgs_b200/_lib.py"
    assert sorted(L.SIGNATURES) == names


def test_version_and_status_strings():
    lib = L.load()
    assert lib.adgs_abi_version() == 1
    assert lib.adgs_status_string(0) == b"ok"
    assert b"argument" in lib.adgs_status_string(-1)


def test_arena_sizes_are_deterministic_and_monotone():
    lib = L.load()
    assert lib.adgs_geometry_bytes(1000) == lib.adgs_geometry_bytes(1000)
    assert lib.adgs_geometry_bytes(2000) > lib.adgs_geometry_bytes(1000) > 0
    assert lib.adgs_binning_bytes(10) < lib.adgs_binning_bytes(10_000_000)
    assert lib.adgs_image_bytes(1242, 375) >= 1242 * 375 * 4
    gl = L.GeometryLayout()
    assert lib.adgs_geometry_offsets(12345, C.byref(gl)) == 0
    offs = [gl.counters, gl.depths, gl.tiles_touched, gl.record, gl.cov3D, gl.clamped]
    assert offs == sorted(offs) and all(o % 128 == 0 for o in offs)
    assert gl.total + 128 <= lib.adgs_geometry_bytes(12345)
    # per-Gaussian forward state stays ~120 B (vs 79 B + CUB temp in the reference, rasterizer_impl.cu:155-170)
    assert lib.adgs_geometry_bytes(1_000_000) < 140 * 1_000_000
    # 16 B per instance (two uint32 ping-pong pairs) vs the reference's 24 B + CUB temp (rasterizer_impl.cu:180-194)
    assert lib.adgs_binning_bytes(3_000_000) < 20 * 3_000_000


def test_struct_sizes_match_header():
    # LP64 layout of the structs that cross the boundary
    assert C.sizeof(L.Camera) == 72
    assert C.sizeof(L.Gaussians) == 16 + 9 * 8
    assert C.sizeof(L.Images) == 48
    assert C.sizeof(L.LinBasis) == 8 + 96 + 192 + 192
    assert C.sizeof(L.QuatBasis) == 16 + 32
    assert C.sizeof(L.TimeBasis) == 4 * 488 + 48 + 16
    assert C.sizeof(L.Model) == 8 + 11 * 8
    assert C.sizeof(L.DensifyParams) == 14 * 4
    assert C.sizeof(L.GatherSegment) == 2 * 8 + 8 * 4


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = L.load()
    assert lib.adgs_mark_visible(-1, None, None, None, None, None) == -1
    assert lib.adgs_sort_pairs(None, None, None, None, 10, 0, 40, None, None) == -1
    assert lib.adgs_rasterize_backward(None, None, None, None, 0, None, None, None, None, None, None, None) == -1


def test_more_than_2_pow_30_instances_is_refused_not_wrapped():
    """The CUB-free sort's look-back words carry a 30-bit count (adgs_b200/csrc/sort.cu); the reference's 64-bit
    sort (rasterizer_impl.cu:310-315) is valid to 2^32. Beyond 2^30 the library must say so, before any CUDA call."""
    lib = L.load()
    assert lib.adgs_sort_pairs(None, None, None, None, 1 << 30, 0, 32, None, None) == -4
    assert lib.adgs_status_string(-4) == b"unsupported configuration"
    assert lib.adgs_rasterize_forward_async(None, None, None, None, None, 1 << 30, None, None) == -4
    assert lib.adgs_rasterize_forward_async(None, None, None, None, None, (1 << 30) - 1, None, None) == -1


def test_densify_and_knn_reject_bad_arguments_without_a_gpu():
    """Argument validation of the densification / K-NN entry points returns before any CUDA call."""
    lib = L.load()
    p = L.DensifyParams(N_scene=10, N_obj=5, mode=L.DENSIFY_AND_PRUNE, n_split=2)
    tot = (C.c_int32 * 8)()
    assert lib.adgs_densify_classify(None, None, None, None, None, None, None, tot, None) == -1
    assert lib.adgs_densify_classify(C.byref(p), None, None, None, None, None, None, tot, None) == -1   # no workspace
    ws = (C.c_char * 4096)()
    base = C.addressof(ws)
    aligned = (base + 15) & ~15
    assert lib.adgs_densify_classify(C.byref(p), None, None, None, None, None, aligned, tot, None) == -1   # inputs missing
    assert lib.adgs_densify_classify(C.byref(p), None, None, None, None, None, aligned + 4, tot, None) == -1  # alignment
    bad = L.DensifyParams(N_scene=10, N_obj=5, mode=7, n_split=2)
    assert lib.adgs_densify_classify(C.byref(bad), None, None, None, None, None, aligned, tot, None) == -1
    pr = L.DensifyParams(N_scene=10, N_obj=5, mode=L.DENSIFY_PRUNE_ONLY, n_split=1)
    assert lib.adgs_densify_classify(C.byref(pr), None, None, None, None, None, aligned, tot, None) == -1   # no mask
    assert lib.adgs_densify_workspace_bytes(1000, 500) >= 1500
    assert lib.adgs_densify_workspace_bytes(10**6, 10**6) > lib.adgs_densify_workspace_bytes(1000, 500)
    assert lib.adgs_densify_plan(C.byref(p), None, tot, None, None, None) == -1
    seg = (L.GatherSegment * 1)()
    assert lib.adgs_densify_gather(seg, L.GATHER_MAX_SEGMENTS + 1, None, None, None) == -1
    seg[0].planes, seg[0].width, seg[0].dst_rows = 1, 5, 4                       # width must be 1..4
    assert lib.adgs_densify_gather(seg, 1, None, None, None) == -1
    seg[0].width, seg[0].dst_rows = 4, 0                                         # nothing to write: no launch
    assert lib.adgs_densify_gather(seg, 1, None, None, None) == 0
    assert lib.adgs_densify_stats(-1, None, None, None, None, None, None) == -1
    assert lib.adgs_densify_stats(0, None, None, None, None, None, None) == 0
    assert lib.adgs_reset_opacity(0, 0.01, None, None, None, None) == 0
    assert lib.adgs_reset_opacity(5, 0.01, None, None, None, None) == -1
    # K-NN: K <= 32, D in {3, 4}, K <= P
    assert lib.adgs_knn_points(4, 100, 3, 33, None, None, None, None, None, None) == -4
    assert lib.adgs_knn_points(4, 100, 5, 8, None, None, None, None, None, None) == -4
    assert lib.adgs_knn_points(4, 6, 3, 8, None, None, None, None, None, None) == -1
    assert lib.adgs_knn_points(0, 100, 3, 8, None, None, None, None, None, None) == 0
    assert lib.adgs_knn_points_workspace_bytes(1000, 10**5, 8) >= 1000 * 8 * 8
