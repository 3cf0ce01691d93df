"""ctypes binding of libadgs_b200.so (the C ABI declared in include/adgs_b200.h).

The product path has NO fallback: if the CUDA library is missing, importing this module raises.
Nothing here imports `oracle/`.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libadgs_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

MAX_TERMS = 48
MAX_QUAT_ORDER = 7
MAX_SEMANTIC = 32


def build(verbose: bool = False) -> str:
    """Compile every kernel for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    res = subprocess.run(["make", "-j8", "-C", CSRC_DIR], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libadgs_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


class Camera(C.Structure):
    _fields_ = [
        ("image_height", C.c_int32),
        ("image_width", C.c_int32),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32),
        ("prefiltered", C.c_int32),
        ("inv_depth", C.c_int32),
        ("debug", C.c_int32),
        ("_pad", C.c_int32),
        ("bg", C.c_void_p),
        ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p),
    ]


class Gaussians(C.Structure):
    _fields_ = [
        ("P", C.c_int32),
        ("M", C.c_int32),
        ("D_S", C.c_int32),
        ("_pad", C.c_int32),
        ("means3D", C.c_void_p),
        ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p),
        ("flow_points", C.c_void_p),
        ("semantic", C.c_void_p),
        ("opacities", C.c_void_p),
        ("scales", C.c_void_p),
        ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p),
    ]


class Images(C.Structure):
    _fields_ = [
        ("color", C.c_void_p),
        ("depth", C.c_void_p),
        ("opacity", C.c_void_p),
        ("flow", C.c_void_p),
        ("semantic", C.c_void_p),
        ("radii", C.c_void_p),
    ]


class ImageGrads(C.Structure):
    _fields_ = [
        ("dL_dcolor", C.c_void_p),
        ("dL_ddepth", C.c_void_p),
        ("dL_dflow", C.c_void_p),
        ("dL_dsemantic", C.c_void_p),
        ("dL_dopacity", C.c_void_p),
    ]


class GaussianGrads(C.Structure):
    _fields_ = [
        ("dL_dmeans2D", C.c_void_p),
        ("dL_dcolors", C.c_void_p),
        ("dL_dopacity", C.c_void_p),
        ("dL_dmeans3D", C.c_void_p),
        ("dL_dcov3D", C.c_void_p),
        ("dL_dsh", C.c_void_p),
        ("dL_dscales", C.c_void_p),
        ("dL_drotations", C.c_void_p),
        ("dL_dflow_points", C.c_void_p),
        ("dL_dsemantic", C.c_void_p),
    ]


class GeometryLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "counters", "depths", "tiles_touched", "record", "cov3D", "clamped", "depth_order",
        "point_offsets", "total")]


class BinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "point_list", "point_list_tile", "point_list_alt", "point_list_tile_alt", "total")]


class ImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("ranges", "n_contrib", "total")]


class LinBasis(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("n_cols", C.c_int32),
        ("col", C.c_int16 * MAX_TERMS),
        ("w0", C.c_float * MAX_TERMS),
        ("w1", C.c_float * MAX_TERMS),
    ]


class QuatBasis(C.Structure):
    _fields_ = [
        ("k", C.c_int32),
        ("n_ctrl", C.c_int32),
        ("start", C.c_int32),
        ("_pad", C.c_int32),
        ("cum", C.c_float * (MAX_QUAT_ORDER + 1)),
    ]


class TimeBasis(C.Structure):
    _fields_ = [
        ("xyz", LinBasis),
        ("background", LinBasis),
        ("shs", LinBasis),
        ("rotation", LinBasis),
        ("quat", QuatBasis),
        ("t", C.c_float),
        ("use_time_mask", C.c_int32),
        ("has_flow", C.c_int32),
        ("sparse_grads", C.c_int32),
    ]


class Model(C.Structure):
    _fields_ = [
        ("N_scene", C.c_int32),
        ("N_obj", C.c_int32),
        ("xyz", C.c_void_p),
        ("scaling", C.c_void_p),
        ("rotation", C.c_void_p),
        ("opacity", C.c_void_p),
        ("sh4", C.c_void_p),
        ("shs_deform4", C.c_void_p),
        ("xyz_deform", C.c_void_p),
        ("rot_deform", C.c_void_p),
        ("background_deform", C.c_void_p),
        ("gs_time", C.c_void_p),
        ("gs_time_sigma", C.c_void_p),
    ]


class Deformed(C.Structure):
    _fields_ = [
        ("xyz", C.c_void_p),
        ("rotation", C.c_void_p),
        ("shs", C.c_void_p),
        ("opacity", C.c_void_p),
        ("scaling", C.c_void_p),
        ("flow_xyz", C.c_void_p),
    ]


class Splats(C.Structure):
    _fields_ = [
        ("P", C.c_int32),
        ("_pad", C.c_int32),
        ("record", C.c_void_p),
        ("depth_keys", C.c_void_p),
        ("tiles_touched", C.c_void_p),
        ("radii", C.c_void_p),
        ("mean_x", C.c_void_p),
        ("mean_y", C.c_void_p),
    ]


class AdamSegment(C.Structure):
    _fields_ = [
        ("param", C.c_void_p),
        ("grad", C.c_void_p),
        ("exp_avg", C.c_void_p),
        ("exp_avg_sq", C.c_void_p),
        ("n", C.c_int64),
        ("split", C.c_int64),
        ("plane", C.c_int64),
        ("active", C.c_uint64 * 2),
        ("lr_a", C.c_double),
        ("lr_b", C.c_double),
        ("lr_rule", C.c_int32),
        ("_pad", C.c_int32),
    ]


class PixelLossInputs(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32),
        ("depth", C.c_void_p), ("gt_depth", C.c_void_p), ("img_semantic", C.c_void_p), ("gt_semantic", C.c_void_p),
        ("img_opacity", C.c_void_p), ("gt_sky", C.c_void_p), ("img_flow", C.c_void_p), ("flow", C.c_void_p),
        ("flow_vis", C.c_void_p), ("flow_opacity", C.c_void_p),
        ("K", C.c_float * 9), ("R", C.c_float * 9), ("T", C.c_float * 3),
        ("flow_dist", C.c_float),
        ("lambda_depth", C.c_float), ("lambda_obj", C.c_float), ("lambda_sky", C.c_float), ("lambda_flow", C.c_float),
    ]


class EnvMap(C.Structure):
    _fields_ = [("R", C.c_int32), ("C", C.c_int32), ("grid", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("touched", C.c_void_p), ("tile_list", C.c_void_p)]


class DensifyParams(C.Structure):
    _fields_ = [("N_scene", C.c_int32), ("N_obj", C.c_int32), ("mode", C.c_int32), ("n_split", C.c_int32),
                ("max_scene_grad", C.c_float), ("max_obj_grad", C.c_float), ("scene_split_size", C.c_float),
                ("obj_split_size", C.c_float), ("min_opacity", C.c_float), ("prune_big", C.c_int32),
                ("scene_big_size", C.c_float), ("obj_big_size", C.c_float), ("inv_split_scale", C.c_float),
                ("_pad", C.c_int32)]


class GatherSegment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("planes", C.c_int32), ("width", C.c_int32),
                ("src_rows", C.c_int32), ("dst_rows", C.c_int32), ("src_row0", C.c_int32), ("dst_row0", C.c_int32),
                ("zero_new", C.c_int32), ("_pad", C.c_int32)]


DENSIFY_AND_PRUNE, DENSIFY_PRUNE_ONLY = 0, 1
DENSIFY_KIND_KEEP, DENSIFY_KIND_CLONE, DENSIFY_KIND_CHILD = 0, 1, 2
GATHER_MAX_SEGMENTS = 40

ADAM_MAX_SEGMENTS = 16
ADAM_LR_UNIFORM, ADAM_LR_SPLIT, ADAM_LR_SH4 = 0, 1, 2

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)

_P = C.POINTER

# name -> (restype, argtypes); every symbol include/adgs_b200.h declares
SIGNATURES = {
    "adgs_abi_version": (C.c_int, []),
    "adgs_status_string": (C.c_char_p, [C.c_int]),
    "adgs_last_cuda_error": (C.c_char_p, []),
    "adgs_geometry_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_binning_bytes": (C.c_size_t, [C.c_int64]),
    "adgs_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "adgs_backward_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_mark_visible": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_rasterize_forward": (C.c_int, [_P(Camera), _P(Gaussians), _P(Images), ALLOC_FN, ALLOC_FN, ALLOC_FN,
                                         C.c_void_p, C.c_void_p]),
    "adgs_rasterize_forward_async": (C.c_int, [_P(Camera), _P(Gaussians), _P(Images), C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p, C.c_void_p]),
    "adgs_read_counters": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "adgs_rasterize_backward": (C.c_int, [_P(Camera), _P(Gaussians), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, _P(ImageGrads), _P(GaussianGrads), C.c_void_p,
                                          C.c_void_p]),
    "adgs_geometry_offsets": (C.c_int, [C.c_int32, _P(GeometryLayout)]),
    "adgs_binning_offsets": (C.c_int, [C.c_int64, _P(BinningLayout)]),
    "adgs_binning_result_in_alt": (C.c_int, [C.c_int32, C.c_int32]),
    "adgs_image_offsets": (C.c_int, [C.c_int32, C.c_int32, _P(ImageLayout)]),
    "adgs_sort_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "adgs_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
    "adgs_trajectory_forward": (C.c_int, [_P(Model), _P(TimeBasis), _P(Deformed), C.c_void_p]),
    "adgs_render_saved_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_render_forward": (C.c_int, [_P(Camera), _P(Model), _P(TimeBasis), C.c_int32, _P(Images), _P(Deformed),
                                      C.c_void_p, C.c_void_p, C.c_int64, ALLOC_FN, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "adgs_render_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "adgs_render_backward": (C.c_int, [_P(Camera), _P(Model), _P(TimeBasis), C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, _P(ImageGrads),
                                       _P(Model), C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_shard_state_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_shard_state_radii_offset": (C.c_size_t, [C.c_int32]),
    "adgs_shard_forward": (C.c_int, [_P(Camera), _P(Model), _P(TimeBasis), C.c_int32, _P(Splats), C.c_void_p,
                                     C.c_void_p]),
    "adgs_splats_forward": (C.c_int, [_P(Camera), _P(Splats), C.c_int32, C.c_int32, _P(Images), C.c_void_p, C.c_void_p,
                                      C.c_int64, ALLOC_FN, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_splats_bin": (C.c_int, [_P(Camera), _P(Splats), C.c_void_p, C.c_void_p, C.c_int64, ALLOC_FN, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "adgs_splats_blend": (C.c_int, [_P(Camera), _P(Splats), C.c_int32, C.c_int32, _P(Images), C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p]),
    "adgs_splats_backward": (C.c_int, [_P(Camera), _P(Splats), C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, _P(ImageGrads), C.c_void_p, C.c_void_p]),
    "adgs_shard_backward": (C.c_int, [_P(Camera), _P(Model), _P(TimeBasis), C.c_void_p, C.c_void_p, C.c_void_p,
                                      _P(Model), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_shard_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "adgs_shard_forward_multi": (C.c_int, [C.c_int32, _P(Camera), _P(Model), _P(TimeBasis), C.c_int32, _P(Splats),
                                           _P(C.c_void_p), C.c_void_p]),
    "adgs_shard_backward_multi": (C.c_int, [C.c_int32, _P(Camera), _P(Model), _P(TimeBasis), _P(C.c_void_p),
                                            _P(C.c_void_p), _P(C.c_void_p), _P(Model), C.c_int32, _P(C.c_void_p),
                                            C.c_void_p, C.c_void_p]),
    "adgs_peer_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p), C.c_char_p]),
    "adgs_peer_open": (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    "adgs_peer_close": (C.c_int, [C.c_void_p]),
    "adgs_peer_free": (C.c_int, [C.c_void_p]),
    "adgs_peer_sum": (C.c_int, [C.c_int32, _P(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p]),
    "adgs_peer_barrier": (C.c_int, [C.c_int32, C.c_int32, _P(C.c_void_p), C.c_uint32, C.c_void_p, C.c_void_p]),
    "adgs_adam_step": (C.c_int, [_P(AdamSegment), C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int64,
                                 C.c_void_p]),
    "adgs_image_loss_partial_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "adgs_image_loss_forward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 6 +
                                [C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "adgs_image_loss_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 6 +
                                 [C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "adgs_pixel_loss_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "adgs_pixel_loss": (C.c_int, [_P(PixelLossInputs), C.c_int32] + [C.c_void_p] * 8),
    "adgs_env_touched_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_env_tile_list_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_env_forward": (C.c_int, [_P(EnvMap), C.c_int32, C.c_int32, C.c_float] + [C.c_void_p] * 6),
    "adgs_env_backward": (C.c_int, [_P(EnvMap), C.c_int32, C.c_int32, C.c_float] + [C.c_void_p] * 6),
    "adgs_env_adam_step": (C.c_int, [_P(EnvMap), C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "adgs_launch_count": (C.c_ulonglong, []),
    "adgs_selftest_exp_pair": (C.c_int, [C.c_void_p, C.c_void_p]),
    "adgs_profile_begin": (C.c_int, []),
    "adgs_profile_num_stages": (C.c_int, []),
    "adgs_profile_stage_name": (C.c_char_p, [C.c_int]),
    "adgs_profile_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "adgs_knn_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "adgs_dist_cuda2": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_densify_stats": (C.c_int, [C.c_int32] + [C.c_void_p] * 6),
    "adgs_densify_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "adgs_densify_classify": (C.c_int, [_P(DensifyParams)] + [C.c_void_p] * 6 + [_P(C.c_int32), C.c_void_p]),
    "adgs_densify_plan": (C.c_int, [_P(DensifyParams), C.c_void_p, _P(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_densify_gather": (C.c_int, [_P(GatherSegment), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_densify_split": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_int32, C.c_void_p, C.c_void_p,
                                                                                C.c_void_p]),
    "adgs_reset_opacity": (C.c_int, [C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "adgs_knn_points_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "adgs_knn_points": (C.c_int, [C.c_int32] * 4 + [C.c_void_p] * 6),
}

_lib = None


def load():
    """Load libadgs_b200.so (raises if it has not been built: there is no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C adgs_b200/csrc`. adgs_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AdgsError(RuntimeError):
    pass


def check(status: int, what: str) -> int:
    if status < 0:
        lib = load()
        msg = lib.adgs_status_string(status).decode()
        if status == -2:
            msg += " (" + lib.adgs_last_cuda_error().decode() + ")"
        raise AdgsError(f"{what}: {msg}")
    return status


def ptr(t):
    """Device pointer of a tensor, or None (null) for None / empty tensors (the reference's sentinel)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()
