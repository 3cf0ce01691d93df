/*
 * adgs_b200.h -- C ABI of the B200-native AD-GS hot path (libadgs_b200.so).
 *
 * Plain pointers and sizes only; no torch / C++ types cross this boundary. Every DEVICE pointer
 * must be valid on the current CUDA device, every launch goes to the `stream` argument (the
 * host passes torch's current stream), and no entry point allocates device memory: the caller
 * owns all storage (directly, or through the adgs_alloc_fn callbacks that mirror the
 * reference's std::function<char*(size_t)> arena allocators).
 *
 * Each entry point names the reference interface it replaces; paths are relative to the
 * reference checkout, RZ/ = submodules/depth-diff-gaussian-rasterization/,
 * KNN/ = submodules/simple-knn/.
 *
 * Return value: >= 0 on success (adgs_rasterize_forward returns num_rendered), < 0 = adgs_status.
 */
#ifndef ADGS_B200_H_INCLUDED
#define ADGS_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADGS_ABI_VERSION 1

#if defined(__GNUC__)
#define ADGS_API __attribute__((visibility("default")))
#else
#define ADGS_API
#endif

typedef struct CUstream_st* adgs_stream_t;

/* Arena allocator: must return a DEVICE pointer to at least `bytes` bytes, 128-byte aligned.
 * Replaces the three std::function<char*(size_t)> of RZ/cuda_rasterizer/rasterizer.h:34-36
 * (bound to torch tensors by resizeFunctional, RZ/rasterize_points.cu:27-33). */
typedef char* (*adgs_alloc_fn)(size_t bytes, void* user);

typedef enum adgs_status {
    ADGS_OK = 0,
    ADGS_ERR_ARG = -1,         /* bad argument (null pointer, bad shape) */
    ADGS_ERR_CUDA = -2,        /* a CUDA call failed; adgs_last_cuda_error() has the text */
    ADGS_ERR_CAPACITY = -3,    /* binning arena too small in the *_async path */
    ADGS_ERR_UNSUPPORTED = -4, /* e.g. D_S > 32, SH degree > 3 (same limits as the reference) */
    ADGS_ERR_ALLOC = -5        /* an adgs_alloc_fn returned null */
} adgs_status;

#define ADGS_MAX_SEMANTIC 32 /* RZ/cuda_rasterizer/config.h:18 */
#define ADGS_TILE 16         /* RZ/cuda_rasterizer/config.h:16-17 */

ADGS_API int adgs_abi_version(void);
ADGS_API const char* adgs_status_string(int status);
ADGS_API const char* adgs_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Camera / raster settings: the fields of GaussianRasterizationSettings
 * (RZ/diff_gaussian_rasterization/__init__.py:176-189). Matrices are the 16 floats of the
 * torch tensors as they lie in memory, i.e. the transposed W2C / full projection of
 * scene/cameras.py:77-79, read as m[col*4+row] (RZ/cuda_rasterizer/auxiliary.h:58-77).
 * ---------------------------------------------------------------------------------------- */
typedef struct adgs_camera {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
    int32_t sh_degree;
    int32_t prefiltered;
    int32_t inv_depth;
    int32_t debug;           /* !=0: synchronise + check after every stage (CHECK_CUDA, auxiliary.h:166) */
    int32_t _pad;
    const float* bg;         /* device, 3 */
    const float* viewmatrix; /* device, 16 */
    const float* projmatrix; /* device, 16 */
    const float* campos;     /* device, 3 */
} adgs_camera;

/* Per-Gaussian inputs of CudaRasterizer::Rasterizer::forward (RZ/cuda_rasterizer/rasterizer.h:33-65).
 * A null pointer means "not provided" exactly like the reference's empty-tensor sentinel
 * (RZ/diff_gaussian_rasterization/__init__.py:220-236). */
typedef struct adgs_gaussians {
    int32_t P;                /* number of Gaussians */
    int32_t M;                /* SH coefficients per Gaussian as laid out in `shs` (0 if none) */
    int32_t D_S;              /* semantic channels (0 if none), <= 32 */
    int32_t _pad;
    const float* means3D;     /* (P,3) */
    const float* shs;         /* (P,M,3) or null */
    const float* colors_precomp; /* (P,3) or null */
    const float* flow_points; /* (P,3) or null */
    const float* semantic;    /* (P,D_S) or null */
    const float* opacities;   /* (P,1) */
    const float* scales;      /* (P,3) or null */
    const float* rotations;   /* (P,4) wxyz, NOT normalised by the rasterizer (forward.cu:127) */
    const float* cov3D_precomp; /* (P,6) or null */
} adgs_gaussians;

/* Outputs of the forward: the six tensors of _RasterizeGaussians.forward
 * (RZ/diff_gaussian_rasterization/__init__.py:107), planar CHW float32. The library
 * writes every element (no pre-zeroing needed). */
typedef struct adgs_images {
    float* color;    /* (3,H,W) */
    float* depth;    /* (1,H,W) */
    float* opacity;  /* (1,H,W)  = 1 - T_final */
    float* flow;     /* (3,H,W) */
    float* semantic; /* (D_S,H,W) */
    int32_t* radii;  /* (P,) */
} adgs_images;

/* Arena sizes. Layout is a pure function of (P), (R), (W,H) so the backward can re-derive it
 * from the same base pointers (same contract as RZ/cuda_rasterizer/rasterizer_impl.cu:398-400). */
ADGS_API size_t adgs_geometry_bytes(int32_t P);
ADGS_API size_t adgs_binning_bytes(int64_t R);
ADGS_API size_t adgs_image_bytes(int32_t width, int32_t height);
ADGS_API size_t adgs_backward_scratch_bytes(int32_t P);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:24-29; _C.mark_visible). */
ADGS_API int adgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, adgs_stream_t stream);

/* Replaces CudaRasterizer::Rasterizer::forward (rasterizer.h:31-65; _C.rasterize_gaussians,
 * RZ/rasterize_points.cu:35-140). Returns num_rendered (>= 0). Like the reference it performs
 * exactly one blocking 4-byte device-to-host read (rasterizer_impl.cu:288) to size the
 * binning arena. */
ADGS_API int adgs_rasterize_forward(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out,
                           adgs_alloc_fn geometry_alloc, adgs_alloc_fn binning_alloc,
                           adgs_alloc_fn image_alloc, void* alloc_user, adgs_stream_t stream);

/* Same computation with caller-provided arenas and NO host synchronisation: `binning` must hold
 * adgs_binning_bytes(capacity) bytes. The number of rendered instances stays on the device
 * (adgs_geometry_layout.num_rendered); if it exceeds `capacity` the overflow flag is raised and
 * the images are undefined -- the caller checks adgs_read_counters() later and retries bigger. */
ADGS_API int adgs_rasterize_forward_async(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out,
                                 char* geometry, char* binning, int64_t capacity, char* image,
                                 adgs_stream_t stream);

/* Async 8-byte copy of {num_rendered, overflow} from the geometry arena into pinned host memory. */
ADGS_API int adgs_read_counters(const char* geometry, int32_t P, uint32_t* host_pinned_2, adgs_stream_t stream);

/* Cotangents of the five images (grad_radii is ignored, as in the reference). Null = zero. */
typedef struct adgs_image_grads {
    const float* dL_dcolor;    /* (3,H,W) */
    const float* dL_ddepth;    /* (1,H,W) */
    const float* dL_dflow;     /* (3,H,W) */
    const float* dL_dsemantic; /* (D_S,H,W) */
    const float* dL_dopacity;  /* (1,H,W) grad_img_opacity */
} adgs_image_grads;

/* The ten gradients of _C.rasterize_gaussians_backward (RZ/rasterize_points.cu:253). The library
 * writes every element of every non-null tensor (zeros for culled Gaussians): no pre-zeroing. */
typedef struct adgs_gaussian_grads {
    float* dL_dmeans2D;     /* (P,3), z unused (=0) */
    float* dL_dcolors;      /* (P,3) */
    float* dL_dopacity;     /* (P,1) */
    float* dL_dmeans3D;     /* (P,3) */
    float* dL_dcov3D;       /* (P,6) */
    float* dL_dsh;          /* (P,M,3) */
    float* dL_dscales;      /* (P,3) */
    float* dL_drotations;   /* (P,4) */
    float* dL_dflow_points; /* (P,3) */
    float* dL_dsemantic;    /* (P,D_S) */
} adgs_gaussian_grads;

/* Replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:67-101;
 * _C.rasterize_gaussians_backward, RZ/rasterize_points.cu:142-254). `img_opacity` is the
 * forward's opacity image (the reference's final_T buffer, forward.cu:389 / backward.cu:473). */
ADGS_API int adgs_rasterize_backward(const adgs_camera* cam, const adgs_gaussians* g, const int32_t* radii,
                            const char* geometry, int64_t R, const char* binning, const char* image,
                            const float* img_opacity, const adgs_image_grads* dpix,
                            const adgs_gaussian_grads* grads, char* scratch, adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Introspection (tests / parity only): byte offsets of the internal arrays inside the arenas,
 * the analogue of re-running GeometryState/BinningState/ImageState::fromChunk
 * (rasterizer_impl.cu:155-194) on a returned buffer.
 * ---------------------------------------------------------------------------------------- */
typedef struct adgs_geometry_layout {
    size_t counters;       /* uint32[4]: num_rendered, overflow, -, -            */
    size_t depths;         /* uint32 [P]  float bits of view-space z, SORTED ascending (culled = ~0) */
    size_t tiles_touched;  /* uint32 [P]                                          */
    size_t record;         /* float  [P][16] packed blend record                  */
    size_t cov3D;          /* float  [P][6]                                       */
    size_t clamped;        /* uint8  [P]  bit c set = channel c clamped           */
    size_t depth_order;    /* uint32 [P]  Gaussian ids sorted by (depth, id)      */
    size_t point_offsets;  /* uint32 [P]  inclusive scan of tiles_touched in depth order */
    size_t total;
} adgs_geometry_layout;

typedef struct adgs_binning_layout {
    /* The sorted lists live in the primary or the alternate buffers depending on the parity of
     * the tile-sort pass count, a function of the image size only: adgs_binning_result_in_alt(). */
    size_t point_list;          /* uint32 [R] Gaussian id per instance, sorted by (tile, depth, id) */
    size_t point_list_tile;     /* uint32 [R] tile id per sorted instance                           */
    size_t point_list_alt;
    size_t point_list_tile_alt;
    size_t total;
} adgs_binning_layout;

typedef struct adgs_image_layout {
    size_t ranges;    /* uint32 [tiles][2] */
    size_t n_contrib; /* uint32 [H*W]      */
    size_t total;
} adgs_image_layout;

ADGS_API int adgs_geometry_offsets(int32_t P, adgs_geometry_layout* out);
ADGS_API int adgs_binning_offsets(int64_t R, adgs_binning_layout* out);
ADGS_API int adgs_binning_result_in_alt(int32_t width, int32_t height);
ADGS_API int adgs_image_offsets(int32_t width, int32_t height, adgs_image_layout* out);

/* Stand-alone stable LSD radix sort of (uint32 key, uint32 value) pairs on key bits
 * [begin_bit, end_bit) -- the CUB-free onesweep that replaces cub::DeviceRadixSort::SortPairs
 * (rasterizer_impl.cu:310-315; KNN/simple_knn.cu:210-213). Exposed for tests and for distCUDA2.
 * Both buffer pairs are clobbered; returns 0 when the sorted pairs are in (keys_out, vals_out),
 * 1 when they are in (keys_in, vals_in) (even number of 8-bit passes), < 0 on error.
 * workspace: adgs_sort_workspace_bytes(n). */
ADGS_API size_t adgs_sort_workspace_bytes(int64_t n);
ADGS_API int adgs_sort_pairs(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                    int64_t n, int32_t begin_bit, int32_t end_bit, char* workspace,
                    adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Trajectory (object-aware B-spline / Fourier / polynomial / cumulative quaternion B-spline)
 * fused with the rasterizer front end. Replaces, for one time t (and optionally a second time
 * flow_t), GaussianModel.get_deformed_pkg / get_deformed_xyz (scene/gaussian_model.py:173-231),
 * get_scaling (:88-91), get_obj_mask (:154-159) and utils/func_utils.py:get_func_result
 * (:121-173) followed by the per-Gaussian preprocess (forward.cu:155-256).
 *
 * The host reduces get_func_result's four linear bases (B-spline window, polynomial, Fourier;
 * func_utils.py:127-153) to one sparse weight list per attribute: value = sum_j
 * param[..., col[j]] * w[j]. `w1` carries the weights of the second time (flow_t) over the same
 * column list so that both evaluations read each coefficient once.
 * ---------------------------------------------------------------------------------------- */
#define ADGS_MAX_TERMS 48
#define ADGS_MAX_QUAT_ORDER 7

typedef struct adgs_lin_basis {
    int32_t n;                    /* number of terms (0 = attribute not deformed) */
    int32_t n_cols;               /* C: total parameter columns of the attribute */
    int16_t col[ADGS_MAX_TERMS];
    float w0[ADGS_MAX_TERMS];     /* weights at t */
    float w1[ADGS_MAX_TERMS];     /* weights at flow_t (xyz / background only) */
} adgs_lin_basis;

typedef struct adgs_quat_basis {
    int32_t k;                    /* spline order k_q (0 with n_ctrl==0 => no quaternion spline) */
    int32_t n_ctrl;               /* n_q */
    int32_t start;                /* first control quaternion of the window (column index) */
    int32_t _pad;
    float cum[ADGS_MAX_QUAT_ORDER + 1]; /* cum[i] = sum_{j>=i} B_j(u), i = 1..k (func_utils.py:163) */
} adgs_quat_basis;

typedef struct adgs_time_basis {
    adgs_lin_basis xyz;        /* order_args['xyz'] */
    adgs_lin_basis background; /* order_args['background'] */
    adgs_lin_basis shs;        /* order_args['shs'] */
    adgs_lin_basis rotation;   /* linear part of order_args['rotation'] */
    adgs_quat_basis quat;      /* quaternion-spline part of order_args['rotation'] */
    float t;                   /* camera time (time mask, gaussian_model.py:207-214) */
    int32_t use_time_mask;
    int32_t has_flow;          /* evaluate xyz at the second time too (gaussian_renderer/__init__.py:54-57) */
    int32_t sparse_grads;      /* backward only, != 0: leave the gradient planes of control-point columns that are
                                * inactive at this time UNWRITTEN instead of zero-filling them; the caller's
                                * optimizer must then be window-aware (adgs_adam_segment.active). Not valid
                                * together with accumulate != 0. */
} adgs_time_basis;

/* Parameter storage of the B200-native model (host container: adgs_b200/gaussian_model.py).
 * Gaussians are ordered [scene (N_scene) ; object (N_obj)], N = N_scene + N_obj, i.e. the order
 * of every torch.cat in scene/gaussian_model.py. Per-Gaussian rows stay AoS where a row is one
 * vector load; the wide per-Gaussian blocks are planar so that a warp reads 128-byte lines:
 *   xyz (N,3) scaling (N,3) rotation (N,4) opacity (N,)                    raw, pre-activation
 *   sh4       (12,N,4)   the (16,3) SH block of Gaussian g, flattened, in float4 chunks
 *   shs_deform4 (ceil(3*Cs/4),N,4)  (3,Cs) block flattened, in float4 chunks
 *   xyz_deform (Cx,3,N_obj)   rot_deform (Cr,N_obj,4)   background_deform (3,Cb)
 *   gs_time (N_obj,)  gs_time_sigma (N_obj,2)
 * The same struct describes the gradient buffers (same layouts). */
typedef struct adgs_model {
    int32_t N_scene;
    int32_t N_obj;
    float* xyz;
    float* scaling;
    float* rotation;
    float* opacity;
    float* sh4;
    float* shs_deform4;
    float* xyz_deform;
    float* rot_deform;
    float* background_deform;
    float* gs_time;        /* not a trainable parameter in the reference; null in gradient structs */
    float* gs_time_sigma;
} adgs_model;

/* Optional materialised outputs of the trajectory (the `deform_pkg` of
 * gaussian_renderer/__init__.py:61-66 in reference shapes); any pointer may be null. */
typedef struct adgs_deformed {
    float* xyz;      /* (N,3) */
    float* rotation; /* (N,4) normalised */
    float* shs;      /* (N,16,3) */
    float* opacity;  /* (N,1) */
    float* scaling;  /* (N,3) exp-activated */
    float* flow_xyz; /* (N,3) xyz at flow_t */
} adgs_deformed;

/* get_deformed_pkg(t) alone (no rasterisation): fills `out`. */
ADGS_API int adgs_trajectory_forward(const adgs_model* model, const adgs_time_basis* basis,
                            const adgs_deformed* out, adgs_stream_t stream);

/* Fused render forward: trajectory + preprocess + binning + blend.
 * semantic = object mask (render_objmask=True, gaussian_renderer/__init__.py:71-73) when
 * `render_objmask` != 0. Two binning modes:
 *   - `binning` != null: caller-provided arena of adgs_binning_bytes(capacity); NO host
 *     synchronisation; returns 0. Overflow is reported through adgs_read_counters().
 *   - `binning` == null: `binning_alloc` is called with the exact size after one blocking 4-byte
 *     read of num_rendered (the reference's behaviour, rasterizer_impl.cu:288); returns num_rendered.
 * `saved` (adgs_render_saved_bytes(N)) keeps what the backward needs from the trajectory
 * (xyz(t), activated opacity, normalised rotation, deformed SH DC). */
ADGS_API size_t adgs_render_saved_bytes(int32_t N);
ADGS_API size_t adgs_render_scratch_bytes(int32_t N, int32_t N_obj);
ADGS_API int adgs_render_forward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                        int32_t render_objmask, const adgs_images* out, const adgs_deformed* deformed,
                        char* geometry, char* binning, int64_t capacity, adgs_alloc_fn binning_alloc,
                        void* alloc_user, char* image, char* saved, adgs_stream_t stream);

/* Fused render backward: blend backward + preprocess backward + trajectory backward, writing
 * DENSE parameter gradients in the model layouts (zeros outside the B-spline windows, which is
 * what autograd produces in the reference -- SURVEY section 7 hard part 6) and the screen-space
 * gradient dL_dmeans2D (N,3) used by densification (gaussian_model.py:863-867). */
ADGS_API int adgs_render_backward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                         int32_t render_objmask, const int32_t* radii, const char* geometry,
                         const char* binning, int64_t capacity, const char* image, const char* saved,
                         const float* img_opacity, const adgs_image_grads* dpix,
                         const adgs_model* grads, float* dL_dmeans2D, char* scratch,
                         adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Splat exchange (multi-GPU, no reference counterpart: the reference is single-GPU). The fused
 * render split at its two natural seams so that Gaussians can be sharded across ranks while every
 * view is blended on one rank:
 *   adgs_shard_forward   trajectory + preprocess of one model shard for one view -> adgs_splats
 *                        (the only state binning/blend need: 76 B per Gaussian) + shard-local
 *                        backward state (adgs_shard_state_bytes);
 *   adgs_splats_forward  binning + blend of any concatenation of adgs_splats (e.g. all shards of a view
 *                        after an all-to-all); binning modes as in adgs_render_forward;
 *   adgs_splats_backward blend backward -> packed 64-byte gradient records, one per splat;
 *   adgs_shard_backward  per-Gaussian backward of a shard from its gradient records; `accumulate` != 0
 *                        adds into `grads` (second and later views of a batch).
 * adgs_render_forward/backward are exactly these four stages back to back on one device.
 * scratch for adgs_shard_backward: adgs_render_scratch_bytes(0, N_obj).
 * ---------------------------------------------------------------------------------------- */
typedef struct adgs_splats {
    int32_t P;
    int32_t _pad;
    float* record;           /* (P,16) packed blend records */
    uint32_t* depth_keys;    /* (P) float bits of view-space z, ~0 if culled; clobbered by the sort */
    uint32_t* tiles_touched; /* (P) */
    int32_t* radii;          /* (P) */
    float* mean_x;           /* (P) optional: pixel-space means as separate planes, so that binning can */
    float* mean_y;           /* (P)           start before the (4x larger) records have arrived         */
} adgs_splats;

ADGS_API size_t adgs_shard_state_bytes(int32_t N);
/* byte offset, inside a 128-byte aligned shard-state chunk, of the int32 radii (N) that adgs_shard_forward_multi keeps
 * for the owner of the Gaussians (the adgs_splats copy may live in the blending rank's memory); adgs_shard_backward_multi
 * reads it when its radii[v] is null */
ADGS_API size_t adgs_shard_state_radii_offset(int32_t N);
ADGS_API int adgs_shard_forward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                       int32_t render_objmask, const adgs_splats* out, char* shard_state, adgs_stream_t stream);
ADGS_API int adgs_splats_forward(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                        const adgs_images* out, char* geometry, char* binning, int64_t capacity,
                        adgs_alloc_fn binning_alloc, void* alloc_user, char* image, adgs_stream_t stream);
/* adgs_splats_forward split in two, so that the all-to-all of the records overlaps the binning:
 * adgs_splats_bin needs depth_keys / tiles_touched / radii / mean_x / mean_y only; adgs_splats_blend
 * needs the records and the arenas adgs_splats_bin filled. */
ADGS_API int adgs_splats_bin(const adgs_camera* cam, const adgs_splats* splats, char* geometry, char* binning,
                    int64_t capacity, adgs_alloc_fn binning_alloc, void* alloc_user, char* image,
                    adgs_stream_t stream);
ADGS_API int adgs_splats_blend(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                      const adgs_images* out, char* geometry, char* binning, int64_t capacity, char* image,
                      adgs_stream_t stream);
ADGS_API int adgs_splats_backward(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                         const char* binning, int64_t capacity, const char* image, const float* img_opacity,
                         const adgs_image_grads* dpix, float* grad_record, adgs_stream_t stream);
ADGS_API int adgs_shard_backward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                        const int32_t* radii, const char* shard_state, const float* grad_record,
                        const adgs_model* grads, int32_t accumulate, float* dL_dmeans2D, char* scratch,
                        adgs_stream_t stream);

/* Multi-view variants (up to 8 views per call): ONE launch evaluates a shard for all views of a
 * round -- parameters are read once, the backward sums the views in registers and writes every dense
 * gradient once. Array arguments hold `num_views` entries. scratch: adgs_shard_scratch_bytes().
 * adgs_shard_backward_multi `accumulate`: bit 0 = add into `grads` (later rounds of a batch), bit 1 = the caller has
 * already zero-filled grads->xyz_deform / rot_deform (their windows differ between views and are accumulated). */
#define ADGS_MAX_VIEWS 8
ADGS_API size_t adgs_shard_scratch_bytes(int32_t num_views, int32_t N_obj);
ADGS_API int adgs_shard_forward_multi(int32_t num_views, const adgs_camera* cams, const adgs_model* model,
                             const adgs_time_basis* bases, int32_t render_objmask, const adgs_splats* outs,
                             char* const* shard_states, adgs_stream_t stream);
ADGS_API int adgs_shard_backward_multi(int32_t num_views, const adgs_camera* cams, const adgs_model* model,
                              const adgs_time_basis* bases, const int32_t* const* radii,
                              char* const* shard_states, const float* const* grad_records,
                              const adgs_model* grads, int32_t accumulate, float* const* dL_dmeans2D,
                              char* scratch, adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Peer memory for the splat exchange (one process per GPU, all GPUs of one NVSwitch box). The reference has no
 * distributed code (train.py:55-61 renders one view per iteration on one GPU), so there is no counterpart.
 *   adgs_peer_alloc    device buffer (zero-filled) that other processes of the box can map + its 64-byte handle
 *   adgs_peer_open     map a peer's buffer from its handle -> a device pointer valid in THIS process; kernels of
 *                      this library may store to / load from it like local memory (the traffic is NVLink)
 *   adgs_peer_barrier  stream-ordered barrier between the `world` ranks: flag_arrays[p] = rank p's flag array
 *                      (ADGS_MAX_PEERS words inside a peer buffer, zero at start) as mapped in this process;
 *                      `epoch` must grow by one per barrier and be the same on every rank. When the barrier
 *                      completes on a rank's stream, everything every rank queued before ITS barrier has completed
 *                      (peer stores included). status (local device word): set to 1 if a peer did not arrive within ~20 s; the
 *                      barrier kernel then traps (CUDA error on this rank) instead of continuing with half-exchanged data.
 * ---------------------------------------------------------------------------------------- */
#define ADGS_MAX_PEERS 8
#define ADGS_PEER_HANDLE_BYTES 64
ADGS_API int adgs_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
ADGS_API int adgs_peer_open(const unsigned char* handle64, void** ptr);
ADGS_API int adgs_peer_close(void* ptr);
ADGS_API int adgs_peer_free(void* ptr);
/* out[i] = sum over ranks p < world of partials[p][i] (rank order: identical bits on every rank): the all-reduce of a
 * SMALL vector (the 3 x C_bg background-trajectory gradient every Gaussian shares) with plain peer loads, after an
 * adgs_peer_barrier -- a latency-bound NCCL call replaced by ~10 us of kernel. */
ADGS_API int adgs_peer_sum(int32_t world, const float* const* partials, int32_t n, float* out, adgs_stream_t stream);
ADGS_API int adgs_peer_barrier(int32_t world, int32_t rank, uint32_t* const* flag_arrays, uint32_t epoch,
                               uint32_t* status, adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step: replaces `gaussians.optimizer.step()` of train.py:163-167
 * for the torch.optim.Adam(l, lr=0.0, eps=1e-15) that GaussianModel.training_setup builds over 18
 * parameter groups (scene/gaussian_model.py:346-372). One launch updates every array of adgs_model;
 * groups that share an array (scene / object rows of xyz, DC / rest coefficients of the SH block)
 * differ only in their learning rate, expressed as a per-element rule:
 *   ADGS_ADAM_LR_UNIFORM  lr_a everywhere
 *   ADGS_ADAM_LR_SPLIT    lr_a for elements [0, split), lr_b behind            (xyz: split = 3 * N_scene)
 *   ADGS_ADAM_LR_SH4      lr_a for elements i < split with i % 4 != 3, else lr_b (sh4: split = 4 * N; the
 *                         DC coefficient is floats 0..2 of chunk 0 of the (12,N,4) layout)
 * `plane` > 0 declares the array as column planes of `plane` floats (control-point arrays): bit c of
 * `active` set = the backward wrote column c; for every other column (c < 128) the gradient is taken
 * as zero without being read (dense-Adam semantics without the zero-fill). plane = 0: read everything.
 * `step` is the 1-based step count (bias correction), shared by all segments.
 * ---------------------------------------------------------------------------------------- */
#define ADGS_ADAM_MAX_SEGMENTS 16
enum { ADGS_ADAM_LR_UNIFORM = 0, ADGS_ADAM_LR_SPLIT = 1, ADGS_ADAM_LR_SH4 = 2 };
typedef struct adgs_adam_segment {
    float* param;        /* device, n floats, updated in place */
    const float* grad;   /* device, n floats */
    float* exp_avg;      /* device, n floats (first moment), updated in place */
    float* exp_avg_sq;   /* device, n floats (second moment), updated in place */
    int64_t n;
    int64_t split;
    int64_t plane;
    uint64_t active[2];
    double lr_a;
    double lr_b;
    int32_t lr_rule;
    int32_t _pad;
} adgs_adam_segment;
ADGS_API int adgs_adam_step(const adgs_adam_segment* segments, int32_t num_segments, double beta1, double beta2,
                            double eps, int64_t step, adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Loss front-end: the image term of train.py:79-80,113,
 *   (1 - lambda_dssim) * lambda_l1 * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt)),
 * replacing utils/loss_utils.py:l1_loss (:20-21) and ssim/_ssim (:35-58; window 11, sigma 1.5,
 * zero padding 5, size_average=True). forward writes out3 = {mean |img - gt|, mean ssim,
 * w_l1 * out3[0] + w_dssim * (1 - out3[1])} (device) and, when the three derivative planes (C,H,W each)
 * are given, what backward needs; `partial` is scratch of adgs_image_loss_partial_floats(C,H,W) floats.
 * backward takes the upstream gradients of the two means from DEVICE memory (0-d tensors; they may alias),
 * each multiplied by a host-side weight: dL/d l1 = grad_l1[0] * w_l1, dL/d ssim = grad_ssim[0] * w_ssim,
 * and writes dL/d img.
 * ---------------------------------------------------------------------------------------- */
ADGS_API size_t adgs_image_loss_partial_floats(int32_t C, int32_t H, int32_t W);
ADGS_API int adgs_image_loss_forward(int32_t C, int32_t H, int32_t W, const float* img, const float* gt,
                                     float* f_mu, float* f_e1, float* f_e12, float* partial, float w_l1,
                                     float w_dssim, float* out3, adgs_stream_t stream);
ADGS_API int adgs_image_loss_backward(int32_t C, int32_t H, int32_t W, const float* img, const float* gt,
                                      const float* f_mu, const float* f_e1, const float* f_e12,
                                      const float* grad_l1, float w_l1, const float* grad_ssim, float w_ssim,
                                      float* d_img, adgs_stream_t stream);

/* The per-pixel terms of train.py:82-100: lambda_depth * get_depth_loss(depth, gt_depth)
 * (utils/loss_utils.py:60-65, utils/depth_utils.py:3-45, mask = None) + lambda_obj * BCE(clip(img_semantic[0]),
 * gt_semantic > 0) + lambda_sky * BCE(1 - clip(img_opacity), gt_sky) (train.py:91-99) + lambda_flow *
 * get_flow_loss(img_flow, flow_pkg, img_opacity, dist) (utils/loss_utils.py:88-108, utils/flow_utils.py:5-10).
 * A term is skipped when its ground-truth pointer is null. Three grid-stride passes, no host synchronisation
 * (the reference's flow loss blocks on torch.nonzero): `phases` bit 0 = the two reduction passes (results stay
 * in `scratch`), bit 1 = the plane pass; the forward of an autograd function runs phases = 1 (scalars only,
 * planes null), its backward phases = 2 on the same scratch with the upstream gradient. Outputs: out6 = {depth, obj, sky, flow
 * losses, their lambda-weighted sum, number of selected flow pixels}; and, for every non-null plane, the
 * cotangent of the weighted sum times grad_total[0] (device scalar; null = 1): d_depth (H,W), d_semantic
 * (H,W), d_opacity (H,W; sky + flow-weight paths), d_flow (3,H,W). Planes are (H,W) row-major, img_flow (3,H,W),
 * flow (2,H,W) target pixel coordinates; K, R, T are HOST values (row-major 3x3, 3x3, 3).
 * scratch: adgs_pixel_loss_scratch_bytes(H,W), 8-byte aligned. */
typedef struct adgs_pixel_loss_inputs {
    int32_t H, W;
    const float* depth;         /* rendered (inverse) depth */
    const float* gt_depth;
    const float* img_semantic;  /* channel 0 of the rendered object mask */
    const float* gt_semantic;
    const float* img_opacity;   /* sky term */
    const float* gt_sky;
    const float* img_flow;      /* rendered flow points (3,H,W) */
    const float* flow;          /* (2,H,W) */
    const float* flow_vis;      /* (H,W) */
    const float* flow_opacity;  /* (H,W) optional weight of the flow term (img_opacity), may be null */
    float K[9], R[9], T[3];
    float flow_dist;
    float lambda_depth, lambda_obj, lambda_sky, lambda_flow;
} adgs_pixel_loss_inputs;
ADGS_API size_t adgs_pixel_loss_scratch_bytes(int32_t H, int32_t W);
ADGS_API int adgs_pixel_loss(const adgs_pixel_loss_inputs* in, int32_t phases, char* scratch, const float* grad_total,
                             float* d_depth, float* d_semantic, float* d_opacity, float* d_flow, float* out6,
                             adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Environment map: replaces EnvironmentMap.get_image_background /
 * get_env_color (scene/env.py:44-76, rays :11-27, utils/graphics_utils.py:95-100), the composite
 * `foreground + (1 - img_opacity) * background` (gaussian_renderer/__init__.py:92-94) and
 * `env_map.optimizer.step()` (scene/env.py:78-83, train.py:165) for the (1,C,R,R) grid_map.
 *   forward : background (C,H,W) and / or rendered = foreground + (1 - img_opacity) * background
 *   backward: d img_opacity (H,W) written; texel gradients are ADDED into the persistent dense buffer
 *             env->grad (zero-initialised once by the caller) and the 32x32-texel tiles they fall into are
 *             marked in env->touched (adgs_env_touched_bytes(R) bytes, zero-initialised once);
 *             d foreground = g_rendered (no kernel needed)
 *   step    : Adam (torch.optim.Adam arithmetic, eps = 1e-15 in the reference) over the tiles ever touched,
 *             zeroing their gradient in the same pass -- identical to the dense step, since a texel that
 *             never received a gradient has zero moments and a zero update.
 * `world_view_transform` = the camera's (4,4) tensor on the DEVICE; focal = W / (2 tan(FoVx / 2)).
 * ---------------------------------------------------------------------------------------- */
typedef struct adgs_env_map {
    int32_t R;            /* resolution (scene/env.py:31) */
    int32_t C;            /* channels */
    float* grid;          /* (1,C,R,R) parameter */
    float* grad;          /* (1,C,R,R) persistent gradient buffer (backward, step) */
    float* exp_avg;       /* (1,C,R,R) (step) */
    float* exp_avg_sq;    /* (1,C,R,R) (step) */
    uint8_t* touched;     /* adgs_env_touched_bytes(R) (backward, step) */
    uint32_t* tile_list;  /* adgs_env_tile_list_bytes(R): scratch of the step (compacted touched tiles) */
} adgs_env_map;
ADGS_API size_t adgs_env_touched_bytes(int32_t R);
ADGS_API size_t adgs_env_tile_list_bytes(int32_t R);
ADGS_API int adgs_env_forward(const adgs_env_map* env, int32_t H, int32_t W, float focal,
                              const float* world_view_transform, const float* foreground, const float* img_opacity,
                              float* background, float* rendered, adgs_stream_t stream);
ADGS_API int adgs_env_backward(const adgs_env_map* env, int32_t H, int32_t W, float focal,
                               const float* world_view_transform, const float* img_opacity, const float* g_rendered,
                               const float* g_background, float* d_opacity, adgs_stream_t stream);
ADGS_API int adgs_env_adam_step(const adgs_env_map* env, double lr, double beta1, double beta2, double eps,
                                int64_t step, adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): cumulative number of kernels launched by the library, and
 * optional CUDA-event timing of each pipeline stage on the caller's stream.
 * adgs_profile_begin() arms it; adgs_profile_end() synchronises the recorded events and returns,
 * per stage, the summed milliseconds and the number of timed scopes.
 * ---------------------------------------------------------------------------------------- */
ADGS_API unsigned long long adgs_launch_count(void);

/* Device self-test of the blend kernels' packed (FP32x2) exponential against CUDA's expf() -- the function the
 * reference's renderCUDA calls (RZ/cuda_rasterizer/forward.cu:345, backward.cu:560) -- over every float bit pattern
 * of its domain [-87, 87] plus NaN. out (device, 2 x uint64): [0] mismatching bit patterns, [1] patterns checked. */
ADGS_API int adgs_selftest_exp_pair(unsigned long long* out, adgs_stream_t stream);
ADGS_API int adgs_profile_begin(void);
ADGS_API int adgs_profile_num_stages(void);
ADGS_API const char* adgs_profile_stage_name(int stage);
ADGS_API int adgs_profile_end(float* ms_per_stage, int32_t* scopes_per_stage);

/* ------------------------------------------------------------------------------------------
 * simple-knn: replaces SimpleKNN::knn / distCUDA2 (KNN/simple_knn.h:18, KNN/spatial.cu:16-26):
 * mean squared distance to the 3 nearest neighbours of every point.
 * ---------------------------------------------------------------------------------------- */
ADGS_API size_t adgs_knn_workspace_bytes(int32_t P);
ADGS_API int adgs_dist_cuda2(int32_t P, const float* points, float* mean_dist2, char* workspace,
                    adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Densification / pruning over the planar storage of adgs_model.
 * Replaces GaussianModel.densify_and_prune / densify_and_clone / densify_and_split / prune_points /
 * _prune_optimizer / cat_tensors_to_optimizer / densification_postfix / add_densification_stats /
 * reset_opacity (scene/gaussian_model.py:463-467, 547-861, 863-867) and the max_radii2D update of
 * train.py:151. Rows are ordered [scene ; object] like adgs_model; the result of clone -> split -> prune
 * is, per partition, [originals neither split nor pruned] ++ [clones] ++ [children copy 0] ++ ...
 * ++ [children copy n_split-1], each in source order -- the order the reference's torch.cat / mask
 * chain produces.
 *
 *   adgs_densify_stats     visible (radii > 0) rows: max_radii2D = max(max_radii2D, radii),
 *                          xyz_gradient_accum += |grad_means2D[:, :2]|, denom += 1. accum/denom or
 *                          max_radii2D may be null (skipped).
 *   adgs_densify_classify  per-row decisions + counts. host_totals8 (HOST memory, may be null = no
 *                          synchronisation) receives, per partition p in {scene, object}:
 *                          [4p+0] originals kept, [4p+1] clones, [4p+2] split sources whose children
 *                          survive, [4p+3] split sources selected (= rows of unit normals to draw / n_split).
 *                          New row count of partition p = [4p+0] + [4p+1] + n_split * [4p+2].
 *                          mode ADGS_DENSIFY_PRUNE_ONLY: keep the rows whose prune_mask byte is 0
 *                          (prune_points with explicit masks); the other inputs may be null.
 *   adgs_densify_plan      src[d] = source row of output row d; tag[d] = kind | (sample row << 2).
 *   adgs_densify_gather    ONE launch over up to ADGS_GATHER_MAX_SEGMENTS arrays: dst[plane][r][0..width)
 *                          = src[plane][src[dst_row0 + r] - src_row0][0..width), or zeros when zero_new is
 *                          set and the row is a clone / child (Adam moments of new points,
 *                          cat_tensors_to_optimizer).
 *   adgs_densify_split     rows tagged CHILD: new_xyz = R(q/|q|) (z * exp(scaling)) + xyz,
 *                          new_scaling = log(exp(scaling) / (0.8 n_split)); z_scene / z_obj are the unit
 *                          normals (n_split * selected, 3), i.e. torch.normal(0, stds) = randn * stds.
 *   adgs_reset_opacity     opacity = inverse_sigmoid(min(sigmoid(opacity), cap)); moments zeroed.
 * Thresholds are floats: the reference compares float tensors with python scalars, which torch
 * evaluates in float.
 * ---------------------------------------------------------------------------------------- */
enum { ADGS_DENSIFY_AND_PRUNE = 0, ADGS_DENSIFY_PRUNE_ONLY = 1 };
enum { ADGS_DENSIFY_KIND_KEEP = 0, ADGS_DENSIFY_KIND_CLONE = 1, ADGS_DENSIFY_KIND_CHILD = 2 };
typedef struct adgs_densify_params {
    int32_t N_scene, N_obj;
    int32_t mode;            /* ADGS_DENSIFY_AND_PRUNE | ADGS_DENSIFY_PRUNE_ONLY */
    int32_t n_split;         /* N of densify_and_split (2) */
    float max_scene_grad;    /* densify_scene_grad_threshold */
    float max_obj_grad;      /* densify_obj_grad_threshold */
    float scene_split_size;  /* scene_extent * percent_dense */
    float obj_split_size;    /* object_extent * percent_dense */
    float min_opacity;       /* 0.005 */
    int32_t prune_big;       /* prune_big_points */
    float scene_big_size;    /* scene_extent * 0.05 */
    float obj_big_size;      /* object_extent * 0.1 */
    float inv_split_scale;   /* filled by the library: 1 / (float)(0.8 n_split) */
    int32_t _pad;
} adgs_densify_params;

ADGS_API int adgs_densify_stats(int32_t N, const float* grad_means2D, const int32_t* radii,
                                float* xyz_gradient_accum, float* denom, float* max_radii2D, adgs_stream_t stream);
ADGS_API size_t adgs_densify_workspace_bytes(int32_t N_scene, int32_t N_obj);
ADGS_API int adgs_densify_classify(const adgs_densify_params* params, const float* xyz_gradient_accum,
                                   const float* denom, const float* scaling, const float* opacity,
                                   const uint8_t* prune_mask, char* workspace, int32_t* host_totals8,
                                   adgs_stream_t stream);
ADGS_API int adgs_densify_plan(const adgs_densify_params* params, const char* workspace, const int32_t* host_totals8,
                               int32_t* src, int32_t* tag, adgs_stream_t stream);

#define ADGS_GATHER_MAX_SEGMENTS 40
typedef struct adgs_gather_segment {
    const float* src;  /* (planes, src_rows, width) */
    float* dst;        /* (planes, dst_rows, width) */
    int32_t planes;
    int32_t width;     /* 1..4 floats per row and plane */
    int32_t src_rows;
    int32_t dst_rows;
    int32_t src_row0;  /* global index of this array's first row: 0, or the OLD N_scene for object-only arrays */
    int32_t dst_row0;  /* 0, or the NEW N_scene */
    int32_t zero_new;  /* 1: rows that are clones / children are written as zeros */
    int32_t _pad;
} adgs_gather_segment;
ADGS_API int adgs_densify_gather(const adgs_gather_segment* segments, int32_t num_segments, const int32_t* src,
                                 const int32_t* tag, adgs_stream_t stream);
ADGS_API int adgs_densify_split(int32_t n_dst, int32_t n_scene_dst, const int32_t* src, const int32_t* tag,
                                const float* xyz, const float* scaling, const float* rotation, const float* z_scene,
                                const float* z_obj, int32_t n_split, float* new_xyz, float* new_scaling,
                                adgs_stream_t stream);
ADGS_API int adgs_reset_opacity(int32_t N, float cap, float* opacity, float* exp_avg, float* exp_avg_sq,
                                adgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K nearest neighbours: replaces pytorch3d.ops.knn_points(anchor[None], xyz[None], K).idx of
 * GaussianModel.set_obj_near_idx (scene/gaussian_model.py:825-833). anchors (A,D), points (P,D),
 * D = 3 or 4, K <= 32 and K <= P. idx (A,K) int64 ascending by squared distance, ties by smaller index;
 * dists (A,K) squared distances, may be null.
 * ---------------------------------------------------------------------------------------- */
ADGS_API size_t adgs_knn_points_workspace_bytes(int32_t A, int32_t P, int32_t K);
ADGS_API int adgs_knn_points(int32_t A, int32_t P, int32_t D, int32_t K, const float* anchors, const float* points,
                             int64_t* idx, float* dists, char* workspace, adgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADGS_B200_H_INCLUDED */
