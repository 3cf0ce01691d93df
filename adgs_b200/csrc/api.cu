// C-ABI entry points of libadgs_b200.so: strict drop-in rasterizer (forward / backward /
// mark_visible), arena sizing + introspection, stand-alone sort. See include/adgs_b200.h.
#include <cstdio>
#include <cstring>
#include "api_internal.cuh"

namespace adgs {

static thread_local char g_cuda_error[256] = "";

int record_cuda_error(cudaError_t e, const char* where)
{
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s", where, cudaGetErrorString(e));
    return ADGS_ERR_CUDA;
}

int check_stage(const char* where, bool debug, cudaStream_t stream)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return record_cuda_error(e, where);
    if (debug) {
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return record_cuda_error(e, where);
    }
    return ADGS_OK;
}

RasterParams make_raster_params(const adgs_camera* cam)
{
    RasterParams rp;
    rp.W = cam->image_width;
    rp.H = cam->image_height;
    rp.grid_x = (rp.W + ADGS_BLOCK_X - 1) / ADGS_BLOCK_X;
    rp.grid_y = (rp.H + ADGS_BLOCK_Y - 1) / ADGS_BLOCK_Y;
    rp.tan_fovx = cam->tanfovx;
    rp.tan_fovy = cam->tanfovy;
    rp.focal_y = rp.H / (2.0f * cam->tanfovy);
    rp.focal_x = rp.W / (2.0f * cam->tanfovx);
    rp.scale_modifier = cam->scale_modifier;
    rp.sh_degree = cam->sh_degree;
    rp.inv_depth = cam->inv_depth;
    rp.prefiltered = cam->prefiltered;
    return rp;
}

// Binning + blend, shared by the drop-in path and the fused path. Expects the per-Gaussian
// state (radii, depth keys, tiles_touched, records) to be in place.
int bin_and_blend(const adgs_camera* cam, int P, int D_S, bool has_flow, const float* semantic,
                  const adgs_images* out, const int32_t* radii, GeometryState& gs, char* binning,
                  adgs_alloc_fn binning_alloc, void* alloc_user, int64_t capacity, ImageState& is,
                  bool sync_for_count, int* num_rendered, cudaStream_t stream, int stages, const float* mean_x,
                  const float* mean_y)
{
    const bool debug = cam->debug != 0;
    const RasterParams rp = make_raster_params(cam);
    const int num_tiles = rp.grid_x * rp.grid_y;
    int st;

    const bool do_bin = (stages & 1) != 0, do_blend = (stages & 2) != 0;
    // The onesweep look-back words carry a 30-bit count and the offset scan is 32-bit (sort.cu): the reference's
    // 64-bit-key sort works up to 2^32 instances, this path to 2^30 - 1 Gaussians / instances. Refuse, never wrap.
    if (P >= kMaxSortItems || capacity >= kMaxSortItems) return ADGS_ERR_UNSUPPORTED;
    // (1) Gaussians by (depth bits, id): 4 onesweep passes over 8 B/Gaussian.
    if (do_bin) {
        StageScope sc(kStageDepthSort, stream);
        sort_pairs_async(gs.depth_keys, gs.depth_keys_alt, gs.order_a, gs.order_b, (size_t)P, nullptr, 0, 32, gs.sort,
                         /*iota*/ true, /*clear*/ true, stream);
    }
    if ((st = check_stage("depth sort", debug, stream))) return st;

    // (2) offsets of every Gaussian's tile instances, in depth order; total -> counters[0]
    if (do_bin) {
        StageScope sc(kStageScan, stream);
        inclusive_scan_gather_async(gs.tiles_touched, gs.depth_order, gs.point_offsets, (size_t)P, gs.scan_status,
                                    gs.counters, stream);
    }
    if ((st = check_stage("scan", debug, stream))) return st;

    if (sync_for_count && do_bin) {
        uint32_t R = 0;
        cudaError_t e = cudaMemcpyAsync(&R, gs.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return record_cuda_error(e, "num_rendered readback");
        capacity = (int64_t)R;
        if (capacity >= kMaxSortItems) return ADGS_ERR_UNSUPPORTED;
        if (num_rendered) *num_rendered = (int)R;
        binning = binning_alloc(adgs_binning_bytes(capacity), alloc_user);
        if (!binning) return ADGS_ERR_ALLOC;
    }
    BinningState bs = BinningState::from_chunk(binning, (size_t)capacity);

    if (do_bin) cudaMemsetAsync(is.ranges, 0, (size_t)num_tiles * 2 * sizeof(uint32_t), stream);
    const uint32_t* point_list = sorted_point_list(bs, num_tiles);
    if (capacity > 0 && do_bin) {
        // (3) instances in depth order, (4) stable sort by tile id only
        {
            StageScope sc(kStageEmit, stream);
            launch_emit(P, gs.depth_order, gs.point_offsets, gs.tiles_touched,
                        reinterpret_cast<const float4*>(gs.record), radii, rp.grid_x, rp.grid_y, bs.keys_a, bs.vals_a,
                        (uint32_t)capacity, gs.counters, stream, mean_x, mean_y);
        }
        if ((st = check_stage("emit", debug, stream))) return st;
        int passes;
        {
            StageScope sc(kStageTileSort, stream);
            passes = sort_pairs_async(bs.keys_a, bs.keys_b, bs.vals_a, bs.vals_b, (size_t)capacity, gs.counters, 0,
                                      tile_id_bits(num_tiles), bs.sort, false, true, stream);
        }
        if ((st = check_stage("tile sort", debug, stream))) return st;
        const uint32_t* sorted_tiles = (passes & 1) ? bs.keys_b : bs.keys_a;
        point_list = (passes & 1) ? bs.vals_b : bs.vals_a;
        StageScope sc(kStageTileRanges, stream);
        launch_tile_ranges(sorted_tiles, gs.counters, (uint32_t)capacity, is.ranges, stream);
        if ((st = check_stage("tile ranges", debug, stream))) return st;
    }

    if (!do_blend) return ADGS_OK;
    BlendFwdArgs b;
    b.ranges = is.ranges;
    b.point_list = point_list;
    b.record = reinterpret_cast<const float4*>(gs.record);
    b.semantic = semantic;
    b.bg = cam->bg;
    b.W = rp.W;
    b.H = rp.H;
    b.D_S = D_S;
    b.n_contrib = is.n_contrib;
    b.out_color = out->color;
    b.out_depth = out->depth;
    b.out_opacity = out->opacity;
    b.out_flow = out->flow;
    b.out_semantic = out->semantic;
    b.counters = gs.counters;
    b.capacity = (uint32_t)capacity;
    b.cull_mask = bs.cull_mask;
    {
        StageScope sc(kStageBlendFwd, stream);
        launch_blend_forward(b, has_flow, stream);
    }
    return check_stage("blend forward", debug, stream);
}

const uint32_t* sorted_point_list(const BinningState& bs, int num_tiles)
{
    const int passes = sort_num_passes(0, tile_id_bits(num_tiles));
    return (passes & 1) ? bs.vals_b : bs.vals_a;
}

const uint32_t* sorted_tile_ids(const BinningState& bs, int num_tiles)
{
    const int passes = sort_num_passes(0, tile_id_bits(num_tiles));
    return (passes & 1) ? bs.keys_b : bs.keys_a;
}

static int validate(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out)
{
    if (!cam || !g || !out) return ADGS_ERR_ARG;
    if (g->P < 0 || cam->image_width <= 0 || cam->image_height <= 0) return ADGS_ERR_ARG;
    if (g->D_S < 0 || g->D_S > ADGS_MAX_SEMANTIC) return ADGS_ERR_UNSUPPORTED;
    if (!out->depth || !out->opacity) return ADGS_ERR_ARG;
    if (g->P > 0) {
        if (!g->means3D || !g->opacities) return ADGS_ERR_ARG;
        if (!g->cov3D_precomp && (!g->scales || !g->rotations)) return ADGS_ERR_ARG;
        if (!cam->viewmatrix || !cam->projmatrix || !cam->bg) return ADGS_ERR_ARG;
        if (g->shs && !g->colors_precomp && !cam->campos) return ADGS_ERR_ARG;
    }
    return ADGS_OK;
}

static void zero_images(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out, cudaStream_t stream)
{
    const size_t HW = (size_t)cam->image_width * cam->image_height;
    if (out->color) cudaMemsetAsync(out->color, 0, 3 * HW * 4, stream);
    if (out->depth) cudaMemsetAsync(out->depth, 0, HW * 4, stream);
    if (out->opacity) cudaMemsetAsync(out->opacity, 0, HW * 4, stream);
    if (out->flow) cudaMemsetAsync(out->flow, 0, 3 * HW * 4, stream);
    if (out->semantic && g->D_S) cudaMemsetAsync(out->semantic, 0, (size_t)g->D_S * HW * 4, stream);
}

static int forward_common(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out, char* geometry,
                          char* binning, adgs_alloc_fn binning_alloc, void* alloc_user, int64_t capacity, char* image,
                          bool sync_for_count, int* num_rendered, cudaStream_t stream)
{
    const int P = g->P;
    GeometryState gs = GeometryState::from_chunk(geometry, (size_t)P);
    ImageState is = ImageState::from_chunk(image, cam->image_width, cam->image_height);
    cudaMemsetAsync(gs.counters, 0, 32 * sizeof(uint32_t), stream);

    int32_t* radii = out->radii ? out->radii : gs.radii;
    PreprocessArgs a;
    a.P = P;
    a.M = g->shs ? g->M : 0;
    a.D_S = g->semantic ? g->D_S : 0;
    a.rp = make_raster_params(cam);
    a.means3D = g->means3D;
    a.scales = g->scales;
    a.rotations = g->rotations;
    a.opacities = g->opacities;
    a.shs = g->shs;
    a.colors_precomp = g->colors_precomp;
    a.cov3D_precomp = g->cov3D_precomp;
    a.flow_points = g->flow_points;
    a.semantic = g->semantic;
    a.view = cam->viewmatrix;
    a.proj = cam->projmatrix;
    a.campos = cam->campos;
    a.radii = radii;
    a.depth_keys = gs.depth_keys;
    a.tiles_touched = gs.tiles_touched;
    a.record = reinterpret_cast<float4*>(gs.record);
    a.cov3D = gs.cov3D;
    a.clamped = gs.clamped;
    {
        StageScope sc(kStagePerGaussianFwd, stream);
        launch_preprocess(a, stream);
    }
    int st = check_stage("preprocess", cam->debug != 0, stream);
    if (st) return st;
    return bin_and_blend(cam, P, a.D_S, g->flow_points != nullptr, g->semantic, out, radii, gs, binning,
                         binning_alloc, alloc_user, capacity, is, sync_for_count, num_rendered, stream, 3, nullptr,
                         nullptr);
}

}  // namespace adgs

using namespace adgs;

extern "C" {

int adgs_abi_version(void)
{
    return ADGS_ABI_VERSION;
}

const char* adgs_status_string(int status)
{
    switch (status) {
        case ADGS_OK: return "ok";
        case ADGS_ERR_ARG: return "invalid argument";
        case ADGS_ERR_CUDA: return "CUDA error";
        case ADGS_ERR_CAPACITY: return "binning arena capacity exceeded";
        case ADGS_ERR_UNSUPPORTED: return "unsupported configuration";
        case ADGS_ERR_ALLOC: return "arena allocator returned null";
        default: return status > 0 ? "ok" : "unknown error";
    }
}

const char* adgs_last_cuda_error(void)
{
    return g_cuda_error;
}

size_t adgs_geometry_bytes(int32_t P)
{
    char* c = nullptr;
    GeometryState::from_chunk(c, (size_t)(P > 0 ? P : 0));
    return (size_t)c + 128;
}

size_t adgs_binning_bytes(int64_t R)
{
    char* c = nullptr;
    BinningState::from_chunk(c, (size_t)(R > 0 ? R : 0));
    return (size_t)c + 128;
}

size_t adgs_image_bytes(int32_t width, int32_t height)
{
    char* c = nullptr;
    ImageState::from_chunk(c, (size_t)width, (size_t)height);
    return (size_t)c + 128;
}

size_t adgs_backward_scratch_bytes(int32_t P)
{
    return (size_t)(P > 0 ? P : 0) * ADGS_GRAD_FLOATS * sizeof(float) + 256;
}

int adgs_geometry_offsets(int32_t P, adgs_geometry_layout* o)
{
    if (!o || P < 0) return ADGS_ERR_ARG;
    char* c = nullptr;
    GeometryState g = GeometryState::from_chunk(c, (size_t)P);
    o->counters = (size_t)g.counters;
    o->depths = (size_t)g.depth_keys;  // sorted float bits after the forward
    o->tiles_touched = (size_t)g.tiles_touched;
    o->record = (size_t)g.record;
    o->cov3D = (size_t)g.cov3D;
    o->clamped = (size_t)g.clamped;
    o->depth_order = (size_t)g.depth_order;
    o->point_offsets = (size_t)g.point_offsets;
    o->total = (size_t)c;
    return ADGS_OK;
}

int adgs_binning_offsets(int64_t R, adgs_binning_layout* o)
{
    // the sorted buffers depend on the pass parity, i.e. on the tile count: report both candidates
    if (!o || R < 0) return ADGS_ERR_ARG;
    char* c = nullptr;
    BinningState b = BinningState::from_chunk(c, (size_t)R);
    o->point_list = (size_t)b.vals_a;
    o->point_list_tile = (size_t)b.keys_a;
    o->point_list_alt = (size_t)b.vals_b;
    o->point_list_tile_alt = (size_t)b.keys_b;
    o->total = (size_t)c;
    return ADGS_OK;
}

int adgs_binning_result_in_alt(int32_t width, int32_t height)
{
    const int tiles = ((width + 15) / 16) * ((height + 15) / 16);
    return sort_num_passes(0, tile_id_bits(tiles)) & 1;
}

int adgs_image_offsets(int32_t width, int32_t height, adgs_image_layout* o)
{
    if (!o) return ADGS_ERR_ARG;
    char* c = nullptr;
    ImageState s = ImageState::from_chunk(c, (size_t)width, (size_t)height);
    o->ranges = (size_t)s.ranges;
    o->n_contrib = (size_t)s.n_contrib;
    o->total = (size_t)c;
    return ADGS_OK;
}

int adgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      uint8_t* present, adgs_stream_t stream)
{
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !projmatrix || !present))) return ADGS_ERR_ARG;
    launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream);
    return check_stage("mark_visible", false, (cudaStream_t)stream);
}

int adgs_rasterize_forward(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out,
                           adgs_alloc_fn geometry_alloc, adgs_alloc_fn binning_alloc, adgs_alloc_fn image_alloc,
                           void* alloc_user, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate(cam, g, out);
    if (st) return st;
    if (!geometry_alloc || !binning_alloc || !image_alloc) return ADGS_ERR_ARG;
    if (g->P == 0) {
        zero_images(cam, g, out, stream);
        return 0;
    }
    char* geometry = geometry_alloc(adgs_geometry_bytes(g->P), alloc_user);
    char* image = image_alloc(adgs_image_bytes(cam->image_width, cam->image_height), alloc_user);
    if (!geometry || !image) return ADGS_ERR_ALLOC;
    int R = 0;
    st = forward_common(cam, g, out, geometry, nullptr, binning_alloc, alloc_user, 0, image, true, &R, stream);
    return st ? st : R;
}

int adgs_rasterize_forward_async(const adgs_camera* cam, const adgs_gaussians* g, const adgs_images* out,
                                 char* geometry, char* binning, int64_t capacity, char* image,
                                 adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (capacity >= kMaxSortItems) return ADGS_ERR_UNSUPPORTED;
    int st = validate(cam, g, out);
    if (st) return st;
    if (g->P == 0) {
        zero_images(cam, g, out, stream);
        return 0;
    }
    if (!geometry || !binning || !image || capacity < 0) return ADGS_ERR_ARG;
    return forward_common(cam, g, out, geometry, binning, nullptr, nullptr, capacity, image, false, nullptr, stream);
}

int adgs_read_counters(const char* geometry, int32_t P, uint32_t* host_pinned_2, adgs_stream_t stream)
{
    if (!geometry || !host_pinned_2) return ADGS_ERR_ARG;
    char* c = const_cast<char*>(geometry);
    GeometryState gs = GeometryState::from_chunk(c, (size_t)P);
    cudaError_t e = cudaMemcpyAsync(host_pinned_2, gs.counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    return e == cudaSuccess ? ADGS_OK : record_cuda_error(e, "read_counters");
}

int adgs_rasterize_backward(const adgs_camera* cam, const adgs_gaussians* g, const int32_t* radii,
                            const char* geometry, int64_t R, const char* binning, const char* image,
                            const float* img_opacity, const adgs_image_grads* dpix,
                            const adgs_gaussian_grads* grads, char* scratch, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cam || !g || !dpix || !grads) return ADGS_ERR_ARG;
    if (g->D_S < 0 || g->D_S > ADGS_MAX_SEMANTIC) return ADGS_ERR_UNSUPPORTED;
    const int P = g->P;
    if (P == 0) return ADGS_OK;
    if (!geometry || !binning || !image || !img_opacity || !scratch) return ADGS_ERR_ARG;
    const bool debug = cam->debug != 0;
    const RasterParams rp = make_raster_params(cam);
    char* gc = const_cast<char*>(geometry);
    char* bc = const_cast<char*>(binning);
    char* ic = const_cast<char*>(image);
    GeometryState gs = GeometryState::from_chunk(gc, (size_t)P);
    BinningState bs = BinningState::from_chunk(bc, (size_t)R);
    ImageState is = ImageState::from_chunk(ic, rp.W, rp.H);
    if (!radii) radii = gs.radii;
    const int D_S = g->semantic ? g->D_S : 0;

    float* grad_record = nullptr;
    char* sc = scratch;
    carve(sc, grad_record, (size_t)P * ADGS_GRAD_FLOATS);
    cudaMemsetAsync(grad_record, 0, (size_t)P * ADGS_GRAD_FLOATS * sizeof(float), stream);
    if (D_S > 1 && grads->dL_dsemantic)
        cudaMemsetAsync(grads->dL_dsemantic, 0, (size_t)P * D_S * sizeof(float), stream);

    BlendBwdArgs b;
    b.ranges = is.ranges;
    b.point_list = sorted_point_list(bs, rp.grid_x * rp.grid_y);
    b.record = reinterpret_cast<const float4*>(gs.record);
    b.semantic = g->semantic;
    b.bg = cam->bg;
    b.W = rp.W;
    b.H = rp.H;
    b.D_S = D_S;
    b.n_contrib = is.n_contrib;
    b.img_opacity = img_opacity;
    b.dL_dcolor = (g->colors_precomp || g->shs) ? dpix->dL_dcolor : nullptr;
    b.dL_ddepth = dpix->dL_ddepth;
    b.dL_dflow = g->flow_points ? dpix->dL_dflow : nullptr;
    b.dL_dsemantic = g->semantic ? dpix->dL_dsemantic : nullptr;
    b.dL_dopacity = dpix->dL_dopacity;
    b.grad_record = grad_record;
    b.dL_dsemantic_g = grads->dL_dsemantic;
    b.counters = nullptr;  // exact binning: the arena was sized from num_rendered
    b.cull_mask = bs.cull_mask;
    if (D_S > 1 && !grads->dL_dsemantic) return ADGS_ERR_ARG;
    if (R > 0) {
        {
            StageScope sc(kStageBlendBwd, stream);
            launch_blend_backward(b, g->flow_points != nullptr, stream);
        }
        int st = check_stage("blend backward", debug, stream);
        if (st) return st;
    }

    PreprocessBwdArgs a;
    a.P = P;
    a.M = g->shs ? g->M : 0;
    a.D_S = D_S;
    a.rp = rp;
    a.means3D = g->means3D;
    a.scales = g->scales;
    a.rotations = g->rotations;
    a.shs = g->shs;
    a.colors_precomp = g->colors_precomp;
    a.cov3D_precomp = g->cov3D_precomp;
    a.view = cam->viewmatrix;
    a.proj = cam->projmatrix;
    a.campos = cam->campos;
    a.radii = radii;
    a.cov3D = gs.cov3D;
    a.clamped = gs.clamped;
    a.grad_record = grad_record;
    a.dL_dmeans2D = grads->dL_dmeans2D;
    a.dL_dcolors = grads->dL_dcolors;
    a.dL_dopacity = grads->dL_dopacity;
    a.dL_dmeans3D = grads->dL_dmeans3D;
    a.dL_dcov3D = grads->dL_dcov3D;
    a.dL_dsh = grads->dL_dsh;
    a.dL_dscales = grads->dL_dscales;
    a.dL_drotations = grads->dL_drotations;
    a.dL_dflow_points = grads->dL_dflow_points;
    a.dL_dsemantic = grads->dL_dsemantic;
    {
        StageScope sc(kStagePerGaussianBwd, stream);
        launch_preprocess_backward(a, stream);
    }
    return check_stage("preprocess backward", debug, stream);
}

size_t adgs_sort_workspace_bytes(int64_t n)
{
    char* c = nullptr;
    SortWorkspace::from_chunk(c, (size_t)(n > 0 ? n : 0));
    return (size_t)c + 128;
}

int adgs_sort_pairs(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                    int32_t begin_bit, int32_t end_bit, char* workspace, adgs_stream_t stream)
{
    if (n < 0 || begin_bit < 0 || end_bit > 32 || end_bit <= begin_bit) return ADGS_ERR_ARG;
    if (n >= kMaxSortItems) return ADGS_ERR_UNSUPPORTED;
    if (n == 0) return 1;
    if (!keys_in || !vals_in || !keys_out || !vals_out || !workspace) return ADGS_ERR_ARG;
    char* c = workspace;
    SortWorkspace ws = SortWorkspace::from_chunk(c, (size_t)n);
    const int passes = sort_pairs_async(keys_in, keys_out, vals_in, vals_out, (size_t)n, nullptr, begin_bit, end_bit,
                                        ws, false, true, (cudaStream_t)stream);
    if (passes < 0) return ADGS_ERR_ARG;
    int st = check_stage("sort_pairs", false, (cudaStream_t)stream);
    if (st) return st;
    return (passes & 1) ? 0 : 1;
}

}  // extern "C"
