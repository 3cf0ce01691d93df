// TEST INFRASTRUCTURE ONLY -- never linked into or imported by the product path.
//
// C-ABI shim over the UNMODIFIED reference rasterizer / simple-knn sources, which are
// compiled where they lie under /root/reference by oracle/Makefile into oracle/_ref/.
// The shim itself is ours; it only forwards raw pointers to the reference's public statics
//   CudaRasterizer::Rasterizer::{markVisible,forward,backward}
//     (submodules/depth-diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:24-101)
//   SimpleKNN::knn (submodules/simple-knn/simple_knn.h:18)
// and re-runs the reference's own arena carve-up (rasterizer_impl.h:22-68, fromChunk) so
// tests can look inside geomBuffer / binningBuffer / imgBuffer (tiles_touched, point_offsets,
// sorted keys, point_list, ranges, n_contrib) for the bit-exact comparisons.
#include <cstdint>
#include <cstddef>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer_impl.h"
#include "simple_knn.h"

typedef char* (*ref_alloc_fn)(size_t bytes, void* user);

extern "C" {

int ref_forward(
    ref_alloc_fn geom_alloc, ref_alloc_fn binning_alloc, ref_alloc_fn img_alloc, void* user,
    int P, int D, int M, int D_S,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* flow_points, const float* semantic, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_depth, float* img_opacity, float* img_flow,
    float* img_semantic, int inv_depth, int* radii, int debug)
{
    std::function<char*(size_t)> g = [=](size_t n) { return geom_alloc(n, user); };
    std::function<char*(size_t)> b = [=](size_t n) { return binning_alloc(n, user); };
    std::function<char*(size_t)> i = [=](size_t n) { return img_alloc(n, user); };
    try {
        return CudaRasterizer::Rasterizer::forward(
            g, b, i, P, D, M, D_S, background, width, height, means3D, shs, colors_precomp,
            flow_points, semantic, opacities, scales, scale_modifier, rotations, cov3D_precomp,
            viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered != 0,
            out_color, out_depth, img_opacity, img_flow, img_semantic, inv_depth != 0,
            radii, debug != 0);
    } catch (...) {
        return -1;
    }
}

int ref_backward(
    int P, int D, int M, int R, int D_S,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* flow_points, const float* semantic, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_flow,
    const float* dL_dpix_semantic,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, float* dL_dflow, float* dL_dsemantic,
    float* grad_img_opacity, float* img_opacity, int inv_depth, int debug)
{
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, R, D_S, background, width, height, means3D, shs, colors_precomp,
            flow_points, semantic, scales, scale_modifier, rotations, cov3D_precomp,
            viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii,
            geom_buffer, binning_buffer, image_buffer,
            dL_dpix, dL_dpix_depth, dL_dpix_flow, dL_dpix_semantic,
            dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D,
            dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dflow, dL_dsemantic,
            grad_img_opacity, img_opacity, inv_depth != 0, debug != 0);
        return 0;
    } catch (...) {
        return -1;
    }
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, unsigned char* present)
{
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
}

void ref_dist_cuda2(int P, float* points, float* mean_dists)
{
    SimpleKNN::knn(P, (float3*)points, mean_dists);
}

// Where the reference put things inside its three arenas (same carve-up it re-runs in backward,
// rasterizer_impl.cu:398-400). Offsets are in bytes from the arena base.
struct ref_geom_layout {
    size_t depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb, tiles_touched, point_offsets, total;
};
struct ref_binning_layout {
    size_t point_list, point_list_unsorted, point_list_keys, point_list_keys_unsorted, total;
};
struct ref_image_layout {
    size_t n_contrib, ranges, total;
};

void ref_geom_offsets(char* base, size_t P, ref_geom_layout* o)
{
    char* c = base;
    CudaRasterizer::GeometryState g = CudaRasterizer::GeometryState::fromChunk(c, P);
    o->depths = (char*)g.depths - base;
    o->clamped = (char*)g.clamped - base;
    o->internal_radii = (char*)g.internal_radii - base;
    o->means2D = (char*)g.means2D - base;
    o->cov3D = (char*)g.cov3D - base;
    o->conic_opacity = (char*)g.conic_opacity - base;
    o->rgb = (char*)g.rgb - base;
    o->tiles_touched = (char*)g.tiles_touched - base;
    o->point_offsets = (char*)g.point_offsets - base;
    o->total = c - base;
}

void ref_binning_offsets(char* base, size_t R, ref_binning_layout* o)
{
    char* c = base;
    CudaRasterizer::BinningState b = CudaRasterizer::BinningState::fromChunk(c, R);
    o->point_list = (char*)b.point_list - base;
    o->point_list_unsorted = (char*)b.point_list_unsorted - base;
    o->point_list_keys = (char*)b.point_list_keys - base;
    o->point_list_keys_unsorted = (char*)b.point_list_keys_unsorted - base;
    o->total = c - base;
}

void ref_image_offsets(char* base, size_t N, ref_image_layout* o)
{
    char* c = base;
    CudaRasterizer::ImageState im = CudaRasterizer::ImageState::fromChunk(c, N);
    o->n_contrib = (char*)im.n_contrib - base;
    o->ranges = (char*)im.ranges - base;
    o->total = c - base;
}

}  // extern "C"
