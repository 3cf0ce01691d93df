// Per-Gaussian front end (strict drop-in, AoS inputs), binning and per-Gaussian backward.
#pragma once
#include "common.cuh"
#include "gaussian_math.cuh"

namespace adgs {

struct PreprocessArgs {
    int P, M, D_S;
    RasterParams rp;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* colors_precomp;
    const float* cov3D_precomp;
    const float* flow_points;
    const float* semantic;
    const float* view;
    const float* proj;
    const float* campos;
    int32_t* radii;
    uint32_t* depth_keys;
    uint32_t* tiles_touched;
    float4* record;
    float* cov3D;
    uint8_t* clamped;
};

struct PreprocessBwdArgs {
    int P, M, D_S;
    RasterParams rp;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* shs;
    const float* colors_precomp;
    const float* cov3D_precomp;
    const float* view;
    const float* proj;
    const float* campos;
    const int32_t* radii;
    const float* cov3D;       // geometry state (when not precomputed)
    const uint8_t* clamped;
    const float* grad_record; // [P][16]
    float* dL_dmeans2D;
    float* dL_dcolors;
    float* dL_dopacity;
    float* dL_dmeans3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscales;
    float* dL_drotations;
    float* dL_dflow_points;
    float* dL_dsemantic;  // D_S == 1 only; D_S > 1 is accumulated by the blend kernel directly
};

void launch_preprocess(const PreprocessArgs& a, cudaStream_t stream);
void launch_preprocess_backward(const PreprocessBwdArgs& a, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                         cudaStream_t stream);

// Binning: instances are emitted in (depth, id) order -- one stable 32-bit sort of the Gaussians --
// and then stably sorted by tile id only, which yields exactly the (tile, depth, id) order of the
// reference's 64-bit key sort (rasterizer_impl.cu:70-111, 310-315) at a fraction of the traffic.
void launch_emit(int P, const uint32_t* depth_order, const uint32_t* point_offsets, const uint32_t* tiles_touched,
                 const float4* record, const int32_t* radii, int grid_x, int grid_y, uint32_t* keys,
                 uint32_t* vals, uint32_t capacity, uint32_t* counters, cudaStream_t stream,
                 const float* mean_x = nullptr, const float* mean_y = nullptr);
void launch_tile_ranges(const uint32_t* sorted_tiles, const uint32_t* counters, uint32_t capacity,
                        uint32_t* ranges, cudaStream_t stream);

inline int tile_id_bits(int num_tiles)
{
    int bits = 1;
    while ((1 << bits) < num_tiles) ++bits;
    return bits;
}

}  // namespace adgs
