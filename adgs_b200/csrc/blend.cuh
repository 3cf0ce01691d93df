// Tile blend (forward + backward) shared pieces.
#pragma once
#include "common.cuh"

namespace adgs {

struct BlendFwdArgs {
    const uint32_t* ranges;      // [tiles][2]
    const uint32_t* point_list;  // [R]
    const float4* record;        // [P][4] packed blend records
    const float* semantic;       // (P,D_S) only for D_S > 1
    const float* bg;             // device (3)
    int W, H, D_S;
    uint32_t* n_contrib;
    float* out_color;
    float* out_depth;
    float* out_opacity;
    float* out_flow;
    float* out_semantic;
    const uint32_t* counters;    // [0]=num_rendered [1]=overflow (range clamp in async mode)
    uint32_t capacity;
    uint8_t* cull_mask;          // [R] out: which of the tile's eight 8x4 sub-tiles each instance can reach
};

struct BlendBwdArgs {
    const uint32_t* ranges;
    const uint32_t* point_list;
    const float4* record;
    const float* semantic;  // (P,D_S) for D_S > 1
    const float* bg;
    int W, H, D_S;
    const uint32_t* n_contrib;
    const float* img_opacity;
    const float* dL_dcolor;
    const float* dL_ddepth;
    const float* dL_dflow;
    const float* dL_dsemantic;
    const float* dL_dopacity;
    float* grad_record;    // [P][16] zero-initialised accumulation target
    float* dL_dsemantic_g; // (P,D_S) for D_S > 1 (zero-initialised)
    const uint8_t* cull_mask;  // [R] the forward's per-instance sub-tile masks: the backward culls by table look-up
    const uint32_t* counters;  // [1] = binning overflow of the forward (sync-free mode): leave the
                               // gradient record zero instead of reading images that were never written
};

// Conservative test: can ANY pixel centre of the rectangle [X0,X1]x[Y0,Y1] receive
// alpha >= 1/255 from this splat? thresh = -log(255*opacity): power(d) must reach it.
// power(d) = -0.5*(A dx^2 + C dy^2) - B dx dy is concave for a positive-definite conic, so its
// maximum over the rectangle is 0 if the mean lies inside, and otherwise sits on an edge that
// FACES the mean (the segment from the mean to any better interior point would cross such an
// edge inside the same super-level set). At most two edges face the mean; on each the maximiser
// is the clamped vertex of a 1-D parabola. Non-definite conics (numerical garbage) are kept.
__device__ __forceinline__ float power_at(float A, float B, float C, float dx, float dy)
{
    return -0.5f * (A * dx * dx + C * dy * dy) - B * dx * dy;
}

__device__ __forceinline__ bool splat_may_touch_rect(float mx, float my, float A, float B, float C, float thresh,
                                                     float X0, float Y0, float X1, float Y1)
{
    if (!(thresh <= 0.01f)) return false;  // opacity < 1/255 (or non-positive): alpha can never pass
    if (!(A > 0.f && C > 0.f && A * C - B * B > 0.f)) return true;
    // d = mean - pixel; nearest rectangle coordinates to the mean
    const float dx0 = mx - fminf(fmaxf(mx, X0), X1);  // 0 when the mean is inside the x-range
    const float dy0 = my - fminf(fmaxf(my, Y0), Y1);
    if (dx0 == 0.f && dy0 == 0.f) return true;
    const float dx_lo = mx - X1, dx_hi = mx - X0;
    const float dy_lo = my - Y1, dy_hi = my - Y0;
    float best = -3.0e38f;
    if (dx0 != 0.f) {  // vertical edge facing the mean: dx fixed, dy free in [dy_lo, dy_hi]
        const float v = fminf(fmaxf(__fdividef(-B * dx0, C), dy_lo), dy_hi);
        best = power_at(A, B, C, dx0, v);
    }
    if (dy0 != 0.f) {  // horizontal edge facing the mean
        const float v = fminf(fmaxf(__fdividef(-B * dy0, A), dx_lo), dx_hi);
        best = fmaxf(best, power_at(A, B, C, v, dy0));
    }
    // rounding slack: covers the evaluation error of this test and of the per-pixel formula
    const float dxm = fmaxf(fabsf(dx_lo), fabsf(dx_hi)), dym = fmaxf(fabsf(dy_lo), fabsf(dy_hi));
    const float eps = 0.01f + 2e-6f * (A * dxm * dxm + C * dym * dym);
    return !(best < thresh - eps);  // NaN compares false -> keep
}

// Level-1 cull: does the splat's bounding box (packed half extents, pack_splat_extent) reach the rectangle of
// pixel centres with centre (cx, cy) and half size (wx, wy)? NaN extents never hit, +inf always.
__device__ __forceinline__ bool splat_bbox_hits_rect(float mx, float my, float packed_extent, float cx, float cy,
                                                     float wx, float wy)
{
    const uint32_t bits = __float_as_uint(packed_extent);
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&bits));
    return fabsf(mx - cx) <= h.x + wx && fabsf(my - cy) <= h.y + wy;
}

// cull threshold of the exact test from the opacity: -log(255 op), 1e30 if op <= 0
__device__ __forceinline__ float splat_cull_threshold(float op)
{
    return op > 0.f ? -__logf(255.f * op) : 1e30f;
}

void launch_blend_forward(const BlendFwdArgs& a, bool has_flow, cudaStream_t stream);
void launch_blend_backward(const BlendBwdArgs& a, bool has_flow, cudaStream_t stream);

}  // namespace adgs
