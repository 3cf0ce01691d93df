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
};

// Conservative test: can ANY pixel centre of the rectangle [X0,X1]x[Y0,Y1] receive
// alpha >= 1/255 from this splat? thresh = -log(255*opacity) (power must reach it).
// The quadratic power(d) = -0.5*(A dx^2 + C dy^2) - B dx dy attains its maximum over the
// rectangle either at d = 0 (inside) or on the boundary; each edge is a 1-D quadratic.
__device__ __forceinline__ float power_at(float A, float B, float C, float dx, float dy)
{
    return -0.5f * (A * dx * dx + C * dy * dy) - B * dx * dy;
}

__device__ __forceinline__ bool splat_may_touch_rect(float mx, float my, float A, float B, float C, float thresh,
                                                     float X0, float Y0, float X1, float Y1)
{
    if (!(thresh <= 0.01f)) return false;  // opacity < 1/255 (or non-positive): alpha can never pass
    // d = mean - pixel
    const float dx_lo = mx - X1, dx_hi = mx - X0;
    const float dy_lo = my - Y1, dy_hi = my - Y0;
    if (dx_lo <= 0.f && dx_hi >= 0.f && dy_lo <= 0.f && dy_hi >= 0.f) return true;
    float best = power_at(A, B, C, dx_lo, dy_lo);
    best = fmaxf(best, power_at(A, B, C, dx_lo, dy_hi));
    best = fmaxf(best, power_at(A, B, C, dx_hi, dy_lo));
    best = fmaxf(best, power_at(A, B, C, dx_hi, dy_hi));
    if (C > 0.f) {
        const float inv = 1.f / C;
        float v = fminf(fmaxf(-B * dx_lo * inv, dy_lo), dy_hi);
        best = fmaxf(best, power_at(A, B, C, dx_lo, v));
        v = fminf(fmaxf(-B * dx_hi * inv, dy_lo), dy_hi);
        best = fmaxf(best, power_at(A, B, C, dx_hi, v));
    }
    if (A > 0.f) {
        const float inv = 1.f / A;
        float v = fminf(fmaxf(-B * dy_lo * inv, dx_lo), dx_hi);
        best = fmaxf(best, power_at(A, B, C, v, dy_lo));
        v = fminf(fmaxf(-B * dy_hi * inv, dx_lo), dx_hi);
        best = fmaxf(best, power_at(A, B, C, v, dy_hi));
    }
    // rounding slack: covers the evaluation error of both this test and the per-pixel formula
    const float dxm = fmaxf(fabsf(dx_lo), fabsf(dx_hi)), dym = fmaxf(fabsf(dy_lo), fabsf(dy_hi));
    const float mag = 0.5f * (fabsf(A) * dxm * dxm + fabsf(C) * dym * dym) + fabsf(B) * dxm * dym;
    const float eps = 0.01f + 2e-6f * mag;
    // NaNs compare false -> keep (never skip on garbage)
    return !(best < thresh - eps);
}

void launch_blend_forward(const BlendFwdArgs& a, bool has_flow, cudaStream_t stream);
void launch_blend_backward(const BlendBwdArgs& a, bool has_flow, cudaStream_t stream);

}  // namespace adgs
