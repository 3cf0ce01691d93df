// Shared device helpers for libadgs_b200 (sm_100a only).
//
// Numerical contract: radii, tile rectangles, depth keys and n_contrib must be BIT-EXACT against
// the reference kernels (RZ/cuda_rasterizer/forward.cu, auxiliary.h), so the small linear-algebra
// helpers here keep the summation order of the reference's column-major 3x3 products
// (entry = a0*b0 + a1*b1 + a2*b2, left to right) and are compiled with the same nvcc defaults
// (FMA contraction on, IEEE division / sqrt, no fast-math).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../include/adgs_b200.h"

#define ADGS_BLOCK_X 16
#define ADGS_BLOCK_Y 16
#define ADGS_BLOCK_SIZE 256
#define ADGS_REC_FLOATS 16 /* packed per-Gaussian blend record, 64 B */
#define ADGS_GRAD_FLOATS 16 /* packed per-Gaussian blend gradient record, 64 B */

// Record layout (floats), one 16-byte quad per consumer so that every field group is a single LDS.128:
//   q0 = 0 x, 1 y, 2 conic.x, 3 conic.y          (cull test + alpha evaluation)
//   q1 = 4 conic.z, 5 opacity, 6 depth, 7 half2 {hx, hy}: conservative half extents (pixels) of the region
//        where the splat can reach alpha >= 1/255 (NaN = nowhere, +inf = unknown / keep everywhere)
//   q2 = 8..10 rgb, 11 depth feature (depth or 1/(depth+1e-7))
//   q3 = 12..14 flow point, 15 semantic[0]
// Gradient record layout: 0,1 dmean2D  2,3,4 dconic(x,y,w)  5 dopacity  6..8 dcolor
// 9 ddepthfeat  10..12 dflow  13 dsemantic[0]

namespace adgs {

// Real spherical-harmonics constants (same values as RZ/cuda_rasterizer/auxiliary.h:22-39 and
// utils/sh_utils.py:26-54).
#define ADGS_SH_C0 0.28209479177387814f
#define ADGS_SH_C1 0.4886025119029199f
#define ADGS_SH_C2_0 1.0925484305920792f
#define ADGS_SH_C2_1 -1.0925484305920792f
#define ADGS_SH_C2_2 0.31539156525252005f
#define ADGS_SH_C2_3 -1.0925484305920792f
#define ADGS_SH_C2_4 0.5462742152960396f
#define ADGS_SH_C3_0 -0.5900435899266435f
#define ADGS_SH_C3_1 2.890611442640554f
#define ADGS_SH_C3_2 -0.4570457994644658f
#define ADGS_SH_C3_3 0.3731763325901154f
#define ADGS_SH_C3_4 -0.4570457994644658f
#define ADGS_SH_C3_5 1.445305721320277f
#define ADGS_SH_C3_6 -0.5900435899266435f

// Column-major 3x3 (c[col][row]), product with the reference's accumulation order.
struct Mat3 {
    float c[3][3];
};

// a0 b0 + a1 b1 + a2 b2 with the FMA contraction PINNED to what nvcc emits for the reference's expression
// (left product fused, middle product rounded on its own): fma(a2, b2, fma(a0, b0, a1 * b1)). Explicit intrinsics,
// because the compiler's choice of which product to fuse depends on the code around the expression, and the strict
// and the fused front ends must agree bit for bit with each other and with the reference kernels.
__device__ __forceinline__ float dot3_pinned(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

__device__ __forceinline__ Mat3 mat3_mul(const Mat3& a, const Mat3& b)
{
    Mat3 r;
#pragma unroll
    for (int col = 0; col < 3; ++col) {
#pragma unroll
        for (int row = 0; row < 3; ++row) {
            r.c[col][row] = dot3_pinned(a.c[0][row], b.c[col][0], a.c[1][row], b.c[col][1], a.c[2][row], b.c[col][2]);
        }
    }
    return r;
}

__device__ __forceinline__ Mat3 mat3_transpose(const Mat3& a)
{
    Mat3 r;
#pragma unroll
    for (int col = 0; col < 3; ++col)
#pragma unroll
        for (int row = 0; row < 3; ++row) r.c[col][row] = a.c[row][col];
    return r;
}

// Camera block shared by a CTA: view (16), proj (16), campos (3), bg (3).
struct CamSmem {
    float view[16];
    float proj[16];
    float campos[3];
    float bg[3];
};

__device__ __forceinline__ void load_camera(CamSmem& s, const float* view, const float* proj,
                                            const float* campos, const float* bg)
{
    int t = threadIdx.x + threadIdx.y * blockDim.x;
    if (t < 16) {
        s.view[t] = view[t];
        s.proj[t] = proj[t];
    } else if (t < 19) {
        s.campos[t - 16] = campos ? campos[t - 16] : 0.f;
    } else if (t < 22) {
        s.bg[t - 19] = bg ? bg[t - 19] : 0.f;
    }
    __syncthreads();
}

// ((v + 1) * S - 1) / 2 evaluated in double, as the reference does (auxiliary.h:41-44).
__device__ __forceinline__ float ndc_to_pix(float v, int S)
{
    return ((v + 1.0) * S - 1.0) * 0.5;
}

// Tile rectangle of a splat (auxiliary.h:46-56): C float->int truncation happens before the clamp.
__device__ __forceinline__ void tile_rect(float px, float py, int max_radius, int grid_x, int grid_y,
                                          uint32_t& x0, uint32_t& y0, uint32_t& x1, uint32_t& y1)
{
    x0 = min((unsigned)grid_x, (unsigned)max((int)0, (int)((px - max_radius) / ADGS_BLOCK_X)));
    y0 = min((unsigned)grid_y, (unsigned)max((int)0, (int)((py - max_radius) / ADGS_BLOCK_Y)));
    x1 = min((unsigned)grid_x, (unsigned)max((int)0, (int)((px + max_radius + ADGS_BLOCK_X - 1) / ADGS_BLOCK_X)));
    y1 = min((unsigned)grid_y, (unsigned)max((int)0, (int)((py + max_radius + ADGS_BLOCK_Y - 1) / ADGS_BLOCK_Y)));
}

__device__ __forceinline__ uint32_t lane_id()
{
    return threadIdx.x & 31;
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Conservative axis-aligned half extents of {d : alpha(d) >= 1/255} for a splat with conic (A, B, C) and
// opacity op: alpha >= 1/255 <=> -0.5 d'Qd >= -log(255 op) =: -t, an ellipse whose bounding box is
// hx = sqrt(2 t C / det), hy = sqrt(2 t A / det). Slack: t is raised by 0.02 and by the rounding error of the
// per-pixel evaluation (which grows with the anisotropy A C / det), det is lowered by its own rounding error,
// and the result is rounded UP to half precision. Packed as the bits of a half2 in a float slot.
__device__ __forceinline__ float pack_splat_extent(float A, float B, float C, float op)
{
    const uint32_t kNever = 0x7FFF7FFFu, kAlways = 0x7C007C00u;  // half NaN / +inf in both lanes
    if (!(op > 0.f)) return __uint_as_float(kNever);
    const float t0 = __logf(255.f * op) + 0.02f;
    if (!(t0 > 0.f)) return __uint_as_float(t0 != t0 ? kAlways : kNever);
    const float AC = A * C, BB = B * B;
    const float det = (AC - BB) - 4e-7f * (AC + BB);
    if (!(A > 0.f && C > 0.f && det > 0.f)) return __uint_as_float(kAlways);
    const float inv = 1.0f / det;
    const float t = 2.0f * t0 * (1.0f + 1e-5f * AC * inv);
    const float hx = sqrtf(t * C * inv) * 1.0001f + 0.01f, hy = sqrtf(t * A * inv) * 1.0001f + 0.01f;
    const __half2 h = __halves2half2(__float2half_ru(hx), __float2half_ru(hy));
    return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
}

// The 64-byte blend record of one splat (layout above) as four quads.
__device__ __forceinline__ void make_blend_record(float4* q, float px, float py, float conic_x, float conic_y,
                                                  float conic_z, float opacity, float depth, const float* rgb,
                                                  float depth_feature, float fx, float fy, float fz, float sem0)
{
    q[0] = make_float4(px, py, conic_x, conic_y);
    q[1] = make_float4(conic_z, opacity, depth, pack_splat_extent(conic_x, conic_y, conic_z, opacity));
    q[2] = make_float4(rgb[0], rgb[1], rgb[2], depth_feature);
    q[3] = make_float4(fx, fy, fz, sem0);
}

__device__ __forceinline__ void store_blend_record(float4* rec, float px, float py, float conic_x, float conic_y,
                                                   float conic_z, float opacity, float depth, const float* rgb,
                                                   float depth_feature, float fx, float fy, float fz, float sem0)
{
    float4 q[4];
    make_blend_record(q, px, py, conic_x, conic_y, conic_z, opacity, depth, rgb, depth_feature, fx, fy, fz, sem0);
#pragma unroll
    for (int i = 0; i < 4; ++i) rec[i] = q[i];
}

// Warp-cooperative, fully coalesced I/O of 32 consecutive 64-byte records (one per lane). A lane storing its own
// four quads issues 16-byte accesses 64 bytes apart: fine for local memory (L2 merges them), ruinous when the
// records live in ANOTHER GPU's memory -- every 16 bytes becomes its own NVLink transaction (measured: ~100 GB/s).
// Here the warp's 2 KB block goes through shared memory and each instruction moves 512 contiguous bytes.
// `stage`: 128 float4 of shared memory owned by the warp. Quads at or beyond `limit_quads` (the end of the array)
// are skipped. Quad (lane, i) sits at stage[(lane * 4 + i) ^ (lane >> 1)]: both phases are bank-conflict free enough.
__device__ __forceinline__ uint32_t record_stage_index(uint32_t lane, uint32_t i)
{
    return (lane * 4 + i) ^ ((lane >> 1) & 3);
}

__device__ __forceinline__ void warp_store_records(float4* warp_base, size_t limit_quads, const float4* q, float4* stage)
{
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) stage[record_stage_index(lane, i)] = q[i];
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t k = i * 32 + lane;  // quad k of the block = quad (k & 3) of lane k >> 2
        if (k < limit_quads) warp_base[k] = stage[record_stage_index(k >> 2, k & 3)];
    }
}

// load phase 1: the block's quads, coalesced, into registers (may be issued long before phase 2)
__device__ __forceinline__ void warp_load_records_issue(const float4* warp_base, size_t limit_quads, float4* r)
{
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t k = i * 32 + lane;
        r[i] = k < limit_quads ? warp_base[k] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// load phase 2: through shared memory to the owning lanes
__device__ __forceinline__ void warp_load_records_finish(const float4* r, float4* q, float4* stage)
{
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t k = i * 32 + lane;
        stage[record_stage_index(k >> 2, k & 3)] = r[i];
    }
    __syncwarp();
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) q[i] = stage[record_stage_index(lane, i)];
}

// fire-and-forget float add (RED.E.ADD.F32)
__device__ __forceinline__ void red_add_f32(float* p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace adgs

// ---- host-side arena carve-up (deterministic in P / R / W,H) -------------------------------
namespace adgs {

template <typename T>
static inline void carve(char*& chunk, T*& ptr, size_t count, size_t alignment = 128)
{
    size_t off = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(off);
    chunk = reinterpret_cast<char*>(ptr + count);
}

// radix sort constants
constexpr int kRadixBits = 8;
constexpr int kRadix = 256;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 pairs per CTA step
constexpr int kMaxPasses = 4;
constexpr long long kMaxSortItems = 1ll << 30;  // {flag, count} look-back words hold a 30-bit count

struct SortWorkspace {
    uint32_t* hist;     // [kMaxPasses][256] digit histograms
    uint32_t* tickets;  // [kMaxPasses] dynamic tile tickets
    uint32_t* status;   // [kMaxPasses][tiles][256] decoupled look-back words
    size_t tiles;
    size_t zero_bytes;  // bytes from hist to the end of status that must be zero before a sort
    static SortWorkspace from_chunk(char*& chunk, size_t n)
    {
        SortWorkspace w;
        w.tiles = (n + kSortTile - 1) / kSortTile;
        if (w.tiles == 0) w.tiles = 1;
        carve(chunk, w.hist, (size_t)kMaxPasses * kRadix);
        char* begin = reinterpret_cast<char*>(w.hist);
        carve(chunk, w.tickets, 32);
        carve(chunk, w.status, (size_t)kMaxPasses * w.tiles * kRadix);
        w.zero_bytes = (size_t)(chunk - begin);
        return w;
    }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct GeometryState {
    uint32_t* counters;       // [4] num_rendered, overflow
    uint32_t* depth_keys;     // [P] float bits of view-space z (0xFFFFFFFF if culled); sort ping
    uint32_t* tiles_touched;  // [P]
    float* record;            // [P][16]
    float* cov3D;             // [P][6]
    uint8_t* clamped;         // [P]
    uint32_t* depth_keys_alt; // [P] sort pong
    uint32_t* order_a;        // [P] sort ping (values); 4 passes => the sorted order ends here
    uint32_t* order_b;        // [P] sort pong
    uint32_t* depth_order;    // == order_a (even number of passes)
    uint32_t* point_offsets;  // [P]
    int32_t* radii;           // [P] internal radii (used when the caller passes no radii tensor)
    uint32_t* scan_status;    // [scan tiles + 1]
    SortWorkspace sort;
    static GeometryState from_chunk(char*& chunk, size_t P)
    {
        GeometryState g;
        carve(chunk, g.counters, 32);
        carve(chunk, g.depth_keys, P);
        carve(chunk, g.tiles_touched, P);
        carve(chunk, g.record, P * ADGS_REC_FLOATS);
        carve(chunk, g.cov3D, P * 6);
        carve(chunk, g.clamped, P);
        carve(chunk, g.depth_keys_alt, P);
        carve(chunk, g.order_a, P);
        carve(chunk, g.order_b, P);
        g.depth_order = g.order_a;
        carve(chunk, g.point_offsets, P);
        carve(chunk, g.radii, P);
        carve(chunk, g.scan_status, (P + kScanTile - 1) / kScanTile + 32);
        g.sort = SortWorkspace::from_chunk(chunk, P);
        return g;
    }
};

struct BinningState {
    uint32_t* keys_a;  // [R] tile id per instance (unsorted, then ping-pong)
    uint32_t* keys_b;
    uint32_t* vals_a;  // [R] gaussian id per instance
    uint32_t* vals_b;
    uint8_t* cull_mask;  // [R] per sorted instance: bit s = sub-tile s of its tile passed the forward's exact cull
    SortWorkspace sort;
    static BinningState from_chunk(char*& chunk, size_t R)
    {
        BinningState b;
        if (R == 0) R = 1;
        carve(chunk, b.keys_a, R);
        carve(chunk, b.keys_b, R);
        carve(chunk, b.vals_a, R);
        carve(chunk, b.vals_b, R);
        carve(chunk, b.cull_mask, R + 256);
        b.sort = SortWorkspace::from_chunk(chunk, R);
        return b;
    }
};

struct ImageState {
    uint32_t* ranges;     // [tiles][2]
    uint32_t* n_contrib;  // [H*W]
    static ImageState from_chunk(char*& chunk, size_t W, size_t H)
    {
        ImageState s;
        size_t tiles = ((W + 15) / 16) * ((H + 15) / 16);
        carve(chunk, s.ranges, tiles * 2);
        carve(chunk, s.n_contrib, W * H);
        return s;
    }
};

}  // namespace adgs
