// Per-Gaussian kernels of the strict drop-in rasterizer path + binning kernels. See raster.cuh.
#include "raster.cuh"
#include "sort.cuh"

namespace adgs {
void count_launch(int n);
namespace {

__device__ __forceinline__ int effective_sh_degree(int deg, int M)
{
    int d = deg;
    while (d > 0 && (d + 1) * (d + 1) > M) --d;
    return d;
}

// Load the first `n` floats of an AoS SH row into registers (rest zero).
__device__ __forceinline__ void load_sh_row(const float* row, int M, int deg, float* sh)
{
#pragma unroll
    for (int i = 0; i < 48; ++i) sh[i] = 0.f;
    if (M == 16 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
        const int chunks = sh_chunks_for_degree(deg);
        const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            if (q < chunks) {
                const float4 v = __ldg(r4 + q);
                sh[4 * q + 0] = v.x;
                sh[4 * q + 1] = v.y;
                sh[4 * q + 2] = v.z;
                sh[4 * q + 3] = v.w;
            }
        }
    } else {
        const int n = 3 * (deg + 1) * (deg + 1);
#pragma unroll
        for (int i = 0; i < 48; ++i)
            if (i < n) sh[i] = __ldg(row + i);
    }
}

__global__ void __launch_bounds__(256) preprocess_aos_kernel(const PreprocessArgs a)
{
    __shared__ CamSmem cam;
    load_camera(cam, a.view, a.proj, a.campos, nullptr);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.P) return;

    const float3 p = make_float3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
    float scale[3] = {0.f, 0.f, 0.f}, rot[4] = {1.f, 0.f, 0.f, 0.f};
    const float* cov_pre = nullptr;
    if (a.cov3D_precomp) {
        cov_pre = a.cov3D_precomp + (size_t)idx * 6;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) scale[i] = a.scales[3 * idx + i];
        const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
        rot[0] = q.x;
        rot[1] = q.y;
        rot[2] = q.z;
        rot[3] = q.w;
    }

    SplatGeom g;
    const bool visible = splat_geometry(p, scale, rot, cov_pre, a.rp, cam.view, cam.proj, g);
    if (!visible) {
        a.radii[idx] = 0;
        a.tiles_touched[idx] = 0;
        a.depth_keys[idx] = 0xFFFFFFFFu;
        return;
    }

    float rgb[3] = {0.f, 0.f, 0.f};
    uint32_t clamped = 0;
    if (a.colors_precomp) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = a.colors_precomp[3 * idx + c];
    } else if (a.shs) {
        float sh[48];
        const int deg = effective_sh_degree(a.rp.sh_degree, a.M);
        load_sh_row(a.shs + (size_t)idx * a.M * 3, a.M, deg, sh);
        sh_to_rgb(deg, p, cam.campos, sh, rgb, clamped);
    }
    a.clamped[idx] = (uint8_t)clamped;

    if (!a.cov3D_precomp) {
        float2* c2 = reinterpret_cast<float2*>(a.cov3D + (size_t)idx * 6);
        c2[0] = make_float2(g.cov3D[0], g.cov3D[1]);
        c2[1] = make_float2(g.cov3D[2], g.cov3D[3]);
        c2[2] = make_float2(g.cov3D[4], g.cov3D[5]);
    }

    float fl[3] = {0.f, 0.f, 0.f};
    if (a.flow_points) {
#pragma unroll
        for (int c = 0; c < 3; ++c) fl[c] = a.flow_points[3 * idx + c];
    }
    const float sem0 = (a.D_S == 1 && a.semantic) ? a.semantic[idx] : 0.f;
    const float dfeat = a.rp.inv_depth ? (1.0f / (g.depth + 0.0000001f)) : g.depth;

    store_blend_record(a.record + (size_t)idx * 4, g.px, g.py, g.conic_x, g.conic_y, g.conic_z, a.opacities[idx], g.depth,
                       rgb, dfeat, fl[0], fl[1], fl[2], sem0);

    a.radii[idx] = g.radius;
    a.tiles_touched[idx] = g.tiles;
    a.depth_keys[idx] = __float_as_uint(g.depth);
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* means3D, const float* view,
                                                           const float* proj, uint8_t* present)
{
    __shared__ CamSmem cam;
    load_camera(cam, view, proj, nullptr, nullptr);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 pv = xform_point_4x3(p, cam.view);
    present[idx] = pv.z > 0.2f ? 1 : 0;
}

// One thread per depth-ordered Gaussian; splats that cover many tiles are written by the whole warp.
__global__ void __launch_bounds__(256) emit_kernel(int P, const uint32_t* __restrict__ depth_order,
                                                   const uint32_t* __restrict__ point_offsets,
                                                   const uint32_t* __restrict__ tiles_touched,
                                                   const float4* __restrict__ record,
                                                   const int32_t* __restrict__ radii, int grid_x, int grid_y,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                   uint32_t capacity, uint32_t* counters,
                                                   const float* __restrict__ mean_x, const float* __restrict__ mean_y)
{
    constexpr uint32_t kCoop = 32;  // splats with more tiles than this are emitted cooperatively
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t gid = 0, n = 0, off = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (i < P) {
        gid = depth_order[i];
        n = tiles_touched[gid];
        if (n) {
            const uint32_t end = point_offsets[i];
            off = end - n;
            if (end > capacity) {
                counters[1] = 1;  // overflow: binning arena too small
                n = 0;
            } else {
                float mx, my;
                if (mean_x) {  // splat exchange: the pixel means travel apart from the (larger) records
                    mx = mean_x[gid];
                    my = mean_y[gid];
                } else {
                    const float4 q0 = record[(size_t)gid * 4];
                    mx = q0.x;
                    my = q0.y;
                }
                tile_rect(mx, my, radii[gid], grid_x, grid_y, x0, y0, x1, y1);
            }
        }
    }
    if (n && n <= kCoop) {
        uint32_t o = off;
        for (uint32_t y = y0; y < y1; ++y)
            for (uint32_t x = x0; x < x1; ++x) {
                keys[o] = y * grid_x + x;
                vals[o] = gid;
                ++o;
            }
    }
    uint32_t big = __ballot_sync(0xffffffffu, n > kCoop);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t b_gid = __shfl_sync(0xffffffffu, gid, src);
        const uint32_t b_n = __shfl_sync(0xffffffffu, n, src);
        const uint32_t b_off = __shfl_sync(0xffffffffu, off, src);
        const uint32_t b_x0 = __shfl_sync(0xffffffffu, x0, src);
        const uint32_t b_y0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t b_w = __shfl_sync(0xffffffffu, x1 - x0, src);
        for (uint32_t k = lane; k < b_n; k += 32) {
            const uint32_t ty = b_y0 + k / b_w, tx = b_x0 + k % b_w;
            keys[b_off + k] = ty * grid_x + tx;
            vals[b_off + k] = b_gid;
        }
    }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ sorted_tiles,
                                                          const uint32_t* __restrict__ counters, uint32_t capacity,
                                                          uint32_t* __restrict__ ranges)
{
    // overflow (sync-free mode): the tail of the list was never emitted and holds arbitrary tile ids;
    // the blend kernels skip such a frame, so no range is needed -- and none may be derived from garbage
    if (counters[1]) return;
    const uint32_t L = min(counters[0], capacity);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < L; idx += stride) {
        const uint32_t cur = sorted_tiles[idx];
        if (idx == 0) {
            ranges[2 * cur] = 0;
        } else {
            const uint32_t prev = sorted_tiles[idx - 1];
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)idx;
                ranges[2 * cur] = (uint32_t)idx;
            }
        }
        if (idx == L - 1) ranges[2 * cur + 1] = L;
    }
}

__global__ void __launch_bounds__(256) preprocess_bwd_aos_kernel(const PreprocessBwdArgs a)
{
    __shared__ CamSmem cam;
    load_camera(cam, a.view, a.proj, a.campos, nullptr);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.P) return;

    const float4* gr = reinterpret_cast<const float4*>(a.grad_record) + (size_t)idx * 4;
    const float4 g0 = gr[0], g1 = gr[1], g2 = gr[2], g3 = gr[3];
    const bool visible = a.radii[idx] > 0;

    if (a.dL_dmeans2D) {
        a.dL_dmeans2D[3 * idx + 0] = g0.x;
        a.dL_dmeans2D[3 * idx + 1] = g0.y;
        a.dL_dmeans2D[3 * idx + 2] = 0.f;
    }
    if (a.dL_dcolors) {
        a.dL_dcolors[3 * idx + 0] = g1.z;
        a.dL_dcolors[3 * idx + 1] = g1.w;
        a.dL_dcolors[3 * idx + 2] = g2.x;
    }
    if (a.dL_dopacity) a.dL_dopacity[idx] = g1.y;
    if (a.dL_dflow_points) {
        a.dL_dflow_points[3 * idx + 0] = g2.z;
        a.dL_dflow_points[3 * idx + 1] = g2.w;
        a.dL_dflow_points[3 * idx + 2] = g3.x;
    }
    if (a.dL_dsemantic && a.D_S == 1) a.dL_dsemantic[idx] = g3.y;

    float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    float dsh[48];
    const bool want_sh = a.shs && a.dL_dsh;
    if (want_sh) {
#pragma unroll
        for (int i = 0; i < 48; ++i) dsh[i] = 0.f;
    }

    if (visible) {
        const float3 p = make_float3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
        const float* cov3D = a.cov3D_precomp ? a.cov3D_precomp + (size_t)idx * 6 : a.cov3D + (size_t)idx * 6;
        float cv[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cv[i] = cov3D[i];
        float3 dm = cov2d_bwd(p, a.rp, cv, cam.view, g0.z, g0.w, g1.x, dcov);
        const float3 dm2 = mean_proj_depth_bwd(p, cam.view, cam.proj, g0.x, g0.y, g2.y, a.rp.inv_depth);
        dm.x += dm2.x;
        dm.y += dm2.y;
        dm.z += dm2.z;
        if (a.shs) {
            float sh[48];
            const int deg = effective_sh_degree(a.rp.sh_degree, a.M);
            load_sh_row(a.shs + (size_t)idx * a.M * 3, a.M, deg, sh);
            const float dcol[3] = {g1.z, g1.w, g2.x};
            float dsh_local[48];
            const float3 dm3 = sh_to_rgb_bwd(deg, p, cam.campos, sh, a.clamped[idx], dcol, dsh_local);
            if (want_sh) {
#pragma unroll
                for (int i = 0; i < 48; ++i) dsh[i] = dsh_local[i];
            }
            dm.x += dm3.x;
            dm.y += dm3.y;
            dm.z += dm3.z;
        }
        if (a.scales && !a.cov3D_precomp) {
            const float sc[3] = {a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]};
            const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
            const float rt[4] = {q.x, q.y, q.z, q.w};
            cov3d_bwd(sc, a.rp.scale_modifier, rt, dcov, dscale, drot);
        }
        dmean[0] = dm.x;
        dmean[1] = dm.y;
        dmean[2] = dm.z;
    }

    if (a.dL_dmeans3D) {
#pragma unroll
        for (int i = 0; i < 3; ++i) a.dL_dmeans3D[3 * idx + i] = dmean[i];
    }
    if (a.dL_dcov3D) {
        float2* o = reinterpret_cast<float2*>(a.dL_dcov3D + (size_t)idx * 6);
        o[0] = make_float2(dcov[0], dcov[1]);
        o[1] = make_float2(dcov[2], dcov[3]);
        o[2] = make_float2(dcov[4], dcov[5]);
    }
    if (a.dL_dscales) {
#pragma unroll
        for (int i = 0; i < 3; ++i) a.dL_dscales[3 * idx + i] = dscale[i];
    }
    if (a.dL_drotations) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
    if (want_sh) {
        float* o = a.dL_dsh + (size_t)idx * a.M * 3;
        if (a.M == 16 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int q = 0; q < 12; ++q) o4[q] = make_float4(dsh[4 * q], dsh[4 * q + 1], dsh[4 * q + 2], dsh[4 * q + 3]);
        } else {
            const int n = a.M * 3;
#pragma unroll
            for (int i = 0; i < 48; ++i)
                if (i < n) o[i] = dsh[i];
        }
    }
}

}  // namespace

void launch_preprocess(const PreprocessArgs& a, cudaStream_t stream)
{
    if (a.P <= 0) return;
    count_launch(1);
    preprocess_aos_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(a);
}

void launch_preprocess_backward(const PreprocessBwdArgs& a, cudaStream_t stream)
{
    if (a.P <= 0) return;
    count_launch(1);
    preprocess_bwd_aos_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(a);
}

void launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                         cudaStream_t stream)
{
    if (P <= 0) return;
    count_launch(1);
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, view, proj, present);
}

void launch_emit(int P, const uint32_t* depth_order, const uint32_t* point_offsets, const uint32_t* tiles_touched,
                 const float4* record, const int32_t* radii, int grid_x, int grid_y, uint32_t* keys, uint32_t* vals,
                 uint32_t capacity, uint32_t* counters, cudaStream_t stream, const float* mean_x, const float* mean_y)
{
    if (P <= 0) return;
    count_launch(1);
    emit_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, depth_order, point_offsets, tiles_touched, record, radii,
                                                     grid_x, grid_y, keys, vals, capacity, counters, mean_x, mean_y);
}

void launch_tile_ranges(const uint32_t* sorted_tiles, const uint32_t* counters, uint32_t capacity, uint32_t* ranges,
                        cudaStream_t stream)
{
    const int sms = device_info().sm_count;
    size_t want = ((size_t)capacity + 255) / 256;
    int grid = (int)min((size_t)sms * 8, want);
    if (grid < 1) grid = 1;
    count_launch(1);
    tile_ranges_kernel<<<grid, 256, 0, stream>>>(sorted_tiles, counters, capacity, ranges);
}

}  // namespace adgs
