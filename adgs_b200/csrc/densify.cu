// Densification / pruning of the planar model storage.
//
// Replaces, for the planar layout of adgs_model, what the reference does with ~250 boolean-mask
// indexing / torch.cat launches and a dozen host synchronisations every 200 iterations
// (scene/gaussian_model.py:560-861, train.py:149-160):
//
//   adgs_densify_stats      max_radii2D update + add_densification_stats    train.py:151-152, gaussian_model.py:863-867
//   adgs_densify_classify   the masks of densify_and_prune / _clone / _split / the final prune, per source Gaussian
//   adgs_densify_plan       where every surviving / new Gaussian comes from (one int pair per output row)
//   adgs_densify_gather     ONE launch that builds every new parameter and Adam-moment array
//   adgs_densify_split      positions and scales of the split children
//   adgs_reset_opacity      reset_opacity + replace_tensor_to_optimizer      gaussian_model.py:463-467, 547-558
//
// The reference applies clone -> split -> prune sequentially, re-allocating all 17 per-Gaussian tensors
// and their 34 moments three times. Composed, the result per partition (scene rows, object rows) is
//   [originals that are neither split nor pruned] ++ [clones, not pruned] ++ [children copy 0] ++ [children copy 1],
// each list in source order, and every decision is a function of the SOURCE Gaussian alone (a clone has its
// source's opacity and scale, a child its source's opacity and scale / (0.8 N)). So one classification
// pass, two exclusive scans and one gather write each output byte exactly once.
//
// Arithmetic follows the torch CUDA kernels the reference runs, so that the integer outputs (which rows
// survive, in which order) are identical: exp / log are the IEEE libm versions (no fast-math),
// sigmoid = 1 / (1 + exp(-x)), thresholds are rounded to float once on the host (a python scalar compared with a
// float tensor is compared in float), `tensor / python_scalar` is a multiplication by the float reciprocal
// (ATen BinaryDivTrueKernel.cu).
#include "api_internal.cuh"
#include <cstring>

namespace adgs {
namespace {

constexpr int kThreads = 256;

enum : uint8_t { kKeep = 1, kClone = 2, kChild = 4, kSplitSel = 8 };

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
densify_stats_kernel(int N, const float* __restrict__ grad_means2D, const int32_t* __restrict__ radii,
                     float* __restrict__ accum, float* __restrict__ denom, float* __restrict__ max_radii)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    const int r = radii[i];
    if (r <= 0) return;  // visibility_filter = radii > 0 (gaussian_renderer/__init__.py:104)
    if (max_radii) max_radii[i] = fmaxf(max_radii[i], (float)r);
    if (accum) {
        const float gx = grad_means2D[3 * (size_t)i], gy = grad_means2D[3 * (size_t)i + 1];
        accum[i] += sqrtf(gx * gx + gy * gy);  // torch.norm(grad[:, :2], dim=-1)
        denom[i] += 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// classification
// ------------------------------------------------------------------------------------------------
struct ClassifyArgs {
    adgs_densify_params p;
    const float* accum;
    const float* denom;
    const float* scaling;
    const float* opacity;
    const uint8_t* prune_mask;  // mode 1
    uint8_t* flags;
    int32_t* block_counts;  // [num_blocks][4]
    int blocks_scene;
};

// CTAs never straddle the scene / object boundary: blocks [0, blocks_scene) cover the scene rows.
__device__ __forceinline__ int block_row0(const ClassifyArgs& a, int blk, int& n_rows, int& part)
{
    if (blk < a.blocks_scene) {
        part = 0;
        n_rows = a.p.N_scene;
        return blk * kThreads;
    }
    part = 1;
    n_rows = a.p.N_obj;
    return (blk - a.blocks_scene) * kThreads;
}

__global__ void __launch_bounds__(kThreads) densify_classify_kernel(const __grid_constant__ ClassifyArgs a)
{
    int n_rows, part;
    const int local = block_row0(a, blockIdx.x, n_rows, part) + threadIdx.x;
    const bool valid = local < n_rows;
    const int i = local + (part ? a.p.N_scene : 0);
    uint8_t f = 0;
    if (valid) {
        if (a.p.mode == ADGS_DENSIFY_PRUNE_ONLY) {
            f = a.prune_mask[i] ? 0 : kKeep;
        } else {
            // grads = xyz_gradient_accum / denom; grads[isnan] = 0; norm over the size-1 last dim
            float g = a.accum[i] / a.denom[i];
            if (isnan(g)) g = 0.0f;
            g = fabsf(g);
            const bool sel = g >= (part ? a.p.max_obj_grad : a.p.max_scene_grad);
            const float e0 = expf(a.scaling[3 * (size_t)i]), e1 = expf(a.scaling[3 * (size_t)i + 1]),
                        e2 = expf(a.scaling[3 * (size_t)i + 2]);
            const float smax = fmaxf(fmaxf(e0, e1), e2);
            const float split_size = part ? a.p.obj_split_size : a.p.scene_split_size;
            const bool clone = sel && smax <= split_size;
            const bool split = sel && smax > split_size;
            const float big = part ? a.p.obj_big_size : a.p.scene_big_size;
            const bool low = sigmoid_ref(a.opacity[i]) < a.p.min_opacity;
            const bool pruned = low || (a.p.prune_big && smax > big);
            // children: _scaling = log(get_scaling / (0.8 N)), tested through get_scaling = exp(_scaling)
            const float inv = a.p.inv_split_scale;
            const float c0 = expf(logf(e0 * inv)), c1 = expf(logf(e1 * inv)), c2 = expf(logf(e2 * inv));
            const float cmax = fmaxf(fmaxf(c0, c1), c2);
            const bool child_pruned = low || (a.p.prune_big && cmax > big);
            if (!split && !pruned) f |= kKeep;
            if (clone && !pruned) f |= kClone;
            if (split && !child_pruned) f |= kChild;
            if (split) f |= kSplitSel;
        }
        a.flags[i] = f;
    }
    const int c_keep = __syncthreads_count(f & kKeep);
    const int c_clone = __syncthreads_count(f & kClone);
    const int c_child = __syncthreads_count(f & kChild);
    const int c_sel = __syncthreads_count(f & kSplitSel);
    if (threadIdx.x == 0) {
        int32_t* o = a.block_counts + 4 * (size_t)blockIdx.x;
        o[0] = c_keep;
        o[1] = c_clone;
        o[2] = c_child;
        o[3] = c_sel;
    }
}

// One CTA: exclusive scan of the per-block counts, separately over the scene and the object blocks.
// block_counts is overwritten with the exclusive bases; totals[part*4 + k] receives the sums.
__global__ void __launch_bounds__(1024) densify_scan_kernel(int32_t* block_counts, int blocks_scene, int blocks_total,
                                                            int32_t* totals)
{
    __shared__ int warp_sums[32][4];
    __shared__ int carry[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int part = 0; part < 2; ++part) {
        const int b0 = part ? blocks_scene : 0, b1 = part ? blocks_total : blocks_scene;
        if (threadIdx.x < 4) carry[threadIdx.x] = 0;
        __syncthreads();
        for (int base = b0; base < b1; base += 1024) {
            const int b = base + threadIdx.x;
            int v[4] = {0, 0, 0, 0};
            if (b < b1) {
                const int4 q = *reinterpret_cast<const int4*>(block_counts + 4 * (size_t)b);
                v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
            }
            int inc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int x = v[k];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                inc[k] = x;
                if (lane == 31) warp_sums[warp][k] = x;
            }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int x = warp_sums[lane][k];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int y = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= d) x += y;
                    }
                    warp_sums[lane][k] = x;  // inclusive over warps
                }
            }
            __syncthreads();
            int4 out;
            int excl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                excl[k] = carry[k] + (warp ? warp_sums[warp - 1][k] : 0) + inc[k] - v[k];
            out.x = excl[0], out.y = excl[1], out.z = excl[2], out.w = excl[3];
            if (b < b1) *reinterpret_cast<int4*>(block_counts + 4 * (size_t)b) = out;
            __syncthreads();
            if (threadIdx.x < 4) carry[threadIdx.x] += warp_sums[31][threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x < 4) totals[part * 4 + threadIdx.x] = carry[threadIdx.x];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// plan: src[dst] = source row (global index), tag[dst] = kind | sample_row << 2
// ------------------------------------------------------------------------------------------------
struct PlanArgs {
    int N_scene, N_obj, blocks_scene, n_split;
    const uint8_t* flags;
    const int32_t* block_base;  // exclusive bases from the scan
    int32_t totals[8];          // host copy of the totals
    int32_t* src;
    int32_t* tag;
};

__global__ void __launch_bounds__(kThreads) densify_plan_kernel(const __grid_constant__ PlanArgs a)
{
    __shared__ int warp_tot[kThreads / 32][4];
    const int blk = blockIdx.x;
    const int part = blk >= a.blocks_scene;
    const int n_rows = part ? a.N_obj : a.N_scene;
    const int local = (part ? blk - a.blocks_scene : blk) * kThreads + threadIdx.x;
    const bool valid = local < n_rows;
    const int i = local + (part ? a.N_scene : 0);
    const uint8_t f = valid ? a.flags[i] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int rank[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, (f >> k) & 1);
        rank[k] = __popc(m & lt);
        if (lane == 0) warp_tot[warp][k] = __popc(m);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += warp_tot[w][k];
        rank[k] += before + a.block_base[4 * (size_t)blk + k];
    }
    if (!valid) return;
    const int* t = a.totals + 4 * part;
    const int kept = t[0], clones = t[1], children = t[2], sel = t[3];
    // first output row of this partition
    const int out0 = part ? a.totals[0] + a.totals[1] + a.n_split * a.totals[2] : 0;
    if (f & kKeep) {
        const int d = out0 + rank[0];
        a.src[d] = i;
        a.tag[d] = ADGS_DENSIFY_KIND_KEEP;
    }
    if (f & kClone) {
        const int d = out0 + kept + rank[1];
        a.src[d] = i;
        a.tag[d] = ADGS_DENSIFY_KIND_CLONE;
    }
    if (f & kChild) {
        for (int c = 0; c < a.n_split; ++c) {
            const int d = out0 + kept + clones + c * children + rank[2];
            a.src[d] = i;
            // samples were drawn for every SELECTED source (before the final prune): row c * sel + rank among selected
            a.tag[d] = ADGS_DENSIFY_KIND_CHILD | ((c * sel + rank[3]) << 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gather
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
    adgs_gather_segment seg[ADGS_GATHER_MAX_SEGMENTS];
    long long first_block[ADGS_GATHER_MAX_SEGMENTS + 1];
    int n_seg;
    const int32_t* src;
    const int32_t* tag;
};

template <int W>
__device__ __forceinline__ void copy_row(float* __restrict__ d, const float* __restrict__ s, bool zero)
{
    if (W == 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!zero) v = *reinterpret_cast<const float4*>(s);
        *reinterpret_cast<float4*>(d) = v;
    } else if (W == 2) {
        float2 v = make_float2(0.f, 0.f);
        if (!zero) v = *reinterpret_cast<const float2*>(s);
        *reinterpret_cast<float2*>(d) = v;
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) d[k] = zero ? 0.0f : s[k];
    }
}

__global__ void __launch_bounds__(kThreads) densify_gather_kernel(const __grid_constant__ GatherArgs a)
{
    int si = 0;
    while (si + 1 < a.n_seg && (long long)blockIdx.x >= a.first_block[si + 1]) ++si;
    const adgs_gather_segment& s = a.seg[si];
    const long long rel = (long long)blockIdx.x - a.first_block[si];
    const int blocks_per_plane = (s.dst_rows + kThreads - 1) / kThreads;
    const int plane = (int)(rel / blocks_per_plane);
    const int r = (int)(rel % blocks_per_plane) * kThreads + threadIdx.x;
    if (r >= s.dst_rows) return;
    const int d_global = s.dst_row0 + r;
    const int sr = a.src[d_global] - s.src_row0;
    const bool zero = s.zero_new && (a.tag[d_global] & 3) != ADGS_DENSIFY_KIND_KEEP;
    const float* sp = s.src + ((size_t)plane * s.src_rows + sr) * s.width;
    float* dp = s.dst + ((size_t)plane * s.dst_rows + r) * s.width;
    switch (s.width) {
    case 1: copy_row<1>(dp, sp, zero); break;
    case 2: copy_row<2>(dp, sp, zero); break;
    case 3: copy_row<3>(dp, sp, zero); break;
    default: copy_row<4>(dp, sp, zero); break;
    }
}

// ------------------------------------------------------------------------------------------------
// split children: new_xyz = R(q / |q|) (z * exp(s)) + xyz, new _scaling = log(exp(s) / (0.8 N))
// (gaussian_model.py:719-725, utils/general_utils.py:77-94)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
densify_split_kernel(int n_dst, int n_scene_dst, const int32_t* __restrict__ src, const int32_t* __restrict__ tag,
                     const float* __restrict__ xyz, const float* __restrict__ scaling, const float* __restrict__ rotation,
                     const float* __restrict__ z_scene, const float* __restrict__ z_obj, float inv_split_scale,
                     float* __restrict__ new_xyz, float* __restrict__ new_scaling)
{
    const int d = blockIdx.x * kThreads + threadIdx.x;
    if (d >= n_dst) return;
    const int t = tag[d];
    if ((t & 3) != ADGS_DENSIFY_KIND_CHILD) return;
    const size_t i = (size_t)src[d];
    const float* z = (d < n_scene_dst ? z_scene : z_obj) + 3 * (size_t)(t >> 2);
    const float e0 = expf(scaling[3 * i]), e1 = expf(scaling[3 * i + 1]), e2 = expf(scaling[3 * i + 2]);
    const float4 q4 = *reinterpret_cast<const float4*>(rotation + 4 * i);
    const float norm = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
    const float r = q4.x / norm, x = q4.y / norm, y = q4.z / norm, zq = q4.w / norm;
    const float s0 = z[0] * e0, s1 = z[1] * e1, s2 = z[2] * e2;
    const float R00 = 1.f - 2.f * (y * y + zq * zq), R01 = 2.f * (x * y - r * zq), R02 = 2.f * (x * zq + r * y);
    const float R10 = 2.f * (x * y + r * zq), R11 = 1.f - 2.f * (x * x + zq * zq), R12 = 2.f * (y * zq - r * x);
    const float R20 = 2.f * (x * zq - r * y), R21 = 2.f * (y * zq + r * x), R22 = 1.f - 2.f * (x * x + y * y);
    new_xyz[3 * (size_t)d] = R00 * s0 + R01 * s1 + R02 * s2 + xyz[3 * i];
    new_xyz[3 * (size_t)d + 1] = R10 * s0 + R11 * s1 + R12 * s2 + xyz[3 * i + 1];
    new_xyz[3 * (size_t)d + 2] = R20 * s0 + R21 * s1 + R22 * s2 + xyz[3 * i + 2];
    new_scaling[3 * (size_t)d] = logf(e0 * inv_split_scale);
    new_scaling[3 * (size_t)d + 1] = logf(e1 * inv_split_scale);
    new_scaling[3 * (size_t)d + 2] = logf(e2 * inv_split_scale);
}

// reset_opacity: inverse_sigmoid(min(sigmoid(o), 0.01)), moments zeroed (replace_tensor_to_optimizer)
__global__ void __launch_bounds__(kThreads)
reset_opacity_kernel(int N, float cap, float* __restrict__ opacity, float* __restrict__ exp_avg,
                     float* __restrict__ exp_avg_sq)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    const float x = fminf(sigmoid_ref(opacity[i]), cap);
    opacity[i] = logf(x / (1.0f - x));  // utils/general_utils.py:inverse_sigmoid
    if (exp_avg) exp_avg[i] = 0.0f;
    if (exp_avg_sq) exp_avg_sq[i] = 0.0f;
}

inline int num_blocks(int n) { return (n + kThreads - 1) / kThreads; }

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

int adgs_densify_stats(int32_t N, const float* grad_means2D, const int32_t* radii, float* xyz_gradient_accum,
                       float* denom, float* max_radii2D, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || (N > 0 && !radii)) return ADGS_ERR_ARG;
    if ((xyz_gradient_accum != nullptr) != (denom != nullptr)) return ADGS_ERR_ARG;
    if (xyz_gradient_accum && !grad_means2D) return ADGS_ERR_ARG;
    if (N == 0) return ADGS_OK;
    densify_stats_kernel<<<num_blocks(N), kThreads, 0, stream>>>(N, grad_means2D, radii, xyz_gradient_accum, denom,
                                                                  max_radii2D);
    count_launch(1);
    return check_stage("densify stats", false, stream);
}

size_t adgs_densify_workspace_bytes(int32_t N_scene, int32_t N_obj)
{
    const size_t blocks = (size_t)num_blocks(N_scene) + (size_t)num_blocks(N_obj);
    // flags (N bytes, padded) + block counts (16 B per block) + totals (8 ints)
    const size_t flags = (((size_t)N_scene + (size_t)N_obj) + 127) / 128 * 128;
    return flags + (blocks * 16 + 127) / 128 * 128 + 128;
}

static void carve(char* ws, int N_scene, int N_obj, uint8_t** flags, int32_t** counts, int32_t** totals)
{
    const size_t blocks = (size_t)num_blocks(N_scene) + (size_t)num_blocks(N_obj);
    const size_t fbytes = (((size_t)N_scene + (size_t)N_obj) + 127) / 128 * 128;
    *flags = reinterpret_cast<uint8_t*>(ws);
    *counts = reinterpret_cast<int32_t*>(ws + fbytes);
    *totals = reinterpret_cast<int32_t*>(ws + fbytes + (blocks * 16 + 127) / 128 * 128);
}

int adgs_densify_classify(const adgs_densify_params* p, const float* xyz_gradient_accum, const float* denom,
                          const float* scaling, const float* opacity, const uint8_t* prune_mask, char* workspace,
                          int32_t* host_totals8, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !workspace || p->N_scene < 0 || p->N_obj < 0) return ADGS_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return ADGS_ERR_ARG;
    const long long N = (long long)p->N_scene + p->N_obj;
    if (N > 0x7fffffffLL / 4) return ADGS_ERR_UNSUPPORTED;
    if (p->mode == ADGS_DENSIFY_PRUNE_ONLY) {
        if (N > 0 && !prune_mask) return ADGS_ERR_ARG;
    } else if (p->mode == ADGS_DENSIFY_AND_PRUNE) {
        if (N > 0 && (!xyz_gradient_accum || !denom || !scaling || !opacity)) return ADGS_ERR_ARG;
        if (p->n_split < 1 || p->n_split > 8) return ADGS_ERR_ARG;
    } else {
        return ADGS_ERR_ARG;
    }
    ClassifyArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.p.inv_split_scale = 1.0f / (float)(0.8 * (p->n_split > 0 ? p->n_split : 1));
    a.accum = xyz_gradient_accum;
    a.denom = denom;
    a.scaling = scaling;
    a.opacity = opacity;
    a.prune_mask = prune_mask;
    a.blocks_scene = num_blocks(p->N_scene);
    const int blocks = a.blocks_scene + num_blocks(p->N_obj);
    int32_t* totals;
    carve(workspace, p->N_scene, p->N_obj, &a.flags, &a.block_counts, &totals);
    if (blocks > 0) {
        densify_classify_kernel<<<blocks, kThreads, 0, stream>>>(a);
        count_launch(1);
    }
    densify_scan_kernel<<<1, 1024, 0, stream>>>(a.block_counts, a.blocks_scene, blocks, totals);
    count_launch(1);
    if (host_totals8) {
        // the one host round-trip of a densification: the new array sizes
        cudaError_t e = cudaMemcpyAsync(host_totals8, totals, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return record_cuda_error(e, "densify totals copy");
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return record_cuda_error(e, "densify totals sync");
    }
    return check_stage("densify classify", false, stream);
}

int adgs_densify_plan(const adgs_densify_params* p, const char* workspace, const int32_t* host_totals8, int32_t* src,
                      int32_t* tag, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!p || !workspace || !host_totals8) return ADGS_ERR_ARG;
    const int n_split = p->mode == ADGS_DENSIFY_PRUNE_ONLY ? 1 : p->n_split;
    long long n_dst = 0;
    for (int part = 0; part < 2; ++part)
        n_dst += (long long)host_totals8[4 * part] + host_totals8[4 * part + 1] + (long long)n_split * host_totals8[4 * part + 2];
    if (n_dst > 0x7fffffffLL / 4) return ADGS_ERR_UNSUPPORTED;
    if (n_dst > 0 && (!src || !tag)) return ADGS_ERR_ARG;
    // sample rows are packed above the two kind bits
    for (int part = 0; part < 2; ++part)
        if ((long long)n_split * host_totals8[4 * part + 3] >= (1LL << 29)) return ADGS_ERR_UNSUPPORTED;
    PlanArgs a;
    memset(&a, 0, sizeof(a));
    a.N_scene = p->N_scene;
    a.N_obj = p->N_obj;
    a.blocks_scene = num_blocks(p->N_scene);
    a.n_split = n_split;
    uint8_t* flags;
    int32_t *counts, *totals;
    carve(const_cast<char*>(workspace), p->N_scene, p->N_obj, &flags, &counts, &totals);
    a.flags = flags;
    a.block_base = counts;
    memcpy(a.totals, host_totals8, sizeof(a.totals));
    a.src = src;
    a.tag = tag;
    const int blocks = a.blocks_scene + num_blocks(p->N_obj);
    if (blocks == 0) return ADGS_OK;
    densify_plan_kernel<<<blocks, kThreads, 0, stream>>>(a);
    count_launch(1);
    return check_stage("densify plan", false, stream);
}

int adgs_densify_gather(const adgs_gather_segment* segments, int32_t num_segments, const int32_t* src,
                        const int32_t* tag, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_segments < 0 || num_segments > ADGS_GATHER_MAX_SEGMENTS || (num_segments > 0 && !segments)) return ADGS_ERR_ARG;
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    long long blocks = 0;
    int n = 0;
    for (int i = 0; i < num_segments; ++i) {
        const adgs_gather_segment& s = segments[i];
        if (s.planes < 0 || s.dst_rows < 0 || s.src_rows < 0 || s.width < 1 || s.width > 4) return ADGS_ERR_ARG;
        if (s.planes == 0 || s.dst_rows == 0) continue;
        if (!s.src || !s.dst || !src || !tag) return ADGS_ERR_ARG;
        if ((s.width == 4 && ((reinterpret_cast<uintptr_t>(s.src) | reinterpret_cast<uintptr_t>(s.dst)) & 15)) ||
            (s.width == 2 && ((reinterpret_cast<uintptr_t>(s.src) | reinterpret_cast<uintptr_t>(s.dst)) & 7)))
            return ADGS_ERR_ARG;
        a.seg[n] = s;
        a.first_block[n] = blocks;
        blocks += (long long)s.planes * num_blocks(s.dst_rows);
        ++n;
    }
    a.n_seg = n;
    a.first_block[n] = blocks;
    a.src = src;
    a.tag = tag;
    if (blocks == 0) return ADGS_OK;
    if (blocks > 0x7fffffffLL) return ADGS_ERR_UNSUPPORTED;
    densify_gather_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(a);
    count_launch(1);
    return check_stage("densify gather", false, stream);
}

int adgs_densify_split(int32_t n_dst, int32_t n_scene_dst, const int32_t* src, const int32_t* tag, const float* xyz,
                       const float* scaling, const float* rotation, const float* z_scene, const float* z_obj,
                       int32_t n_split, float* new_xyz, float* new_scaling, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_dst < 0 || n_scene_dst < 0 || n_scene_dst > n_dst || n_split < 1) return ADGS_ERR_ARG;
    if (n_dst == 0) return ADGS_OK;
    if (!src || !tag || !xyz || !scaling || !rotation || !new_xyz || !new_scaling) return ADGS_ERR_ARG;
    const float inv = 1.0f / (float)(0.8 * n_split);
    densify_split_kernel<<<num_blocks(n_dst), kThreads, 0, stream>>>(n_dst, n_scene_dst, src, tag, xyz, scaling, rotation,
                                                                     z_scene, z_obj, inv, new_xyz, new_scaling);
    count_launch(1);
    return check_stage("densify split", false, stream);
}

int adgs_reset_opacity(int32_t N, float cap, float* opacity, float* exp_avg, float* exp_avg_sq, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || (N > 0 && !opacity)) return ADGS_ERR_ARG;
    if (N == 0) return ADGS_OK;
    reset_opacity_kernel<<<num_blocks(N), kThreads, 0, stream>>>(N, cap, opacity, exp_avg, exp_avg_sq);
    count_launch(1);
    return check_stage("reset opacity", false, stream);
}

}  // extern "C"
