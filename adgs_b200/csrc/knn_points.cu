// K nearest neighbours of a set of anchors in a point cloud: replaces the
// `pytorch3d.ops.knn_points(anchor[None], xyz[None], K=K).idx` call of GaussianModel.set_obj_near_idx
// (scene/gaussian_model.py:825-833; re-run every near_idx_reset_interval = 10 iterations, train.py:156-157).
//
// Contract (pytorch3d 0.7 knn_points, brute-force path): squared Euclidean distance accumulated over the D
// coordinates in order, the K smallest per anchor returned in ascending order of distance (an anchor that is
// itself a member of the cloud finds itself first, at distance 0). Ties keep the smaller point index.
// D = 3 (positions) or 4 (positions + gs_time * scene_extent, use_time_mask = True).
//
// Design: exact brute force, A x P pairs. With A = P / K anchors there are too few anchors to fill 148 SMs with
// one thread per anchor, so the cloud is cut into S slices: CTA (x, s) answers 128 anchors against slice s,
// staging the slice through shared memory (points in pairs for the packed FP32x2 pipe, broadcast reads) and keeping
// each anchor's K best in registers; a second pass merges the S sorted partial lists of an anchor.
#include <cfloat>
#include "api_internal.cuh"

namespace adgs {
namespace {

constexpr int kAnchors = 128;  // threads per CTA = anchors per CTA
constexpr int kTile = 1024;    // points staged per step (16 KB)

template <int K>
__device__ __forceinline__ void knn_insert(float (&bd)[K], int (&bi)[K], float d, int idx)
{
    // bd ascending; strict '<' keeps the earlier (smaller-index) point on ties
    if (!(d < bd[K - 1])) return;
    bd[K - 1] = d;
    bi[K - 1] = idx;
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        if (bd[j] < bd[j - 1]) {
            const float td = bd[j];
            bd[j] = bd[j - 1];
            bd[j - 1] = td;
            const int ti = bi[j];
            bi[j] = bi[j - 1];
            bi[j - 1] = ti;
        }
    }
}

// Points are staged in PAIRS so that the distance arithmetic runs on the packed FP32x2 pipe (FADD2 / FMUL2 / FFMA2,
// two points per instruction): pair jp of a tile is {x0, x1, y0, y1}, {z0, z1, w0, w1}. Every component is the same
// IEEE operation sequence as the scalar form (q - v as one rounding, then one multiply and D - 1 fused
// multiply-adds), so distances and therefore the returned indices are bit-identical to the scalar kernel.
template <int D, int K>
__global__ void __launch_bounds__(kAnchors)
knn_partial_kernel(int A, int P, const float* __restrict__ anchors, const float* __restrict__ points, int slice_len,
                   float* __restrict__ part_d, int* __restrict__ part_i)
{
    __shared__ float4 tile_xy[kTile / 2];
    __shared__ float4 tile_zw[kTile / 2];
    const int a = blockIdx.x * kAnchors + threadIdx.x;
    const int s = blockIdx.y, S = gridDim.y;
    const int p0 = s * slice_len, p1 = min(P, p0 + slice_len);
    float q[4] = {0.f, 0.f, 0.f, 0.f};
    if (a < A)
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = anchors[(size_t)a * D + d];
    const float2 qx = make_float2(q[0], q[0]), qy = make_float2(q[1], q[1]), qz = make_float2(q[2], q[2]),
                 qw = make_float2(q[3], q[3]);
    const float2 neg1 = make_float2(-1.f, -1.f);
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = FLT_MAX;
        bi[j] = -1;
    }
    float* sxy = reinterpret_cast<float*>(tile_xy);
    float* szw = reinterpret_cast<float*>(tile_zw);
    for (int base = p0; base < p1; base += kTile) {
        const int n = min(kTile, p1 - base);
        const int n_pairs = (n + 1) >> 1;
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * n_pairs; j += kAnchors) {
            // an odd tail is padded with a point at infinity: its distance is +inf and never beats a candidate
            float x = __int_as_float(0x7f800000), y = 0.f, z = 0.f, w = 0.f;
            if (j < n) {
                const float* src = points + (size_t)(base + j) * D;
                x = src[0];
                y = src[1];
                z = src[2];
                if (D == 4) w = src[3];
            }
            const int o = 4 * (j >> 1) + (j & 1);
            sxy[o] = x;
            sxy[o + 2] = y;
            szw[o] = z;
            szw[o + 2] = w;
        }
        __syncthreads();
        if (a < A) {
#pragma unroll 4
            for (int jp = 0; jp < n_pairs; ++jp) {
                const float4 vxy = tile_xy[jp];
                const float4 vzw = tile_zw[jp];
                float2 diff = __ffma2_rn(make_float2(vxy.x, vxy.y), neg1, qx);  // q - v, one rounding
                float2 dist = __fmul2_rn(diff, diff);
                diff = __ffma2_rn(make_float2(vxy.z, vxy.w), neg1, qy);
                dist = __ffma2_rn(diff, diff, dist);
                diff = __ffma2_rn(make_float2(vzw.x, vzw.y), neg1, qz);
                dist = __ffma2_rn(diff, diff, dist);
                if (D == 4) {
                    diff = __ffma2_rn(make_float2(vzw.z, vzw.w), neg1, qw);
                    dist = __ffma2_rn(diff, diff, dist);
                }
                knn_insert<K>(bd, bi, dist.x, base + 2 * jp);      // lower index first: ties keep it
                knn_insert<K>(bd, bi, dist.y, base + 2 * jp + 1);
            }
        }
    }
    if (a < A) {
        float* od = part_d + ((size_t)a * S + s) * K;
        int* oi = part_i + ((size_t)a * S + s) * K;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            od[j] = bd[j];
            oi[j] = bi[j];
        }
    }
}

// Merge the S partial lists of an anchor (slices are in increasing index order, so on equal distance the
// candidate met first has the smaller index and strict '<' keeps it).
template <int K>
__global__ void __launch_bounds__(128)
knn_merge_kernel(int A, int S, int K_out, const float* __restrict__ part_d, const int* __restrict__ part_i,
                 long long* __restrict__ idx, float* __restrict__ dists)
{
    const int a = blockIdx.x * 128 + threadIdx.x;
    if (a >= A) return;
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = FLT_MAX;
        bi[j] = -1;
    }
    const float* pd = part_d + (size_t)a * S * K;
    const int* pi = part_i + (size_t)a * S * K;
    for (int c = 0; c < S * K; ++c) {
        const int i = pi[c];
        if (i >= 0) knn_insert<K>(bd, bi, pd[c], i);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < K_out) {
            idx[(size_t)a * K_out + j] = bi[j];
            if (dists) dists[(size_t)a * K_out + j] = bi[j] >= 0 ? bd[j] : 0.0f;
        }
    }
}

inline int padded_k(int K) { return K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 32; }

inline int num_slices(int A, int P)
{
    const int ctas_x = (A + kAnchors - 1) / kAnchors;
    const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
    int S = (sms * 8 + ctas_x - 1) / ctas_x;                    // ~8 CTAs of 128 threads per SM
    const int max_by_len = (P + 4 * kTile - 1) / (4 * kTile);    // a slice is at least 4 tiles long
    if (S > max_by_len) S = max_by_len;
    if (S > 64) S = 64;
    if (S < 1) S = 1;
    return S;
}

template <int D, int K>
void launch_knn(int A, int P, int K_out, const float* anchors, const float* points, int S, float* part_d, int* part_i,
                long long* idx, float* dists, cudaStream_t stream)
{
    const int slice_len = (P + S - 1) / S;
    dim3 grid((A + kAnchors - 1) / kAnchors, S);
    knn_partial_kernel<D, K><<<grid, kAnchors, 0, stream>>>(A, P, anchors, points, slice_len, part_d, part_i);
    knn_merge_kernel<K><<<(A + 127) / 128, 128, 0, stream>>>(A, S, K_out, part_d, part_i, idx, dists);
    count_launch(2);
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

size_t adgs_knn_points_workspace_bytes(int32_t A, int32_t P, int32_t K)
{
    if (A <= 0 || P <= 0 || K <= 0 || K > 32) return 128;
    // the slice count depends on the device's SM count only through an upper bound of 64
    return (size_t)A * 64 * padded_k(K) * 8 + 256;
}

int adgs_knn_points(int32_t A, int32_t P, int32_t D, int32_t K, const float* anchors, const float* points,
                    int64_t* idx, float* dists, char* workspace, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (A < 0 || P < 0 || K < 1) return ADGS_ERR_ARG;
    if (K > 32 || (D != 3 && D != 4)) return ADGS_ERR_UNSUPPORTED;
    if (A == 0) return ADGS_OK;
    if (K > P) return ADGS_ERR_ARG;  // pytorch3d pads with -1; set_obj_near_idx never asks for it
    if (!anchors || !points || !idx || !workspace) return ADGS_ERR_ARG;
    const int Kp = padded_k(K);
    const int S = num_slices(A, P);
    char* c = workspace;
    c = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(c) + 127) & ~(uintptr_t)127);
    float* part_d = reinterpret_cast<float*>(c);
    int* part_i = reinterpret_cast<int*>(c + (size_t)A * 64 * Kp * 4);
    long long* out = reinterpret_cast<long long*>(idx);
#define ADGS_KNN_CASE(DD, KK) \
    launch_knn<DD, KK>(A, P, K, anchors, points, S, part_d, part_i, out, dists, stream)
    if (D == 3) {
        if (Kp == 4) ADGS_KNN_CASE(3, 4);
        else if (Kp == 8) ADGS_KNN_CASE(3, 8);
        else if (Kp == 16) ADGS_KNN_CASE(3, 16);
        else ADGS_KNN_CASE(3, 32);
    } else {
        if (Kp == 4) ADGS_KNN_CASE(4, 4);
        else if (Kp == 8) ADGS_KNN_CASE(4, 8);
        else if (Kp == 16) ADGS_KNN_CASE(4, 16);
        else ADGS_KNN_CASE(4, 32);
    }
#undef ADGS_KNN_CASE
    return check_stage("knn_points", false, stream);
}

}  // extern "C"
