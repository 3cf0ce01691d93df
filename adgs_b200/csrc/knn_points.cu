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
// staging the slice through shared memory (one float4 per point, broadcast reads) and keeping each anchor's
// K best in registers; a second pass merges the S sorted partial lists of an anchor.
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

template <int D, int K>
__global__ void __launch_bounds__(kAnchors)
knn_partial_kernel(int A, int P, const float* __restrict__ anchors, const float* __restrict__ points, int slice_len,
                   float* __restrict__ part_d, int* __restrict__ part_i)
{
    __shared__ float4 tile[kTile];
    const int a = blockIdx.x * kAnchors + threadIdx.x;
    const int s = blockIdx.y, S = gridDim.y;
    const int p0 = s * slice_len, p1 = min(P, p0 + slice_len);
    float q[4] = {0.f, 0.f, 0.f, 0.f};
    if (a < A)
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = anchors[(size_t)a * D + d];
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = FLT_MAX;
        bi[j] = -1;
    }
    for (int base = p0; base < p1; base += kTile) {
        const int n = min(kTile, p1 - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kAnchors) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* src = points + (size_t)(base + j) * D;
            v.x = src[0];
            v.y = src[1];
            v.z = src[2];
            if (D == 4) v.w = src[3];
            tile[j] = v;
        }
        __syncthreads();
        if (a < A) {
#pragma unroll 4
            for (int j = 0; j < n; ++j) {
                const float4 v = tile[j];
                float diff = q[0] - v.x;
                float dist = diff * diff;
                diff = q[1] - v.y;
                dist += diff * diff;
                diff = q[2] - v.z;
                dist += diff * diff;
                if (D == 4) {
                    diff = q[3] - v.w;
                    dist += diff * diff;
                }
                knn_insert<K>(bd, bi, dist, base + j);
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
