// distCUDA2: mean squared distance of every point to its 3 nearest neighbours.
//
// Replaces SimpleKNN::knn (KNN/simple_knn.cu:185-221 and its kernels :54-183). Same contract --
// EXACT 3-NN, a point is only excluded against itself (by index, so coincident points count with
// distance 0), result = (d0 + d1 + d2) / 3 -- and the same per-pair distance expression, so the
// output is bit-identical to the reference whatever the traversal order. The search itself is
// re-designed: points are Morton-sorted with the library's onesweep sort, grouped in boxes of 256,
// and one CTA answers the 256 queries of a box together: it prunes whole candidate boxes against
// the CTA's own bounding box and stages surviving candidates through shared memory, instead of
// every thread walking every box on its own.
#include <cfloat>
#include "api_internal.cuh"

namespace adgs {
namespace {

constexpr int kBox = 256;

struct Aabb {
    float lo[3], hi[3];
};

__device__ __forceinline__ uint32_t spread3(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) bounds_kernel(int P, const float* __restrict__ pts, float* __restrict__ mm)
{
    // mm[0..2] = min (as ordered ints), mm[3..5] = max; initialised by the host to +/-FLT_MAX bits
    __shared__ float s_lo[8][3], s_hi[8][3];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = pts[3 * i + d];
            lo[d] = fminf(lo[d], v);
            hi[d] = fmaxf(hi[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], off));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], off));
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_lo[warp][d] = lo[d];
            s_hi[warp][d] = hi[d];
        }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        float a = FLT_MAX, b = -FLT_MAX;
        for (int w = 0; w < 8; ++w) {
            a = fminf(a, s_lo[w][d]);
            b = fmaxf(b, s_hi[w][d]);
        }
        // float atomics via the sign-aware integer trick
        int* ilo = reinterpret_cast<int*>(mm + d);
        int* ihi = reinterpret_cast<int*>(mm + 3 + d);
        if (a >= 0) atomicMin(ilo, __float_as_int(a)); else atomicMax(reinterpret_cast<unsigned*>(ilo), __float_as_uint(a));
        if (b >= 0) atomicMax(ihi, __float_as_int(b)); else atomicMin(reinterpret_cast<unsigned*>(ihi), __float_as_uint(b));
    }
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ mm,
                                                     uint32_t* __restrict__ codes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float lo = mm[d], hi = mm[3 + d];
        const float ext = hi - lo;
        float u = ext > 0.f ? (pts[3 * (size_t)i + d] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        c[d] = spread3((uint32_t)(u * 1023.f));
    }
    codes[i] = c[0] | (c[1] << 1) | (c[2] << 2);
}

__global__ void __launch_bounds__(kBox) box_bounds_kernel(int P, const float* __restrict__ pts,
                                                          const uint32_t* __restrict__ order, Aabb* __restrict__ boxes)
{
    __shared__ float s_lo[kBox / 32][3], s_hi[kBox / 32][3];
    const int i = blockIdx.x * kBox + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const size_t g = order[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) lo[d] = hi[d] = pts[3 * g + d];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], off));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], off));
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_lo[warp][d] = lo[d];
            s_hi[warp][d] = hi[d];
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        Aabb b;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            b.lo[d] = FLT_MAX;
            b.hi[d] = -FLT_MAX;
            for (int w = 0; w < kBox / 32; ++w) {
                b.lo[d] = fminf(b.lo[d], s_lo[w][d]);
                b.hi[d] = fmaxf(b.hi[d], s_hi[w][d]);
            }
        }
        boxes[blockIdx.x] = b;
    }
}

__device__ __forceinline__ void insert3(float dist, float* best)
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
        }
    }
}

// squared distance, with the reference's expression shape (simple_knn.cu:131-135)
__device__ __forceinline__ float pair_dist2(const float3& ref, const float3& point)
{
    const float3 d = make_float3(point.x - ref.x, point.y - ref.y, point.z - ref.z);
    return d.x * d.x + d.y * d.y + d.z * d.z;
}

__device__ __forceinline__ float box_point_dist2(const Aabb& b, const float3& p)
{
    float s = 0.f;
    const float pv[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float e = 0.f;
        if (pv[d] < b.lo[d]) e = b.lo[d] - pv[d];
        else if (pv[d] > b.hi[d]) e = pv[d] - b.hi[d];
        s += e * e;
    }
    return s;
}

__device__ __forceinline__ float box_box_dist2(const Aabb& a, const Aabb& b)
{
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float e = 0.f;
        if (a.hi[d] < b.lo[d]) e = b.lo[d] - a.hi[d];
        else if (b.hi[d] < a.lo[d]) e = a.lo[d] - b.hi[d];
        s += e * e;
    }
    return s;
}

__global__ void __launch_bounds__(kBox) knn3_kernel(int P, const float* __restrict__ pts,
                                                    const uint32_t* __restrict__ order, const Aabb* __restrict__ boxes,
                                                    int num_boxes, float* __restrict__ out)
{
    __shared__ float3 s_pts[kBox];
    __shared__ float s_red[kBox / 32];
    __shared__ float s_reject;
    const int tid = threadIdx.x;
    const int my_box = blockIdx.x;
    const int i = my_box * kBox + tid;
    const bool valid = i < P;
    float3 p = make_float3(0.f, 0.f, 0.f);
    uint32_t g = 0;
    if (valid) {
        g = order[i];
        p = make_float3(pts[3 * (size_t)g], pts[3 * (size_t)g + 1], pts[3 * (size_t)g + 2]);
    }
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};

    // candidate boxes by increasing Morton distance from home: own box first (tight bounds early)
    const Aabb home = boxes[my_box];
    bool dirty = true;  // CTA-uniform: somebody's best[] may have changed since s_reject was computed
    for (int step = 0; step < 2 * num_boxes + 1; ++step) {
        const int k = (step + 1) >> 1;
        const int b = (step & 1) ? my_box - k : my_box + k;
        if (b < 0 || b >= num_boxes) {
            if (my_box - k < 0 && my_box + k >= num_boxes) break;
            continue;
        }
        if (step > 0) {
            if (dirty) {  // refresh the CTA's reject radius = max over its queries of the 3rd-best distance
                float r = valid ? best[2] : 0.f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, off));
                __syncthreads();
                if ((tid & 31) == 0) s_red[tid >> 5] = r;
                __syncthreads();
                if (tid == 0) {
                    float m = 0.f;
                    for (int w = 0; w < kBox / 32; ++w) m = fmaxf(m, s_red[w]);
                    s_reject = m;
                }
                __syncthreads();
                dirty = false;
            }
            if (box_box_dist2(home, boxes[b]) > s_reject) continue;  // nobody here can be improved by box b
        }
        __syncthreads();
        const int j = b * kBox + tid;
        if (j < P) {
            const size_t gj = order[j];
            s_pts[tid] = make_float3(pts[3 * gj], pts[3 * gj + 1], pts[3 * gj + 2]);
        }
        __syncthreads();
        const int cnt = min(kBox, P - b * kBox);
        if (valid && box_point_dist2(boxes[b], p) <= best[2]) {
            for (int q = 0; q < cnt; ++q) {
                if (b == my_box && q == tid) continue;
                insert3(pair_dist2(p, s_pts[q]), best);
            }
        }
        dirty = true;
    }
    if (valid) out[g] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnWorkspace {
    float* mm;
    uint32_t *codes_a, *codes_b, *order_a, *order_b;
    Aabb* boxes;
    SortWorkspace sort;
    static KnnWorkspace from_chunk(char*& chunk, size_t P)
    {
        KnnWorkspace w;
        carve(chunk, w.mm, 32);
        carve(chunk, w.codes_a, P);
        carve(chunk, w.codes_b, P);
        carve(chunk, w.order_a, P);
        carve(chunk, w.order_b, P);
        carve(chunk, w.boxes, (P + kBox - 1) / kBox + 1);
        w.sort = SortWorkspace::from_chunk(chunk, P);
        return w;
    }
};

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

size_t adgs_knn_workspace_bytes(int32_t P)
{
    char* c = nullptr;
    KnnWorkspace::from_chunk(c, (size_t)(P > 0 ? P : 0));
    return (size_t)c + 128;
}

int adgs_dist_cuda2(int32_t P, const float* points, float* mean_dist2, char* workspace, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0) return ADGS_ERR_ARG;
    if (P == 0) return ADGS_OK;
    if (!points || !mean_dist2 || !workspace) return ADGS_ERR_ARG;
    char* c = workspace;
    KnnWorkspace w = KnnWorkspace::from_chunk(c, (size_t)P);
    const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    cudaMemcpyAsync(w.mm, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    const int sms = device_info().sm_count;
    bounds_kernel<<<min(sms * 4, (P + 255) / 256), 256, 0, stream>>>(P, points, w.mm);
    morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, w.mm, w.codes_a);
    count_launch(2);
    const int passes = sort_pairs_async(w.codes_a, w.codes_b, w.order_a, w.order_b, (size_t)P, nullptr, 0, 30, w.sort,
                                        true, true, stream);
    const uint32_t* order = (passes & 1) ? w.order_b : w.order_a;
    const int num_boxes = (P + kBox - 1) / kBox;
    box_bounds_kernel<<<num_boxes, kBox, 0, stream>>>(P, points, order, w.boxes);
    knn3_kernel<<<num_boxes, kBox, 0, stream>>>(P, points, order, w.boxes, num_boxes, mean_dist2);
    count_launch(2);
    return check_stage("dist_cuda2", false, stream);
}

}  // extern "C"
