// Peer memory for the splat exchange (multi-GPU, one process per GPU on one NVSwitch box): buffers that every
// rank of the box can address with plain device pointers, and a stream-ordered barrier over them.
//
// The exchange kernels (adgs_shard_forward_multi / adgs_shard_backward_multi) take one pointer set PER VIEW, so
// handing them pointers into the blending rank's memory turns their ordinary stores / loads into NVLink traffic
// that overlaps the arithmetic -- no collective, no staging copy. What is needed around them is (1) memory that a
// peer process can map (CUDA IPC) and (2) "every rank's kernel before this point has finished" on the stream:
// a one-CTA kernel that raises this rank's flag in every peer's flag array and spins until all peers raised theirs.
#include <cstdio>
#include "api_internal.cuh"

namespace adgs {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct BarrierArgs {
    uint32_t* flags[ADGS_MAX_PEERS];  // flags[p] = base of rank p's flag array ([ADGS_MAX_PEERS] words, zero-initialised)
    int world, rank;
    uint32_t epoch;
    uint32_t* status;  // local: [0] = 1 if a peer did not arrive within the timeout (the kernel then traps)
};

// Thread p: tell rank p that `rank` has reached `epoch` (everything this rank queued before the barrier is
// complete: the kernel runs after it in stream order), then wait until rank p said the same to us.
__global__ void peer_barrier_kernel(const BarrierArgs a)
{
    const int p = threadIdx.x;
    if (p >= a.world) return;
    __threadfence_system();
    st_release_sys(a.flags[p] + a.rank, a.epoch);
    const uint32_t* mine = a.flags[a.rank] + p;
    const long long t0 = clock64();
    // epochs only grow; a signed difference tolerates wrap-around
    while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
        if (clock64() - t0 > 40000000000ll) {  // ~20 s at 2 GHz: a peer died. Neither hang the box nor carry on
            a.status[0] = 1;                   // with half-exchanged data: record it and fail the context loudly
            __threadfence_system();
            __trap();
        }
        __nanosleep(64);
    }
}

struct SumArgs {
    const float* src[ADGS_MAX_PEERS];
    float* out;
    int world, n;
};

// out[i] = sum over ranks of src[rank][i] in rank order (the same order on every rank: identical results everywhere)
__global__ void peer_sum_kernel(const SumArgs a)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < a.world; ++p) s += a.src[p][i];
        a.out[i] = s;
    }
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

int adgs_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64)
{
    if (!ptr || !handle64 || bytes == 0) return ADGS_ERR_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return record_cuda_error(e, "peer_alloc");
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return record_cuda_error(e, "peer_alloc memset");
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return record_cuda_error(e, "peer_alloc ipc handle");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == ADGS_PEER_HANDLE_BYTES, "IPC handle size");
    memcpy(handle64, &h, sizeof(h));
    *ptr = p;
    return ADGS_OK;
}

int adgs_peer_open(const unsigned char* handle64, void** ptr)
{
    if (!ptr || !handle64) return ADGS_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return record_cuda_error(e, "peer_open");
    *ptr = p;
    return ADGS_OK;
}

int adgs_peer_close(void* ptr)
{
    if (!ptr) return ADGS_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? ADGS_OK : record_cuda_error(e, "peer_close");
}

int adgs_peer_free(void* ptr)
{
    if (!ptr) return ADGS_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? ADGS_OK : record_cuda_error(e, "peer_free");
}

int adgs_peer_sum(int32_t world, const float* const* partials, int32_t n, float* out, adgs_stream_t stream)
{
    if (world < 1 || world > ADGS_MAX_PEERS || !partials || n < 0 || (n > 0 && !out)) return ADGS_ERR_ARG;
    if (n == 0) return ADGS_OK;
    SumArgs a;
    for (int p = 0; p < ADGS_MAX_PEERS; ++p) a.src[p] = p < world ? partials[p] : nullptr;
    for (int p = 0; p < world; ++p)
        if (!a.src[p]) return ADGS_ERR_ARG;
    a.out = out;
    a.world = world;
    a.n = n;
    count_launch(1);
    peer_sum_kernel<<<(n + 255) / 256 > 64 ? 64 : (n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    return check_stage("peer_sum", false, (cudaStream_t)stream);
}

int adgs_peer_barrier(int32_t world, int32_t rank, uint32_t* const* flag_arrays, uint32_t epoch, uint32_t* status,
                      adgs_stream_t stream)
{
    if (world < 1 || world > ADGS_MAX_PEERS || rank < 0 || rank >= world || !flag_arrays || !status) return ADGS_ERR_ARG;
    BarrierArgs a;
    for (int p = 0; p < ADGS_MAX_PEERS; ++p) a.flags[p] = p < world ? flag_arrays[p] : nullptr;
    for (int p = 0; p < world; ++p)
        if (!a.flags[p]) return ADGS_ERR_ARG;
    a.world = world;
    a.rank = rank;
    a.epoch = epoch;
    a.status = status;
    count_launch(1);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    return check_stage("peer_barrier", false, (cudaStream_t)stream);
}

}  // extern "C"
