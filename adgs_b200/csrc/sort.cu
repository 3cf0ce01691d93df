// Onesweep radix sort + decoupled-look-back scan for sm_100a. See sort.cuh.
//
// Design (B200): 148 SMs x 2 resident CTAs of 256 threads pull 4096-pair tiles from an atomic
// ticket (so a CTA only ever waits on tiles that are already running); keys are read
// warp-striped (128-byte lines per warp instruction), ranked with __match_any_sync against
// per-warp digit counters in shared memory (stable within the tile), exchanged through shared
// memory, and written back in digit-contiguous runs. One 32-bit status word per (tile, digit)
// carries {flag, count} so the look-back needs no fences. HBM traffic per pass is the
// algorithmic 16 B/pair (8 read + 8 written) plus 1 KB of status per 4096 pairs.
#include "sort.cuh"

namespace adgs {
void count_launch(int n);

const DeviceInfo& device_info()
{
    static DeviceInfo info = [] {
        DeviceInfo d;
        int dev = 0;
        cudaGetDevice(&dev);
        d.sm_count = 148;
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (d.sm_count <= 0) d.sm_count = 148;
        return d;
    }();
    return info;
}

namespace {

constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
{
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= (uint32_t)off) v += o;
    }
    return v;
}

// Exclusive scan over the 256 threads of a CTA. s_warp: >= 8 words. Returns the exclusive prefix
// and the CTA total through `total`.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_prefix = 0, t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t c = s_warp[w];
        if ((uint32_t)w < warp) warp_prefix += c;
        t += c;
    }
    total = t;
    __syncthreads();
    return warp_prefix + inc - v;
}

__global__ void __launch_bounds__(256) radix_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n_max,
                                                              const uint32_t* __restrict__ d_n, int begin_bit,
                                                              int end_bit, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    const int passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = d_n ? min(*d_n, n_max) : n_max;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p) {
            if (p < passes) {
                const int shift = begin_bit + p * kRadixBits;
                const int bits = min(kRadixBits, end_bit - shift);
                const uint32_t d = (k >> shift) & ((1u << bits) - 1u);
                atomicAdd(&s_hist[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <bool IOTA>
__global__ void __launch_bounds__(kSortThreads, 4) onesweep_pass_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n_max,
    const uint32_t* __restrict__ d_n, int shift, uint32_t digit_mask, const uint32_t* __restrict__ hist,
    uint32_t* __restrict__ status, uint32_t* __restrict__ ticket)
{
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_warp_hist[kSortThreads / 32][kRadix];
    __shared__ uint32_t s_digit_start[kRadix];
    __shared__ uint32_t s_global_base[kRadix];
    __shared__ uint32_t s_hist_excl[kRadix];
    __shared__ uint32_t s_warp_scan[8];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = lanemask_lt();

    {
        uint32_t total;
        const uint32_t h = hist[tid];
        s_hist_excl[tid] = block_exclusive_scan_256(h, s_warp_scan, total);
    }

    const uint32_t n = d_n ? min(*d_n, n_max) : n_max;
    const uint32_t num_tiles = (n + kSortTile - 1) / kSortTile;

    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;
        const uint32_t base = tile * kSortTile;
        const uint32_t valid = min((uint32_t)kSortTile, n - base);

        uint32_t key[kSortItems], val[kSortItems];
        const uint32_t warp_base = warp * (kSortItems * 32);
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = warp_base + i * 32 + lane;
            if (idx < valid) {
                key[i] = keys_in[base + idx];
                val[i] = IOTA ? (base + idx) : vals_in[base + idx];
            } else {
                key[i] = 0xFFFFFFFFu;
                val[i] = 0;
            }
        }
#pragma unroll
        for (int j = 0; j < kRadix / 32; ++j) s_warp_hist[warp][j * 32 + lane] = 0;
        __syncwarp();

        uint32_t rank[kSortItems];
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const uint32_t d = (idx < valid) ? ((key[i] >> shift) & digit_mask) : digit_mask;
            const uint32_t m = __match_any_sync(0xffffffffu, d);
            const uint32_t lower = __popc(m & lt_mask);
            const uint32_t prev = s_warp_hist[warp][d];
            __syncwarp();
            if (lower == 0) s_warp_hist[warp][d] = prev + __popc(m);
            __syncwarp();
            rank[i] = prev + lower;
        }
        __syncthreads();

        // digit `tid`: exclusive scan of the per-warp counts, CTA count
        uint32_t block_count = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const uint32_t c = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = block_count;
            block_count += c;
        }
        uint32_t total;
        const uint32_t dstart = block_exclusive_scan_256(block_count, s_warp_scan, total);
        s_digit_start[tid] = dstart;

        // decoupled look-back for digit `tid`
        uint32_t* my_status = status + (size_t)tile * kRadix + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_volatile_u32(my_status, kFlagPrefix | block_count);
        } else {
            st_volatile_u32(my_status, kFlagAgg | block_count);
            // look back with four predecessor words in flight: the walk is a chain of L2 round trips,
            // and a tile deep in the grid may have to sum many aggregates before it meets a prefix
            int back = (int)tile - 1;  // nearest predecessor not yet accounted for
            bool done = false;
            while (!done) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = (back - j >= 0) ? ld_volatile_u32(status + (size_t)(back - j) * kRadix + tid) : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (done) break;
                    const uint32_t f = w[j] & kFlagMask;
                    if (f == 0) break;  // not published yet: reload from here
                    excl += w[j] & kValueMask;
                    --back;
                    if (f == kFlagPrefix) done = true;
                }
            }
            st_volatile_u32(my_status, kFlagPrefix | (excl + block_count));
        }
        s_global_base[tid] = s_hist_excl[tid] + excl - dstart;
        __syncthreads();

#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            const uint32_t idx = warp_base + i * 32 + lane;
            const uint32_t d = (idx < valid) ? ((key[i] >> shift) & digit_mask) : digit_mask;
            const uint32_t pos = s_digit_start[d] + s_warp_hist[warp][d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
        __syncthreads();

#pragma unroll
        for (int j = 0; j < kSortItems; ++j) {
            const uint32_t idx = j * kSortThreads + tid;
            if (idx < valid) {
                const uint32_t k = s_keys[idx];
                const uint32_t d = (k >> shift) & digit_mask;
                const uint32_t dst = s_global_base[d] + idx;
                keys_out[dst] = k;
                vals_out[dst] = s_vals[idx];
            }
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_gather_kernel(const uint32_t* __restrict__ in,
                                                                   const uint32_t* __restrict__ order,
                                                                   uint32_t* __restrict__ out, uint32_t n,
                                                                   uint32_t* __restrict__ status,
                                                                   uint32_t* __restrict__ total_out)
{
    __shared__ uint32_t s_warp_scan[8];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_prefix;
    const uint32_t tid = threadIdx.x;
    uint32_t* ticket = status;       // word 0
    uint32_t* st = status + 1;       // per-tile status words
    const uint32_t num_tiles = (n + kScanTile - 1) / kScanTile;
    if (n == 0) {
        if (blockIdx.x == 0 && tid == 0 && total_out) *total_out = 0;
        return;
    }
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;
        const uint32_t base = tile * kScanTile + tid * kScanItems;
        uint32_t v[kScanItems];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint32_t i = base + j;
            uint32_t x = 0;
            if (i < n) x = in[order ? order[i] : i];
            sum += x;
            v[j] = sum;
        }
        uint32_t total;
        const uint32_t excl = block_exclusive_scan_256(sum, s_warp_scan, total);
        if (tid < 32) {
            // warp-parallel look-back: lane l inspects tile - 1 - l of the current 32-tile window
            uint32_t prefix = 0;
            if (tile == 0) {
                if (tid == 0) st_volatile_u32(st + tile, kFlagPrefix | total);
            } else {
                if (tid == 0) st_volatile_u32(st + tile, kFlagAgg | total);
                int p = (int)tile - 1;
                while (true) {
                    const int q = p - (int)tid;
                    uint32_t sv = kFlagPrefix;  // lanes before tile 0 act as an empty prefix
                    if (q >= 0) {
                        do {
                            sv = ld_volatile_u32(st + q);
                        } while ((sv & kFlagMask) == 0);
                    }
                    const uint32_t has_prefix = __ballot_sync(0xffffffffu, (sv & kFlagMask) == kFlagPrefix);
                    const int stop = __ffs(has_prefix) - 1;  // nearest predecessor with an inclusive prefix (or -1)
                    uint32_t v = (stop < 0 || (int)tid <= stop) ? (sv & kValueMask) : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    prefix += v;
                    if (stop >= 0) break;
                    p -= 32;
                }
                if (tid == 0) st_volatile_u32(st + tile, kFlagPrefix | (prefix + total));
            }
            if (tid == 0) s_prefix = prefix;
        }
        __syncthreads();
        const uint32_t off = s_prefix + excl;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            const uint32_t i = base + j;
            if (i < n) {
                out[i] = off + v[j];
                if (i == n - 1 && total_out) *total_out = off + v[j];
            }
        }
    }
}

}  // namespace

int sort_pairs_async(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n_max,
                     const uint32_t* d_n, int begin_bit, int end_bit, const SortWorkspace& ws, bool iota_values,
                     bool clear_workspace, cudaStream_t stream)
{
    const int passes = sort_num_passes(begin_bit, end_bit);
    if (passes <= 0 || passes > kMaxPasses) return -1;
    if (n_max == 0) return passes;
    if (clear_workspace) cudaMemsetAsync(ws.hist, 0, ws.zero_bytes, stream);
    const int sms = device_info().sm_count;
    const size_t tiles = (n_max + kSortTile - 1) / kSortTile;
    {
        int grid = (int)min((size_t)sms * 4, (n_max + 256 * 16 - 1) / (256 * 16));
        if (grid < 1) grid = 1;
        count_launch(1);
        radix_histogram_kernel<<<grid, 256, 0, stream>>>(keys_a, (uint32_t)n_max, d_n, begin_bit, end_bit, ws.hist);
    }
    uint32_t* kin = keys_a;
    uint32_t* kout = keys_b;
    uint32_t* vin = vals_a;
    uint32_t* vout = vals_b;
    const int grid = (int)min((size_t)sms * 4, tiles);
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = min(kRadixBits, end_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        uint32_t* hist = ws.hist + (size_t)p * kRadix;
        uint32_t* status = ws.status + (size_t)p * ws.tiles * kRadix;
        uint32_t* ticket = ws.tickets + p;
        if (p == 0 && iota_values) {
            count_launch(1);
            onesweep_pass_kernel<true><<<grid, kSortThreads, 0, stream>>>(kin, nullptr, kout, vout, (uint32_t)n_max,
                                                                          d_n, shift, mask, hist, status, ticket);
        } else {
            count_launch(1);
            onesweep_pass_kernel<false><<<grid, kSortThreads, 0, stream>>>(kin, vin, kout, vout, (uint32_t)n_max, d_n,
                                                                           shift, mask, hist, status, ticket);
        }
        uint32_t* t = kin;
        kin = kout;
        kout = t;
        t = vin;
        vin = vout;
        vout = t;
    }
    return passes;
}

void inclusive_scan_gather_async(const uint32_t* in, const uint32_t* order, uint32_t* out, size_t n,
                                 uint32_t* status, uint32_t* total_out, cudaStream_t stream)
{
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    cudaMemsetAsync(status, 0, (tiles + 1) * sizeof(uint32_t), stream);
    const int sms = device_info().sm_count;
    int grid = (int)min((size_t)sms * 4, tiles);
    if (grid < 1) grid = 1;
    count_launch(1);
    scan_gather_kernel<<<grid, kScanThreads, 0, stream>>>(in, order, out, (uint32_t)n, status, total_out);
}

}  // namespace adgs
