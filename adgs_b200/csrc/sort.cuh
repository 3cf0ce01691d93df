// CUB-free stable LSD radix sort (onesweep: one histogram pass + one decoupled-look-back
// scatter pass per 8-bit digit) and a single-pass decoupled-look-back inclusive scan.
//
// Replaces cub::DeviceRadixSort::SortPairs / cub::DeviceScan::InclusiveSum at
// RZ/cuda_rasterizer/rasterizer_impl.cu:284,310-315 and KNN/simple_knn.cu:210-213.
//
// All kernels are persistent (grid sized to the SM count) and take the element count from
// DEVICE memory when `d_n` is non-null, so a sort over `num_rendered` instances can be queued
// without the host ever knowing num_rendered.
#pragma once
#include "common.cuh"

namespace adgs {

// Launch geometry shared by every persistent kernel in the library.
struct DeviceInfo {
    int sm_count;
};
const DeviceInfo& device_info();

// Queue a full sort of n (uint32 key, uint32 value) pairs over key bits [begin_bit, end_bit).
// keys ping-pong between (keys_a, keys_b), values between (vals_a, vals_b); with `iota_values`
// the first pass uses 0..n-1 as values instead of reading vals_a (which is still the pong buffer). `d_n` (device, optional) overrides `n_max` as the
// live element count; n_max still bounds it (arena capacity). Returns the number of passes p:
// the result is in the *_b buffers when p is odd, in *_a when p is even.
int sort_pairs_async(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                     size_t n_max, const uint32_t* d_n, int begin_bit, int end_bit,
                     const SortWorkspace& ws, bool iota_values, bool clear_workspace,
                     cudaStream_t stream);

inline int sort_num_passes(int begin_bit, int end_bit)
{
    return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
}

// Inclusive scan out[i] = sum_{j<=i} in[order ? order[j] : j]. status: >= tiles+1 words, zeroed
// by the call. Also writes the total to *total_out (device) if non-null.
void inclusive_scan_gather_async(const uint32_t* in, const uint32_t* order, uint32_t* out, size_t n,
                                 uint32_t* status, uint32_t* total_out, cudaStream_t stream);

}  // namespace adgs
