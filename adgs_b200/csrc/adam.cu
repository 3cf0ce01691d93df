// Fused Adam over the planar model storage: ONE launch updates every parameter group of
// GaussianModel.training_setup (scene/gaussian_model.py:337-372: torch.optim.Adam(l, lr=0.0,
// eps=1e-15), 18 groups, stepped by train.py:163-167).
//
// The reference's 18 tensors are 10 arrays here (adgs_model); groups that share an array differ
// only in their learning rate, which becomes a per-element rule of the segment:
//   xyz            rows [0, N_scene) -> scene_xyz lr, rest -> obj_xyz lr            (ADGS_ADAM_LR_SPLIT)
//   sh4 (12,N,4)   flattened (16,3) per Gaussian: floats 0..2 of chunk 0 are the DC
//                  coefficient (shs_dc lr), everything else shs_rest lr               (ADGS_ADAM_LR_SH4)
//   everything else one lr                                                           (ADGS_ADAM_LR_UNIFORM)
//
// Arithmetic = torch.optim.Adam (default betas, no weight decay, no amsgrad), per element:
//   m += (g - m) (1 - b1);  v = v b2 + (1 - b2) g g;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//
// HBM-bound: 16 B read + 12 B written per element. Window-aware gradients: for the control-point
// arrays (column-major planes) the caller may pass the set of columns the backward actually wrote
// (the B-spline window + Fourier columns of the views of this step); every other plane has an
// all-zero gradient under the reference's dense-Adam semantics, so the kernel takes g = 0 there
// WITHOUT reading it -- the result is bit-identical to reading a zero-filled gradient, and neither
// the zero-fill nor its read touch HBM.
#include "api_internal.cuh"

namespace adgs {
namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamVec = 4;                                   // float4 per thread
constexpr int kAdamChunk = kAdamThreads * kAdamVec * 4;       // floats per CTA

struct AdamArgs {
    adgs_adam_segment seg[ADGS_ADAM_MAX_SEGMENTS];
    long long first_chunk[ADGS_ADAM_MAX_SEGMENTS + 1];
    float neg_step_a[ADGS_ADAM_MAX_SEGMENTS];  // -(lr_a / (1 - b1^t)), evaluated in double on the host like torch
    float neg_step_b[ADGS_ADAM_MAX_SEGMENTS];
    int n_seg;
    float w1;        // 1 - b1
    float b2, w2;    // b2, 1 - b2
    float eps;
    float bc2_sqrt;  // sqrt(1 - b2^t)
};

// -step_size of element idx of segment si
__device__ __forceinline__ float seg_step(const AdamArgs& a, int si, long long idx)
{
    const adgs_adam_segment& s = a.seg[si];
    bool first = true;
    if (s.lr_rule == ADGS_ADAM_LR_SPLIT) first = idx < s.split;
    if (s.lr_rule == ADGS_ADAM_LR_SH4) first = idx < s.split && (idx & 3) != 3;
    return first ? a.neg_step_a[si] : a.neg_step_b[si];
}

// torch/optim/adam.py (_multi_tensor_adam): lerp_, mul_ + addcmul_, sqrt / bc2_sqrt + eps, addcdiv_
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float neg_step, const AdamArgs& a)
{
    m = fmaf(a.w1, g - m, m);
    v = fmaf(a.w2 * g, g, v * a.b2);
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = fmaf(neg_step, m / denom, p);
}

__global__ void __launch_bounds__(kAdamThreads) fused_adam_kernel(const __grid_constant__ AdamArgs a)
{
    // which segment does this CTA work on (<= 16 segments: a linear scan of kernel parameters)
    int si = 0;
    while (si + 1 < a.n_seg && (long long)blockIdx.x >= a.first_chunk[si + 1]) ++si;
    const adgs_adam_segment& s = a.seg[si];
    const long long base = ((long long)blockIdx.x - a.first_chunk[si]) * kAdamChunk;
    const bool aligned = ((reinterpret_cast<uintptr_t>(s.param) | reinterpret_cast<uintptr_t>(s.grad) |
                           reinterpret_cast<uintptr_t>(s.exp_avg) | reinterpret_cast<uintptr_t>(s.exp_avg_sq)) & 15) == 0;
#pragma unroll
    for (int it = 0; it < kAdamVec; ++it) {
        const long long i = base + ((long long)it * kAdamThreads + threadIdx.x) * 4;
        if (i >= s.n) break;
        // which of the four elements have a gradient: a plane the backward did not write holds arbitrary bytes and
        // counts as zero. With plane % 4 != 0 a float4 straddles two planes: decided per element, and the vector
        // load happens if ANY of them is live (the dead lanes are dropped by selection, never by arithmetic).
        bool hg[4] = {true, true, true, true};
        bool has_grad = true;
        if (s.plane > 0) {
            auto live = [&](long long c) -> bool { return c >= 128 || ((s.active[c >> 6] >> (c & 63)) & 1ull); };
            const long long c0 = i / s.plane, c3 = (i + 3) / s.plane;
            if (c0 == c3) {
                has_grad = live(c0);
                hg[0] = hg[1] = hg[2] = hg[3] = has_grad;
            } else {
                has_grad = false;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hg[e] = live((i + e) / s.plane);
                    has_grad |= hg[e];
                }
            }
        }
        if (aligned && i + 4 <= s.n) {
            float4 p = *reinterpret_cast<const float4*>(s.param + i);
            float4 m = *reinterpret_cast<const float4*>(s.exp_avg + i);
            float4 v = *reinterpret_cast<const float4*>(s.exp_avg_sq + i);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_grad) {
                g = __ldcs(reinterpret_cast<const float4*>(s.grad + i));  // read once: streaming
                g.x = hg[0] ? g.x : 0.f;
                g.y = hg[1] ? g.y : 0.f;
                g.z = hg[2] ? g.z : 0.f;
                g.w = hg[3] ? g.w : 0.f;
            }
            adam_update(p.x, g.x, m.x, v.x, seg_step(a, si, i), a);
            adam_update(p.y, g.y, m.y, v.y, seg_step(a, si, i + 1), a);
            adam_update(p.z, g.z, m.z, v.z, seg_step(a, si, i + 2), a);
            adam_update(p.w, g.w, m.w, v.w, seg_step(a, si, i + 3), a);
            *reinterpret_cast<float4*>(s.param + i) = p;
            *reinterpret_cast<float4*>(s.exp_avg + i) = m;
            *reinterpret_cast<float4*>(s.exp_avg_sq + i) = v;
        } else {
            for (long long e = i; e < i + 4 && e < s.n; ++e) {
                float p = s.param[e], m = s.exp_avg[e], v = s.exp_avg_sq[e];
                bool hg = true;
                if (s.plane > 0) {
                    const long long c = e / s.plane;
                    if (c < 128) hg = (s.active[c >> 6] >> (c & 63)) & 1ull;
                }
                const float g = hg ? s.grad[e] : 0.f;
                adam_update(p, g, m, v, seg_step(a, si, e), a);
                s.param[e] = p;
                s.exp_avg[e] = m;
                s.exp_avg_sq[e] = v;
            }
        }
    }
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

int adgs_adam_step(const adgs_adam_segment* segments, int32_t num_segments, double beta1, double beta2, double eps,
                   int64_t step, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_segments < 0 || num_segments > ADGS_ADAM_MAX_SEGMENTS || (num_segments > 0 && !segments)) return ADGS_ERR_ARG;
    if (step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0)) return ADGS_ERR_ARG;
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    // scalars in double on the host, rounded to float once, as torch does with python-number hyper-parameters
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    long long chunks = 0;
    int n = 0;
    for (int i = 0; i < num_segments; ++i) {
        const adgs_adam_segment& s = segments[i];
        if (s.n < 0 || s.plane < 0) return ADGS_ERR_ARG;
        if (s.n == 0) continue;
        if (!s.param || !s.grad || !s.exp_avg || !s.exp_avg_sq) return ADGS_ERR_ARG;
        if (s.lr_rule < ADGS_ADAM_LR_UNIFORM || s.lr_rule > ADGS_ADAM_LR_SH4) return ADGS_ERR_ARG;
        a.seg[n] = s;
        a.neg_step_a[n] = (float)(s.lr_a / bc1 * -1.0);
        a.neg_step_b[n] = (float)(s.lr_b / bc1 * -1.0);
        a.first_chunk[n] = chunks;
        chunks += (s.n + kAdamChunk - 1) / kAdamChunk;
        ++n;
    }
    a.n_seg = n;
    a.first_chunk[n] = chunks;
    if (chunks == 0) return ADGS_OK;
    if (chunks > 0x7fffffffLL) return ADGS_ERR_UNSUPPORTED;
    a.w1 = (float)(1.0 - beta1);
    a.b2 = (float)beta2;
    a.w2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.bc2_sqrt = (float)sqrt(bc2);
    fused_adam_kernel<<<(unsigned)chunks, kAdamThreads, 0, stream>>>(a);
    count_launch(1);
    return check_stage("fused adam", false, stream);
}

}  // extern "C"
