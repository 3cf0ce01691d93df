// Loss front-end: the image term of train.py:79-80,113
//   (1 - lambda_dssim) * lambda_l1 * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
// (utils/loss_utils.py:20-58) as two fused kernels instead of five grouped 11x11 convolutions, a dozen
// element-wise kernels and their autograd mirror images.
//
// forward : one CTA per 16x16 pixel tile and channel. The image and ground-truth tiles (+5 pixel halo,
//           zero padded like F.conv2d(padding=5)) are staged in shared memory once; the separable
//           Gaussian window is applied horizontally to the five moment maps (x, y, x^2, y^2, xy) and then
//           vertically; SSIM and its partial derivatives w.r.t. the three x-dependent window moments
//           (mu1, E[x^2], E[xy]) are evaluated in registers. Outputs: the three derivative maps (what the
//           backward needs -- 12 B/pixel/channel instead of the ~60 B autograd saves) and per-CTA partial
//           sums of |x - y| and of the SSIM map (deterministic two-stage reduction).
// backward: d ssim_mean / d x(p) = sum_q w(q - p) [ f_mu(q) + 2 x(p) f_e1(q) + y(p) f_e12(q) ] / N, i.e. the
//           same separable window over the three derivative maps, plus the L1 sign term; the upstream
//           gradients of the two scalar outputs are read from DEVICE memory (no host round trip).
// HBM-bound: forward reads 8 and writes 12 B/pixel/channel, backward reads 20 and writes 4.
#include "api_internal.cuh"

namespace adgs {
namespace {

constexpr int kLT = 16;            // tile edge
constexpr int kLH = 5;             // halo = window_size / 2
constexpr int kLS = kLT + 2 * kLH; // staged edge (26)
constexpr float kC1 = 0.01f * 0.01f;
constexpr float kC2 = 0.03f * 0.03f;

struct Window {
    float g[11];
};

struct ImageLossFwdArgs {
    const float* img;
    const float* gt;
    float* f_mu;   // d ssim / d mu1      (C,H,W), may be null (evaluation only)
    float* f_e1;   // d ssim / d E[x^2]
    float* f_e12;  // d ssim / d E[xy]
    float* partial;  // [num_ctas][2]: sum |x - y|, sum ssim
    int C, H, W;
    Window w;
};

__global__ void __launch_bounds__(kLT * kLT) image_loss_forward_kernel(const ImageLossFwdArgs a)
{
    __shared__ float s_x[kLS][kLS + 1];
    __shared__ float s_y[kLS][kLS + 1];
    __shared__ float s_h[5][kLS][kLT + 1];  // horizontally filtered moments
    __shared__ float s_red[2][8];

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kLT + tx;
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const size_t plane = (size_t)a.H * a.W;
    const float* img = a.img + c * plane;
    const float* gt = a.gt + c * plane;

    for (int i = tid; i < kLS * kLS; i += kLT * kLT) {
        const int ly = i / kLS, lx = i - ly * kLS;
        const int gx = x0 + lx - kLH, gy = y0 + ly - kLH;
        float vx = 0.f, vy = 0.f;
        if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
            vx = img[(size_t)gy * a.W + gx];
            vy = gt[(size_t)gy * a.W + gx];
        }
        s_x[ly][lx] = vx;
        s_y[ly][lx] = vy;
    }
    __syncthreads();

    for (int i = tid; i < kLS * kLT; i += kLT * kLT) {
        const int ly = i / kLT, lx = i - ly * kLT;
        float m1 = 0.f, m2 = 0.f, e1 = 0.f, e2 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = a.w.g[k];
            const float x = s_x[ly][lx + k], y = s_y[ly][lx + k];
            m1 = fmaf(w, x, m1);
            m2 = fmaf(w, y, m2);
            e1 = fmaf(w, x * x, e1);
            e2 = fmaf(w, y * y, e2);
            e12 = fmaf(w, x * y, e12);
        }
        s_h[0][ly][lx] = m1;
        s_h[1][ly][lx] = m2;
        s_h[2][ly][lx] = e1;
        s_h[3][ly][lx] = e2;
        s_h[4][ly][lx] = e12;
    }
    __syncthreads();

    float mu1 = 0.f, mu2 = 0.f, e1 = 0.f, e2 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float w = a.w.g[k];
        mu1 = fmaf(w, s_h[0][ty + k][tx], mu1);
        mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
        e1 = fmaf(w, s_h[2][ty + k][tx], e1);
        e2 = fmaf(w, s_h[3][ty + k][tx], e2);
        e12 = fmaf(w, s_h[4][ty + k][tx], e12);
    }
    const int gx = x0 + tx, gy = y0 + ty;
    const bool inside = gx < a.W && gy < a.H;
    float ssim_v = 0.f, l1_v = 0.f;
    if (inside) {
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e1 - mu1_sq, s2 = e2 - mu2_sq, s12 = e12 - mu12;
        const float A = 2.f * mu12 + kC1, B = 2.f * s12 + kC2;
        const float Cc = mu1_sq + mu2_sq + kC1, D = s1 + s2 + kC2;
        const float inv_cd = 1.f / (Cc * D);
        ssim_v = A * B * inv_cd;
        l1_v = fabsf(s_x[ty + kLH][tx + kLH] - s_y[ty + kLH][tx + kLH]);
        if (a.f_mu) {
            const size_t o = c * plane + (size_t)gy * a.W + gx;
            // total derivative w.r.t. mu1 with s1 = e1 - mu1^2 and s12 = e12 - mu1 mu2 substituted
            a.f_mu[o] = 2.f * mu2 * (B - A) * inv_cd - ssim_v * 2.f * mu1 * (D - Cc) * inv_cd;
            a.f_e1[o] = -ssim_v / D;
            a.f_e12[o] = 2.f * A * inv_cd;
        }
    }
    // CTA sums
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ssim_v += __shfl_xor_sync(0xffffffffu, ssim_v, off);
        l1_v += __shfl_xor_sync(0xffffffffu, l1_v, off);
    }
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = l1_v;
        s_red[1][tid >> 5] = ssim_v;
    }
    __syncthreads();
    if (tid < 2) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_red[tid][w];
        const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        a.partial[cta * 2 + tid] = s;
    }
}

// out[0] = mean |x - y|, out[1] = mean ssim, out[2] = w_l1 out[0] + w_dssim (1 - out[1])
// (double accumulation, fixed order: deterministic)
__global__ void __launch_bounds__(256) image_loss_finalize_kernel(const float* partial, int num_ctas, double inv_n,
                                                                 float w_l1, float w_dssim, float* out)
{
    __shared__ double s_sum[2][8];
    double l1 = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < num_ctas; i += blockDim.x) {
        l1 += (double)partial[2 * i];
        ss += (double)partial[2 * i + 1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((threadIdx.x & 31) == 0) {
        s_sum[0][threadIdx.x >> 5] = l1;
        s_sum[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0, q = 0.0;
        for (int w = 0; w < 8; ++w) {
            l += s_sum[0][w];
            q += s_sum[1][w];
        }
        const float l1_mean = (float)(l * inv_n), ssim_mean = (float)(q * inv_n);
        out[0] = l1_mean;
        out[1] = ssim_mean;
        out[2] = w_l1 * l1_mean + w_dssim * (1.0f - ssim_mean);
    }
}

struct ImageLossBwdArgs {
    const float* img;
    const float* gt;
    const float* f_mu;
    const float* f_e1;
    const float* f_e12;
    const float* g_l1;    // device scalars: upstream gradients, multiplied by the host-side weights below
    const float* g_ssim;
    float w_l1, w_ssim;
    float* d_img;
    int C, H, W;
    float inv_n;
    Window w;
};

__global__ void __launch_bounds__(kLT * kLT) image_loss_backward_kernel(const ImageLossBwdArgs a)
{
    __shared__ float s_f[3][kLS][kLS + 1];
    __shared__ float s_h[3][kLS][kLT + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kLT + tx;
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const size_t plane = (size_t)a.H * a.W;
    const float* f0 = a.f_mu + c * plane;
    const float* f1 = a.f_e1 + c * plane;
    const float* f2 = a.f_e12 + c * plane;

    for (int i = tid; i < kLS * kLS; i += kLT * kLT) {
        const int ly = i / kLS, lx = i - ly * kLS;
        const int gx = x0 + lx - kLH, gy = y0 + ly - kLH;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
            const size_t o = (size_t)gy * a.W + gx;
            v0 = f0[o];
            v1 = f1[o];
            v2 = f2[o];
        }
        s_f[0][ly][lx] = v0;
        s_f[1][ly][lx] = v1;
        s_f[2][ly][lx] = v2;
    }
    __syncthreads();
    for (int i = tid; i < kLS * kLT; i += kLT * kLT) {
        const int ly = i / kLT, lx = i - ly * kLT;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = a.w.g[k];
            h0 = fmaf(w, s_f[0][ly][lx + k], h0);
            h1 = fmaf(w, s_f[1][ly][lx + k], h1);
            h2 = fmaf(w, s_f[2][ly][lx + k], h2);
        }
        s_h[0][ly][lx] = h0;
        s_h[1][ly][lx] = h1;
        s_h[2][ly][lx] = h2;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= a.W || gy >= a.H) return;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float w = a.w.g[k];
        c0 = fmaf(w, s_h[0][ty + k][tx], c0);
        c1 = fmaf(w, s_h[1][ty + k][tx], c1);
        c2 = fmaf(w, s_h[2][ty + k][tx], c2);
    }
    const size_t o = c * plane + (size_t)gy * a.W + gx;
    const float x = a.img[o], y = a.gt[o];
    const float g_l1 = a.g_l1[0] * a.w_l1, g_ssim = a.g_ssim[0] * a.w_ssim;
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    a.d_img[o] = (g_l1 * sgn + g_ssim * (c0 + 2.f * x * c1 + y * c2)) * a.inv_n;
}

Window make_window()
{
    // gaussian(11, 1.5) of utils/loss_utils.py:26-28 in the same precision: float32 values of the double
    // exponentials, normalised by their float32 sum
    Window w;
    float sum = 0.f;
    for (int k = 0; k < 11; ++k) {
        w.g[k] = (float)exp(-(double)((k - 5) * (k - 5)) / (2.0 * 1.5 * 1.5));
        sum += w.g[k];
    }
    for (int k = 0; k < 11; ++k) w.g[k] /= sum;
    return w;
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

size_t adgs_image_loss_partial_floats(int32_t C, int32_t H, int32_t W)
{
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)C * ((H + kLT - 1) / kLT) * ((W + kLT - 1) / kLT) * 2;
}

int adgs_image_loss_forward(int32_t C, int32_t H, int32_t W, const float* img, const float* gt, float* f_mu,
                            float* f_e1, float* f_e12, float* partial, float w_l1, float w_dssim, float* out3,
                            adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return ADGS_ERR_ARG;
    if (!img || !gt || !partial || !out3) return ADGS_ERR_ARG;
    if ((f_mu || f_e1 || f_e12) && !(f_mu && f_e1 && f_e12)) return ADGS_ERR_ARG;
    ImageLossFwdArgs a;
    a.img = img;
    a.gt = gt;
    a.f_mu = f_mu;
    a.f_e1 = f_e1;
    a.f_e12 = f_e12;
    a.partial = partial;
    a.C = C;
    a.H = H;
    a.W = W;
    a.w = make_window();
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, C);
    image_loss_forward_kernel<<<grid, dim3(kLT, kLT, 1), 0, stream>>>(a);
    const int ctas = (int)(grid.x * grid.y * grid.z);
    image_loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, ctas, 1.0 / ((double)C * H * W), w_l1, w_dssim, out3);
    count_launch(2);
    return check_stage("image loss forward", false, stream);
}

int adgs_image_loss_backward(int32_t C, int32_t H, int32_t W, const float* img, const float* gt, const float* f_mu,
                             const float* f_e1, const float* f_e12, const float* grad_l1, float w_l1,
                             const float* grad_ssim, float w_ssim, float* d_img, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return ADGS_ERR_ARG;
    if (!img || !gt || !f_mu || !f_e1 || !f_e12 || !grad_l1 || !grad_ssim || !d_img) return ADGS_ERR_ARG;
    ImageLossBwdArgs a;
    a.img = img;
    a.gt = gt;
    a.f_mu = f_mu;
    a.f_e1 = f_e1;
    a.f_e12 = f_e12;
    a.g_l1 = grad_l1;
    a.g_ssim = grad_ssim;
    a.w_l1 = w_l1;
    a.w_ssim = w_ssim;
    a.d_img = d_img;
    a.C = C;
    a.H = H;
    a.W = W;
    a.inv_n = (float)(1.0 / ((double)C * H * W));
    a.w = make_window();
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, C);
    image_loss_backward_kernel<<<grid, dim3(kLT, kLT, 1), 0, stream>>>(a);
    count_launch(1);
    return check_stage("image loss backward", false, stream);
}

}  // extern "C"

// ==========================================================================================
// Per-pixel terms of train.py:82-100,113-115: scale/shift-invariant depth L1 (utils/loss_utils.py:60-65 with
// utils/depth_utils.py:3-45, mask = None as train.py:86 calls it), object-mask and sky binary cross entropy on
// clipped predictions (train.py:91-99), and the optical-flow reprojection loss (utils/loss_utils.py:88-108 with
// utils/flow_utils.py:5-10). Three grid-stride passes over the H*W pixels, all sums in double with a fixed
// reduction order (deterministic), no host round trip: the reference's flow loss blocks on torch.nonzero.
//   pass 1: a00 = sum p^2, a01 = sum p, b0 = sum p g, b1 = sum g (depth least squares), the two BCE sums, the
//           flow sum and the number of selected flow pixels
//   pass 2: scale / shift from pass 1; sums of sign(r), sign(r) p, |r| for the residual r = s p + t - g
//   pass 3: the four cotangent planes (the gradient of the depth term includes the paths through s and t) and
//           the loss scalars
// ==========================================================================================
namespace adgs {
namespace {

constexpr int kPT = 256;          // threads per CTA
constexpr int kP1 = 8, kP2 = 3;   // doubles per CTA partial in pass 1 / pass 2

struct PixelLossArgs {
    adgs_pixel_loss_inputs in;
    double* partial1;   // [ctas][8]
    double* partial2;   // [ctas][3]
    const float* g_up;  // device scalar: upstream gradient of the weighted total
    float* d_depth;
    float* d_semantic;
    float* d_opacity;
    float* d_flow;
    float* out;         // 6 floats: depth, obj, sky, flow, weighted total, selected flow pixels
    int ctas;
};

template <int NV>
__device__ __forceinline__ void cta_reduce_store(double (&v)[NV], double* dst)
{
    __shared__ double s_part[NV][kPT / 32];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) s_part[i][threadIdx.x >> 5] = v[i];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kPT / 32; ++w) s += s_part[threadIdx.x][w];
        dst[threadIdx.x] = s;
    }
    __syncthreads();
}

// every CTA sums the per-CTA partials in the same fixed order -> identical, deterministic totals
template <int NV>
__device__ __forceinline__ void sum_partials(const double* partial, int ctas, double* s_tot)
{
    __shared__ double s_acc[NV][kPT / 32];
    double v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.0;
    for (int c = threadIdx.x; c < ctas; c += kPT)
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += partial[(size_t)c * NV + i];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) s_acc[i][threadIdx.x >> 5] = v[i];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kPT / 32; ++w) s += s_acc[threadIdx.x][w];
        s_tot[threadIdx.x] = s;
    }
    __syncthreads();
}

struct FlowTerm {
    bool selected;   // counted in the mean's denominator
    float weight;    // flow_vis (0/1) [* opacity] * (z > dist)
    float base;      // same without the opacity factor
    float term;      // |u - f0| / W + |v - f1| / H   (unweighted)
    float gu, gv;    // sign(u - f0) / W, sign(v - f1) / H
    float Px, Py, Pz;
};

__device__ __forceinline__ float sgnf(float d)
{
    return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
}

__device__ __forceinline__ FlowTerm flow_term(const adgs_pixel_loss_inputs& in, size_t i, size_t HW)
{
    FlowTerm f;
    const float f0 = in.flow[i], f1 = in.flow[HW + i];
    f.selected = (in.flow_vis[i] > 0.5f) && (f0 <= (float)in.W - 1.0f) && (f0 >= 0.0f) && (f1 <= (float)in.H - 1.0f) &&
                 (f1 >= 0.0f);
    f.weight = f.base = f.term = f.gu = f.gv = 0.f;
    f.Px = f.Py = 0.f;
    f.Pz = 1.f;
    if (!f.selected) return f;
    const float px = in.img_flow[i], py = in.img_flow[HW + i], pz = in.img_flow[2 * HW + i];
    // K @ (R @ p + T), utils/flow_utils.py:7
    const float qx = in.R[0] * px + in.R[1] * py + in.R[2] * pz + in.T[0];
    const float qy = in.R[3] * px + in.R[4] * py + in.R[5] * pz + in.T[1];
    const float qz = in.R[6] * px + in.R[7] * py + in.R[8] * pz + in.T[2];
    f.Px = in.K[0] * qx + in.K[1] * qy + in.K[2] * qz;
    f.Py = in.K[3] * qx + in.K[4] * qy + in.K[5] * qz;
    f.Pz = in.K[6] * qx + in.K[7] * qy + in.K[8] * qz;
    const bool front = f.Pz > in.flow_dist;
    const float z = fmaxf(f.Pz, in.flow_dist);
    const float u = f.Px / z, v = f.Py / z;
    f.base = front ? 1.f : 0.f;
    f.weight = f.base * (in.flow_opacity ? in.flow_opacity[i] : 1.f);
    const float du = u - f0, dv = v - f1;
    f.term = fabsf(du) / (float)in.W + fabsf(dv) / (float)in.H;
    f.gu = sgnf(du) / (float)in.W;
    f.gv = sgnf(dv) / (float)in.H;
    return f;
}

constexpr float kClipLo = 1e-3f, kClipHi = 1.0f - 1e-3f;

__global__ void __launch_bounds__(kPT) pixel_loss_pass1_kernel(const PixelLossArgs a)
{
    const adgs_pixel_loss_inputs& in = a.in;
    const size_t HW = (size_t)in.H * in.W;
    double v[kP1] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * kPT + threadIdx.x; i < HW; i += (size_t)gridDim.x * kPT) {
        if (in.gt_depth) {
            const double p = in.depth[i], g = in.gt_depth[i];
            v[0] += p * p;
            v[1] += p;
            v[2] += p * g;
            v[3] += g;
        }
        if (in.gt_semantic) {
            const float p = fminf(fmaxf(in.img_semantic[i], kClipLo), kClipHi);
            const bool t = in.gt_semantic[i] > 0.f;
            v[4] += (double)(t ? -fmaxf(logf(p), -100.f) : -fmaxf(logf(1.f - p), -100.f));
        }
        if (in.gt_sky) {
            const float p = 1.0f - fminf(fmaxf(in.img_opacity[i], kClipLo), kClipHi);
            const float y = in.gt_sky[i];
            v[5] += (double)(-(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f)));
        }
        if (in.flow) {
            const FlowTerm f = flow_term(in, i, HW);
            if (f.selected) {
                v[6] += (double)(f.term * f.weight);
                v[7] += 1.0;
            }
        }
    }
    cta_reduce_store<kP1>(v, a.partial1 + (size_t)blockIdx.x * kP1);
}

struct DepthFit {
    double s, t, det, a00, a01, a11, b0, b1;
};

__device__ __forceinline__ DepthFit depth_fit(const double* tot, double n)
{
    DepthFit d;
    d.a00 = tot[0];
    d.a01 = tot[1];
    d.a11 = n;
    d.b0 = tot[2];
    d.b1 = tot[3];
    d.det = d.a00 * d.a11 - d.a01 * d.a01;
    if (d.det == 0.0) {
        d.s = d.t = 0.0;  // utils/depth_utils.py:36-37
    } else {
        d.s = (d.a11 * d.b0 - d.a01 * d.b1) / d.det;
        d.t = (-d.a01 * d.b0 + d.a00 * d.b1) / d.det;
    }
    return d;
}

__global__ void __launch_bounds__(kPT) pixel_loss_pass2_kernel(const PixelLossArgs a)
{
    __shared__ double s_tot[kP1];
    const adgs_pixel_loss_inputs& in = a.in;
    const size_t HW = (size_t)in.H * in.W;
    sum_partials<kP1>(a.partial1, a.ctas, s_tot);
    double v[kP2] = {0, 0, 0};
    if (in.gt_depth) {
        const DepthFit d = depth_fit(s_tot, (double)HW);
        const float s = (float)d.s, t = (float)d.t;
        for (size_t i = (size_t)blockIdx.x * kPT + threadIdx.x; i < HW; i += (size_t)gridDim.x * kPT) {
            const float p = in.depth[i];
            const float r = s * p + t - in.gt_depth[i];
            const float sg = sgnf(r);
            v[0] += (double)sg;
            v[1] += (double)(sg * p);
            v[2] += (double)fabsf(r);
        }
    }
    cta_reduce_store<kP2>(v, a.partial2 + (size_t)blockIdx.x * kP2);
}

__global__ void __launch_bounds__(kPT) pixel_loss_pass3_kernel(const PixelLossArgs a)
{
    __shared__ double s_tot[kP1];
    __shared__ double s_tot2[kP2];
    const adgs_pixel_loss_inputs& in = a.in;
    const size_t HW = (size_t)in.H * in.W;
    sum_partials<kP1>(a.partial1, a.ctas, s_tot);
    sum_partials<kP2>(a.partial2, a.ctas, s_tot2);
    const double n = (double)HW;
    const DepthFit d = depth_fit(s_tot, n);
    const double S0 = s_tot2[0], S1 = s_tot2[1];
    const double n_sel = s_tot[7];
    const float g = a.g_up ? a.g_up[0] : 1.f;
    const float w_depth = g * in.lambda_depth, w_obj = g * in.lambda_obj, w_sky = g * in.lambda_sky;
    const float w_flow = (n_sel > 0.0) ? g * in.lambda_flow / (float)n_sel : 0.f;
    const float inv_hw = (float)(1.0 / n);

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float depth_l = in.gt_depth ? (float)(s_tot2[2] / n) : 0.f;
        const float obj_l = in.gt_semantic ? (float)(s_tot[4] / n) : 0.f;
        const float sky_l = in.gt_sky ? (float)(s_tot[5] / n) : 0.f;
        const float flow_l = (in.flow && n_sel > 0.0) ? (float)(s_tot[6] / n_sel) : 0.f;
        a.out[0] = depth_l;
        a.out[1] = obj_l;
        a.out[2] = sky_l;
        a.out[3] = flow_l;
        a.out[4] = in.lambda_depth * depth_l + in.lambda_obj * obj_l + in.lambda_sky * sky_l + in.lambda_flow * flow_l;
        a.out[5] = (float)n_sel;
    }

    if (!a.d_depth && !a.d_semantic && !a.d_opacity && !a.d_flow) return;
    for (size_t i = (size_t)blockIdx.x * kPT + threadIdx.x; i < HW; i += (size_t)gridDim.x * kPT) {
        if (a.d_depth) {
            float gd = 0.f;
            if (in.gt_depth && d.det != 0.0) {
                const double p = in.depth[i], gt = in.gt_depth[i];
                const float r = (float)d.s * (float)p + (float)d.t - (float)gt;
                const double ddet = 2.0 * p * d.a11 - 2.0 * d.a01;
                const double ds = (d.a11 * gt - d.b1 - d.s * ddet) / d.det;
                const double dt = (-d.b0 - d.a01 * gt + 2.0 * p * d.b1 - d.t * ddet) / d.det;
                gd = (float)((d.s * (double)sgnf(r) + ds * S1 + dt * S0) / n);
            }
            a.d_depth[i] = w_depth * gd;
        }
        if (a.d_semantic) {
            float gs = 0.f;
            if (in.gt_semantic) {
                const float x = in.img_semantic[i];
                if (x >= kClipLo && x <= kClipHi) gs = (in.gt_semantic[i] > 0.f) ? -1.f / x : 1.f / (1.f - x);
            }
            a.d_semantic[i] = w_obj * gs * inv_hw;
        }
        float go = 0.f;
        if (in.gt_sky) {
            const float x = in.img_opacity[i];
            if (x >= kClipLo && x <= kClipHi) {
                const float y = in.gt_sky[i];
                const float q = 1.0f - x;  // the "probability" handed to binary_cross_entropy
                // d/dq [-(y log q + (1-y) log(1-q))] = -y/q + (1-y)/(1-q);  dq/dx = -1
                go = w_sky * inv_hw * (y / q - (1.f - y) / (1.f - q));
            }
        }
        if (in.flow) {
            const FlowTerm f = flow_term(in, i, HW);
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (f.selected && f.base != 0.f) {
                if (in.flow_opacity) go += w_flow * f.term;  // the opacity is only a weight: d/d opacity = term
                const float gu = w_flow * f.weight * f.gu, gv = w_flow * f.weight * f.gv;
                const float inv_z = 1.f / f.Pz;
                const float gPx = gu * inv_z, gPy = gv * inv_z, gPz = -(gu * f.Px + gv * f.Py) * inv_z * inv_z;
                // q = R p + T, P = K q  =>  grad_p = R^T K^T grad_P
                const float gqx = in.K[0] * gPx + in.K[3] * gPy + in.K[6] * gPz;
                const float gqy = in.K[1] * gPx + in.K[4] * gPy + in.K[7] * gPz;
                const float gqz = in.K[2] * gPx + in.K[5] * gPy + in.K[8] * gPz;
                gx = in.R[0] * gqx + in.R[3] * gqy + in.R[6] * gqz;
                gy = in.R[1] * gqx + in.R[4] * gqy + in.R[7] * gqz;
                gz = in.R[2] * gqx + in.R[5] * gqy + in.R[8] * gqz;
            }
            if (a.d_flow) {
                a.d_flow[i] = gx;
                a.d_flow[HW + i] = gy;
                a.d_flow[2 * HW + i] = gz;
            }
        } else if (a.d_flow) {
            a.d_flow[i] = 0.f;
            a.d_flow[HW + i] = 0.f;
            a.d_flow[2 * HW + i] = 0.f;
        }
        if (a.d_opacity) a.d_opacity[i] = go;
    }
}

int pixel_loss_ctas(int H, int W)
{
    const long long px = (long long)H * W;
    long long c = (px + kPT * 4 - 1) / (kPT * 4);
    const long long cap = (long long)device_info().sm_count * 4;
    if (c > cap) c = cap;
    if (c < 1) c = 1;
    return (int)c;
}

}  // namespace
}  // namespace adgs

extern "C" {

size_t adgs_pixel_loss_scratch_bytes(int32_t H, int32_t W)
{
    if (H <= 0 || W <= 0) return 0;
    return (size_t)pixel_loss_ctas(H, W) * (kP1 + kP2) * sizeof(double) + 64;
}

int adgs_pixel_loss(const adgs_pixel_loss_inputs* in, int32_t phases, char* scratch, const float* grad_total,
                    float* d_depth, float* d_semantic, float* d_opacity, float* d_flow, float* out6,
                    adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!in || !scratch || !out6 || (phases & 3) == 0) return ADGS_ERR_ARG;
    if (in->H <= 0 || in->W <= 0) return ADGS_ERR_ARG;
    if (in->gt_depth && !in->depth) return ADGS_ERR_ARG;
    if (in->gt_semantic && !in->img_semantic) return ADGS_ERR_ARG;
    if (in->gt_sky && !in->img_opacity) return ADGS_ERR_ARG;
    if (in->flow && (!in->img_flow || !in->flow_vis)) return ADGS_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(scratch) & 7) != 0) return ADGS_ERR_ARG;
    PixelLossArgs a;
    a.in = *in;
    a.ctas = pixel_loss_ctas(in->H, in->W);
    a.partial1 = reinterpret_cast<double*>(scratch);
    a.partial2 = a.partial1 + (size_t)a.ctas * kP1;
    a.g_up = grad_total;
    a.d_depth = d_depth;
    a.d_semantic = d_semantic;
    a.d_opacity = d_opacity;
    a.d_flow = d_flow;
    a.out = out6;
    if (phases & 1) {
        pixel_loss_pass1_kernel<<<a.ctas, kPT, 0, stream>>>(a);
        pixel_loss_pass2_kernel<<<a.ctas, kPT, 0, stream>>>(a);
        count_launch(2);
    }
    const bool planes = d_depth || d_semantic || d_opacity || d_flow;
    // the scalars are written by pass 3 as well: without planes a single CTA is enough
    pixel_loss_pass3_kernel<<<((phases & 2) && planes) ? a.ctas : 1, kPT, 0, stream>>>(a);
    count_launch(1);
    return check_stage("pixel loss", false, stream);
}

}  // extern "C"
