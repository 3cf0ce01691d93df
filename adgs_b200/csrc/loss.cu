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
