// Environment map: the sky / far-field colour behind the Gaussians,
//   rendered = foreground + (1 - img_opacity) * sigmoid(grid_sample(grid_map, angles(ray)))
// replacing scene/env.py:get_image_cam_rays (:11-27), EnvironmentMap.get_image_background / get_env_color
// (:44-76), utils/graphics_utils.py:vector_to_theta (:95-100) and the composite of
// gaussian_renderer/__init__.py:92-94 -- and the two costs that dominate the reference's handling of its
// 8192 x 8192 x 3 map (arguments/__init__.py:66-67): the dense zero-filled gradient that grid_sample's
// autograd allocates every iteration (805 MB) and the dense torch.optim.Adam step over 201 M texels
// (scene/env.py:78-83, train.py:165), about 5.6 GB of traffic per iteration for a few hundred thousand
// texels that actually receive a gradient.
//
// forward : one thread per pixel: ray through the pixel centre, rotated by world_view_transform[:3,:3],
//           azimuth / elevation -> bilinear sample (align_corners=True, zero padding) of C channels ->
//           sigmoid -> background; optionally the composite in the same pass.
// backward: recomputes the sample position (nothing is saved), d foreground = g, d opacity = -sum_c g_c bg_c,
//           and scatters the texel gradients with RED.ADD into a PERSISTENT dense gradient buffer, marking the
//           32x32-texel tiles it touches.
// step    : the touched-tile map is compacted into a list; Adam runs over the tiles that have EVER been touched
//           (one CTA per tile), zeroing their gradient in the same pass. This is
//           exactly dense Adam: a texel that never received a gradient has m = v = g = 0, so its update is
//           0 / (0 + eps) = 0 and skipping it changes nothing; a tile touched once keeps being stepped (its
//           moments decay like in the dense optimizer).
#include "api_internal.cuh"

namespace adgs {
namespace {

constexpr int kEnvTile = 32;  // texels per tile edge (touched-tile bookkeeping)
constexpr float kPi = 3.14159265358979323846f;

struct EnvPixel {
    int x0, y0;        // top-left texel of the bilinear footprint
    float wx, wy;      // weights of the right / bottom neighbours
    bool in[4];        // footprint texels inside the map (zero padding outside): (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

// scene/env.py:11-27 + :60 + :65-71 + torch grid_sample(align_corners=True) un-normalisation
__device__ __forceinline__ EnvPixel env_pixel(int px, int py, int H, int W, float focal, const float* rot, int R)
{
    // K^-1 [x, y, 1], normalised
    float rx = ((float)px - 0.5f * (float)W) / focal, ry = ((float)py - 0.5f * (float)H) / focal, rz = 1.f;
    float inv = 1.f / fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);
    rx *= inv;
    ry *= inv;
    rz *= inv;
    // world_view_transform[:3,:3] @ ray   (row-major 4x4)
    float vx = rot[0] * rx + rot[1] * ry + rot[2] * rz;
    float vy = rot[4] * rx + rot[5] * ry + rot[6] * rz;
    float vz = rot[8] * rx + rot[9] * ry + rot[10] * rz;
    inv = 1.f / fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-12f);
    vx *= inv;
    vy *= inv;
    vz *= inv;
    const float az = atan2f(vy, vx);
    const float el = atan2f(vz, hypotf(vx, vy));
    const float gx = az * (1.0f / kPi), gy = el * (2.0f / kPi);  // self.scale, scene/env.py:35
    const float fx = (gx + 1.f) * 0.5f * (float)(R - 1), fy = (gy + 1.f) * 0.5f * (float)(R - 1);
    EnvPixel e;
    const float flx = floorf(fx), fly = floorf(fy);
    e.x0 = (int)flx;
    e.y0 = (int)fly;
    e.wx = fx - flx;
    e.wy = fy - fly;
    const bool xin0 = e.x0 >= 0 && e.x0 < R, xin1 = e.x0 + 1 >= 0 && e.x0 + 1 < R;
    const bool yin0 = e.y0 >= 0 && e.y0 < R, yin1 = e.y0 + 1 >= 0 && e.y0 + 1 < R;
    e.in[0] = yin0 && xin0;
    e.in[1] = yin0 && xin1;
    e.in[2] = yin1 && xin0;
    e.in[3] = yin1 && xin1;
    return e;
}

struct EnvFwdArgs {
    adgs_env_map env;
    int H, W;
    float focal;
    const float* view;        // device, 16 floats (world_view_transform)
    const float* foreground;  // (C,H,W) or null
    const float* opacity;     // (H,W) or null
    float* background;        // (C,H,W) or null
    float* rendered;          // (C,H,W) or null
};

__global__ void __launch_bounds__(256) env_forward_kernel(const EnvFwdArgs a)
{
    __shared__ float s_rot[12];
    if (threadIdx.x < 12) s_rot[threadIdx.x] = a.view[threadIdx.x];
    __syncthreads();
    const size_t HW = (size_t)a.H * a.W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int py = (int)(i / a.W), px = (int)(i - (size_t)py * a.W);
    const int R = a.env.R;
    const EnvPixel e = env_pixel(px, py, a.H, a.W, a.focal, s_rot, R);
    const float w00 = (1.f - e.wx) * (1.f - e.wy), w01 = e.wx * (1.f - e.wy), w10 = (1.f - e.wx) * e.wy, w11 = e.wx * e.wy;
    const float om = a.opacity ? 1.f - a.opacity[i] : 1.f;
    const size_t RR = (size_t)R * R;
    for (int c = 0; c < a.env.C; ++c) {
        const float* g = a.env.grid + c * RR;
        const size_t o = (size_t)e.y0 * R + e.x0;
        float v = 0.f;
        if (e.in[0]) v += w00 * g[o];
        if (e.in[1]) v += w01 * g[o + 1];
        if (e.in[2]) v += w10 * g[o + R];
        if (e.in[3]) v += w11 * g[o + R + 1];
        const float bg = 1.f / (1.f + expf(-v));
        if (a.background) a.background[c * HW + i] = bg;
        if (a.rendered) a.rendered[c * HW + i] = (a.foreground ? a.foreground[c * HW + i] : 0.f) + om * bg;
    }
}

struct EnvBwdArgs {
    adgs_env_map env;
    int H, W;
    float focal;
    const float* view;
    const float* opacity;      // (H,W) or null
    const float* g_rendered;   // (C,H,W) or null: cotangent of foreground + (1 - opacity) * background
    const float* g_background; // (C,H,W) or null: cotangent of the background itself
    float* d_opacity;          // (H,W) or null
};

__global__ void __launch_bounds__(256) env_backward_kernel(const EnvBwdArgs a)
{
    __shared__ float s_rot[12];
    if (threadIdx.x < 12) s_rot[threadIdx.x] = a.view[threadIdx.x];
    __syncthreads();
    const size_t HW = (size_t)a.H * a.W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const int py = (int)(i / a.W), px = (int)(i - (size_t)py * a.W);
    const int R = a.env.R;
    const EnvPixel e = env_pixel(px, py, a.H, a.W, a.focal, s_rot, R);
    const float w00 = (1.f - e.wx) * (1.f - e.wy), w01 = e.wx * (1.f - e.wy), w10 = (1.f - e.wx) * e.wy, w11 = e.wx * e.wy;
    const float om = a.opacity ? 1.f - a.opacity[i] : 1.f;
    const size_t RR = (size_t)R * R;
    const size_t o = (size_t)e.y0 * R + e.x0;
    float dop = 0.f;
    bool any = false;
    for (int c = 0; c < a.env.C; ++c) {
        const float* g = a.env.grid + c * RR;
        float v = 0.f;
        if (e.in[0]) v += w00 * g[o];
        if (e.in[1]) v += w01 * g[o + 1];
        if (e.in[2]) v += w10 * g[o + R];
        if (e.in[3]) v += w11 * g[o + R + 1];
        const float bg = 1.f / (1.f + expf(-v));
        float gbg = 0.f;
        if (a.g_rendered) {
            const float gr = a.g_rendered[c * HW + i];
            gbg = gr * om;
            dop -= gr * bg;
        }
        if (a.g_background) gbg += a.g_background[c * HW + i];
        const float gv = gbg * bg * (1.f - bg);
        if (gv != 0.f) {
            float* d = a.env.grad + c * RR;
            if (e.in[0]) red_add_f32(d + o, gv * w00);
            if (e.in[1]) red_add_f32(d + o + 1, gv * w01);
            if (e.in[2]) red_add_f32(d + o + R, gv * w10);
            if (e.in[3]) red_add_f32(d + o + R + 1, gv * w11);
            any = true;
        }
    }
    if (a.d_opacity) a.d_opacity[i] = dop;
    if (any) {
        const int tiles = (R + kEnvTile - 1) / kEnvTile;
        // benign race: every writer stores the same byte
        if (e.in[0]) a.env.touched[(size_t)(e.y0 / kEnvTile) * tiles + e.x0 / kEnvTile] = 1;
        if (e.in[1]) a.env.touched[(size_t)(e.y0 / kEnvTile) * tiles + (e.x0 + 1) / kEnvTile] = 1;
        if (e.in[2]) a.env.touched[(size_t)((e.y0 + 1) / kEnvTile) * tiles + e.x0 / kEnvTile] = 1;
        if (e.in[3]) a.env.touched[(size_t)((e.y0 + 1) / kEnvTile) * tiles + (e.x0 + 1) / kEnvTile] = 1;
    }
}

struct EnvAdamArgs {
    adgs_env_map env;
    float w1, b2, w2, eps, bc2_sqrt, neg_step;
};

// Compaction of the touched-tile map: tile_list[0] = count, tile_list[1 + i] = tile id (any order).
__global__ void __launch_bounds__(256) env_compact_kernel(const uint8_t* __restrict__ touched, int num_tiles,
                                                          uint32_t* __restrict__ tile_list)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = t < num_tiles && touched[t] != 0;
    const uint32_t mask = __ballot_sync(0xffffffffu, on);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(tile_list, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) tile_list[1 + base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)t;
}

// One CTA per touched tile (grid-stride over the compacted list, whose length lives on the device): warp w
// takes texel rows w, w+8, w+16, w+24 of the tile, lanes are the 32 columns -> 128-byte lines, and the
// 4 rows x C channels of a thread are independent loads (memory-level parallelism instead of a serial walk).
__global__ void __launch_bounds__(256) env_adam_kernel(const EnvAdamArgs a)
{
    const int R = a.env.R;
    const int tiles = (R + kEnvTile - 1) / kEnvTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t RR = (size_t)R * R;
    const uint32_t count = a.env.tile_list[0];
    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        const uint32_t t = a.env.tile_list[1 + i];
        const int ty = (int)(t / tiles), tx = (int)(t - (uint32_t)ty * tiles);
        const int x = tx * kEnvTile + lane;
        if (x >= R) continue;
#pragma unroll
        for (int rr = 0; rr < kEnvTile / 8; ++rr) {
            const int y = ty * kEnvTile + warp + 8 * rr;
            if (y >= R) continue;
            for (int c = 0; c < a.env.C; ++c) {
                const size_t o = c * RR + (size_t)y * R + x;
                const float g = a.env.grad[o];
                float m = a.env.exp_avg[o], v = a.env.exp_avg_sq[o];
                if (g == 0.f && m == 0.f && v == 0.f) continue;  // never-touched texel inside a touched tile: exact no-op
                float p = a.env.grid[o];
                m = fmaf(a.w1, g - m, m);
                v = fmaf(a.w2 * g, g, v * a.b2);
                const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
                p = fmaf(a.neg_step, m / denom, p);
                a.env.grid[o] = p;
                a.env.exp_avg[o] = m;
                a.env.exp_avg_sq[o] = v;
                if (g != 0.f) a.env.grad[o] = 0.f;
            }
        }
    }
}

int check_env(const adgs_env_map* env)
{
    if (!env || env->R < 2 || env->C < 1 || env->C > 16 || !env->grid) return ADGS_ERR_ARG;
    return ADGS_OK;
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

size_t adgs_env_touched_bytes(int32_t R)
{
    if (R <= 0) return 0;
    const size_t t = (size_t)((R + kEnvTile - 1) / kEnvTile);
    return t * t;
}

size_t adgs_env_tile_list_bytes(int32_t R)
{
    return (adgs_env_touched_bytes(R) + 1) * sizeof(uint32_t);
}

int adgs_env_forward(const adgs_env_map* env, int32_t H, int32_t W, float focal, const float* world_view_transform,
                     const float* foreground, const float* img_opacity, float* background, float* rendered,
                     adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = check_env(env);
    if (st) return st;
    if (H <= 0 || W <= 0 || !(focal > 0.f) || !world_view_transform || (!background && !rendered)) return ADGS_ERR_ARG;
    EnvFwdArgs a;
    a.env = *env;
    a.H = H;
    a.W = W;
    a.focal = focal;
    a.view = world_view_transform;
    a.foreground = foreground;
    a.opacity = img_opacity;
    a.background = background;
    a.rendered = rendered;
    const size_t HW = (size_t)H * W;
    env_forward_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(a);
    count_launch(1);
    return check_stage("env forward", false, stream);
}

int adgs_env_backward(const adgs_env_map* env, int32_t H, int32_t W, float focal, const float* world_view_transform,
                      const float* img_opacity, const float* g_rendered, const float* g_background, float* d_opacity,
                      adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = check_env(env);
    if (st) return st;
    if (H <= 0 || W <= 0 || !(focal > 0.f) || !world_view_transform || !env->grad || !env->touched) return ADGS_ERR_ARG;
    if (!g_rendered && !g_background) return ADGS_ERR_ARG;
    EnvBwdArgs a;
    a.env = *env;
    a.H = H;
    a.W = W;
    a.focal = focal;
    a.view = world_view_transform;
    a.opacity = img_opacity;
    a.g_rendered = g_rendered;
    a.g_background = g_background;
    a.d_opacity = d_opacity;
    const size_t HW = (size_t)H * W;
    env_backward_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, stream>>>(a);
    count_launch(1);
    return check_stage("env backward", false, stream);
}

int adgs_env_adam_step(const adgs_env_map* env, double lr, double beta1, double beta2, double eps, int64_t step,
                       adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = check_env(env);
    if (st) return st;
    if (!env->grad || !env->exp_avg || !env->exp_avg_sq || !env->touched || !env->tile_list) return ADGS_ERR_ARG;
    if (step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0)) return ADGS_ERR_ARG;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    EnvAdamArgs a;
    a.env = *env;
    a.w1 = (float)(1.0 - beta1);
    a.b2 = (float)beta2;
    a.w2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.bc2_sqrt = (float)sqrt(bc2);
    a.neg_step = (float)(lr / bc1 * -1.0);
    const int tiles = (env->R + kEnvTile - 1) / kEnvTile;
    const int num_tiles = tiles * tiles;
    cudaMemsetAsync(env->tile_list, 0, sizeof(uint32_t), stream);
    env_compact_kernel<<<(num_tiles + 255) / 256, 256, 0, stream>>>(env->touched, num_tiles, env->tile_list);
    const int grid = min(num_tiles, device_info().sm_count * 8);
    env_adam_kernel<<<grid, 256, 0, stream>>>(a);
    count_launch(2);
    return check_stage("env adam", false, stream);
}

}  // extern "C"
