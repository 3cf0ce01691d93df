// Per-Gaussian geometry, SH colour and their backward passes as register-level device functions,
// shared by the strict drop-in rasterizer kernels (AoS inputs) and the fused
// trajectory+projection kernels (planar model storage).
//
// Equations follow SURVEY.md Appendix A.2 / A.6, i.e. the behaviour of
//   RZ/cuda_rasterizer/forward.cu:20-256 (computeColorFromSH, computeCov2D, computeCov3D, preprocessCUDA)
//   RZ/cuda_rasterizer/backward.cu:20-414 (their backward counterparts)
//   RZ/cuda_rasterizer/auxiliary.h:41-164 (ndc2Pix, getRect, transforms, in_frustum)
// The integer-deciding chain (cull, radius, tile rectangle, depth key) keeps the reference's
// floating-point evaluation order so those outputs are bit-exact.
#pragma once
#include "common.cuh"

namespace adgs {

struct RasterParams {
    int W, H;
    int grid_x, grid_y;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    float scale_modifier;
    int sh_degree;
    int inv_depth;
    int prefiltered;
};

struct SplatGeom {
    float depth;     // view-space z
    float px, py;    // pixel-space mean
    float conic_x, conic_y, conic_z;
    int radius;
    uint32_t tiles;
    float cov3D[6];
};

// y = M[:, :3] x + M[:, 3] for the transposed-in-memory 4x4 (m[col*4+row]); contraction pinned (dot3_pinned).
__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* m)
{
    float3 r;
    r.x = __fadd_rn(dot3_pinned(m[0], p.x, m[4], p.y, m[8], p.z), m[12]);
    r.y = __fadd_rn(dot3_pinned(m[1], p.x, m[5], p.y, m[9], p.z), m[13]);
    r.z = __fadd_rn(dot3_pinned(m[2], p.x, m[6], p.y, m[10], p.z), m[14]);
    return r;
}

__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* m)
{
    float4 r;
    r.x = __fadd_rn(dot3_pinned(m[0], p.x, m[4], p.y, m[8], p.z), m[12]);
    r.y = __fadd_rn(dot3_pinned(m[1], p.x, m[5], p.y, m[9], p.z), m[13]);
    r.z = __fadd_rn(dot3_pinned(m[2], p.x, m[6], p.y, m[10], p.z), m[14]);
    r.w = __fadd_rn(dot3_pinned(m[3], p.x, m[7], p.y, m[11], p.z), m[15]);
    return r;
}

// Rotation matrix of the UN-normalised quaternion (w,x,y,z) in column-major storage, with the
// element placement of the reference's constructor call (forward.cu:134-138). Every product of two
// quaternion components feeds two entries (x y in 2(xy - rz) and 2(xy + rz), ...); which of the two products of
// an entry nvcc rounds on its own and which it fuses is pinned here to what it emits for the reference's
// expression in preprocessCUDA (read off the SASS; tests/test_parity_gpu.py holds cov3D / conic bit-exact).
__device__ __forceinline__ Mat3 quat_to_mat3(float r, float x, float y, float z)
{
    const float zx = __fmul_rn(z, x), xr = __fmul_rn(x, r), zr = __fmul_rn(z, r);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    auto twice = [](float v) { return __fadd_rn(v, v); };
    auto one_minus = [](float v) { return __fadd_rn(-v, 1.f); };
    Mat3 R;
    R.c[0][0] = one_minus(twice(__fadd_rn(yy, zz)));
    R.c[0][1] = twice(__fmaf_rn(y, x, -zr));
    R.c[0][2] = twice(__fmaf_rn(y, r, zx));
    R.c[1][0] = twice(__fmaf_rn(y, x, zr));
    R.c[1][1] = one_minus(twice(__fmaf_rn(x, x, zz)));
    R.c[1][2] = twice(__fmaf_rn(z, y, -xr));
    R.c[2][0] = twice(__fmaf_rn(y, -r, zx));
    R.c[2][1] = twice(__fmaf_rn(z, y, xr));
    R.c[2][2] = one_minus(twice(__fmaf_rn(x, x, yy)));
    return R;
}

__device__ __forceinline__ Mat3 scale_mat3(float sx, float sy, float sz)
{
    Mat3 S;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) S.c[c][r] = 0.f;
    S.c[0][0] = sx;
    S.c[1][1] = sy;
    S.c[2][2] = sz;
    return S;
}

// Sigma = (S R)^T (S R), upper triangle (forward.cu:118-152).
__device__ __forceinline__ void cov3d_from_scale_rot(const float* scale, float mod, const float* rot, float* cov3D)
{
    const Mat3 S = scale_mat3(mod * scale[0], mod * scale[1], mod * scale[2]);
    const Mat3 R = quat_to_mat3(rot[0], rot[1], rot[2], rot[3]);
    const Mat3 M = mat3_mul(S, R);
    const Mat3 Sigma = mat3_mul(mat3_transpose(M), M);
    cov3D[0] = Sigma.c[0][0];
    cov3D[1] = Sigma.c[0][1];
    cov3D[2] = Sigma.c[0][2];
    cov3D[3] = Sigma.c[1][1];
    cov3D[4] = Sigma.c[1][2];
    cov3D[5] = Sigma.c[2][2];
}

// EWA projection intermediates shared by forward and backward.
struct Cov2DCtx {
    float3 t;          // clamped view-space point
    float txtz, tytz;  // unclamped ratios
    Mat3 T;            // W * J
    Mat3 Vrk;
    Mat3 Wm;
    float a, b, c;     // cov2D + low-pass
};

__device__ __forceinline__ void cov2d_project(const float3& mean, const RasterParams& rp, const float* cov3D,
                                              const float* view, Cov2DCtx& o)
{
    float3 t = xform_point_4x3(mean, view);
    const float limx = 1.3f * rp.tan_fovx;
    const float limy = 1.3f * rp.tan_fovy;
    o.txtz = t.x / t.z;
    o.tytz = t.y / t.z;
    t.x = min(limx, max(-limx, o.txtz)) * t.z;
    t.y = min(limy, max(-limy, o.tytz)) * t.z;
    o.t = t;

    Mat3 J;
    J.c[0][0] = rp.focal_x / t.z;
    J.c[0][1] = 0.0f;
    J.c[0][2] = -(rp.focal_x * t.x) / (t.z * t.z);
    J.c[1][0] = 0.0f;
    J.c[1][1] = rp.focal_y / t.z;
    J.c[1][2] = -(rp.focal_y * t.y) / (t.z * t.z);
    J.c[2][0] = 0.f;
    J.c[2][1] = 0.f;
    J.c[2][2] = 0.f;

    Mat3& Wm = o.Wm;
    Wm.c[0][0] = view[0];
    Wm.c[0][1] = view[4];
    Wm.c[0][2] = view[8];
    Wm.c[1][0] = view[1];
    Wm.c[1][1] = view[5];
    Wm.c[1][2] = view[9];
    Wm.c[2][0] = view[2];
    Wm.c[2][1] = view[6];
    Wm.c[2][2] = view[10];

    o.T = mat3_mul(Wm, J);

    Mat3& V = o.Vrk;
    V.c[0][0] = cov3D[0];
    V.c[0][1] = cov3D[1];
    V.c[0][2] = cov3D[2];
    V.c[1][0] = cov3D[1];
    V.c[1][1] = cov3D[3];
    V.c[1][2] = cov3D[4];
    V.c[2][0] = cov3D[2];
    V.c[2][1] = cov3D[4];
    V.c[2][2] = cov3D[5];

    const Mat3 cov = mat3_mul(mat3_mul(mat3_transpose(o.T), mat3_transpose(V)), o.T);
    o.a = cov.c[0][0] + 0.3f;
    o.b = cov.c[0][1];
    o.c = cov.c[1][1] + 0.3f;
}

// Everything of preprocessCUDA (forward.cu:155-256) that decides visibility, radius and tiles.
// `scale`/`rot` may be null when cov3D_precomp is given. Returns false if culled.
__device__ __forceinline__ bool splat_geometry(const float3& p, const float* scale, const float* rot,
                                               const float* cov3D_precomp, const RasterParams& rp,
                                               const float* view, const float* proj, SplatGeom& g)
{
    const float4 p_hom = xform_point_4x4(p, proj);
    const float p_w = 1.0f / (p_hom.w + 0.0000001f);
    const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
    const float3 p_view = xform_point_4x3(p, view);
    if (p_view.z <= 0.2f) {
        if (rp.prefiltered) __trap();
        return false;
    }
    if (cov3D_precomp) {
#pragma unroll
        for (int i = 0; i < 6; ++i) g.cov3D[i] = cov3D_precomp[i];
    } else {
        cov3d_from_scale_rot(scale, rp.scale_modifier, rot, g.cov3D);
    }
    Cov2DCtx ctx;
    cov2d_project(p, rp, g.cov3D, view, ctx);
    const float cx = ctx.a, cy = ctx.b, cz = ctx.c;

    const float det = __fmaf_rn(cx, cz, -__fmul_rn(cy, cy));  // cx cz - cy cy, contraction pinned
    if (det == 0.0f) return false;
    const float det_inv = 1.f / det;
    g.conic_x = cz * det_inv;
    g.conic_y = -cy * det_inv;
    g.conic_z = cx * det_inv;

    const float mid = 0.5f * (cx + cz);
    const float disc = sqrtf(max(0.1f, __fmaf_rn(mid, mid, -det)));
    const float lambda1 = mid + disc;
    const float lambda2 = mid - disc;
    const float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
    g.px = ndc_to_pix(p_proj.x, rp.W);
    g.py = ndc_to_pix(p_proj.y, rp.H);
    uint32_t x0, y0, x1, y1;
    tile_rect(g.px, g.py, (int)my_radius, rp.grid_x, rp.grid_y, x0, y0, x1, y1);
    const uint32_t area = (x1 - x0) * (y1 - y0);
    if (area == 0) return false;
    g.depth = p_view.z;
    g.radius = (int)my_radius;
    g.tiles = area;
    return true;
}

// Degree <= 3 real SH -> RGB (+0.5, clamp at 0) with the per-channel clamp mask
// (forward.cu:20-71). sh: 48 floats, coefficient-major (l*3 + channel).
__device__ __forceinline__ void sh_to_rgb(int deg, const float3& pos, const float* campos, const float* sh,
                                          float* rgb, uint32_t& clamped_bits)
{
    // Sums of two or more products are written with explicit intrinsics (dot3_pinned & co.): which product nvcc
    // fuses into an FMA depends on the code around the expression, and the strict, fused and multi-view front ends
    // must produce the same colour bits as each other and as the reference's computeColorFromSH. The patterns are
    // the ones nvcc 12.9 emits for the reference expression (read off the SASS of the strict kernel).
    float dx = pos.x - campos[0], dy = pos.y - campos[1], dz = pos.z - campos[2];
    const float len = sqrtf(dot3_pinned(dx, dx, dy, dy, dz, dz));
    dx = dx / len;
    dy = dy / len;
    dz = dz / len;
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) res[c] = ADGS_SH_C0 * sh[c];
    if (deg > 0) {
        const float x = dx, y = dy, z = dz;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            res[c] = res[c] - ADGS_SH_C1 * y * sh[3 + c] + ADGS_SH_C1 * z * sh[6 + c] - ADGS_SH_C1 * x * sh[9 + c];
        if (deg > 1) {
            const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
            const float zz2 = __fadd_rn(zz, zz);                                    // 2 zz (exact)
            const float p20 = __fadd_rn(__fadd_rn(zz2, -xx), -yy);                  // 2zz - xx - yy
            const float xx_yy = __fadd_rn(xx, -yy);                                 // xx - yy
#pragma unroll
            for (int c = 0; c < 3; ++c)
                res[c] = res[c] + ADGS_SH_C2_0 * xy * sh[12 + c] + ADGS_SH_C2_1 * yz * sh[15 + c] +
                         ADGS_SH_C2_2 * p20 * sh[18 + c] + ADGS_SH_C2_3 * xz * sh[21 + c] +
                         ADGS_SH_C2_4 * xx_yy * sh[24 + c];
            if (deg > 2) {
                const float p3a = __fmaf_rn(xx, 3.0f, -yy);                         // 3xx - yy
                const float p3b = __fadd_rn(__fmaf_rn(zz, 4.0f, -xx), -yy);         // 4zz - xx - yy
                const float p3c = __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2));  // 2zz - 3xx - 3yy
                const float p3d = __fmaf_rn(yy, -3.0f, xx);                         // xx - 3yy
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    res[c] = res[c] + ADGS_SH_C3_0 * y * p3a * sh[27 + c] +
                             ADGS_SH_C3_1 * xy * z * sh[30 + c] +
                             ADGS_SH_C3_2 * y * p3b * sh[33 + c] +
                             ADGS_SH_C3_3 * z * p3c * sh[36 + c] +
                             ADGS_SH_C3_4 * x * p3b * sh[39 + c] +
                             ADGS_SH_C3_5 * z * xx_yy * sh[42 + c] +
                             ADGS_SH_C3_6 * x * p3d * sh[45 + c];
            }
        }
    }
    clamped_bits = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        res[c] += 0.5f;
        if (res[c] < 0) clamped_bits |= (1u << c);
        rgb[c] = max(res[c], 0.0f);
    }
}

// number of float4 chunks of the flattened (16,3) SH block needed for a degree
__device__ __host__ __forceinline__ int sh_chunks_for_degree(int deg)
{
    const int n = 3 * (deg + 1) * (deg + 1);
    return (n + 3) / 4;
}

// d(v/|v|)/dv applied to dv (auxiliary.h:107-117)
__device__ __forceinline__ float3 normalize_vjp(const float3& v, const float3& dv)
{
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// Backward of sh_to_rgb (backward.cu:20-139): writes dL_dsh[0..3*M) (zeros above the active
// degree) and returns the view-direction contribution to dL_dmean.
// dsh: 48 floats out (coefficient-major). M: coefficients present (<=16).
template <bool ACCUM>
__device__ __forceinline__ void sh_put(float* dsh, int i, float v)
{
    if (ACCUM) dsh[i] += v; else dsh[i] = v;
}

// ACCUM = false: dsh is overwritten (zeros above the active degree); ACCUM = true: dsh += (multi-view sums).
// BASIS = true: dsh receives the 16 basis values instead (zeros above the active degree) and the caller forms
// dL/dsh[3k + c] = dsh[k] * g[c] (g = clamp-masked dL_dcolor) when it stores, keeping 16 registers live instead of 48.
template <bool ACCUM = false, bool BASIS = false>
__device__ __forceinline__ float3 sh_to_rgb_bwd(int deg, const float3& pos, const float* campos, const float* sh,
                                                uint32_t clamped_bits, const float* dL_dcolor, float* dsh)
{
    const float3 dir_orig = make_float3(pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]);
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;

    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = dL_dcolor[c] * ((clamped_bits >> c) & 1u ? 0.f : 1.f);

    float dx[3] = {0.f, 0.f, 0.f}, dy[3] = {0.f, 0.f, 0.f}, dz[3] = {0.f, 0.f, 0.f};
    if (!ACCUM) {
#pragma unroll
        for (int i = 0; i < (BASIS ? 16 : 48); ++i) dsh[i] = 0.f;
    }
    // coefficient k, channel c, basis value bk
    auto emit = [&](int k, int c, float bk) {
        if (BASIS) {
            if (c == 0) dsh[k] = bk;
        } else {
            sh_put<ACCUM>(dsh, 3 * k + c, bk * g[c]);
        }
    };

#pragma unroll
    for (int c = 0; c < 3; ++c) emit(0, c, ADGS_SH_C0);
    if (deg > 0) {
        const float b1 = -ADGS_SH_C1 * y, b2 = ADGS_SH_C1 * z, b3 = -ADGS_SH_C1 * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            emit(1, c, b1);
            emit(2, c, b2);
            emit(3, c, b3);
            dx[c] = -ADGS_SH_C1 * sh[9 + c];
            dy[c] = -ADGS_SH_C1 * sh[3 + c];
            dz[c] = ADGS_SH_C1 * sh[6 + c];
        }
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
            const float b4 = ADGS_SH_C2_0 * xy, b5 = ADGS_SH_C2_1 * yz, b6 = ADGS_SH_C2_2 * (2.f * zz - xx - yy),
                        b7 = ADGS_SH_C2_3 * xz, b8 = ADGS_SH_C2_4 * (xx - yy);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                emit(4, c, b4);
                emit(5, c, b5);
                emit(6, c, b6);
                emit(7, c, b7);
                emit(8, c, b8);
                dx[c] += ADGS_SH_C2_0 * y * sh[12 + c] + ADGS_SH_C2_2 * 2.f * -x * sh[18 + c] +
                         ADGS_SH_C2_3 * z * sh[21 + c] + ADGS_SH_C2_4 * 2.f * x * sh[24 + c];
                dy[c] += ADGS_SH_C2_0 * x * sh[12 + c] + ADGS_SH_C2_1 * z * sh[15 + c] +
                         ADGS_SH_C2_2 * 2.f * -y * sh[18 + c] + ADGS_SH_C2_4 * 2.f * -y * sh[24 + c];
                dz[c] += ADGS_SH_C2_1 * y * sh[15 + c] + ADGS_SH_C2_2 * 2.f * 2.f * z * sh[18 + c] +
                         ADGS_SH_C2_3 * x * sh[21 + c];
            }
            if (deg > 2) {
                const float b9 = ADGS_SH_C3_0 * y * (3.f * xx - yy), b10 = ADGS_SH_C3_1 * xy * z,
                            b11 = ADGS_SH_C3_2 * y * (4.f * zz - xx - yy),
                            b12 = ADGS_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy),
                            b13 = ADGS_SH_C3_4 * x * (4.f * zz - xx - yy), b14 = ADGS_SH_C3_5 * z * (xx - yy),
                            b15 = ADGS_SH_C3_6 * x * (xx - 3.f * yy);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    emit(9, c, b9);
                    emit(10, c, b10);
                    emit(11, c, b11);
                    emit(12, c, b12);
                    emit(13, c, b13);
                    emit(14, c, b14);
                    emit(15, c, b15);
                    dx[c] += (ADGS_SH_C3_0 * sh[27 + c] * 3.f * 2.f * xy + ADGS_SH_C3_1 * sh[30 + c] * yz +
                              ADGS_SH_C3_2 * sh[33 + c] * -2.f * xy + ADGS_SH_C3_3 * sh[36 + c] * -3.f * 2.f * xz +
                              ADGS_SH_C3_4 * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) +
                              ADGS_SH_C3_5 * sh[42 + c] * 2.f * xz + ADGS_SH_C3_6 * sh[45 + c] * 3.f * (xx - yy));
                    dy[c] += (ADGS_SH_C3_0 * sh[27 + c] * 3.f * (xx - yy) + ADGS_SH_C3_1 * sh[30 + c] * xz +
                              ADGS_SH_C3_2 * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                              ADGS_SH_C3_3 * sh[36 + c] * -3.f * 2.f * yz + ADGS_SH_C3_4 * sh[39 + c] * -2.f * xy +
                              ADGS_SH_C3_5 * sh[42 + c] * -2.f * yz + ADGS_SH_C3_6 * sh[45 + c] * -3.f * 2.f * xy);
                    dz[c] += (ADGS_SH_C3_1 * sh[30 + c] * xy + ADGS_SH_C3_2 * sh[33 + c] * 4.f * 2.f * yz +
                              ADGS_SH_C3_3 * sh[36 + c] * 3.f * (2.f * zz - xx - yy) +
                              ADGS_SH_C3_4 * sh[39 + c] * 4.f * 2.f * xz + ADGS_SH_C3_5 * sh[42 + c] * (xx - yy));
                }
            }
        }
    }
    const float3 dL_ddir = make_float3(dx[0] * g[0] + dx[1] * g[1] + dx[2] * g[2],
                                       dy[0] * g[0] + dy[1] * g[1] + dy[2] * g[2],
                                       dz[0] * g[0] + dz[1] * g[1] + dz[2] * g[2]);
    return normalize_vjp(dir_orig, dL_ddir);
}

// conic -> cov2D -> cov3D / mean (backward.cu:144-274). dconic = (x, y, w) of the float4 record.
// Returns the covariance-path part of dL_dmean; writes dL_dcov3D[6].
__device__ __forceinline__ float3 cov2d_bwd(const float3& mean, const RasterParams& rp, const float* cov3D,
                                            const float* view, float dconic_x, float dconic_y, float dconic_w,
                                            float* dL_dcov)
{
    Cov2DCtx k;
    cov2d_project(mean, rp, cov3D, view, k);
    const float limx = 1.3f * rp.tan_fovx, limy = 1.3f * rp.tan_fovy;
    const float x_grad_mul = (k.txtz < -limx || k.txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (k.tytz < -limy || k.tytz > limy) ? 0.f : 1.f;
    const float a = k.a, b = k.b, c = k.c;
    const Mat3& T = k.T;
    const Mat3& V = k.Vrk;
    const Mat3& Wm = k.Wm;

    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dconic_x + 2 * b * c * dconic_y + (denom - a * c) * dconic_w);
        dL_dc = denom2inv * (-a * a * dconic_w + 2 * a * b * dconic_y + (denom - a * c) * dconic_x);
        dL_db = denom2inv * 2 * (b * c * dconic_x - (denom + 2 * b * b) * dconic_y + a * b * dconic_w);

        dL_dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
        dL_dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
        dL_dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
        dL_dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db +
                     2 * T.c[1][0] * T.c[1][1] * dL_dc;
        dL_dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db +
                     2 * T.c[1][0] * T.c[1][2] * dL_dc;
        dL_dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db +
                     2 * T.c[1][1] * T.c[1][2] * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) dL_dcov[i] = 0;
    }

    // gradient w.r.t. the upper 2x3 block of T
    float dT0[3], dT1[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float r0 = T.c[0][0] * V.c[j][0] + T.c[0][1] * V.c[j][1] + T.c[0][2] * V.c[j][2];
        const float r1 = T.c[1][0] * V.c[j][0] + T.c[1][1] * V.c[j][1] + T.c[1][2] * V.c[j][2];
        dT0[j] = 2 * r0 * dL_da + r1 * dL_db;
        dT1[j] = 2 * r1 * dL_dc + r0 * dL_db;
    }
    const float dJ00 = Wm.c[0][0] * dT0[0] + Wm.c[0][1] * dT0[1] + Wm.c[0][2] * dT0[2];
    const float dJ02 = Wm.c[2][0] * dT0[0] + Wm.c[2][1] * dT0[1] + Wm.c[2][2] * dT0[2];
    const float dJ11 = Wm.c[1][0] * dT1[0] + Wm.c[1][1] * dT1[1] + Wm.c[1][2] * dT1[2];
    const float dJ12 = Wm.c[2][0] * dT1[0] + Wm.c[2][1] * dT1[1] + Wm.c[2][2] * dT1[2];

    const float tz = 1.f / k.t.z;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;
    const float h_x = rp.focal_x, h_y = rp.focal_y;
    const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
    const float dty = y_grad_mul * -h_y * tz2 * dJ12;
    const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * k.t.x) * tz3 * dJ02 +
                      (2 * h_y * k.t.y) * tz3 * dJ12;
    // transpose of the 3x3 part of the view matrix applied to (dtx, dty, dtz)
    float3 r;
    r.x = view[0] * dtx + view[1] * dty + view[2] * dtz;
    r.y = view[4] * dtx + view[5] * dty + view[6] * dtz;
    r.z = view[8] * dtx + view[9] * dty + view[10] * dtz;
    return r;
}

// Projection + depth paths of dL_dmean (backward.cu:346-405).
__device__ __forceinline__ float3 mean_proj_depth_bwd(const float3& m, const float* view, const float* proj,
                                                      float dmean2D_x, float dmean2D_y, float ddepth, int inv_depth)
{
    const float4 m_hom = xform_point_4x4(m, proj);
    const float m_w = 1.0f / (m_hom.w + 0.0000001f);
    const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
    float3 d;
    d.x = (proj[0] * m_w - proj[3] * mul1) * dmean2D_x + (proj[1] * m_w - proj[3] * mul2) * dmean2D_y;
    d.y = (proj[4] * m_w - proj[7] * mul1) * dmean2D_x + (proj[5] * m_w - proj[7] * mul2) * dmean2D_y;
    d.z = (proj[8] * m_w - proj[11] * mul1) * dmean2D_x + (proj[9] * m_w - proj[11] * mul2) * dmean2D_y;

    const float mul3 = view[2] * m.x + view[6] * m.y + view[10] * m.z + view[14];
    const float demon = (inv_depth ? (-1.0f / (mul3 * mul3 + 0.0000001f)) : 1.0f);
    float3 d2;
    d2.x = (view[2] - view[3] * mul3) * ddepth * demon;
    d2.y = (view[6] - view[7] * mul3) * ddepth * demon;
    d2.z = (view[10] - view[11] * mul3) * ddepth * demon;
    return make_float3(d.x + d2.x, d.y + d2.y, d.z + d2.z);
}

// dL_dcov3D -> dL_dscale (3), dL_drot (4, w.r.t. the un-normalised quaternion the rasterizer was
// given; no normalisation Jacobian) (backward.cu:278-341).
__device__ __forceinline__ void cov3d_bwd(const float* scale, float mod, const float* rot, const float* dL_dcov3D,
                                          float* dL_dscale, float* dL_drot)
{
    const float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    const Mat3 R = quat_to_mat3(r, x, y, z);
    const float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    const Mat3 S = scale_mat3(s[0], s[1], s[2]);
    const Mat3 M = mat3_mul(S, R);

    Mat3 dSigma;
    dSigma.c[0][0] = dL_dcov3D[0];
    dSigma.c[0][1] = 0.5f * dL_dcov3D[1];
    dSigma.c[0][2] = 0.5f * dL_dcov3D[2];
    dSigma.c[1][0] = 0.5f * dL_dcov3D[1];
    dSigma.c[1][1] = dL_dcov3D[3];
    dSigma.c[1][2] = 0.5f * dL_dcov3D[4];
    dSigma.c[2][0] = 0.5f * dL_dcov3D[2];
    dSigma.c[2][1] = 0.5f * dL_dcov3D[4];
    dSigma.c[2][2] = dL_dcov3D[5];

    Mat3 M2;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) M2.c[c][rr] = 2.0f * M.c[c][rr];
    const Mat3 dM = mat3_mul(M2, dSigma);
    const Mat3 Rt = mat3_transpose(R);
    Mat3 dMt = mat3_transpose(dM);

#pragma unroll
    for (int j = 0; j < 3; ++j)
        dL_dscale[j] = Rt.c[j][0] * dMt.c[j][0] + Rt.c[j][1] * dMt.c[j][1] + Rt.c[j][2] * dMt.c[j][2];

#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) dMt.c[j][i] *= s[j];

    dL_drot[0] = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) +
                 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
    dL_drot[1] = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) +
                 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
    dL_drot[2] = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) +
                 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
    dL_drot[3] = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) +
                 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
}

}  // namespace adgs
