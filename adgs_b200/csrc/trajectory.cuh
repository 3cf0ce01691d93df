// Trajectory math in registers: sparse linear time bases (B-spline window + polynomial + Fourier
// collapsed by the host into (column, weight) lists) and the cumulative quaternion B-spline with
// its hand-derived reverse mode.
//
// Behaviour follows utils/func_utils.py:121-173 (get_func_result) and, for the quaternion maps,
// roma 1.5.1's quat_product / quat_conjugation / unitquat_to_rotvec(shortest_arc=True) /
// rotvec_to_unitquat (call sites func_utils.py:164-169): xyzw convention, Taylor branches at
// |angle| <= 1e-3.
#pragma once
#include "common.cuh"

namespace adgs {

struct Quat {
    float x, y, z, w;
};

// Sums of products on the forward path are written with explicit intrinsics: the single-view and the multi-view
// kernels inline this code into different surroundings, and nvcc's choice of which product to fuse into an FMA
// must not depend on that (the two paths are held bit-identical by tests/test_exchange_multi_gpu.py).
__device__ __forceinline__ float sumsq4_pinned(float a, float b, float c, float d)
{
    return __fmaf_rn(d, d, __fmaf_rn(c, c, __fmaf_rn(b, b, __fmul_rn(a, a))));
}

__device__ __forceinline__ Quat qmul(const Quat& p, const Quat& q)
{
    Quat r;
    // p.w q.v + q.w p.v + p.v x q.v  |  p.w q.w - p.v . q.v
    r.x = __fadd_rn(__fmaf_rn(q.w, p.x, __fmul_rn(p.w, q.x)), __fmaf_rn(p.y, q.z, -__fmul_rn(p.z, q.y)));
    r.y = __fadd_rn(__fmaf_rn(q.w, p.y, __fmul_rn(p.w, q.y)), __fmaf_rn(p.z, q.x, -__fmul_rn(p.x, q.z)));
    r.z = __fadd_rn(__fmaf_rn(q.w, p.z, __fmul_rn(p.w, q.z)), __fmaf_rn(p.x, q.y, -__fmul_rn(p.y, q.x)));
    r.w = __fmaf_rn(p.w, q.w, -dot3_pinned(p.x, q.x, p.y, q.y, p.z, q.z));
    return r;
}

__device__ __forceinline__ Quat qconj(const Quat& q)
{
    return Quat{-q.x, -q.y, -q.z, q.w};
}

__device__ __forceinline__ Quat qadd(const Quat& a, const Quat& b)
{
    return Quat{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
}

// rotation vector of a unit quaternion, shortest arc.  sin(angle/2) is taken from the quaternion
// itself (sin(atan2(|v|, w)) = |v| / |q|), which is the same quantity roma evaluates with sinf.
__device__ __noinline__ float3 qlog(Quat q)
{
    if (q.w < 0.f) q = Quat{-q.x, -q.y, -q.z, -q.w};
    const float nv2 = q.x * q.x + q.y * q.y + q.z * q.z;
    const float nv = sqrtf(nv2);
    const float angle = 2.f * atan2f(nv, q.w);
    float scale;
    if (fabsf(angle) <= 1e-3f) {
        const float a2 = angle * angle;
        scale = 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f;
    } else {
        scale = angle * sqrtf(nv2 + q.w * q.w) / nv;
    }
    return make_float3(scale * q.x, scale * q.y, scale * q.z);
}

// reverse mode of qlog: given g = dL/d(rotvec), returns dL/dq
__device__ __noinline__ Quat qlog_bwd(Quat q, const float3& g)
{
    const float sgn = (q.w < 0.f) ? -1.f : 1.f;
    q = Quat{sgn * q.x, sgn * q.y, sgn * q.z, sgn * q.w};
    const float nv2 = q.x * q.x + q.y * q.y + q.z * q.z;
    const float nv = sqrtf(nv2);
    const float den = nv2 + q.w * q.w;
    const float angle = 2.f * atan2f(nv, q.w);
    float scale, dscale;  // scale(angle), d scale / d angle
    if (fabsf(angle) <= 1e-3f) {
        const float a2 = angle * angle;
        scale = 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f;
        dscale = angle / 6.f + 7.f * a2 * angle / 720.f;
    } else {
        const float inv_rho = rsqrtf(den);
        const float sh = nv * inv_rho, ch = q.w * inv_rho;  // sin, cos of angle/2
        scale = angle / sh;
        dscale = (sh - 0.5f * angle * ch) / (sh * sh);
    }
    const float g_scale = g.x * q.x + g.y * q.y + g.z * q.z;
    const float g_half = 2.f * g_scale * dscale;  // d/d(half angle)
    const float g_nv = g_half * q.w / den;
    const float g_w = -g_half * nv / den;
    const float inv_nv = nv > 0.f ? 1.f / nv : 0.f;
    Quat r;
    r.x = sgn * (scale * g.x + g_nv * q.x * inv_nv);
    r.y = sgn * (scale * g.y + g_nv * q.y * inv_nv);
    r.z = sgn * (scale * g.z + g_nv * q.z * inv_nv);
    r.w = sgn * g_w;
    return r;
}

__device__ __noinline__ Quat qexp(const float3& r)
{
    const float n = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z);
    float sh, ch;
    sincosf(n * 0.5f, &sh, &ch);
    float scale;
    if (n <= 1e-3f) {
        const float n2 = n * n;
        scale = 0.5f - n2 / 48.f + n2 * n2 / 3840.f;
    } else {
        scale = sh / n;
    }
    return Quat{scale * r.x, scale * r.y, scale * r.z, ch};
}

// reverse mode of qexp: given g = dL/dq and the forward value e = qexp(r), returns dL/dr.
// cos(n/2) = e.w and sin(n/2) = |e.xyz| come from the forward value: no trigonometry here.
__device__ __noinline__ float3 qexp_bwd(const float3& r, const Quat& e, const Quat& g)
{
    const float n2 = r.x * r.x + r.y * r.y + r.z * r.z;
    const float n = sqrtf(n2);
    float scale, dscale_over_n, half_sin_over_n;  // s(n), s'(n)/n, sin(n/2)/(2n)
    if (n <= 1e-3f) {
        scale = 0.5f - n2 / 48.f + n2 * n2 / 3840.f;
        dscale_over_n = -1.f / 24.f + n2 / 960.f;
        half_sin_over_n = 0.5f * scale;
    } else {
        scale = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z) / n;
        dscale_over_n = (0.5f * e.w - scale) / n2;
        half_sin_over_n = 0.5f * scale;
    }
    const float gv_dot_r = g.x * r.x + g.y * r.y + g.z * r.z;
    const float coef = dscale_over_n * gv_dot_r - half_sin_over_n * g.w;
    return make_float3(scale * g.x + coef * r.x, scale * g.y + coef * r.y, scale * g.z + coef * r.z);
}

// control quaternion from the stored (wxyz) parameter: normalize(param + [1,0,0,0]) -> xyzw
__device__ __forceinline__ Quat ctrl_quat(const float4& p, float& norm)
{
    const float w = p.x + 1.0f, x = p.y, y = p.z, z = p.w;
    norm = fmaxf(sqrtf(sumsq4_pinned(w, x, y, z)), 1e-12f);
    const float inv = 1.f / norm;
    return Quat{x * inv, y * inv, z * inv, w * inv};
}

// q(t) = q0 * prod_i exp(cum_i * log(conj(q_{i-1}) q_i)), i = 1..k
__device__ __forceinline__ Quat quat_spline(const Quat* qt, int k, const float* cum)
{
    Quat out = qt[0];
#pragma unroll
    for (int i = 1; i <= ADGS_MAX_QUAT_ORDER; ++i) {
        if (i <= k) {
            const Quat rel = qmul(qconj(qt[i - 1]), qt[i]);
            const float3 om = qlog(rel);
            const float c = cum[i];
            out = qmul(out, qexp(make_float3(c * om.x, c * om.y, c * om.z)));
        }
    }
    return out;
}

// Forward sweep that caches what the reverse sweep needs (log of every relative rotation, every
// exponential, the prefix products); returns q(t).
__device__ __forceinline__ Quat quat_spline_cached(const Quat* qt, int k, const float* cum, Quat* P, Quat* E,
                                                   float3* om)
{
    P[0] = qt[0];
#pragma unroll
    for (int i = 1; i <= ADGS_MAX_QUAT_ORDER; ++i) {
        if (i <= k) {
            om[i] = qlog(qmul(qconj(qt[i - 1]), qt[i]));
            const float c = cum[i];
            E[i] = qexp(make_float3(c * om[i].x, c * om[i].y, c * om[i].z));
            P[i] = qmul(P[i - 1], E[i]);
        } else {
            P[i] = P[i - 1];
            E[i] = Quat{0.f, 0.f, 0.f, 1.f};
            om[i] = make_float3(0.f, 0.f, 0.f);
        }
    }
    return P[ADGS_MAX_QUAT_ORDER];
}

// reverse mode of quat_spline: gqt[0..k] receive dL/dq_i (xyzw, w.r.t. the NORMALISED controls)
__device__ __forceinline__ void quat_spline_bwd(const Quat* qt, int k, const float* cum, const Quat* P, const Quat* E,
                                                const float3* om, const Quat& gout, Quat* gqt)
{
#pragma unroll
    for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) gqt[i] = Quat{0.f, 0.f, 0.f, 0.f};
    Quat gP = gout;
#pragma unroll
    for (int i = ADGS_MAX_QUAT_ORDER; i >= 1; --i) {
        if (i <= k) {
            const Quat a = qt[i - 1], b = qt[i];
            const float c = cum[i];
            const float3 r = make_float3(c * om[i].x, c * om[i].y, c * om[i].z);
            const Quat ge = qmul(qconj(P[i - 1]), gP);
            gP = qmul(gP, qconj(E[i]));
            const float3 gr = qexp_bwd(r, E[i], ge);
            const float3 gom = make_float3(c * gr.x, c * gr.y, c * gr.z);
            const Quat grel = qlog_bwd(qmul(qconj(a), b), gom);
            const Quat gca = qmul(grel, qconj(b));
            gqt[i - 1] = qadd(gqt[i - 1], qconj(gca));
            gqt[i] = qadd(gqt[i], qmul(a, grel));
        }
    }
    gqt[0] = qadd(gqt[0], gP);
}

// d(v / max(|v|, eps)) applied to g, for a 4-vector given its normalised value n and norm
__device__ __forceinline__ float4 normalize4_bwd(const float4& n, float norm, const float4& g)
{
    const float d = n.x * g.x + n.y * g.y + n.z * g.z + n.w * g.w;
    const float inv = 1.f / norm;
    return make_float4((g.x - n.x * d) * inv, (g.y - n.y * d) * inv, (g.z - n.z * d) * inv, (g.w - n.w * d) * inv);
}

}  // namespace adgs
