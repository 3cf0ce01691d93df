// Fused trajectory + projection kernels over the planar model storage (adgs_model), their
// backward, and the C-ABI entry points adgs_trajectory_forward / adgs_render_forward /
// adgs_render_backward.
//
// One thread per Gaussian. Forward reads every parameter once -- coalesced: AoS rows that are a
// single vector (xyz, scale, quaternion), float4 planes for the wide blocks (SH, SH-deform,
// control quaternions), scalar planes for the position control points -- evaluates
//   position  = xyz + sum_j xyz_deform[col_j] w_j (+ the same columns with the flow-time weights)
//               + background(t)                                  (gaussian_model.py:173-185)
//   rotation  = normalize(scene quaternion | cumulative quaternion B-spline)   (:187-196)
//   SH DC    += sum_j shs_deform[col_j] w_j                       (:198-205)
//   opacity   = sigmoid(o) * exp(-0.5 ((t - tau)/sigma_+-)^2)      (:207-214)
//   scale     = exp(s)                                            (:88-91)
// and runs the per-Gaussian rasterizer front end (forward.cu:155-256) in the same registers, so
// the deformed tensors never round-trip HBM. 48 bytes per Gaussian are kept for the backward.
#include <mutex>
#include "api_internal.cuh"
#include "trajectory.cuh"

namespace adgs {
namespace {

constexpr int kSavedFloats = 12;  // xyz_t(3) op_act(1) | q(4) | dc_t(3) pad(1)

struct FusedFwdArgs {
    adgs_model m;
    adgs_time_basis tb;
    adgs_deformed out;  // optional materialised tensors
    int render;         // 0: trajectory only
    int render_objmask;
    RasterParams rp;
    const float* view;
    const float* proj;
    const float* campos;
    int32_t* radii;
    uint32_t* depth_keys;
    uint32_t* tiles_touched;
    float4* record;
    float* cov3D;
    uint8_t* clamped;
    float4* saved;
};

__device__ __forceinline__ void lin_eval3_planar(const float* base, int n_obj, int j, const adgs_lin_basis& b,
                                                 float* v0, float* v1, bool second)
{
    // base: (C,3,n_obj); value_d += p[col][d][j] * w
    for (int t = 0; t < b.n; ++t) {
        const float* p = base + ((size_t)b.col[t] * 3) * n_obj + j;
        const float w0 = b.w0[t];
        const float w1 = b.w1[t];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float x = __ldg(p + (size_t)d * n_obj);
            v0[d] += x * w0;
            if (second) v1[d] += x * w1;
        }
    }
}

// Object rotation before the final normalisation: [static quaternion if no spline] + linear terms
// + quaternion spline, in wxyz.
__device__ __forceinline__ float4 object_rotation_raw(const adgs_model& m, const adgs_time_basis& tb, int g, int j,
                                                      Quat* qt_out, float* norms_out)
{
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* rd = reinterpret_cast<const float4*>(m.rot_deform);
    if (tb.quat.n_ctrl == 0) q = reinterpret_cast<const float4*>(m.rotation)[g];
    for (int t = 0; t < tb.rotation.n; ++t) {
        const float4 p = __ldg(rd + (size_t)tb.rotation.col[t] * m.N_obj + j);
        const float w = tb.rotation.w0[t];
        q.x += p.x * w;
        q.y += p.y * w;
        q.z += p.z * w;
        q.w += p.w * w;
    }
    if (tb.quat.n_ctrl != 0) {
        Quat qt[ADGS_MAX_QUAT_ORDER + 1];
        const int k = tb.quat.k;
#pragma unroll
        for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
            float nrm = 1.f;
            if (i <= k) {
                const float4 p = __ldg(rd + (size_t)(tb.quat.start + i) * m.N_obj + j);
                qt[i] = ctrl_quat(p, nrm);
            } else {
                qt[i] = Quat{0.f, 0.f, 0.f, 1.f};
            }
            if (qt_out) {
                qt_out[i] = qt[i];
                norms_out[i] = nrm;
            }
        }
        const Quat r = quat_spline(qt, k, tb.quat.cum);
        q.x += r.w;
        q.y += r.x;
        q.z += r.y;
        q.w += r.z;
    }
    return q;
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) fused_forward_kernel(const __grid_constant__ FusedFwdArgs a)
{
    __shared__ CamSmem cam;
    __shared__ float s_bg[6];
    const adgs_model& m = a.m;
    const adgs_time_basis& tb = a.tb;
    const int N = m.N_scene + m.N_obj;
    if (threadIdx.x < 6) {
        const int d = threadIdx.x % 3;
        const bool second = threadIdx.x >= 3;
        float v = 0.f;
        for (int t = 0; t < tb.background.n; ++t)
            v += m.background_deform[d * tb.background.n_cols + tb.background.col[t]] *
                 (second ? tb.background.w1[t] : tb.background.w0[t]);
        s_bg[threadIdx.x] = v;
    }
    if (a.render) {
        load_camera(cam, a.view, a.proj, a.campos, nullptr);
    } else {
        __syncthreads();
    }
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const bool is_obj = g >= m.N_scene;
    const int j = g - m.N_scene;
    const bool flow = tb.has_flow != 0;
    // ---- position ------------------------------------------------------------------------
    float xt[3], xf[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xt[d] = xf[d] = m.xyz[3 * (size_t)g + d];
    if (is_obj && tb.xyz.n) {
        float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f};
        lin_eval3_planar(m.xyz_deform, m.N_obj, j, tb.xyz, d0, d1, flow);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            xt[d] += d0[d];
            xf[d] += d1[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        xt[d] += s_bg[d];
        xf[d] += s_bg[3 + d];
    }

    // ---- rotation ------------------------------------------------------------------------
    float4 qraw;
    if (is_obj) {
        qraw = object_rotation_raw(m, tb, g, j, nullptr, nullptr);
    } else {
        qraw = reinterpret_cast<const float4*>(m.rotation)[g];
    }
    const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
    const float rot[4] = {qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn};

    // ---- opacity, scale ------------------------------------------------------------------
    float op = 1.0f / (1.0f + expf(-m.opacity[g]));
    if (is_obj && tb.use_time_mask) {
        const float delta = tb.t - m.gs_time[j];
        const float2 sg = reinterpret_cast<const float2*>(m.gs_time_sigma)[j];
        const float sigma = expf(delta < 0.0f ? sg.x : sg.y);
        const float z = delta / sigma;
        op *= expf(-0.5f * z * z);
    }
    float scale[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) scale[d] = expf(m.scaling[3 * (size_t)g + d]);

    // ---- SH DC deformation ---------------------------------------------------------------
    float dc[3];
    const float4* sh4 = reinterpret_cast<const float4*>(m.sh4);
    const float4 shq0 = __ldg(sh4 + g);
    dc[0] = shq0.x;
    dc[1] = shq0.y;
    dc[2] = shq0.z;
    if (tb.shs.n) {
        const int Cs = tb.shs.n_cols;
        // (3,Cs) block in float4 chunks; gather the term columns
        const float* sd = m.shs_deform4;
        for (int t = 0; t < tb.shs.n; ++t) {
            const float w = tb.shs.w0[t];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int e = c * Cs + tb.shs.col[t];
                dc[c] += __ldg(sd + ((size_t)(e >> 2) * N + g) * 4 + (e & 3)) * w;
            }
        }
    }

    // ---- optional materialised outputs (get_deformed_pkg shapes) ---------------------------
    if (a.out.xyz) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a.out.xyz[3 * (size_t)g + d] = xt[d];
    }
    if (a.out.flow_xyz) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a.out.flow_xyz[3 * (size_t)g + d] = xf[d];
    }
    if (a.out.rotation) reinterpret_cast<float4*>(a.out.rotation)[g] = make_float4(rot[0], rot[1], rot[2], rot[3]);
    if (a.out.opacity) a.out.opacity[g] = op;
    if (a.out.scaling) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a.out.scaling[3 * (size_t)g + d] = scale[d];
    }
    if (a.out.shs) {
        float4* o = reinterpret_cast<float4*>(a.out.shs + (size_t)g * 48);
        o[0] = make_float4(dc[0], dc[1], dc[2], shq0.w);
#pragma unroll
        for (int q = 1; q < 12; ++q) o[q] = __ldg(sh4 + (size_t)q * N + g);
    }
    if (!a.render) return;

    // ---- rasterizer front end --------------------------------------------------------------
    const float3 p = make_float3(xt[0], xt[1], xt[2]);
    SplatGeom sg;
    const bool visible = splat_geometry(p, scale, rot, nullptr, a.rp, cam.view, cam.proj, sg);
    a.saved[(size_t)g * 3 + 0] = make_float4(xt[0], xt[1], xt[2], op);
    a.saved[(size_t)g * 3 + 1] = make_float4(rot[0], rot[1], rot[2], rot[3]);
    a.saved[(size_t)g * 3 + 2] = make_float4(dc[0], dc[1], dc[2], 0.f);
    if (!visible) {
        a.radii[g] = 0;
        a.tiles_touched[g] = 0;
        a.depth_keys[g] = 0xFFFFFFFFu;
        return;
    }
    float sh[48];
    const int deg = a.rp.sh_degree;
    const int chunks = sh_chunks_for_degree(deg);
    sh[0] = dc[0];
    sh[1] = dc[1];
    sh[2] = dc[2];
    sh[3] = shq0.w;
#pragma unroll
    for (int q = 1; q < 12; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < chunks) v = __ldg(sh4 + (size_t)q * N + g);
        sh[4 * q + 0] = v.x;
        sh[4 * q + 1] = v.y;
        sh[4 * q + 2] = v.z;
        sh[4 * q + 3] = v.w;
    }
    float rgb[3];
    uint32_t clamped;
    sh_to_rgb(deg, p, cam.campos, sh, rgb, clamped);
    a.clamped[g] = (uint8_t)clamped;
    float2* c2 = reinterpret_cast<float2*>(a.cov3D + (size_t)g * 6);
    c2[0] = make_float2(sg.cov3D[0], sg.cov3D[1]);
    c2[1] = make_float2(sg.cov3D[2], sg.cov3D[3]);
    c2[2] = make_float2(sg.cov3D[4], sg.cov3D[5]);
    const float dfeat = a.rp.inv_depth ? (1.0f / (sg.depth + 0.0000001f)) : sg.depth;
    const float sem0 = (a.render_objmask && is_obj) ? 1.f : 0.f;
    store_blend_record(a.record + (size_t)g * 4, sg.px, sg.py, sg.conic_x, sg.conic_y, sg.conic_z, op, sg.depth, rgb, dfeat,
                       flow ? xf[0] : 0.f, flow ? xf[1] : 0.f, flow ? xf[2] : 0.f, sem0);
    a.radii[g] = sg.radius;
    a.tiles_touched[g] = sg.tiles;
    a.depth_keys[g] = __float_as_uint(sg.depth);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct FusedBwdArgs {
    adgs_model m;
    adgs_model g;  // gradient buffers, same layouts
    adgs_time_basis tb;
    RasterParams rp;
    const float* view;
    const float* proj;
    const float* campos;
    const int32_t* radii;
    const float* cov3D;
    const uint8_t* clamped;
    const float4* saved;
    const float* grad_record;
    float* dL_dmeans2D;
    float4* dq_scratch;  // (N_obj) dL/d(normalised object quaternion)
    float* bg_scratch;   // 6 floats: sum dxyz_t, sum dflow
    float4* dm3_scratch; // (N) dL/d(position) through the view direction of the SH colour, written by sh_backward_kernel
                         // (null: the SH block is handled inside fused_backward_kernel)
    int accumulate;      // != 0: add into the gradient buffers (second and later views of a batch)
};

__device__ __forceinline__ void put(float* p, float v, int acc)
{
    *p = acc ? *p + v : v;
}

__device__ __forceinline__ void put2(float2* p, float2 v, int acc)
{
    if (acc) {
        const float2 o = *p;
        v.x += o.x;
        v.y += o.y;
    }
    *p = v;
}

__device__ __forceinline__ void put4(float4* p, float4 v, int acc)
{
    if (acc) {
        const float4 o = *p;
        v.x += o.x;
        v.y += o.y;
        v.z += o.z;
        v.w += o.w;
    }
    *p = v;
}

// shs_deform planes are float4 quads over the flattened (channel, column) index e = c * Cs + col; the channel e / Cs
// of every element is tabulated once per CTA (Cs is a run-time divisor)
constexpr int kShsDeformElems = (3 * 2 * ADGS_MAX_TERMS + 3) / 4 * 4;

// SH_SPLIT: the SH colour block (192 B read + 336 B written per Gaussian, 96 live registers) is handled by
// sh_backward_kernel, which ran before and left dL/d(position) through the view direction in dm3_scratch: without the
// two 48-float arrays this kernel fits in fewer registers and more warps cover its HBM latency.
template <int TPB, int MINB, bool SH_SPLIT>
__global__ void __launch_bounds__(TPB, MINB) fused_backward_kernel(const __grid_constant__ FusedBwdArgs a)
{
    __shared__ CamSmem cam;
    __shared__ float s_wshs[ADGS_MAX_TERMS * 2];  // dense SH-deform weights per column
    __shared__ uint8_t s_chan[SH_SPLIT ? 4 : kShsDeformElems];  // channel of every flattened shs_deform element
    __shared__ float s_red[TPB / 32][6];
    const adgs_model& m = a.m;
    const adgs_time_basis& tb = a.tb;
    const int N = m.N_scene + m.N_obj;
    const int Cs = tb.shs.n_cols;
    for (int i = threadIdx.x; i < Cs && i < ADGS_MAX_TERMS * 2; i += blockDim.x) s_wshs[i] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 0; t < tb.shs.n; ++t) s_wshs[tb.shs.col[t]] += tb.shs.w0[t];
    if (!SH_SPLIT && Cs > 0)
        for (int e = threadIdx.x; e < (3 * Cs + 3) / 4 * 4; e += blockDim.x) s_chan[e] = (uint8_t)(e / Cs);
    load_camera(cam, a.view, a.proj, a.campos, nullptr);

    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < N;
    const bool is_obj = valid && g >= m.N_scene;
    const int j = g - m.N_scene;
    const bool flow = tb.has_flow != 0;
    float dxt[3] = {0.f, 0.f, 0.f}, dfl[3] = {0.f, 0.f, 0.f};

    if (valid) {
        const float4* gr = reinterpret_cast<const float4*>(a.grad_record) + (size_t)g * 4;
        const float4 g0 = gr[0], g1 = gr[1], g2 = gr[2], g3 = gr[3];
        if (a.dL_dmeans2D) {
            a.dL_dmeans2D[3 * (size_t)g + 0] = g0.x;
            a.dL_dmeans2D[3 * (size_t)g + 1] = g0.y;
            a.dL_dmeans2D[3 * (size_t)g + 2] = 0.f;
        }
        const bool visible = a.radii[g] > 0;
        const float4 sv0 = a.saved[(size_t)g * 3 + 0];
        const float4 sv1 = a.saved[(size_t)g * 3 + 1];
        const float rot[4] = {sv1.x, sv1.y, sv1.z, sv1.w};
        // SH_SPLIT: every load of the thread is issued here, none behind the visibility branch (one DRAM latency
        // per thread instead of three)
        float4 dm3_pre = make_float4(0.f, 0.f, 0.f, 0.f), qraw_pre = make_float4(0.f, 0.f, 0.f, 0.f);
        float opacity_pre = 0.f;
        if (SH_SPLIT) {
            dm3_pre = a.dm3_scratch[g];
            opacity_pre = m.opacity[g];
            if (!is_obj) qraw_pre = reinterpret_cast<const float4*>(m.rotation)[g];
        }
        float scale[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) scale[d] = expf(m.scaling[3 * (size_t)g + d]);

        float dscale[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
        float dsh[SH_SPLIT ? 1 : 48];
#pragma unroll
        for (int i = 0; i < (SH_SPLIT ? 1 : 48); ++i) dsh[i] = 0.f;
        if (flow) {
            dfl[0] = g2.z;
            dfl[1] = g2.w;
            dfl[2] = g3.x;
        }
        const float3 p = make_float3(sv0.x, sv0.y, sv0.z);
        float3 dm3 = make_float3(0.f, 0.f, 0.f);
        auto sh_block = [&]() {
            if (SH_SPLIT) {
                dm3 = make_float3(dm3_pre.x, dm3_pre.y, dm3_pre.z);
                return;
            }
            const float4 sv2 = a.saved[(size_t)g * 3 + 2];
            const float4* sh4 = reinterpret_cast<const float4*>(m.sh4);
            float sh[48];
            const int deg = a.rp.sh_degree;
            const int chunks = sh_chunks_for_degree(deg);
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < chunks) v = __ldg(sh4 + (size_t)q * N + g);
                sh[4 * q + 0] = v.x;
                sh[4 * q + 1] = v.y;
                sh[4 * q + 2] = v.z;
                sh[4 * q + 3] = v.w;
            }
            sh[0] = sv2.x;
            sh[1] = sv2.y;
            sh[2] = sv2.z;
            const float dcol[3] = {g1.z, g1.w, g2.x};
            dm3 = sh_to_rgb_bwd(deg, p, cam.campos, sh, a.clamped[g], dcol, dsh);
        };
        auto sh_store = [&]() {
            if (SH_SPLIT) return;
            float4* gsh4 = reinterpret_cast<float4*>(a.g.sh4);
#pragma unroll
            for (int q = 0; q < 12; ++q)
                put4(gsh4 + (size_t)q * N + g, make_float4(dsh[4 * q], dsh[4 * q + 1], dsh[4 * q + 2], dsh[4 * q + 3]),
                     a.accumulate);
            if (a.g.shs_deform4 && Cs > 0) {
                const int nq = (3 * Cs + 3) / 4;
                float4* gsd = reinterpret_cast<float4*>(a.g.shs_deform4);
                // the DC gradient by channel as scalars: a run-time index into dsh[] would push the whole array
                // into local memory
                const float dc0 = dsh[0], dc1 = dsh[1], dc2 = dsh[2];
                for (int q = 0; q < nq; ++q) {
                    float v[4];
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const int e = 4 * q + e4;
                        const int c = s_chan[e];
                        const float dc = c == 0 ? dc0 : (c == 1 ? dc1 : dc2);
                        v[e4] = (c < 3) ? dc * s_wshs[e - c * Cs] : 0.f;
                    }
                    put4(gsd + (size_t)q * N + g, make_float4(v[0], v[1], v[2], v[3]), a.accumulate);
                }
            }
        };
        if (visible) {
            float cv[6], dcov[6];
            if (SH_SPLIT) {
                // the forward's own expression on the same inputs (pinned contraction: the stored bits), cheaper than
                // a dependent 24-byte load
                cov3d_from_scale_rot(scale, a.rp.scale_modifier, rot, cv);
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) cv[i] = a.cov3D[(size_t)g * 6 + i];
            }
            float3 dm = cov2d_bwd(p, a.rp, cv, cam.view, g0.z, g0.w, g1.x, dcov);
            const float3 dm2 = mean_proj_depth_bwd(p, cam.view, cam.proj, g0.x, g0.y, g2.y, a.rp.inv_depth);
            sh_block();
            dxt[0] = dm.x + dm2.x + dm3.x;
            dxt[1] = dm.y + dm2.y + dm3.y;
            dxt[2] = dm.z + dm2.z + dm3.z;
            cov3d_bwd(scale, a.rp.scale_modifier, rot, dcov, dscale, dq);
        }

        // ---- leaf gradients ----------------------------------------------------------------
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            put(a.g.xyz + 3 * (size_t)g + d, dxt[d] + dfl[d], a.accumulate);
            put(a.g.scaling + 3 * (size_t)g + d, dscale[d] * scale[d], a.accumulate);
        }
        sh_store();
        // opacity: op_act = sigmoid(o) [* mask]
        {
            const float dop = g1.y;
            const float sig = 1.0f / (1.0f + expf(-(SH_SPLIT ? opacity_pre : m.opacity[g])));
            float mask = 1.f;
            if (is_obj && tb.use_time_mask) {
                const float delta = tb.t - m.gs_time[j];
                const float2 sgm = reinterpret_cast<const float2*>(m.gs_time_sigma)[j];
                const bool neg = delta < 0.0f;
                const float sigma = expf(neg ? sgm.x : sgm.y);
                const float z = delta / sigma;
                mask = expf(-0.5f * z * z);
                const float dside = dop * sig * mask * z * z;
                if (a.g.gs_time_sigma)
                    put2(reinterpret_cast<float2*>(a.g.gs_time_sigma) + j,
                         neg ? make_float2(dside, 0.f) : make_float2(0.f, dside), a.accumulate);
            } else if (is_obj && a.g.gs_time_sigma) {
                put2(reinterpret_cast<float2*>(a.g.gs_time_sigma) + j, make_float2(0.f, 0.f), a.accumulate);
            }
            put(a.g.opacity + g, dop * mask * sig * (1.f - sig), a.accumulate);
        }
        // rotation
        if (!is_obj) {
            const float4 qraw = SH_SPLIT ? qraw_pre : reinterpret_cast<const float4*>(m.rotation)[g];
            const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
            put4(reinterpret_cast<float4*>(a.g.rotation) + g,
                 normalize4_bwd(make_float4(rot[0], rot[1], rot[2], rot[3]), qn, make_float4(dq[0], dq[1], dq[2], dq[3])),
                 a.accumulate);
        } else {
            a.dq_scratch[j] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        }
        // position control points (dense elsewhere: zero-filled by the caller)
        if (is_obj && tb.xyz.n && a.g.xyz_deform) {
            for (int t = 0; t < tb.xyz.n; ++t) {
                float* o = a.g.xyz_deform + ((size_t)tb.xyz.col[t] * 3) * m.N_obj + j;
                const float w0 = tb.xyz.w0[t], w1 = tb.xyz.w1[t];
#pragma unroll
                for (int d = 0; d < 3; ++d) put(o + (size_t)d * m.N_obj, dxt[d] * w0 + dfl[d] * w1, a.accumulate);
            }
        }
    }

    // ---- background parameter: grid reduction of sum dxyz_t / sum dflow -----------------------
    if (tb.background.n) {
        float r[6] = {dxt[0], dxt[1], dxt[2], dfl[0], dfl[1], dfl[2]};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) r[i] += __shfl_xor_sync(0xffffffffu, r[i], off);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0)
#pragma unroll
            for (int i = 0; i < 6; ++i) s_red[warp][i] = r[i];
        __syncthreads();
        if (threadIdx.x < 6) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < TPB / 32; ++w) s += s_red[w][threadIdx.x];
            if (s != 0.f) atomicAdd(a.bg_scratch + threadIdx.x, s);
        }
    }
}

// The SH colour block of the per-Gaussian backward as its own streaming kernel (see fused_backward_kernel<SH_SPLIT>):
// dL/d(sh4) (dense: zeros for culled Gaussians), dL/d(shs_deform4) = dL/d(DC) x the time weights, and the gradient of
// the colour w.r.t. the position through the normalised view direction (-> dm3_scratch).
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) sh_backward_kernel(const __grid_constant__ FusedBwdArgs a)
{
    // weight and channel of every flattened shs_deform element (kShsDeformElems above)
    __shared__ float s_wshs[ADGS_MAX_TERMS * 2];
    __shared__ __align__(16) float s_w[kShsDeformElems];
    __shared__ __align__(4) uint8_t s_c[kShsDeformElems];
    __shared__ float s_campos[3];
    const adgs_model& m = a.m;
    const adgs_time_basis& tb = a.tb;
    const int N = m.N_scene + m.N_obj;
    const int Cs = (a.g.shs_deform4 && tb.shs.n_cols > 0) ? tb.shs.n_cols : 0;
    const int nq = (3 * Cs + 3) / 4;
    for (int i = threadIdx.x; i < Cs && i < ADGS_MAX_TERMS * 2; i += blockDim.x) s_wshs[i] = 0.f;
    if (threadIdx.x < 3) s_campos[threadIdx.x] = a.campos[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 0; t < tb.shs.n; ++t) s_wshs[tb.shs.col[t]] += tb.shs.w0[t];
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * nq; e += blockDim.x) {
        const int c = e / Cs;
        s_c[e] = (uint8_t)c;
        s_w[e] = (c < 3) ? s_wshs[e - c * Cs] : 0.f;
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;

    // all the unconditional loads first, the 192-byte coefficient block only for Gaussians on screen
    const float4* gr = reinterpret_cast<const float4*>(a.grad_record) + (size_t)g * 4;
    const bool visible = a.radii[g] > 0;
    const float4 g1 = gr[1], g2 = gr[2];
    const float4 sv0 = a.saved[(size_t)g * 3 + 0];
    const float4 sv2 = a.saved[(size_t)g * 3 + 2];
    const uint32_t clamped = a.clamped[g];
    float basis[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) basis[k] = 0.f;
    float gm[3] = {0.f, 0.f, 0.f};
    float3 dm3 = make_float3(0.f, 0.f, 0.f);
    if (visible) {
        const float4* sh4 = reinterpret_cast<const float4*>(m.sh4);
        float sh[48];
        const int deg = a.rp.sh_degree;
        const int chunks = sh_chunks_for_degree(deg);
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < chunks) v = __ldg(sh4 + (size_t)q * N + g);
            sh[4 * q + 0] = v.x;
            sh[4 * q + 1] = v.y;
            sh[4 * q + 2] = v.z;
            sh[4 * q + 3] = v.w;
        }
        sh[0] = sv2.x;
        sh[1] = sv2.y;
        sh[2] = sv2.z;
        const float dcol[3] = {g1.z, g1.w, g2.x};
        dm3 = sh_to_rgb_bwd<false, true>(deg, make_float3(sv0.x, sv0.y, sv0.z), s_campos, sh, clamped, dcol, basis);
#pragma unroll
        for (int c = 0; c < 3; ++c) gm[c] = dcol[c] * ((clamped >> c) & 1u ? 0.f : 1.f);
    }
    a.dm3_scratch[g] = make_float4(dm3.x, dm3.y, dm3.z, 0.f);
    // dL/dsh[3k + c] = basis[k] * gm[c]
    float4* gsh4 = reinterpret_cast<float4*>(a.g.sh4);
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        float v[4];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) v[e4] = basis[(4 * q + e4) / 3] * gm[(4 * q + e4) % 3];
        put4(gsh4 + (size_t)q * N + g, make_float4(v[0], v[1], v[2], v[3]), a.accumulate);
    }
    if (nq > 0) {
        float4* gsd = reinterpret_cast<float4*>(a.g.shs_deform4);
        const float dc0 = basis[0] * gm[0], dc1 = basis[0] * gm[1], dc2 = basis[0] * gm[2];
        auto term = [&](uint32_t c, float w) {
            const float dc = c == 0 ? dc0 : (c == 1 ? dc1 : dc2);
            return (c < 3) ? dc * w : 0.f;
        };
        for (int q = 0; q < nq; ++q) {
            const float4 w = reinterpret_cast<const float4*>(s_w)[q];
            const uchar4 c = reinterpret_cast<const uchar4*>(s_c)[q];
            put4(gsd + (size_t)q * N + g, make_float4(term(c.x, w.x), term(c.y, w.y), term(c.z, w.z), term(c.w, w.w)),
                 a.accumulate);
        }
    }
}

// Object rotation chain: normalised quaternion gradient -> static quaternion / linear terms /
// control quaternions of the spline window. One forward sweep of the spline (cached) + one reverse.
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) rotation_backward_kernel(const __grid_constant__ FusedBwdArgs a)
{
    const adgs_model& m = a.m;
    const adgs_time_basis& tb = a.tb;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m.N_obj) return;
    const int g = m.N_scene + j;
    const float4* rd = reinterpret_cast<const float4*>(m.rot_deform);
    float4 qraw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tb.quat.n_ctrl == 0) qraw = reinterpret_cast<const float4*>(m.rotation)[g];
    for (int t = 0; t < tb.rotation.n; ++t) {
        const float4 p = __ldg(rd + (size_t)tb.rotation.col[t] * m.N_obj + j);
        const float w = tb.rotation.w0[t];
        qraw.x += p.x * w;
        qraw.y += p.y * w;
        qraw.z += p.z * w;
        qraw.w += p.w * w;
    }
    Quat qt[ADGS_MAX_QUAT_ORDER + 1], P[ADGS_MAX_QUAT_ORDER + 1], E[ADGS_MAX_QUAT_ORDER + 1];
    float3 om[ADGS_MAX_QUAT_ORDER + 1];
    float norms[ADGS_MAX_QUAT_ORDER + 1];
    const int k = tb.quat.k;
    if (tb.quat.n_ctrl != 0) {
#pragma unroll
        for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
            norms[i] = 1.f;
            if (i <= k) {
                qt[i] = ctrl_quat(__ldg(rd + (size_t)(tb.quat.start + i) * m.N_obj + j), norms[i]);
            } else {
                qt[i] = Quat{0.f, 0.f, 0.f, 1.f};
            }
        }
        const Quat r = quat_spline_cached(qt, k, tb.quat.cum, P, E, om);
        qraw.x += r.w;
        qraw.y += r.x;
        qraw.z += r.y;
        qraw.w += r.z;
    }
    const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
    const float4 qhat = make_float4(qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn);
    const float4 graw = normalize4_bwd(qhat, qn, a.dq_scratch[j]);  // wxyz
    put4(reinterpret_cast<float4*>(a.g.rotation) + g, (tb.quat.n_ctrl == 0) ? graw : make_float4(0.f, 0.f, 0.f, 0.f),
         a.accumulate);
    float4* grd = reinterpret_cast<float4*>(a.g.rot_deform);
    if (!grd) return;
    for (int t = 0; t < tb.rotation.n; ++t) {
        const float w = tb.rotation.w0[t];
        put4(grd + (size_t)tb.rotation.col[t] * m.N_obj + j, make_float4(graw.x * w, graw.y * w, graw.z * w, graw.w * w),
             a.accumulate);
    }
    if (tb.quat.n_ctrl != 0) {
        Quat gqt[ADGS_MAX_QUAT_ORDER + 1];
        quat_spline_bwd(qt, k, tb.quat.cum, P, E, om, Quat{graw.y, graw.z, graw.w, graw.x}, gqt);
#pragma unroll
        for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
            if (i <= k) {
                // control = normalize(param + e_w): back through the normalisation, xyzw -> wxyz
                const float4 nq = make_float4(qt[i].w, qt[i].x, qt[i].y, qt[i].z);
                const float4 gg = make_float4(gqt[i].w, gqt[i].x, gqt[i].y, gqt[i].z);
                put4(grd + (size_t)(tb.quat.start + i) * m.N_obj + j, normalize4_bwd(nq, norms[i], gg), a.accumulate);
            }
        }
    }
}

__global__ void background_finalize_kernel(const __grid_constant__ FusedBwdArgs a)
{
    const adgs_lin_basis& b = a.tb.background;
    float* out = a.g.background_deform;
    if (!out) return;
    const int C = b.n_cols;
    if (!a.accumulate)
        for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) out[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        for (int t = 0; t < b.n; ++t) out[d * C + b.col[t]] += a.bg_scratch[d] * b.w0[t] + a.bg_scratch[3 + d] * b.w1[t];
    }
}

int validate_model(const adgs_model* m, const adgs_time_basis* tb)
{
    if (!m || !tb) return ADGS_ERR_ARG;
    if (m->N_scene < 0 || m->N_obj < 0) return ADGS_ERR_ARG;
    const int N = m->N_scene + m->N_obj;
    if (N == 0) return ADGS_OK;
    if (!m->xyz || !m->scaling || !m->rotation || !m->opacity || !m->sh4) return ADGS_ERR_ARG;
    const adgs_lin_basis* bs[4] = {&tb->xyz, &tb->background, &tb->shs, &tb->rotation};
    for (int i = 0; i < 4; ++i) {
        if (bs[i]->n < 0 || bs[i]->n > ADGS_MAX_TERMS) return ADGS_ERR_UNSUPPORTED;
        for (int t = 0; t < bs[i]->n; ++t)
            if (bs[i]->col[t] < 0 || bs[i]->col[t] >= bs[i]->n_cols) return ADGS_ERR_ARG;
    }
    if (tb->shs.n_cols > 2 * ADGS_MAX_TERMS) return ADGS_ERR_UNSUPPORTED;
    if (tb->quat.k < 0 || tb->quat.k > ADGS_MAX_QUAT_ORDER) return ADGS_ERR_UNSUPPORTED;
    if (m->N_obj > 0) {
        if (tb->xyz.n && !m->xyz_deform) return ADGS_ERR_ARG;
        if ((tb->rotation.n || tb->quat.n_ctrl) && !m->rot_deform) return ADGS_ERR_ARG;
        if (tb->use_time_mask && (!m->gs_time || !m->gs_time_sigma)) return ADGS_ERR_ARG;
    }
    if (tb->shs.n && !m->shs_deform4) return ADGS_ERR_ARG;
    if (tb->background.n && !m->background_deform) return ADGS_ERR_ARG;
    return ADGS_OK;
}


// ------------------------------------------------------------------------------------------
// Multi-view variants for the splat-exchange path: one launch evaluates a model shard for up to
// kMaxViews views. Parameters are fetched from HBM once (re-reads of later views hit L1/L2), the
// backward sums the per-view contributions in registers and writes every dense gradient once.
// ------------------------------------------------------------------------------------------
constexpr int kMaxViews = 8;

struct ViewIO {
    adgs_time_basis tb;
    RasterParams rp;
    const float* view;
    const float* proj;
    const float* campos;
    int32_t* radii;
    uint32_t* depth_keys;
    uint32_t* tiles_touched;
    float4* record;
    float* cov3D;
    uint8_t* clamped;
    float4* saved;
    int32_t* radii_state;      // owner-side copy of the radii inside the shard state (forward writes, backward reads)
    float* mean_x;             // optional planes of the pixel-space means (splat exchange)
    float* mean_y;
    const float* grad_record;  // backward only
    float* dL_dmeans2D;        // backward only
    float4* dq_scratch;        // backward only
    float* bg_scratch;         // backward only
};

struct MultiViewArgs {
    adgs_model m;
    adgs_model g;
    int render_objmask;
    int num_views;
    int accumulate;
    int _pad;
    ViewIO v[kMaxViews];
};

// SPLIT: blockIdx.y selects ONE view per CTA (as many threads as the single-GPU kernel has, every CTA's stores going
// to one destination); otherwise each thread loops over all views and reads its parameters once.
template <int TPB, int MINB, bool SPLIT>
__global__ void __launch_bounds__(TPB, MINB) shard_forward_multi_kernel(const __grid_constant__ MultiViewArgs a)
{
    __shared__ CamSmem cam;
    __shared__ float s_bg[6];
    __shared__ float4 s_stage[TPB / 32][128];  // per warp: 32 records on their way out, coalesced (warp_store_records)
    const adgs_model& m = a.m;
    const int N = m.N_scene + m.N_obj;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < N;
    const bool is_obj = valid && g >= m.N_scene;
    const int j = g - m.N_scene;

    // view-independent part: raw parameters -> activations, SH block
    float x0[3] = {0.f, 0.f, 0.f}, scale[3] = {1.f, 1.f, 1.f}, sig = 0.f;
    float4 qscene = make_float4(1.f, 0.f, 0.f, 0.f);
    float sh[48];
    const float4* sh4 = reinterpret_cast<const float4*>(m.sh4);
    if (valid) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            x0[d] = m.xyz[3 * (size_t)g + d];
            scale[d] = expf(m.scaling[3 * (size_t)g + d]);
        }
        sig = 1.0f / (1.0f + expf(-m.opacity[g]));
        if (!is_obj) qscene = reinterpret_cast<const float4*>(m.rotation)[g];
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const float4 v = __ldg(sh4 + (size_t)q * N + g);
            sh[4 * q + 0] = v.x;
            sh[4 * q + 1] = v.y;
            sh[4 * q + 2] = v.z;
            sh[4 * q + 3] = v.w;
        }
    }
    const float dc0[3] = {sh[0], sh[1], sh[2]};

    const int v_begin = SPLIT ? (int)blockIdx.y : 0, v_end = SPLIT ? (int)blockIdx.y + 1 : a.num_views;
    for (int vi = v_begin; vi < v_end; ++vi) {
        const ViewIO& V = a.v[vi];
        const adgs_time_basis& tb = V.tb;
        __syncthreads();
        if (threadIdx.x < 6) {
            const int d = threadIdx.x % 3;
            const bool second = threadIdx.x >= 3;
            float v = 0.f;
            for (int t = 0; t < tb.background.n; ++t)
                v += m.background_deform[d * tb.background.n_cols + tb.background.col[t]] *
                     (second ? tb.background.w1[t] : tb.background.w0[t]);
            s_bg[threadIdx.x] = v;
        }
        load_camera(cam, V.view, V.proj, V.campos, nullptr);
        float4 rq[4];  // this Gaussian's blend record (left zero if it is culled: never read, tiles_touched = 0)
        rq[0] = rq[1] = rq[2] = rq[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        [&]() {
        if (!valid) return;
        const bool flow = tb.has_flow != 0;

        float xt[3], xf[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) xt[d] = xf[d] = x0[d];
        if (is_obj && tb.xyz.n) {
            float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f};
            lin_eval3_planar(m.xyz_deform, m.N_obj, j, tb.xyz, d0, d1, flow);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                xt[d] += d0[d];
                xf[d] += d1[d];
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            xt[d] += s_bg[d];
            xf[d] += s_bg[3 + d];
        }
        const float4 qraw = is_obj ? object_rotation_raw(m, tb, g, j, nullptr, nullptr) : qscene;
        const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
        const float rot[4] = {qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn};
        float op = sig;
        if (is_obj && tb.use_time_mask) {
            const float delta = tb.t - m.gs_time[j];
            const float2 sgm = reinterpret_cast<const float2*>(m.gs_time_sigma)[j];
            const float sigma = expf(delta < 0.0f ? sgm.x : sgm.y);
            const float z = delta / sigma;
            op *= expf(-0.5f * z * z);
        }
        float dc[3] = {dc0[0], dc0[1], dc0[2]};
        if (tb.shs.n) {
            const int Cs = tb.shs.n_cols;
            const float* sd = m.shs_deform4;
            for (int t = 0; t < tb.shs.n; ++t) {
                const float w = tb.shs.w0[t];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int e = c * Cs + tb.shs.col[t];
                    dc[c] += __ldg(sd + ((size_t)(e >> 2) * N + g) * 4 + (e & 3)) * w;
                }
            }
        }
        const float3 p = make_float3(xt[0], xt[1], xt[2]);
        SplatGeom sg;
        const bool visible = splat_geometry(p, scale, rot, nullptr, V.rp, cam.view, cam.proj, sg);
        V.saved[(size_t)g * 3 + 0] = make_float4(xt[0], xt[1], xt[2], op);
        V.saved[(size_t)g * 3 + 1] = make_float4(rot[0], rot[1], rot[2], rot[3]);
        V.saved[(size_t)g * 3 + 2] = make_float4(dc[0], dc[1], dc[2], 0.f);
        if (!visible) {
            V.radii[g] = 0;
            if (V.radii_state) V.radii_state[g] = 0;
            V.tiles_touched[g] = 0;
            V.depth_keys[g] = 0xFFFFFFFFu;
            return;
        }
        if (V.radii_state) V.radii_state[g] = sg.radius;
        if (V.mean_x) {
            V.mean_x[g] = sg.px;
            V.mean_y[g] = sg.py;
        }
        sh[0] = dc[0];
        sh[1] = dc[1];
        sh[2] = dc[2];
        float rgb[3];
        uint32_t clamped;
        sh_to_rgb(V.rp.sh_degree, p, cam.campos, sh, rgb, clamped);
        V.clamped[g] = (uint8_t)clamped;
        float2* c2 = reinterpret_cast<float2*>(V.cov3D + (size_t)g * 6);
        c2[0] = make_float2(sg.cov3D[0], sg.cov3D[1]);
        c2[1] = make_float2(sg.cov3D[2], sg.cov3D[3]);
        c2[2] = make_float2(sg.cov3D[4], sg.cov3D[5]);
        const float dfeat = V.rp.inv_depth ? (1.0f / (sg.depth + 0.0000001f)) : sg.depth;
        const float sem0 = (a.render_objmask && is_obj) ? 1.f : 0.f;
        make_blend_record(rq, sg.px, sg.py, sg.conic_x, sg.conic_y, sg.conic_z, op, sg.depth, rgb, dfeat,
                          flow ? xf[0] : 0.f, flow ? xf[1] : 0.f, flow ? xf[2] : 0.f, sem0);
        V.radii[g] = sg.radius;
        V.tiles_touched[g] = sg.tiles;
        V.depth_keys[g] = __float_as_uint(sg.depth);
        }();
        // the records may live in another GPU's memory: the warp writes its 2 KB block with 512-byte instructions
        const int warp_g0 = g - (int)(threadIdx.x & 31);
        warp_store_records(V.record + (size_t)warp_g0 * 4, warp_g0 < N ? (size_t)(N - warp_g0) * 4 : 0, rq,
                           s_stage[threadIdx.x >> 5]);
    }
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) shard_backward_multi_kernel(const __grid_constant__ MultiViewArgs a)
{
    __shared__ CamSmem cam;
    __shared__ float s_wshs[kMaxViews][ADGS_MAX_TERMS * 2];  // dense SH-deform weights per view and column
    __shared__ float s_red[8][6];
    __shared__ uint8_t s_chan[kShsDeformElems];
    const adgs_model& m = a.m;
    const int N = m.N_scene + m.N_obj;
    const int Cs = a.v[0].tb.shs.n_cols;
    for (int i = threadIdx.x; i < kMaxViews * ADGS_MAX_TERMS * 2; i += blockDim.x) (&s_wshs[0][0])[i] = 0.f;
    if (Cs > 0)
        for (int e = threadIdx.x; e < (3 * Cs + 3) / 4 * 4; e += blockDim.x) s_chan[e] = (uint8_t)(e / Cs);
    __syncthreads();
    if (threadIdx.x < (unsigned)a.num_views) {
        const adgs_lin_basis& b = a.v[threadIdx.x].tb.shs;
        for (int t = 0; t < b.n; ++t) s_wshs[threadIdx.x][b.col[t]] += b.w0[t];
    }

    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < N;
    const bool is_obj = valid && g >= m.N_scene;
    const int j = g - m.N_scene;

    float scale[3] = {1.f, 1.f, 1.f}, sig = 0.f;
    const float4* sh4 = reinterpret_cast<const float4*>(m.sh4);
    if (valid) {
#pragma unroll
        for (int d = 0; d < 3; ++d) scale[d] = expf(m.scaling[3 * (size_t)g + d]);
        sig = 1.0f / (1.0f + expf(-m.opacity[g]));
    }
    // accumulators over the views
    float axyz[3] = {0.f, 0.f, 0.f}, ascale[3] = {0.f, 0.f, 0.f}, adq[4] = {0.f, 0.f, 0.f, 0.f};
    float aop = 0.f, asig0 = 0.f, asig1 = 0.f;
    float adsh[48];
    float ddc[kMaxViews][3];
#pragma unroll
    for (int i = 0; i < 48; ++i) adsh[i] = 0.f;
#pragma unroll
    for (int v = 0; v < kMaxViews; ++v) ddc[v][0] = ddc[v][1] = ddc[v][2] = 0.f;
    float4 rot_scene = make_float4(1.f, 0.f, 0.f, 0.f);

    // The gradient records of view v may live in ANOTHER GPU's memory (peer-memory exchange): the warp loads its 2 KB
    // block with four fully coalesced 512-byte instructions (a lane fetching its own record would issue 16-byte
    // loads 64 bytes apart, each its own NVLink transaction), one view AHEAD of its use so that the round trip of a
    // few microseconds is covered by the current view's arithmetic, and transposes it through shared memory.
    // (A cp.async ring of four views in shared memory was tried: 0.60 ms instead of 0.34 ms at 8 GPUs, removed.)
    __shared__ float4 s_stage[TPB / 32][128];
    const int warp_g0 = g - (int)(threadIdx.x & 31);
    const size_t warp_quads = warp_g0 < N ? (size_t)(N - warp_g0) * 4 : 0;
    float4 nr[4];  // the warp's next block, quad k of the block in lane k % 32, not yet transposed
    nr[0] = nr[1] = nr[2] = nr[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto request_view = [&](int vi) {
        if (vi < a.num_views)
            warp_load_records_issue(reinterpret_cast<const float4*>(a.v[vi].grad_record) + (size_t)warp_g0 * 4, warp_quads,
                                    nr);
    };
    request_view(0);

#pragma unroll
    for (int vi = 0; vi < kMaxViews; ++vi) {
        if (vi >= a.num_views) break;
        const ViewIO& V = a.v[vi];
        const adgs_time_basis& tb = V.tb;
        __syncthreads();
        load_camera(cam, V.view, V.proj, V.campos, nullptr);
        const bool flow = tb.has_flow != 0;
        float dxt[3] = {0.f, 0.f, 0.f}, dfl[3] = {0.f, 0.f, 0.f};
        float4 gq[4];
        warp_load_records_finish(nr, gq, s_stage[threadIdx.x >> 5]);
        const float4 g0 = gq[0], g1 = gq[1], g2 = gq[2], g3 = gq[3];
        request_view(vi + 1);
        const int radius = valid ? V.radii[g] : 0;
        if (valid) {
            if (V.dL_dmeans2D) {
                V.dL_dmeans2D[3 * (size_t)g + 0] = g0.x;
                V.dL_dmeans2D[3 * (size_t)g + 1] = g0.y;
                V.dL_dmeans2D[3 * (size_t)g + 2] = 0.f;
            }
            const bool visible = radius > 0;
            const float4 sv0 = V.saved[(size_t)g * 3 + 0];
            const float4 sv1 = V.saved[(size_t)g * 3 + 1];
            const float rot[4] = {sv1.x, sv1.y, sv1.z, sv1.w};
            if (!is_obj) rot_scene = sv1;
            float dscale[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
            if (flow) {
                dfl[0] = g2.z;
                dfl[1] = g2.w;
                dfl[2] = g3.x;
            }
            if (visible) {
                const float3 p = make_float3(sv0.x, sv0.y, sv0.z);
                float cv[6], dcov[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) cv[i] = V.cov3D[(size_t)g * 6 + i];
                const float3 dm = cov2d_bwd(p, V.rp, cv, cam.view, g0.z, g0.w, g1.x, dcov);
                const float3 dm2 = mean_proj_depth_bwd(p, cam.view, cam.proj, g0.x, g0.y, g2.y, V.rp.inv_depth);
                const float4 sv2 = V.saved[(size_t)g * 3 + 2];
                float sh[48];  // re-read per view: later views hit L1/L2, and no 48 registers stay live
#pragma unroll
                for (int q = 0; q < 12; ++q) {
                    const float4 v = __ldg(sh4 + (size_t)q * N + g);
                    sh[4 * q + 0] = v.x;
                    sh[4 * q + 1] = v.y;
                    sh[4 * q + 2] = v.z;
                    sh[4 * q + 3] = v.w;
                }
                sh[0] = sv2.x;
                sh[1] = sv2.y;
                sh[2] = sv2.z;
                const float dcol[3] = {g1.z, g1.w, g2.x};
                const float before[3] = {adsh[0], adsh[1], adsh[2]};
                const float3 dm3 = sh_to_rgb_bwd<true>(V.rp.sh_degree, p, cam.campos, sh, V.clamped[g], dcol, adsh);
                ddc[vi][0] = adsh[0] - before[0];
                ddc[vi][1] = adsh[1] - before[1];
                ddc[vi][2] = adsh[2] - before[2];
                dxt[0] = dm.x + dm2.x + dm3.x;
                dxt[1] = dm.y + dm2.y + dm3.y;
                dxt[2] = dm.z + dm2.z + dm3.z;
                cov3d_bwd(scale, V.rp.scale_modifier, rot, dcov, dscale, dq);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                axyz[d] += dxt[d] + dfl[d];
                ascale[d] += dscale[d] * scale[d];
            }
            {
                const float dop = g1.y;
                float mask = 1.f;
                if (is_obj && tb.use_time_mask) {
                    const float delta = tb.t - m.gs_time[j];
                    const float2 sgm = reinterpret_cast<const float2*>(m.gs_time_sigma)[j];
                    const bool neg = delta < 0.0f;
                    const float sigma = expf(neg ? sgm.x : sgm.y);
                    const float z = delta / sigma;
                    mask = expf(-0.5f * z * z);
                    const float dside = dop * sig * mask * z * z;
                    if (neg) asig0 += dside; else asig1 += dside;
                }
                aop += dop * mask * sig * (1.f - sig);
            }
            if (!is_obj) {
#pragma unroll
                for (int d = 0; d < 4; ++d) adq[d] += dq[d];
            } else {
                V.dq_scratch[j] = make_float4(dq[0], dq[1], dq[2], dq[3]);
            }
            // windows differ between views: read-modify-write on planes the host zero-filled
            if (is_obj && tb.xyz.n && a.g.xyz_deform) {
                for (int t = 0; t < tb.xyz.n; ++t) {
                    float* o = a.g.xyz_deform + ((size_t)tb.xyz.col[t] * 3) * m.N_obj + j;
                    const float w0 = tb.xyz.w0[t], w1 = tb.xyz.w1[t];
#pragma unroll
                    for (int d = 0; d < 3; ++d) o[(size_t)d * m.N_obj] += dxt[d] * w0 + dfl[d] * w1;
                }
            }
        }
        if (tb.background.n) {
            float r[6] = {dxt[0], dxt[1], dxt[2], dfl[0], dfl[1], dfl[2]};
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) r[i] += __shfl_xor_sync(0xffffffffu, r[i], off);
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane == 0)
#pragma unroll
                for (int i = 0; i < 6; ++i) s_red[warp][i] = r[i];
            __syncthreads();
            if (threadIdx.x < 6) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[w][threadIdx.x];
                if (s != 0.f) atomicAdd(V.bg_scratch + threadIdx.x, s);
            }
        }
    }

    if (!valid) return;
    const int acc = a.accumulate;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        put(a.g.xyz + 3 * (size_t)g + d, axyz[d], acc);
        put(a.g.scaling + 3 * (size_t)g + d, ascale[d], acc);
    }
    float4* gsh4 = reinterpret_cast<float4*>(a.g.sh4);
#pragma unroll
    for (int q = 0; q < 12; ++q)
        put4(gsh4 + (size_t)q * N + g, make_float4(adsh[4 * q], adsh[4 * q + 1], adsh[4 * q + 2], adsh[4 * q + 3]), acc);
    if (a.g.shs_deform4 && Cs > 0) {
        const int nq = (3 * Cs + 3) / 4;
        float4* gsd = reinterpret_cast<float4*>(a.g.shs_deform4);
        for (int q = 0; q < nq; ++q) {
            float v[4];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                const int e = 4 * q + e4;
                const int c = s_chan[e];
                float sum = 0.f;
                if (c < 3) {
#pragma unroll
                    for (int vv = 0; vv < kMaxViews; ++vv)
                        if (vv < a.num_views) sum += (c == 0 ? ddc[vv][0] : (c == 1 ? ddc[vv][1] : ddc[vv][2])) * s_wshs[vv][e - c * Cs];
                }
                v[e4] = sum;
            }
            put4(gsd + (size_t)q * N + g, make_float4(v[0], v[1], v[2], v[3]), acc);
        }
    }
    put(a.g.opacity + g, aop, acc);
    if (is_obj && a.g.gs_time_sigma) put2(reinterpret_cast<float2*>(a.g.gs_time_sigma) + j, make_float2(asig0, asig1), acc);
    if (!is_obj) {
        const float4 qraw = reinterpret_cast<const float4*>(m.rotation)[g];
        const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
        put4(reinterpret_cast<float4*>(a.g.rotation) + g,
             normalize4_bwd(rot_scene, qn, make_float4(adq[0], adq[1], adq[2], adq[3])), acc);
    }
}

// Rotation chain of the object Gaussians for every view of the batch; the control-quaternion windows
// differ between views, so the (host zero-filled) planes are accumulated with read-modify-write.
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) rotation_backward_multi_kernel(const __grid_constant__ MultiViewArgs a)
{
    const adgs_model& m = a.m;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m.N_obj) return;
    const int g = m.N_scene + j;
    const float4* rd = reinterpret_cast<const float4*>(m.rot_deform);
    float4* grd = reinterpret_cast<float4*>(a.g.rot_deform);
    float4 grot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int vi = 0; vi < a.num_views; ++vi) {
        const ViewIO& V = a.v[vi];
        const adgs_time_basis& tb = V.tb;
        float4 qraw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tb.quat.n_ctrl == 0) qraw = reinterpret_cast<const float4*>(m.rotation)[g];
        for (int t = 0; t < tb.rotation.n; ++t) {
            const float4 p = __ldg(rd + (size_t)tb.rotation.col[t] * m.N_obj + j);
            const float w = tb.rotation.w0[t];
            qraw.x += p.x * w;
            qraw.y += p.y * w;
            qraw.z += p.z * w;
            qraw.w += p.w * w;
        }
        Quat qt[ADGS_MAX_QUAT_ORDER + 1], P[ADGS_MAX_QUAT_ORDER + 1], E[ADGS_MAX_QUAT_ORDER + 1];
        float3 om[ADGS_MAX_QUAT_ORDER + 1];
        float norms[ADGS_MAX_QUAT_ORDER + 1];
        const int k = tb.quat.k;
        if (tb.quat.n_ctrl != 0) {
#pragma unroll
            for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
                norms[i] = 1.f;
                if (i <= k) {
                    qt[i] = ctrl_quat(__ldg(rd + (size_t)(tb.quat.start + i) * m.N_obj + j), norms[i]);
                } else {
                    qt[i] = Quat{0.f, 0.f, 0.f, 1.f};
                }
            }
            const Quat r = quat_spline_cached(qt, k, tb.quat.cum, P, E, om);
            qraw.x += r.w;
            qraw.y += r.x;
            qraw.z += r.y;
            qraw.w += r.z;
        }
        const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
        const float4 qhat = make_float4(qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn);
        const float4 graw = normalize4_bwd(qhat, qn, V.dq_scratch[j]);  // wxyz
        if (tb.quat.n_ctrl == 0) {
            grot.x += graw.x;
            grot.y += graw.y;
            grot.z += graw.z;
            grot.w += graw.w;
        }
        if (!grd) continue;
        for (int t = 0; t < tb.rotation.n; ++t) {
            const float w = tb.rotation.w0[t];
            put4(grd + (size_t)tb.rotation.col[t] * m.N_obj + j,
                 make_float4(graw.x * w, graw.y * w, graw.z * w, graw.w * w), 1);
        }
        if (tb.quat.n_ctrl != 0) {
            Quat gqt[ADGS_MAX_QUAT_ORDER + 1];
            quat_spline_bwd(qt, k, tb.quat.cum, P, E, om, Quat{graw.y, graw.z, graw.w, graw.x}, gqt);
#pragma unroll
            for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
                if (i <= k) {
                    const float4 nq = make_float4(qt[i].w, qt[i].x, qt[i].y, qt[i].z);
                    const float4 gg = make_float4(gqt[i].w, gqt[i].x, gqt[i].y, gqt[i].z);
                    put4(grd + (size_t)(tb.quat.start + i) * m.N_obj + j, normalize4_bwd(nq, norms[i], gg), 1);
                }
            }
        }
    }
    put4(reinterpret_cast<float4*>(a.g.rotation) + g, grot, a.accumulate);
}

// View-parallel variant for batches in which every view has a quaternion spline: one thread per
// (object, view) -- blockIdx.y = view -- so a rank that owns N_obj / G objects still fills the GPU with
// N_obj / G * V threads (at G = 8 the per-object loop over 8 views left 244 CTAs of serial work). The
// per-view contributions to the control-quaternion planes meet in global memory with 16-byte vector REDs
// on the host-zero-filled planes.
__device__ __forceinline__ void red_add4(float4* p, const float4& v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) rotation_backward_views_kernel(const __grid_constant__ MultiViewArgs a)
{
    const adgs_model& m = a.m;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m.N_obj) return;
    const int g = m.N_scene + j;
    const int vi = blockIdx.y;
    const float4* rd = reinterpret_cast<const float4*>(m.rot_deform);
    float4* grd = reinterpret_cast<float4*>(a.g.rot_deform);
    // with a spline the static object quaternion is unused: its gradient is zero (written once, by view 0)
    if (vi == 0) put4(reinterpret_cast<float4*>(a.g.rotation) + g, make_float4(0.f, 0.f, 0.f, 0.f), a.accumulate);
    const ViewIO& V = a.v[vi];
    const adgs_time_basis& tb = V.tb;
    float4 qraw = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < tb.rotation.n; ++t) {
        const float4 p = __ldg(rd + (size_t)tb.rotation.col[t] * m.N_obj + j);
        const float w = tb.rotation.w0[t];
        qraw.x += p.x * w;
        qraw.y += p.y * w;
        qraw.z += p.z * w;
        qraw.w += p.w * w;
    }
    Quat qt[ADGS_MAX_QUAT_ORDER + 1], P[ADGS_MAX_QUAT_ORDER + 1], E[ADGS_MAX_QUAT_ORDER + 1];
    float3 om[ADGS_MAX_QUAT_ORDER + 1];
    float norms[ADGS_MAX_QUAT_ORDER + 1];
    const int k = tb.quat.k;
#pragma unroll
    for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
        norms[i] = 1.f;
        if (i <= k) {
            qt[i] = ctrl_quat(__ldg(rd + (size_t)(tb.quat.start + i) * m.N_obj + j), norms[i]);
        } else {
            qt[i] = Quat{0.f, 0.f, 0.f, 1.f};
        }
    }
    const Quat r = quat_spline_cached(qt, k, tb.quat.cum, P, E, om);
    qraw.x += r.w;
    qraw.y += r.x;
    qraw.z += r.y;
    qraw.w += r.z;
    const float qn = fmaxf(sqrtf(sumsq4_pinned(qraw.x, qraw.y, qraw.z, qraw.w)), 1e-12f);
    const float4 qhat = make_float4(qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn);
    const float4 graw = normalize4_bwd(qhat, qn, V.dq_scratch[j]);  // wxyz
    if (!grd) return;
    for (int t = 0; t < tb.rotation.n; ++t) {
        const float w = tb.rotation.w0[t];
        red_add4(grd + (size_t)tb.rotation.col[t] * m.N_obj + j, make_float4(graw.x * w, graw.y * w, graw.z * w, graw.w * w));
    }
    Quat gqt[ADGS_MAX_QUAT_ORDER + 1];
    quat_spline_bwd(qt, k, tb.quat.cum, P, E, om, Quat{graw.y, graw.z, graw.w, graw.x}, gqt);
#pragma unroll
    for (int i = 0; i <= ADGS_MAX_QUAT_ORDER; ++i) {
        if (i <= k) {
            const float4 nq = make_float4(qt[i].w, qt[i].x, qt[i].y, qt[i].z);
            const float4 gg = make_float4(gqt[i].w, gqt[i].x, gqt[i].y, gqt[i].z);
            red_add4(grd + (size_t)(tb.quat.start + i) * m.N_obj + j, normalize4_bwd(nq, norms[i], gg));
        }
    }
}

// background gradient from the per-view sums
__global__ void background_finalize_multi_kernel(const __grid_constant__ MultiViewArgs a)
{
    float* out = a.g.background_deform;
    if (!out) return;
    const int C = a.v[0].tb.background.n_cols;
    if (!a.accumulate)
        for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) out[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        for (int vi = 0; vi < a.num_views; ++vi) {
            const adgs_lin_basis& b = a.v[vi].tb.background;
            const float* sums = a.v[vi].bg_scratch;
            for (int t = 0; t < b.n; ++t) out[d * C + b.col[t]] += sums[d] * b.w0[t] + sums[3 + d] * b.w1[t];
        }
    }
}

// ---- host-side stages shared by the single-GPU entry points and the splat-exchange entry points ----

void launch_fused_forward(const FusedFwdArgs& a, int N, cudaStream_t stream)
{
    // sweep r1w: reading the (3,Cs) colour-deformation block as whole float4 chunks with dense weights instead of
    // the 3 n scalar gathers: 0.145 vs 0.135 ms (slower: the scalar loads hit the same L1 lines) -- dropped
    // measured on B200 (profiles/README.md, sweep r1d): 128 threads x 8 CTAs/SM (64 registers) 0.135 ms,
    // 256 x 3 (80 registers) 0.153 ms -- latency-bound on HBM, so occupancy beats the few spilled words
    static const int v = tune_variant("ADGS_TUNE_FWD", 0);
    switch (v) {
    case 1: fused_forward_kernel<256, 3><<<(N + 255) / 256, 256, 0, stream>>>(a); break;
    case 2: fused_forward_kernel<256, 4><<<(N + 255) / 256, 256, 0, stream>>>(a); break;
    default: fused_forward_kernel<128, 8><<<(N + 127) / 128, 128, 0, stream>>>(a); break;
    }
    count_launch(1);
}

void launch_fused_backward(const FusedBwdArgs& a, int N, cudaStream_t stream)
{
    // sweep r1d: 128 x 4 (128 registers) 0.220 ms, 256 x 2 0.229 ms; 96 / 80 registers spill and lose (0.26 / 0.32 ms)
    static const int v = tune_variant("ADGS_TUNE_PGB", 0);
    switch (v) {
    // sweep r1w: the run-time index dsh[e / Cs] had put the 48-float SH gradient array into local memory; with the DC
    // gradient in scalars 0.221 -> 0.201 ms. Consuming / storing the SH block before the covariance chain: no
    // change (0.200); 5 CTAs/SM at 96 registers 0.219-0.227 ms
    case 1: fused_backward_kernel<256, 2, false><<<(N + 255) / 256, 256, 0, stream>>>(a); break;
    case 2: fused_backward_kernel<128, 4, false><<<(N + 127) / 128, 128, 0, stream>>>(a); break;
    default:
        if (a.dm3_scratch) {
            // SH block as its own streaming kernel (96 registers, no spill); the rest then fits 80 registers
            // (6 CTAs/SM instead of 4). sweep r2sh2, per_gaussian_backward stage: unsplit 0.187 ms, split 0.179;
            // variants 3..5: 80 registers for the SH kernel (0.182) / 96 for the main kernel (0.182) / both (0.186)
            const int grid = (N + 127) / 128;
            if (v == 3 || v == 5) sh_backward_kernel<128, 6><<<grid, 128, 0, stream>>>(a);
            else sh_backward_kernel<128, 5><<<grid, 128, 0, stream>>>(a);
            if (v == 4 || v == 5) fused_backward_kernel<128, 5, true><<<grid, 128, 0, stream>>>(a);
            else fused_backward_kernel<128, 6, true><<<grid, 128, 0, stream>>>(a);
            count_launch(1);
        } else {
            fused_backward_kernel<128, 4, false><<<(N + 127) / 128, 128, 0, stream>>>(a);
        }
        break;
    }
    count_launch(1);
}

struct SplatPtrs {  // per-Gaussian per-view state consumed by binning + blend
    float4* record;
    uint32_t* depth_keys;
    uint32_t* tiles_touched;
    int32_t* radii;
};

int run_per_gaussian_forward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                             int render_objmask, const adgs_deformed* deformed, const SplatPtrs& sp, float* cov3D,
                             uint8_t* clamped, float4* saved, cudaStream_t stream)
{
    const int N = model->N_scene + model->N_obj;
    FusedFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.m = *model;
    a.tb = *basis;
    if (deformed) a.out = *deformed;
    a.render = 1;
    a.render_objmask = render_objmask;
    a.rp = make_raster_params(cam);
    a.view = cam->viewmatrix;
    a.proj = cam->projmatrix;
    a.campos = cam->campos;
    a.radii = sp.radii;
    a.depth_keys = sp.depth_keys;
    a.tiles_touched = sp.tiles_touched;
    a.record = sp.record;
    a.cov3D = cov3D;
    a.clamped = clamped;
    a.saved = saved;
    {
        StageScope sc(kStagePerGaussianFwd, stream);
        launch_fused_forward(a, N, stream);
    }
    return check_stage("fused forward", cam->debug != 0, stream);
}

int run_blend_backward(const adgs_camera* cam, int P, int render_objmask, bool has_flow, const float4* record,
                       const BinningState& bs, const ImageState& is, int64_t capacity, const float* img_opacity,
                       const adgs_image_grads* dpix, float* grad_record, const uint32_t* counters, cudaStream_t stream)
{
    const RasterParams rp = make_raster_params(cam);
    {
        StageScope sc(kStageFills, stream);
        cudaMemsetAsync(grad_record, 0, (size_t)P * ADGS_GRAD_FLOATS * sizeof(float), stream);
    }
    BlendBwdArgs b;
    b.ranges = is.ranges;
    b.point_list = sorted_point_list(bs, rp.grid_x * rp.grid_y);
    b.record = record;
    b.semantic = nullptr;
    b.bg = cam->bg;
    b.W = rp.W;
    b.H = rp.H;
    b.D_S = render_objmask ? 1 : 0;
    b.n_contrib = is.n_contrib;
    b.img_opacity = img_opacity;
    b.dL_dcolor = dpix->dL_dcolor;
    b.dL_ddepth = dpix->dL_ddepth;
    b.dL_dflow = has_flow ? dpix->dL_dflow : nullptr;
    // the object mask itself is a constant, but its image cotangent still reaches alpha (backward.cu:597-603)
    b.dL_dsemantic = render_objmask ? dpix->dL_dsemantic : nullptr;
    b.dL_dopacity = dpix->dL_dopacity;
    b.grad_record = grad_record;
    b.dL_dsemantic_g = nullptr;
    b.counters = counters;
    b.cull_mask = bs.cull_mask;
    if (capacity > 0) {
        StageScope sc(kStageBlendBwd, stream);
        launch_blend_backward(b, has_flow, stream);
    }
    return check_stage("blend backward", cam->debug != 0, stream);
}

// Dense-gradient semantics: everything outside the active control-point columns is zero. The kernels write
// every active plane in full, so only the complement is memset (plane = all objects of a column); when
// accumulating over the views of a batch only the first view fills. `fill_stream` may be a side stream: the
// memsets are pure HBM writes and overlap the issue-bound blend backward (adgs_render_backward).
void issue_gradient_fills(const adgs_model* model, const adgs_time_basis* basis, const adgs_model* grads,
                          int accumulate, float* bg_scratch, cudaStream_t stream)
{
    const int No = model->N_obj;
    cudaMemsetAsync(bg_scratch, 0, 32 * sizeof(float), stream);
    auto zero_inactive = [&](float* base, const adgs_lin_basis& lin, int quat_start, int quat_count,
                             size_t plane_floats) {
        const int C = lin.n_cols;
        if (!base || C <= 0 || No <= 0) return;
        bool active[2 * ADGS_MAX_TERMS + 64];
        const int cap = (int)(sizeof(active) / sizeof(active[0]));
        if (C > cap) {
            cudaMemsetAsync(base, 0, (size_t)C * plane_floats * sizeof(float), stream);
            return;
        }
        for (int c = 0; c < C; ++c) active[c] = false;
        for (int t = 0; t < lin.n; ++t) active[lin.col[t]] = true;
        for (int i = 0; i < quat_count; ++i)
            if (quat_start + i < C) active[quat_start + i] = true;
        int c = 0;
        while (c < C) {
            if (active[c]) {
                ++c;
                continue;
            }
            int e = c;
            while (e < C && !active[e]) ++e;
            cudaMemsetAsync(base + (size_t)c * plane_floats, 0, (size_t)(e - c) * plane_floats * sizeof(float), stream);
            c = e;
        }
    };
    if (!accumulate && !basis->sparse_grads) {
        zero_inactive(grads->xyz_deform, basis->xyz, 0, 0, (size_t)3 * No);
        zero_inactive(grads->rot_deform, basis->rotation, basis->quat.start,
                      basis->quat.n_ctrl ? basis->quat.k + 1 : 0, (size_t)4 * No);
    }
}

// One non-blocking side stream + fork / join events per device, created on first use.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
    std::mutex mu;  // the fork .. join enqueue sequence of one backward must not interleave with another host thread's
};

SideStream& side_stream()
{
    static SideStream table[64];
    static bool tried[64] = {};
    static std::mutex init_mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    SideStream& s = table[dev];
    std::lock_guard<std::mutex> lock(init_mu);
    if (!tried[dev]) {
        tried[dev] = true;
        s.ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
        if (!s.ok) cudaGetLastError();
    }
    return s;
}

int run_per_gaussian_backward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                              const int32_t* radii, const float* cov3D, const uint8_t* clamped, const float4* saved,
                              const float* grad_record, const adgs_model* grads, int accumulate, float* dL_dmeans2D,
                              float4* dq_scratch, float* bg_scratch, cudaStream_t stream, bool fills_done = false,
                              float4* dm3_scratch = nullptr)
{
    const bool debug = cam->debug != 0;
    const int N = model->N_scene + model->N_obj;
    const int No = model->N_obj;
    int st;
    if (basis->sparse_grads && accumulate) return ADGS_ERR_ARG;  // unwritten planes cannot be accumulated into
    if (!fills_done) {
        StageScope sc(kStageFills, stream);
        issue_gradient_fills(model, basis, grads, accumulate, bg_scratch, stream);
    }
    FusedBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.m = *model;
    a.g = *grads;
    a.tb = *basis;
    a.rp = make_raster_params(cam);
    a.view = cam->viewmatrix;
    a.proj = cam->projmatrix;
    a.campos = cam->campos;
    a.radii = radii;
    a.cov3D = cov3D;
    a.clamped = clamped;
    a.saved = saved;
    a.grad_record = grad_record;
    a.dL_dmeans2D = dL_dmeans2D;
    a.dq_scratch = dq_scratch;
    a.bg_scratch = bg_scratch;
    a.dm3_scratch = dm3_scratch;
    a.accumulate = accumulate;
    {
        StageScope sc(kStagePerGaussianBwd, stream);
        launch_fused_backward(a, N, stream);
    }
    if ((st = check_stage("fused backward", debug, stream))) return st;
    if (No > 0) {
        StageScope sc(kStageRotationBwd, stream);
        // sweep r1g (B200): 149 registers 0.086 ms, 128 registers 0.074 ms, 96 registers (some spills) 0.068 ms --
        // a long dependent FP32 / MUFU chain per thread: more resident warps beat the spilled words
        static const int rv = tune_variant("ADGS_TUNE_ROT", 0);
        switch (rv) {
        case 1: rotation_backward_kernel<128, 1><<<(No + 127) / 128, 128, 0, stream>>>(a); break;
        case 2: rotation_backward_kernel<128, 6><<<(No + 127) / 128, 128, 0, stream>>>(a); break;
        case 3: rotation_backward_kernel<128, 8><<<(No + 127) / 128, 128, 0, stream>>>(a); break;
        default: rotation_backward_kernel<128, 5><<<(No + 127) / 128, 128, 0, stream>>>(a); break;
        }
        count_launch(1);
    }
    if ((st = check_stage("rotation backward", debug, stream))) return st;
    if (grads->background_deform && basis->background.n_cols > 0) {
        count_launch(1);
        background_finalize_kernel<<<1, 128, 0, stream>>>(a);
        if ((st = check_stage("background finalize", debug, stream))) return st;
    }
    return ADGS_OK;
}

}  // namespace
}  // namespace adgs

using namespace adgs;

extern "C" {

int adgs_trajectory_forward(const adgs_model* model, const adgs_time_basis* basis, const adgs_deformed* out,
                            adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate_model(model, basis);
    if (st) return st;
    if (!out) return ADGS_ERR_ARG;
    const int N = model->N_scene + model->N_obj;
    if (N == 0) return ADGS_OK;
    FusedFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.m = *model;
    a.tb = *basis;
    a.out = *out;
    a.render = 0;
    launch_fused_forward(a, N, stream);
    return check_stage("trajectory forward", false, stream);
}

size_t adgs_render_saved_bytes(int32_t N)
{
    return (size_t)(N > 0 ? N : 0) * kSavedFloats * sizeof(float) + 256;
}

size_t adgs_render_scratch_bytes(int32_t N, int32_t N_obj)
{
    // gradient records (N x 64 B) + dq (N_obj x 16 B) + background sums + dm3 (N x 16 B, SH kernel -> main kernel)
    return (size_t)(N > 0 ? N : 0) * (ADGS_GRAD_FLOATS * sizeof(float) + 16) + (size_t)(N_obj > 0 ? N_obj : 1) * 16 + 1280;
}

static int check_render_args(const adgs_camera* cam)
{
    if (!cam || !cam->viewmatrix || !cam->projmatrix || !cam->campos || !cam->bg) return ADGS_ERR_ARG;
    if (cam->image_width <= 0 || cam->image_height <= 0) return ADGS_ERR_ARG;
    if (cam->sh_degree < 0 || cam->sh_degree > 3) return ADGS_ERR_UNSUPPORTED;
    return ADGS_OK;
}

int adgs_render_forward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                        int32_t render_objmask, const adgs_images* out, const adgs_deformed* deformed,
                        char* geometry, char* binning, int64_t capacity, adgs_alloc_fn binning_alloc,
                        void* alloc_user, char* image, char* saved, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate_model(model, basis);
    if (st) return st;
    if ((st = check_render_args(cam))) return st;
    if (!out || !geometry || !image || !saved || capacity < 0) return ADGS_ERR_ARG;
    if (!binning && !binning_alloc) return ADGS_ERR_ARG;
    if (!out->depth || !out->opacity) return ADGS_ERR_ARG;
    const int N = model->N_scene + model->N_obj;
    if (N == 0) return ADGS_ERR_ARG;
    GeometryState gs = GeometryState::from_chunk(geometry, (size_t)N);
    ImageState is = ImageState::from_chunk(image, cam->image_width, cam->image_height);
    cudaMemsetAsync(gs.counters, 0, 32 * sizeof(uint32_t), stream);
    int32_t* radii = out->radii ? out->radii : gs.radii;
    char* sc = saved;
    float4* saved4 = nullptr;
    carve(sc, saved4, (size_t)N * 3);
    SplatPtrs sp{reinterpret_cast<float4*>(gs.record), gs.depth_keys, gs.tiles_touched, radii};
    if ((st = run_per_gaussian_forward(cam, model, basis, render_objmask, deformed, sp, gs.cov3D, gs.clamped, saved4,
                                       stream)))
        return st;
    int R = 0;
    st = bin_and_blend(cam, N, render_objmask ? 1 : 0, basis->has_flow != 0, nullptr, out, radii, gs, binning,
                       binning_alloc, alloc_user, capacity, is, binning == nullptr, &R, stream);
    return st ? st : R;
}

int adgs_render_backward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                         int32_t render_objmask, const int32_t* radii, const char* geometry, const char* binning,
                         int64_t capacity, const char* image, const char* saved, const float* img_opacity,
                         const adgs_image_grads* dpix, const adgs_model* grads, float* dL_dmeans2D, char* scratch,
                         adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate_model(model, basis);
    if (st) return st;
    if ((st = check_render_args(cam))) return st;
    if (!geometry || !binning || !image || !saved || !img_opacity || !dpix || !grads || !scratch || capacity < 0)
        return ADGS_ERR_ARG;
    if (!grads->xyz || !grads->scaling || !grads->rotation || !grads->opacity || !grads->sh4) return ADGS_ERR_ARG;
    const int N = model->N_scene + model->N_obj;
    if (N == 0) return ADGS_ERR_ARG;
    char* gc = const_cast<char*>(geometry);
    char* bc = const_cast<char*>(binning);
    char* ic = const_cast<char*>(image);
    char* svc = const_cast<char*>(saved);
    GeometryState gs = GeometryState::from_chunk(gc, (size_t)N);
    BinningState bs = BinningState::from_chunk(bc, (size_t)capacity);
    ImageState is = ImageState::from_chunk(ic, cam->image_width, cam->image_height);
    if (!radii) radii = gs.radii;
    char* sc = scratch;
    float* grad_record = nullptr;
    float4* dq_scratch = nullptr;
    float* bg_scratch = nullptr;
    carve(sc, grad_record, (size_t)N * ADGS_GRAD_FLOATS);
    carve(sc, dq_scratch, (size_t)(model->N_obj > 0 ? model->N_obj : 1));
    carve(sc, bg_scratch, 32);
    float4* dm3_scratch = nullptr;
    carve(sc, dm3_scratch, (size_t)N);
    float4* saved4 = nullptr;
    carve(svc, saved4, (size_t)N * 3);
    // The zero-fill of the inactive control-point planes (~260 MB at 250 k object Gaussians) is pure HBM write
    // traffic and the blend backward is issue-bound: run the fill on a side stream underneath it
    // (fork / join with events; ADGS_TUNE_FILL=1 keeps everything on the caller's stream).
    static const int fill_variant = tune_variant("ADGS_TUNE_FILL", 0);
    bool fills_done = false;
    SideStream* side = nullptr;
    std::unique_lock<std::mutex> side_lock;
    if (fill_variant == 0 && !basis->sparse_grads && model->N_obj > 0 && !cam->debug) {
        side = &side_stream();
        if (side->ok) side_lock = std::unique_lock<std::mutex>(side->mu);
        if (side->ok && cudaEventRecord(side->fork, stream) == cudaSuccess &&
            cudaStreamWaitEvent(side->stream, side->fork, 0) == cudaSuccess) {
            issue_gradient_fills(model, basis, grads, 0, bg_scratch, side->stream);
            fills_done = cudaEventRecord(side->join, side->stream) == cudaSuccess;
        }
        if (!fills_done) cudaGetLastError();
    }
    st = run_blend_backward(cam, N, render_objmask, basis->has_flow != 0, reinterpret_cast<const float4*>(gs.record),
                            bs, is, capacity, img_opacity, dpix, grad_record, gs.counters, stream);
    if (fills_done && cudaStreamWaitEvent(stream, side->join, 0) != cudaSuccess)   // always join, even on error
        return record_cuda_error(cudaGetLastError(), "gradient fill join");
    if (side_lock.owns_lock()) side_lock.unlock();
    if (st) return st;
    return run_per_gaussian_backward(cam, model, basis, radii, gs.cov3D, gs.clamped, saved4, grad_record, grads, 0,
                                     dL_dmeans2D, dq_scratch, bg_scratch, stream, fills_done, dm3_scratch);
}

/* ---- splat exchange: Gaussian-sharded front end / back end, view-sharded blend ------------------------ */

size_t adgs_shard_state_bytes(int32_t N)
{
    // cov3D (6 floats) + clamped (1 byte) + saved (12 floats) per Gaussian of the shard, per view
    const size_t n = (size_t)(N > 0 ? N : 0);
    return n * 6 * sizeof(float) + n + n * kSavedFloats * sizeof(float) + n * sizeof(int32_t) + 4 * 128 + 256;
}

struct ShardState {
    float* cov3D;
    uint8_t* clamped;
    float4* saved;
    int32_t* radii;  // the owner's copy of the radii (the splats' own copy may live in another GPU's memory)
    static ShardState from_chunk(char*& c, size_t N)
    {
        ShardState s;
        carve(c, s.cov3D, N * 6);
        carve(c, s.clamped, N);
        carve(c, s.saved, N * 3);
        carve(c, s.radii, N);
        return s;
    }
};

size_t adgs_shard_state_radii_offset(int32_t N)
{
    char* c = nullptr;
    ShardState s = ShardState::from_chunk(c, (size_t)(N > 0 ? N : 0));
    return (size_t)s.radii;   // relative to a 128-byte aligned chunk base
}

int adgs_shard_forward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                       int32_t render_objmask, const adgs_splats* out, char* shard_state, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate_model(model, basis);
    if (st) return st;
    if ((st = check_render_args(cam))) return st;
    const int N = model->N_scene + model->N_obj;
    if (!out || !shard_state || N == 0 || out->P != N) return ADGS_ERR_ARG;
    if (!out->record || !out->depth_keys || !out->tiles_touched || !out->radii) return ADGS_ERR_ARG;
    char* c = shard_state;
    ShardState ss = ShardState::from_chunk(c, (size_t)N);
    SplatPtrs sp{reinterpret_cast<float4*>(out->record), out->depth_keys, out->tiles_touched, out->radii};
    return run_per_gaussian_forward(cam, model, basis, render_objmask, nullptr, sp, ss.cov3D, ss.clamped, ss.saved,
                                    stream);
}

static int splats_forward_stages(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                                 const adgs_images* out, char* geometry, char* binning, int64_t capacity,
                                 adgs_alloc_fn binning_alloc, void* alloc_user, char* image, cudaStream_t stream,
                                 int stages)
{
    int st = check_render_args(cam);
    if (st) return st;
    if (!splats || !geometry || !image || capacity < 0 || splats->P <= 0) return ADGS_ERR_ARG;
    if (!binning && !binning_alloc) return ADGS_ERR_ARG;
    if (D_S < 0 || D_S > 1) return ADGS_ERR_ARG;
    if ((stages & 2) && (!out || !out->depth || !out->opacity || !splats->record)) return ADGS_ERR_ARG;
    if ((stages & 1) && (!splats->depth_keys || !splats->tiles_touched || !splats->radii)) return ADGS_ERR_ARG;
    if ((stages & 1) && !splats->record && !splats->mean_x) return ADGS_ERR_ARG;
    const int P = splats->P;
    GeometryState gs = GeometryState::from_chunk(geometry, (size_t)P);
    ImageState is = ImageState::from_chunk(image, cam->image_width, cam->image_height);
    if (stages & 1) cudaMemsetAsync(gs.counters, 0, 32 * sizeof(uint32_t), stream);
    // binning + blend read the per-Gaussian state from the caller's (gathered) arrays
    gs.record = splats->record;
    gs.depth_keys = splats->depth_keys;
    gs.tiles_touched = splats->tiles_touched;
    int R = 0;
    adgs_images none;
    memset(&none, 0, sizeof(none));
    st = bin_and_blend(cam, P, D_S, has_flow != 0, nullptr, out ? out : &none, splats->radii, gs, binning,
                       binning_alloc, alloc_user, capacity, is, binning == nullptr, &R, stream, stages,
                       splats->mean_x, splats->mean_y);
    return st ? st : R;
}

int adgs_splats_forward(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                        const adgs_images* out, char* geometry, char* binning, int64_t capacity,
                        adgs_alloc_fn binning_alloc, void* alloc_user, char* image, adgs_stream_t stream_)
{
    return splats_forward_stages(cam, splats, D_S, has_flow, out, geometry, binning, capacity, binning_alloc,
                                 alloc_user, image, (cudaStream_t)stream_, 3);
}

int adgs_splats_bin(const adgs_camera* cam, const adgs_splats* splats, char* geometry, char* binning,
                    int64_t capacity, adgs_alloc_fn binning_alloc, void* alloc_user, char* image,
                    adgs_stream_t stream_)
{
    return splats_forward_stages(cam, splats, 0, 0, nullptr, geometry, binning, capacity, binning_alloc, alloc_user,
                                 image, (cudaStream_t)stream_, 1);
}

int adgs_splats_blend(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                      const adgs_images* out, char* geometry, char* binning, int64_t capacity, char* image,
                      adgs_stream_t stream_)
{
    if (!binning) return ADGS_ERR_ARG;
    int st = splats_forward_stages(cam, splats, D_S, has_flow, out, geometry, binning, capacity, nullptr, nullptr,
                                   image, (cudaStream_t)stream_, 2);
    return st < 0 ? st : ADGS_OK;
}

int adgs_splats_backward(const adgs_camera* cam, const adgs_splats* splats, int32_t D_S, int32_t has_flow,
                         const char* binning, int64_t capacity, const char* image, const float* img_opacity,
                         const adgs_image_grads* dpix, float* grad_record, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = check_render_args(cam);
    if (st) return st;
    if (!splats || !binning || !image || !img_opacity || !dpix || !grad_record || capacity < 0 || splats->P <= 0)
        return ADGS_ERR_ARG;
    char* bc = const_cast<char*>(binning);
    char* ic = const_cast<char*>(image);
    BinningState bs = BinningState::from_chunk(bc, (size_t)capacity);
    ImageState is = ImageState::from_chunk(ic, cam->image_width, cam->image_height);
    return run_blend_backward(cam, splats->P, D_S, has_flow != 0, reinterpret_cast<const float4*>(splats->record), bs,
                              is, capacity, img_opacity, dpix, grad_record, nullptr, stream);
}

int adgs_shard_backward(const adgs_camera* cam, const adgs_model* model, const adgs_time_basis* basis,
                        const int32_t* radii, const char* shard_state, const float* grad_record,
                        const adgs_model* grads, int32_t accumulate, float* dL_dmeans2D, char* scratch,
                        adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = validate_model(model, basis);
    if (st) return st;
    if ((st = check_render_args(cam))) return st;
    if (!radii || !shard_state || !grad_record || !grads || !scratch) return ADGS_ERR_ARG;
    if (!grads->xyz || !grads->scaling || !grads->rotation || !grads->opacity || !grads->sh4) return ADGS_ERR_ARG;
    const int N = model->N_scene + model->N_obj;
    if (N == 0) return ADGS_ERR_ARG;
    char* c = const_cast<char*>(shard_state);
    ShardState ss = ShardState::from_chunk(c, (size_t)N);
    char* sc = scratch;
    float4* dq_scratch = nullptr;
    float* bg_scratch = nullptr;
    carve(sc, dq_scratch, (size_t)(model->N_obj > 0 ? model->N_obj : 1));
    carve(sc, bg_scratch, 32);
    return run_per_gaussian_backward(cam, model, basis, radii, ss.cov3D, ss.clamped, ss.saved, grad_record, grads,
                                     accumulate, dL_dmeans2D, dq_scratch, bg_scratch, stream);
}

/* multi-view variants: one launch per stage for all views of a round (adgs_b200/parallel.py) */
int adgs_shard_forward_multi(int32_t num_views, const adgs_camera* cams, const adgs_model* model,
                             const adgs_time_basis* bases, int32_t render_objmask, const adgs_splats* outs,
                             char* const* shard_states, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_views < 1 || num_views > kMaxViews || !cams || !bases || !outs || !shard_states) return ADGS_ERR_ARG;
    const int N = model ? model->N_scene + model->N_obj : 0;
    if (N == 0) return ADGS_ERR_ARG;
    static MultiViewArgs a;  // 17 KB: keep it off the stack
    memset(&a, 0, sizeof(a));
    a.m = *model;
    a.render_objmask = render_objmask;
    a.num_views = num_views;
    for (int v = 0; v < num_views; ++v) {
        int st = validate_model(model, &bases[v]);
        if (st) return st;
        if ((st = check_render_args(&cams[v]))) return st;
        if (outs[v].P != N || !outs[v].record || !outs[v].depth_keys || !outs[v].tiles_touched || !outs[v].radii ||
            !shard_states[v])
            return ADGS_ERR_ARG;
        char* c = shard_states[v];
        ShardState ss = ShardState::from_chunk(c, (size_t)N);
        ViewIO& V = a.v[v];
        V.tb = bases[v];
        V.rp = make_raster_params(&cams[v]);
        V.view = cams[v].viewmatrix;
        V.proj = cams[v].projmatrix;
        V.campos = cams[v].campos;
        V.radii = outs[v].radii;
        V.depth_keys = outs[v].depth_keys;
        V.tiles_touched = outs[v].tiles_touched;
        V.record = reinterpret_cast<float4*>(outs[v].record);
        V.cov3D = ss.cov3D;
        V.clamped = ss.clamped;
        V.saved = ss.saved;
        V.radii_state = ss.radii;
        V.mean_x = outs[v].mean_x;
        V.mean_y = outs[v].mean_y;
    }
    {
        StageScope sc(kStagePerGaussianFwd, stream);
        // ADGS_TUNE_MFWD: 0 = every thread loops over the views (parameters read once; best for a few large views,
        // e.g. a camera rig on one GPU), 1 = one view per CTA (grid.y = views: the thread count of the single-GPU
        // kernel when the model is sharded over as many ranks as there are views), -1 = pick by shard size
        static const int variant = tune_variant("ADGS_TUNE_MFWD", -1);
        const bool split = variant == 1 || (variant == -1 && num_views > 1 && (long long)N * num_views <= 2000000ll);
        if (split)
            shard_forward_multi_kernel<128, 8, true><<<dim3((N + 127) / 128, num_views), 128, 0, stream>>>(a);
        else
            shard_forward_multi_kernel<256, 2, false><<<(N + 255) / 256, 256, 0, stream>>>(a);
        count_launch(1);
    }
    return check_stage("shard forward (multi-view)", cams[0].debug != 0, stream);
}

int adgs_shard_backward_multi(int32_t num_views, const adgs_camera* cams, const adgs_model* model,
                              const adgs_time_basis* bases, const int32_t* const* radii,
                              char* const* shard_states, const float* const* grad_records,
                              const adgs_model* grads, int32_t accumulate, float* const* dL_dmeans2D,
                              char* scratch, adgs_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_views < 1 || num_views > kMaxViews || !cams || !bases || !radii || !shard_states || !grad_records ||
        !grads || !scratch)
        return ADGS_ERR_ARG;
    if (!grads->xyz || !grads->scaling || !grads->rotation || !grads->opacity || !grads->sh4) return ADGS_ERR_ARG;
    const int N = model ? model->N_scene + model->N_obj : 0;
    if (N == 0) return ADGS_ERR_ARG;
    const int No = model->N_obj;
    const bool debug = cams[0].debug != 0;
    static MultiViewArgs a;
    memset(&a, 0, sizeof(a));
    a.m = *model;
    a.g = *grads;
    a.num_views = num_views;
    // accumulate: bit 0 = add into `grads` (later rounds of a batch); bit 1 = the caller has already zero-filled the
    // control-point planes of grads->xyz_deform / rot_deform (e.g. on a side stream underneath the blend backward)
    const bool planes_zeroed = (accumulate & 2) != 0;
    accumulate &= 1;
    a.accumulate = accumulate;
    // scratch: per view dq (N_obj float4) + 32 floats of background sums
    char* sc = scratch;
    for (int v = 0; v < num_views; ++v) {
        int st = validate_model(model, &bases[v]);
        if (st) return st;
        if ((st = check_render_args(&cams[v]))) return st;
        if (!shard_states[v] || !grad_records[v]) return ADGS_ERR_ARG;
        char* c = shard_states[v];
        ShardState ss = ShardState::from_chunk(c, (size_t)N);
        ViewIO& V = a.v[v];
        V.tb = bases[v];
        V.rp = make_raster_params(&cams[v]);
        V.view = cams[v].viewmatrix;
        V.proj = cams[v].projmatrix;
        V.campos = cams[v].campos;
        V.radii = radii[v] ? const_cast<int32_t*>(radii[v]) : ss.radii;   // null: the copy adgs_shard_forward_multi kept
        V.cov3D = ss.cov3D;
        V.clamped = ss.clamped;
        V.saved = ss.saved;
        V.grad_record = grad_records[v];
        V.dL_dmeans2D = dL_dmeans2D ? dL_dmeans2D[v] : nullptr;
        carve(sc, V.dq_scratch, (size_t)(No > 0 ? No : 1));
    }
    float* bg_all = nullptr;  // the views' background sums in one block: one memset instead of one per view
    carve(sc, bg_all, (size_t)32 * kMaxViews);
    for (int v = 0; v < num_views; ++v) a.v[v].bg_scratch = bg_all + 32 * v;
    {
        StageScope scope(kStageFills, stream);
        cudaMemsetAsync(bg_all, 0, (size_t)32 * num_views * sizeof(float), stream);
        if (!accumulate && !planes_zeroed && No > 0) {
            // windows differ between views and are accumulated with += : start from all-zero planes
            if (grads->xyz_deform && bases[0].xyz.n_cols > 0)
                cudaMemsetAsync(grads->xyz_deform, 0, (size_t)bases[0].xyz.n_cols * 3 * No * sizeof(float), stream);
            if (grads->rot_deform && bases[0].rotation.n_cols > 0)
                cudaMemsetAsync(grads->rot_deform, 0, (size_t)bases[0].rotation.n_cols * 4 * No * sizeof(float), stream);
        }
    }
    int st;
    {
        StageScope scope(kStagePerGaussianBwd, stream);
        // sweep r1l (2 GPUs): 168 registers 0.289 ms, 128 registers 0.244 ms, 230 registers 0.366 ms
        shard_backward_multi_kernel<128, 4><<<(N + 127) / 128, 128, 0, stream>>>(a);
        count_launch(1);
    }
    if ((st = check_stage("shard backward (multi-view)", debug, stream))) return st;
    if (No > 0) {
        StageScope scope(kStageRotationBwd, stream);
        // sweep r1l: 199 registers 0.123 ms, 128 registers 0.084 ms, 96 registers 0.093 ms
        bool all_spline = a.g.rot_deform != nullptr;
        for (int v = 0; v < num_views; ++v) all_spline = all_spline && bases[v].quat.n_ctrl != 0;
        if (all_spline && num_views > 1) {
            // per-(object, view) threads; single-GPU sweep r1g: 96 registers best for this chain
            rotation_backward_views_kernel<128, 5><<<dim3((No + 127) / 128, num_views), 128, 0, stream>>>(a);
        } else {
            rotation_backward_multi_kernel<128, 4><<<(No + 127) / 128, 128, 0, stream>>>(a);
        }
        count_launch(1);
    }
    if ((st = check_stage("rotation backward (multi-view)", debug, stream))) return st;
    if (grads->background_deform && bases[0].background.n_cols > 0) {
        background_finalize_multi_kernel<<<1, 128, 0, stream>>>(a);
        count_launch(1);
        if ((st = check_stage("background finalize (multi-view)", debug, stream))) return st;
    }
    return ADGS_OK;
}

size_t adgs_shard_scratch_bytes(int32_t num_views, int32_t N_obj)
{
    return (size_t)(num_views > 0 ? num_views : 1) * ((size_t)(N_obj > 0 ? N_obj : 1) * 16 + 128 + 256) + 256 +
           (size_t)32 * kMaxViews * sizeof(float) + 128;
}

}  // extern "C"
