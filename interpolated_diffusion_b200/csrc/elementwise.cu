// K2 / K2b: the HBM-bound elementwise epilogues around the two denoisers.
//
// Reference arithmetic being replaced (paths under the reference root):
//   src/diffusion/ddpm.py:15-24, 37-48                 q_sample, ddim_step (eta = 0)
//   src/sample/sample_generate.py:260-280, 319-360     _build_known_mask_values, _build_anchor_conf, _anneal_conf
//   src/sample/sample_generate.py:397-399, 1252-1285   known clamp, Stage-2 residual / soft / hard clamp
//   src/utils/clamp.py:4-32, src/utils/normalize.py:4-20
//
// Every fp32 op is an explicitly rounded intrinsic in the reference's order (bit-identical to the
// eager op-by-op evaluation; IEEE sqrt/div).  All kernels are grid-stride with grids sized in
// multiples of the SM count; traffic is exactly one read of each input and one write of the output.
#include <type_traits>

#include "common.cuh"

namespace idb200 {

constexpr int kThreads = 256;

// schedule-table index of a timestep: torch's indexing wraps negatives once; anything still outside [0, n) is clamped so a
// bad timestep can never read outside the table (the reference raises IndexError: the Python mirror's check_t does too)
__device__ __forceinline__ long long table_index(long long t, int n) {
    if (t < 0) t += n;
    return t < 0 ? 0 : (t >= n ? n - 1 : t);
}

__global__ void __launch_bounds__(kThreads) ddim_step_kernel(
    const float* __restrict__ z, const float* __restrict__ eps, const long long* __restrict__ t,
    const long long* __restrict__ t_prev, const float* __restrict__ alpha_bar, int n_train, float ab_t_s, float ab_p_s, long long n,
    long long row_len, int D, const unsigned char* __restrict__ known_mask, const float* __restrict__ known_values,
    int pos_clip, float clip_min, float clip_max, float* __restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float ab_t = ab_t_s, ab_p = ab_p_s;
        if (t) {
            const long long row = i / row_len;
            ab_t = __ldg(alpha_bar + table_index(t[row], n_train));
            ab_p = __ldg(alpha_bar + table_index(t_prev[row], n_train));
        }
        const float e = eps[i];
        // x0 = (rt - sqrt(1 - ab_t) * eps) / sqrt(ab_t)                       ddpm.py:45
        const float x0 = __fdiv_rn(__fsub_rn(z[i], __fmul_rn(__fsqrt_rn(__fsub_rn(1.0f, ab_t)), e)), __fsqrt_rn(ab_t));
        // rt_prev = sqrt(ab_prev) * x0 + sqrt(1 - ab_prev) * eps               ddpm.py:47
        float r = __fadd_rn(__fmul_rn(__fsqrt_rn(ab_p), x0), __fmul_rn(__fsqrt_rn(__fsub_rn(1.0f, ab_p)), e));
        if (known_mask && known_mask[i]) r = known_values[i];                    // sample_generate.py:398
        if (pos_clip && (i % D) < 2) r = fminf(fmaxf(r, clip_min), clip_max);    // :382-386
        out[i] = r;
    }
}

// Batch-constant timestep (the generation loop, sample_generate.py:394-395): the four square roots are per-launch scalars
// and the row is streamed as float4 / uchar4 (n % 4 == 0, D | 4, 16-byte aligned).  Same operations in the same order as
// ddim_step_kernel, so the result is bit-identical.
__global__ void __launch_bounds__(kThreads) ddim_step_scalar4_kernel(
    const float4* __restrict__ z, const float4* __restrict__ eps, float ab_t, float ab_p, long long n4, int D,
    const uchar4* __restrict__ known_mask, const float4* __restrict__ known_values, int pos_clip, float clip_min, float clip_max,
    float4* __restrict__ out) {
    const float s_1mt = __fsqrt_rn(__fsub_rn(1.0f, ab_t)), s_t = __fsqrt_rn(ab_t);
    const float s_p = __fsqrt_rn(ab_p), s_1mp = __fsqrt_rn(__fsub_rn(1.0f, ab_p));
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 zz = z[i], e = eps[i];
        float r[4];
        const float zv[4] = {zz.x, zz.y, zz.z, zz.w}, ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float x0 = __fdiv_rn(__fsub_rn(zv[j], __fmul_rn(s_1mt, ev[j])), s_t);            // ddpm.py:45
            r[j] = __fadd_rn(__fmul_rn(s_p, x0), __fmul_rn(s_1mp, ev[j]));                       // ddpm.py:47
        }
        if (known_mask) {                                                                        // sample_generate.py:398
            const uchar4 km = known_mask[i];
            if (km.x | km.y | km.z | km.w) {
                const float4 kv = known_values[i];
                if (km.x) r[0] = kv.x;
                if (km.y) r[1] = kv.y;
                if (km.z) r[2] = kv.z;
                if (km.w) r[3] = kv.w;
            }
        }
        if (pos_clip) {                                                                          // :382-386 (dims 0, 1 of every D)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((i * 4 + j) % D) < 2) r[j] = fminf(fmaxf(r[j], clip_min), clip_max);
        }
        out[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void __launch_bounds__(kThreads) q_sample_kernel(const float* __restrict__ r0, const float* __restrict__ noise,
                                                            const long long* __restrict__ t, const float* __restrict__ sab,
                                                            const float* __restrict__ s1m, int n_train, long long n, long long row_len,
                                                            float* __restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long tt = table_index(t[i / row_len], n_train);
        out[i] = __fadd_rn(__fmul_rn(__ldg(sab + tt), r0[i]), __fmul_rn(__ldg(s1m + tt), noise[i]));  // ddpm.py:24
    }
}

__device__ __forceinline__ float logit_clamped(float v, float eps) {
    // normalize.py:9-10: pos.clamp(eps, 1-eps); log(pos / (1 - pos)).  1-eps is the fp32 scalar.
    const float hi = static_cast<float>(1.0 - static_cast<double>(eps));
    const float p = fminf(fmaxf(v, eps), hi);
    return logf(__fdiv_rn(p, __fsub_rn(1.0f, p)));
}

__device__ __forceinline__ float sigmoid_f(float v) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
}

__global__ void __launch_bounds__(kThreads) known_mask_values_kernel(const long long* __restrict__ idx,
                                                                     const float* __restrict__ sg, long long B, int K, int D,
                                                                     int T, int clamp_endpoints, int logit_space,
                                                                     float logit_eps, unsigned char* __restrict__ km,
                                                                     float* __restrict__ kv) {
    const long long n = B * K * D;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = static_cast<int>(i % D);
        const long long bk = i / D;
        const long long b = bk / K;
        unsigned char m = 0;
        float v = 0.0f;
        if (clamp_endpoints && D >= 2 && d < 2) {
            const long long ix = idx[bk];
            if (ix == 0) { m = 1; v = sg[b * 4 + d]; }
            if (ix == T - 1) { m = 1; v = sg[b * 4 + 2 + d]; }          // goal wins when T == 1 (:278-279)
        }
        if (logit_space && D >= 2 && d < 2) v = logit_clamped(v, logit_eps);  // applied to zeros too (harmless, masked)
        km[i] = m;
        kv[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads) pos_transform_kernel(const float* __restrict__ x, long long n, int D, int mode,
                                                                 float eps, float* __restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = x[i];
        if (D >= 2 && (i % D) < 2) v = mode ? sigmoid_f(v) : logit_clamped(v, eps);
        out[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads) stage2_epilogue_kernel(
    const float* __restrict__ x_in, const float* __restrict__ delta, const float* __restrict__ x_ref,
    const float* __restrict__ conf, float lam, int policy, const unsigned char* __restrict__ clamp_mask, int dims_all,
    int pos_clip, float clip_min, float clip_max, long long B, int T, int D, float* __restrict__ out) {
    const long long n = B * T * D;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = static_cast<int>(i % D);
        const long long bt = i / D;
        const int t = static_cast<int>(bt % T);
        float x = x_in[i];
        if (delta) x = __fadd_rn(x, delta[i]);                                  // x_hat = x_pred + delta_hat
        const bool in_dims = dims_all || d < 2;
        if (in_dims) {
            if (conf && lam > 0.0f) {                                            // clamp.py:26-32
                const float w = __fmul_rn(conf[bt], lam);
                x = __fadd_rn(x, __fmul_rn(w, __fsub_rn(x_ref[i], x)));
            }
            const bool hard = (policy == IDB200_CLAMP_ENDPOINTS) ? (t == 0 || t == T - 1)
                              : (policy == IDB200_CLAMP_MASK) ? (clamp_mask[bt] != 0) : false;
            if (hard) x = x_ref[i];                                              // clamp.py:7-10
        }
        if (pos_clip && d < 2) x = fminf(fmaxf(x, clip_min), clip_max);
        out[i] = x;
    }
}

// D = 2 or 4, aligned rows: one thread per token (b, t), the token's D values as one float2 / float4; no per-element integer
// divisions.  Same operations in the same order as stage2_epilogue_kernel (bit-identical).
template <int kD>
__global__ void __launch_bounds__(kThreads) stage2_epilogue_vec_kernel(
    const float* __restrict__ x_in, const float* __restrict__ delta, const float* __restrict__ x_ref,
    const float* __restrict__ conf, float lam, int policy, const unsigned char* __restrict__ clamp_mask, int dims_all,
    int pos_clip, float clip_min, float clip_max, long long BT, int T, float* __restrict__ out) {
    using V = typename std::conditional<kD == 2, float2, float4>::type;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const bool soft = conf != nullptr && lam > 0.0f;
    for (long long bt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; bt < BT; bt += stride) {
        float x[kD], xr[kD];
        {
            const V v = reinterpret_cast<const V*>(x_in)[bt];
            const float* vp = reinterpret_cast<const float*>(&v);
#pragma unroll
            for (int d = 0; d < kD; ++d) x[d] = vp[d];
        }
        if (delta) {
            const V v = reinterpret_cast<const V*>(delta)[bt];
            const float* vp = reinterpret_cast<const float*>(&v);
#pragma unroll
            for (int d = 0; d < kD; ++d) x[d] = __fadd_rn(x[d], vp[d]);
        }
        const int t = static_cast<int>(bt % T);
        const bool hard = (policy == IDB200_CLAMP_ENDPOINTS) ? (t == 0 || t == T - 1) : (policy == IDB200_CLAMP_MASK) ? (clamp_mask[bt] != 0) : false;
        if (soft || hard) {
            const V v = reinterpret_cast<const V*>(x_ref)[bt];
            const float* vp = reinterpret_cast<const float*>(&v);
#pragma unroll
            for (int d = 0; d < kD; ++d) xr[d] = vp[d];
            const float w = soft ? __fmul_rn(conf[bt], lam) : 0.0f;
#pragma unroll
            for (int d = 0; d < kD; ++d) {
                if (dims_all || d < 2) {
                    if (soft) x[d] = __fadd_rn(x[d], __fmul_rn(w, __fsub_rn(xr[d], x[d])));
                    if (hard) x[d] = xr[d];
                }
            }
        }
        if (pos_clip) {
#pragma unroll
            for (int d = 0; d < 2; ++d) x[d] = fminf(fmaxf(x[d], clip_min), clip_max);
        }
        V o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int d = 0; d < kD; ++d) op[d] = x[d];
        reinterpret_cast<V*>(out)[bt] = o;
    }
}

__global__ void __launch_bounds__(kThreads) anchor_conf_kernel(
    const unsigned char* __restrict__ mask_s, const unsigned char* __restrict__ student,
    const unsigned char* __restrict__ mask_prev, const long long* __restrict__ s_row, int s_scalar, int levels, int anneal,
    float c_teacher, float c_student, float c_end, float c_missing, int clamp_endpoints, long long B, int T, int C,
    float* __restrict__ conf, float* __restrict__ mask_in) {
    const long long n = B * T;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long b = i / T;
        const int t = static_cast<int>(i - b * T);
        const bool m = mask_s[i] != 0;
        float c = m ? c_teacher : c_missing;                                     // sample_generate.py:329-330
        if (student && m && student[i]) c = c_student;                           // :331-332
        if (clamp_endpoints && (t == 0 || t == T - 1)) c = c_end;                // :333-335
        if (anneal && levels > 0) {                                              // :350-360 / train :565-576
            float lamv;
            if (s_row) {
                const float frac = __fdiv_rn(static_cast<float>(s_row[b]), static_cast<float>(levels));
                lamv = (anneal == 1) ? __fsub_rn(1.0f, frac)
                                     : __fmul_rn(0.5f, __fadd_rn(1.0f, cosf(__fmul_rn(3.14159274101257324f, frac))));
            } else {
                const double frac = static_cast<double>(s_scalar) / static_cast<double>(levels);
                lamv = static_cast<float>((anneal == 1) ? 1.0 - frac : 0.5 * (1.0 + cos(3.141592653589793 * frac)));
            }
            c = __fadd_rn(c, __fmul_rn(__fsub_rn(1.0f, c), lamv));
        }
        if (conf) conf[i] = c;
        if (mask_in) {
            float* mi = mask_in + i * C;
            mi[0] = m ? 1.0f : 0.0f;
            if (C == 3) { mi[1] = mask_prev[i] ? 1.0f : 0.0f; mi[2] = c; }
            else mi[1] = c;
        }
    }
}

static inline int ew_grid(long long n) { return grid_for(n, kThreads * 4, 8); }

}  // namespace idb200

using namespace idb200;

extern "C" int idb200_ddim_step(const float* z, const float* eps, const int64_t* t, const int64_t* t_prev,
                                const float* alpha_bar, int n_train, float ab_t, float ab_prev, int64_t n_rows,
                                int64_t row_len, int D, const uint8_t* known_mask, const float* known_values, int pos_clip,
                                float clip_min, float clip_max, float* z_out, idb200_stream_t stream) {
    IDB_REQUIRE(z && eps && z_out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(n_rows >= 0 && row_len >= 1 && D >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE((t == nullptr) == (t_prev == nullptr), IDB200_EINVAL, "t and t_prev must both be given or both NULL");
    IDB_REQUIRE(t == nullptr || (alpha_bar != nullptr && n_train > 0), IDB200_EINVAL, "alpha_bar table missing");
    IDB_REQUIRE((known_mask == nullptr) == (known_values == nullptr), IDB200_EINVAL, "known_mask/known_values mismatch");
    const long long n = n_rows * row_len;
    if (n == 0) return IDB200_OK;
    if (t == nullptr && n % 4 == 0 && idb200::aligned(z, 16) && idb200::aligned(eps, 16) && idb200::aligned(z_out, 16) &&
        (!known_mask || (idb200::aligned(known_mask, 4) && idb200::aligned(known_values, 16)))) {
        ddim_step_scalar4_kernel<<<ew_grid(n / 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(eps), ab_t, ab_prev, n / 4, D,
            reinterpret_cast<const uchar4*>(known_mask), reinterpret_cast<const float4*>(known_values), pos_clip, clip_min, clip_max,
            reinterpret_cast<float4*>(z_out));
        return check_launch("ddim_step_scalar4_kernel");
    }
    ddim_step_kernel<<<ew_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        z, eps, reinterpret_cast<const long long*>(t), reinterpret_cast<const long long*>(t_prev), alpha_bar, n_train, ab_t, ab_prev,
        n, row_len, D, known_mask, known_values, pos_clip, clip_min, clip_max, z_out);
    return check_launch("ddim_step_kernel");
}

extern "C" int idb200_q_sample(const float* r0, const float* noise, const int64_t* t, const float* sqrt_ab,
                               const float* sqrt_1m_ab, int n_train, int64_t n_rows, int64_t row_len, float* out,
                               idb200_stream_t stream) {
    IDB_REQUIRE(r0 && noise && t && sqrt_ab && sqrt_1m_ab && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(n_rows >= 0 && row_len >= 1 && n_train > 0, IDB200_EINVAL, "bad shape");
    const long long n = n_rows * row_len;
    if (n == 0) return IDB200_OK;
    q_sample_kernel<<<ew_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        r0, noise, reinterpret_cast<const long long*>(t), sqrt_ab, sqrt_1m_ab, n_train, n, row_len, out);
    return check_launch("q_sample_kernel");
}

extern "C" int idb200_known_mask_values(const int64_t* idx, const float* start_goal, int64_t B, int K, int D, int T,
                                        int clamp_endpoints, int logit_space, float logit_eps, uint8_t* known_mask,
                                        float* known_values, idb200_stream_t stream) {
    IDB_REQUIRE(idx && known_mask && known_values, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(!clamp_endpoints || start_goal, IDB200_EINVAL, "clamp_endpoints=True but start_goal missing from cond");
    IDB_REQUIRE(B >= 0 && K >= 1 && D >= 1 && T >= 1, IDB200_EINVAL, "bad shape");
    if (B == 0) return IDB200_OK;
    known_mask_values_kernel<<<ew_grid(B * K * D), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(idx), start_goal, B, K, D, T, clamp_endpoints, logit_space, logit_eps, known_mask,
        known_values);
    return check_launch("known_mask_values_kernel");
}

extern "C" int idb200_logit_pos(const float* x, int64_t n_rows, int D, float eps, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(x && out && n_rows >= 0 && D >= 1, IDB200_EINVAL, "bad arguments");
    if (n_rows == 0) return IDB200_OK;
    pos_transform_kernel<<<ew_grid(n_rows * D), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n_rows * D, D, 0, eps, out);
    return check_launch("pos_transform_kernel(logit)");
}

extern "C" int idb200_sigmoid_pos(const float* x, int64_t n_rows, int D, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(x && out && n_rows >= 0 && D >= 1, IDB200_EINVAL, "bad arguments");
    if (n_rows == 0) return IDB200_OK;
    pos_transform_kernel<<<ew_grid(n_rows * D), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n_rows * D, D, 1, 0.0f, out);
    return check_launch("pos_transform_kernel(sigmoid)");
}

extern "C" int idb200_stage2_epilogue(const float* x_in, const float* delta, const float* x_ref, const float* conf,
                                      float lam, int policy, const uint8_t* clamp_mask, int clamp_dims_all, int pos_clip,
                                      float clip_min, float clip_max, int64_t B, int T, int D, float* out,
                                      idb200_stream_t stream) {
    IDB_REQUIRE(x_in && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && T >= 1 && D >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(policy >= 0 && policy <= 2, IDB200_EINVAL, "unknown clamp policy %d", policy);
    IDB_REQUIRE(policy != IDB200_CLAMP_MASK || clamp_mask, IDB200_EINVAL, "clamp_mask missing");
    IDB_REQUIRE((policy == IDB200_CLAMP_NONE && !(conf && lam > 0.0f)) || x_ref, IDB200_EINVAL, "x_ref missing");
    if (B == 0) return IDB200_OK;
    const size_t al = D == 2 ? 8 : 16;
    if ((D == 2 || D == 4) && idb200::aligned(x_in, al) && idb200::aligned(out, al) && (!delta || idb200::aligned(delta, al)) &&
        (!x_ref || idb200::aligned(x_ref, al))) {
        const int grid = ew_grid(B * T);
        if (D == 2)
            stage2_epilogue_vec_kernel<2><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                x_in, delta, x_ref, conf, lam, policy, clamp_mask, clamp_dims_all, pos_clip, clip_min, clip_max, B * T, T, out);
        else
            stage2_epilogue_vec_kernel<4><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                x_in, delta, x_ref, conf, lam, policy, clamp_mask, clamp_dims_all, pos_clip, clip_min, clip_max, B * T, T, out);
        return check_launch("stage2_epilogue_vec_kernel");
    }
    stage2_epilogue_kernel<<<ew_grid(B * T * D), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        x_in, delta, x_ref, conf, lam, policy, clamp_mask, clamp_dims_all, pos_clip, clip_min, clip_max, B, T, D, out);
    return check_launch("stage2_epilogue_kernel");
}

extern "C" int idb200_anchor_conf(const uint8_t* mask_s, const uint8_t* student, const uint8_t* mask_prev,
                                  const int64_t* s_row, int s_scalar, int levels, int anneal_mode, float conf_teacher,
                                  float conf_student, float conf_endpoints, float conf_missing, int clamp_endpoints,
                                  int64_t B, int T, int C, float* conf, float* mask_in, idb200_stream_t stream) {
    IDB_REQUIRE(mask_s && (conf || mask_in), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && T >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(!mask_in || C == 2 || (C == 3 && mask_prev), IDB200_EINVAL, "mask_in needs C==2, or C==3 with mask_prev");
    IDB_REQUIRE(anneal_mode >= 0 && anneal_mode <= 2, IDB200_EINVAL, "unknown anneal mode");
    if (B == 0) return IDB200_OK;
    anchor_conf_kernel<<<ew_grid(B * T), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        mask_s, student, mask_prev, reinterpret_cast<const long long*>(s_row), s_scalar, levels, anneal_mode, conf_teacher,
        conf_student, conf_endpoints, conf_missing, clamp_endpoints, B, T, C, conf, mask_in);
    return check_launch("anchor_conf_kernel");
}

// ------------------------------------------------------------------------------------------------
// Batched trajectory metrics (src/eval/metrics.py:13-24, 68-128): one warp per trajectory, lanes stride the timesteps.
//   collision_rate = mean_t [occ[b, i(t), j(t)] > 0.5  or  position outside [0,1]^2],  (i, j) = round-half-even cell
//   goal_dist = ||traj[T-1] - goal||, success = goal_dist < thr (thr = 1 / W as fp32, passed by the host),
//   path_length = sum_t ||traj[t+1] - traj[t]||, smoothness = mean_t ||traj[t+2] - 2 traj[t+1] + traj[t]|| (0 if T < 3),
//   mse_to_gt = mean over (t, d) of (traj - gt)^2.  occ_stride / goal_stride / gt_stride = 0 broadcast one row.
// ------------------------------------------------------------------------------------------------
namespace idb200 {
__global__ void __launch_bounds__(256) traj_metrics_kernel(const float* __restrict__ occ, long long occ_stride, const float* __restrict__ traj,
                                                           const float* __restrict__ goal, long long goal_stride,
                                                           const float* __restrict__ gt, long long gt_stride, long long B, int T, int D,
                                                           int H, int W, float thr, float* __restrict__ out, int n_out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const float sw = static_cast<float>(W - 1 > 1 ? W - 1 : 1), sh = static_cast<float>(H - 1 > 1 ? H - 1 : 1);
    for (long long b = warp; b < B; b += nwarps) {
        const float* x = traj + b * T * D;
        const float* oc = occ + b * occ_stride;
        const float* g = gt ? gt + b * gt_stride : nullptr;
        float coll = 0.0f, plen = 0.0f, smooth = 0.0f, mse = 0.0f;
        for (int t = lane; t < T; t += 32) {
            const float px = x[t * D], py = x[t * D + 1];
            const bool oob = (px < 0.0f) | (px > 1.0f) | (py < 0.0f) | (py > 1.0f);
            int j = static_cast<int>(rintf(px * sw)), i = static_cast<int>(rintf(py * sh));
            i = min(max(i, 0), H - 1);
            j = min(max(j, 0), W - 1);
            coll += (oob || oc[i * W + j] > 0.5f) ? 1.0f : 0.0f;
            float d1 = 0.0f, d2 = 0.0f;
            for (int d = 0; d < D; ++d) {
                const float v = x[t * D + d];
                if (t + 1 < T) {
                    const float e = x[(t + 1) * D + d] - v;
                    d1 += e * e;
                }
                if (t + 2 < T) {
                    const float a = x[(t + 2) * D + d] - 2.0f * x[(t + 1) * D + d] + v;
                    d2 += a * a;
                }
                if (g) {
                    const float e = v - g[t * D + d];
                    mse += e * e;
                }
            }
            if (t + 1 < T) plen += sqrtf(d1);
            if (t + 2 < T) smooth += sqrtf(d2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            coll += __shfl_xor_sync(0xffffffffu, coll, o);
            plen += __shfl_xor_sync(0xffffffffu, plen, o);
            smooth += __shfl_xor_sync(0xffffffffu, smooth, o);
            mse += __shfl_xor_sync(0xffffffffu, mse, o);
        }
        if (lane == 0) {
            const float* gl = goal + b * goal_stride;
            float gd = 0.0f;
            for (int d = 0; d < D; ++d) {
                const float e = x[(T - 1) * D + d] - gl[d];
                gd += e * e;
            }
            gd = sqrtf(gd);
            float* o = out + b * n_out;
            o[0] = coll / static_cast<float>(T);
            o[1] = gd;
            o[2] = gd < thr ? 1.0f : 0.0f;
            o[3] = plen;
            o[4] = T < 3 ? 0.0f : smooth / static_cast<float>(T - 2);
            if (n_out > 5) o[5] = mse / static_cast<float>(T * D);
        }
    }
}
}  // namespace idb200

extern "C" int idb200_traj_metrics(const float* occ, int64_t occ_stride, const float* traj, const float* goal, int64_t goal_stride,
                                   const float* gt, int64_t gt_stride, int64_t B, int T, int D, int H, int W, float success_thr,
                                   float* out, int n_out, idb200_stream_t stream) {
    IDB_REQUIRE(occ && traj && goal && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && T >= 1 && D >= 2 && H >= 1 && W >= 1, IDB200_EINVAL, "bad shape (traj needs at least the two position dims)");
    IDB_REQUIRE(n_out == 5 || (n_out == 6 && gt), IDB200_EINVAL, "n_out must be 5, or 6 with gt");
    if (B == 0) return IDB200_OK;
    idb200::traj_metrics_kernel<<<idb200::grid_for(B, 8, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        occ, occ_stride, traj, goal, goal_stride, gt, gt_stride, B, T, D, H, W, success_thr, out, n_out);
    return idb200::check_launch("traj_metrics_kernel");
}
