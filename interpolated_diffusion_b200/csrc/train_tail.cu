// Tail of the Stage-2 training step (src/train/train_interp_levels.py:1142-1173): weighted-MSE loss and its gradient with
// respect to the denoiser output, global-norm gradient clipping (torch.nn.utils.clip_grad_norm_), fused AdamW + EMA update
// (torch.optim.AdamW defaults of :687, src/utils/ema.py:11-17).  All HBM-bound streaming kernels over flat fp32 buffers;
// reductions are two-level with a fixed order (per-block partials, then one block), so results are deterministic.
#include "common.cuh"

namespace idb200 {
namespace tt {
constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1184;            // 148 SMs x 8

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    __syncthreads();
    return t;                                // valid in thread 0
}

// partial[2*b] = sum over this block's tokens of w * ||delta_hat - target||^2, partial[2*b+1] = sum of w
__global__ void __launch_bounds__(kThreads) loss_partial_kernel(const float* __restrict__ dh, const float* __restrict__ tg,
                                                                const float* __restrict__ conf, const unsigned char* __restrict__ mask,
                                                                float w_anchor, float w_missing, long long BT, int D,
                                                                double* __restrict__ partial) {
    __shared__ double sh[kThreads / 32];
    double num = 0.0, den = 0.0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < BT; i += stride) {
        const float w = conf ? __fadd_rn(w_missing, __fmul_rn(__fsub_rn(w_anchor, w_missing), conf[i])) : (mask[i] ? w_anchor : w_missing);
        float d2 = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float e = __fsub_rn(dh[i * D + d], tg[i * D + d]);
            d2 = __fadd_rn(d2, __fmul_rn(e, e));
        }
        num += static_cast<double>(__fmul_rn(d2, w));
        den += static_cast<double>(w);
    }
    const double n = block_sum(num, sh);
    const double dn = block_sum(den, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = n; partial[2 * blockIdx.x + 1] = dn; }
}

// scal[0] = loss, scal[1] = 1 / (sum(w) * D + 1e-8) / grad_accum  (the factor of the gradient)
__global__ void loss_final_kernel(const double* __restrict__ partial, int nblocks, int D, float grad_accum, float* __restrict__ scal) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double num = 0.0, den = 0.0;
    for (int b = 0; b < nblocks; ++b) { num += partial[2 * b]; den += partial[2 * b + 1]; }
    const float denom = static_cast<float>(den) * static_cast<float>(D) + 1e-8f;
    scal[0] = static_cast<float>(num) / denom / grad_accum;
    scal[1] = 1.0f / denom / grad_accum;
}

// grad[i, d] = 2 * w_i * (delta_hat - target) * scal[1]
__global__ void __launch_bounds__(kThreads) loss_grad_kernel(const float* __restrict__ dh, const float* __restrict__ tg,
                                                             const float* __restrict__ conf, const unsigned char* __restrict__ mask,
                                                             float w_anchor, float w_missing, long long BT, int D,
                                                             const float* __restrict__ scal, float* __restrict__ grad) {
    const float f = 2.0f * scal[1];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < BT; i += stride) {
        const float w = conf ? __fadd_rn(w_missing, __fmul_rn(__fsub_rn(w_anchor, w_missing), conf[i])) : (mask[i] ? w_anchor : w_missing);
        const float wf = w * f;
        for (int d = 0; d < D; ++d) grad[i * D + d] = (dh[i * D + d] - tg[i * D + d]) * wf;
    }
}

__global__ void __launch_bounds__(kThreads) sq_norm_partial_kernel(const float4* __restrict__ g, long long n4, const float* __restrict__ tail,
                                                                   int ntail, double* __restrict__ partial) {
    __shared__ double sh[kThreads / 32];
    double s = 0.0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = g[i];
        s += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < ntail) s += static_cast<double>(tail[threadIdx.x]) * tail[threadIdx.x];
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// scal[0] = total 2-norm, scal[1] = clip coefficient min(1, max_norm / (norm + 1e-6))   (clip_grad_norm_)
__global__ void clip_final_kernel(const double* __restrict__ partial, int nblocks, float max_norm, float* __restrict__ scal) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[b];
    const float norm = static_cast<float>(sqrt(s));
    const float c = max_norm / (norm + 1e-6f);
    scal[0] = norm;
    scal[1] = c < 1.0f ? c : 1.0f;
}

// torch/optim/adamw.py (_single_tensor_adamw) operation order on the clipped gradient, then EMA.update.  kV = 4: float4 streams.
struct AdamScalars {
    float decay_mul, one_m_b1, b2, one_m_b2, step_size, inv_sqrt_bc2, eps, ema_decay, one_m_ema;
};
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float* ema, float c, const AdamScalars& a) {
    const float gi = __fmul_rn(g, c);
    float pi = __fmul_rn(p, a.decay_mul);                                                 // p.mul_(1 - lr * wd)
    const float mi = __fadd_rn(m, __fmul_rn(__fsub_rn(gi, m), a.one_m_b1));               // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(gi, gi), a.one_m_b2));
    const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(vi), a.inv_sqrt_bc2), a.eps);      // sqrt(v) / sqrt(bc2) + eps
    pi = __fsub_rn(pi, __fmul_rn(a.step_size, __fdiv_rn(mi, denom)));                     // p.addcdiv_(m, denom, value=-step_size)
    p = pi;
    m = mi;
    v = vi;
    if (ema) *ema = __fadd_rn(__fmul_rn(*ema, a.ema_decay), __fmul_rn(pi, a.one_m_ema));
}
__global__ void __launch_bounds__(kThreads) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                             float* __restrict__ v, float* __restrict__ ema, long long n, AdamScalars a,
                                                             const float* __restrict__ coef) {
    const float c = coef ? coef[1] : 1.0f;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n4 = n >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 ee = ema ? reinterpret_cast<float4*>(ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        adamw_one(pp.x, gg.x, mm.x, vv.x, ema ? &ee.x : nullptr, c, a);
        adamw_one(pp.y, gg.y, mm.y, vv.y, ema ? &ee.y : nullptr, c, a);
        adamw_one(pp.z, gg.z, mm.z, vv.z, ema ? &ee.z : nullptr, c, a);
        adamw_one(pp.w, gg.w, mm.w, vv.w, ema ? &ee.w : nullptr, c, a);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (ema) reinterpret_cast<float4*>(ema)[i] = ee;
    }
    for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        adamw_one(p[i], g[i], m[i], v[i], ema ? ema + i : nullptr, c, a);
}

}  // namespace tt
}  // namespace idb200

using namespace idb200;

// The reduction kernels write one (loss: two) fp64 partial per block into a caller-provided scratch buffer whose size is a
// CONTRACT of the C ABI (idb200_tail_scratch_doubles): the grid is clamped to that capacity, so a device with more SMs than
// the B200's 148 can never overrun a buffer sized by the exported count.
static constexpr int kTailMaxBlocks = 2048;
static inline int tail_grid(int grid) { return grid < kTailMaxBlocks ? grid : kTailMaxBlocks; }

extern "C" int64_t idb200_tail_scratch_doubles() { return 2 * static_cast<int64_t>(kTailMaxBlocks); }

extern "C" int idb200_stage2_loss(const float* delta_hat, const float* target, const float* conf, const uint8_t* mask, float w_anchor,
                                  float w_missing, float grad_accum, int64_t B, int T, int D, double* scratch, float* loss_scal,
                                  float* grad_out, idb200_stream_t stream) {
    IDB_REQUIRE(delta_hat && target && scratch && loss_scal && ((conf != nullptr) != (mask != nullptr)), IDB200_EINVAL,
                "NULL pointer (exactly one of conf / mask must be given)");
    IDB_REQUIRE(B >= 1 && T >= 1 && D >= 1 && grad_accum > 0.0f, IDB200_EINVAL, "bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long BT = static_cast<long long>(B) * T;
    const int grid = tail_grid(grid_for(BT, tt::kThreads, 8));
    tt::loss_partial_kernel<<<grid, tt::kThreads, 0, st>>>(delta_hat, target, conf, mask, w_anchor, w_missing, BT, D, scratch);
    tt::loss_final_kernel<<<1, 32, 0, st>>>(scratch, grid, D, grad_accum, loss_scal);
    if (grad_out) tt::loss_grad_kernel<<<grid, tt::kThreads, 0, st>>>(delta_hat, target, conf, mask, w_anchor, w_missing, BT, D, loss_scal, grad_out);
    return check_launch("stage2_loss kernels");
}

extern "C" int idb200_grad_clip_coef(const float* grad, int64_t n, float max_norm, double* scratch, float* norm_coef,
                                     idb200_stream_t stream) {
    IDB_REQUIRE(grad && scratch && norm_coef && n >= 1, IDB200_EINVAL, "bad arguments");
    IDB_REQUIRE(aligned(grad, 16), IDB200_EALIGN, "grad must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n4 = n / 4;
    const int grid = tail_grid(grid_for(n4 > 0 ? n4 : 1, tt::kThreads * 4, 8));
    tt::sq_norm_partial_kernel<<<grid, tt::kThreads, 0, st>>>(reinterpret_cast<const float4*>(grad), n4, grad + n4 * 4, static_cast<int>(n - n4 * 4), scratch);
    tt::clip_final_kernel<<<1, 32, 0, st>>>(scratch, grid, max_norm, norm_coef);
    return check_launch("grad_clip kernels");
}

extern "C" int idb200_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                                     float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float ema_decay,
                                     const float* norm_coef, idb200_stream_t stream) {
    IDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 1 && step >= 1, IDB200_EINVAL, "bad arguments");
    // scalar prefactors in double, rounded once (what Python floats give torch)
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    IDB_REQUIRE(aligned(param, 16) && aligned(grad, 16) && aligned(exp_avg, 16) && aligned(exp_avg_sq, 16) && (!ema || aligned(ema, 16)),
                IDB200_EALIGN, "optimizer buffers must be 16-byte aligned");
    tt::AdamScalars a{static_cast<float>(1.0 - static_cast<double>(lr) * weight_decay), static_cast<float>(1.0 - static_cast<double>(beta1)), beta2,
                      static_cast<float>(1.0 - static_cast<double>(beta2)), static_cast<float>(static_cast<double>(lr) / bc1),
                      static_cast<float>(1.0 / sqrt(bc2)), eps, ema_decay, static_cast<float>(1.0 - static_cast<double>(ema_decay))};
    tt::adamw_ema_kernel<<<grid_for(n / 4 + 1, tt::kThreads * 2, 8), tt::kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, ema, n, a, norm_coef);
    return check_launch("adamw_ema_kernel");
}
