// Small (non-GEMM-dominant) kernels of the two denoisers.
//
// Reference modules being replaced (paths under the reference root):
//   src/models/encoders.py:8-71                       MazeEncoder / StartGoalEncoder / MazeConditionEncoder
//   src/models/denoiser_keypoints.py:11-34, 91-111    sinusoid embeddings, token assembly, in_proj, t_embed
//   src/models/denoiser_interp_levels.py:54-82        positional embedding, token assembly, in_proj, level_proj
//   src/models/transformer.py:28-46                   LayerNorm + FiLM (the GEMMs are in gemm.cu)
//
//   sgemm_tn      fp32 SIMT GEMM for the per-trajectory linears (cond_proj, FiLM gamma/beta, t_embed,
//                 level_proj, fc) -- M = B rows, tiny K -- and for the fp32 check mode of the token GEMMs
//   conv_encoder  3x3 pad-1 conv stack + SiLU, global mean pool (fp32, planes resident in shared memory)
//   embed_*       token assembly + in_proj as a gathered small-K product (the sinusoid part of in_proj is
//                 a [T, d] table because idx takes T values; timestep / level / cond terms are per-row adds)
//   ln_film       LayerNorm(eps 1e-5) * (1 + gamma) + beta -> bf16 (or fp32) GEMM operand
//   out_head      h[M,d] . W_out[D,d]^T + b  (D = 2 or 4: memory-bound dot products)
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace idb200 {

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM: out[M,N] (ldo) = act(A[M,K] (lda) * W[N,K]^T + bias[N]) (+ out if accumulate)
// 64x64 tile, 16-deep k-blocks, 256 threads, 4x4 register tile.
// ------------------------------------------------------------------------------------------------
template <typename TA>
__device__ __forceinline__ float to_f32(TA v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TA>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const TA* __restrict__ A, long long lda, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ out, long long ldo,
                                                       long long M, int N, int K, int act, int accumulate) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * 64;
    const int n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            const long long m = m0 + r;
            const int k = k0 + c;
            As[c][r] = (m < M && k < K) ? to_f32<TA>(A[m * lda + k]) : 0.0f;
            const int nn = n0 + r;
            Ws[c][r] = (nn < N && k < K) ? W[static_cast<long long>(nn) * K + k] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[c][ty * 4 + i]; w[i] = Ws[c][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= N) continue;
            float v = acc[i][j] + (bias ? bias[nn] : 0.0f);
            if (act == 1) v = silu_exact(v);
            if (accumulate) v += out[m * ldo + nn];
            out[m * ldo + nn] = v;
        }
    }
}

// Small-problem form of the same product (fewer than one 64 x 64 tile per SM, e.g. the per-trajectory linears of a 512-trajectory
// training batch: [512 x 384] . [384 x 384]^T is 48 tiles on 148 SMs, one latency-bound CTA per SM: 63 us for 0.15 GFLOP):
// 32 x 32 tiles, 2 x 2 outputs per thread, k-steps of 32.  Every output is still ONE fmaf chain over ascending k: bit-identical
// to sgemm_tn_kernel.
template <typename TA>
__global__ void __launch_bounds__(256) sgemm_tn_small_kernel(const TA* __restrict__ A, long long lda, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ out, long long ldo,
                                                             long long M, int N, int K, int act, int accumulate) {
    __shared__ float As[32][34], Ws[32][34];                // [k][row]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * 32;
    const int n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
    for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            const long long m = m0 + r;
            const int k = k0 + c, nn = n0 + r;
            As[c][r] = (m < M && k < K) ? to_f32<TA>(A[m * lda + k]) : 0.0f;
            Ws[c][r] = (nn < N && k < K) ? W[static_cast<long long>(nn) * K + k] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float2 a = *reinterpret_cast<const float2*>(&As[c][ty * 2]);
            const float2 w = *reinterpret_cast<const float2*>(&Ws[c][tx * 2]);
            acc[0][0] = fmaf(a.x, w.x, acc[0][0]);
            acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
            acc[1][0] = fmaf(a.y, w.x, acc[1][0]);
            acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const long long m = m0 + ty * 2 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int nn = n0 + tx * 2 + j;
            if (nn >= N) continue;
            float v = acc[i][j] + (bias ? bias[nn] : 0.0f);
            if (act == 1) v = silu_exact(v);
            if (accumulate) v += out[m * ldo + nn];
            out[m * ldo + nn] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// conv encoder: one launch per 3x3 pad-1 conv layer (+ SiLU).  A CTA owns (trajectory, block of 32 output
// channels); input planes stream through shared memory in chunks of <= 32 channels ([23 x 23] zero-bordered
// planes for a 21 x 21 maze), each thread accumulates 4-pixel strips for its output channels in registers.
// Intermediate layers write [B, C, H, W] fp32 to a caller-provided scratch; the last layer is reduced to the
// spatial mean [B, C_last] with a fixed summation order (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kConvCo = 32;       // output channels per CTA
constexpr int kConvCi = 32;       // input channels per shared-memory chunk
constexpr int kConvItems = 16;    // (channel, strip) items per thread: supports up to 32 * 128 strips / 256 threads

struct ConvLayerParams {
    const float* in0;        // [B, ci0, H, W] (first ci0 input channels)
    const float* in1;        // [B, ci - ci0, H, W] or nullptr (sdf plane of layer 0)
    int ci0;
    const float* w;          // [co, ci, 3, 3]
    const float* bias;       // [co]
    float* out;              // [B, co, H, W] or nullptr when pooling
    float* pooled;           // [B, co] (last layer) or nullptr
    int ci, co, H, W;
    long long B;
};

__global__ void __launch_bounds__(256) conv3x3_silu_kernel(const ConvLayerParams p) {
    extern __shared__ float smem_f[];
    const int PW = p.W + 2, plane = (p.H + 2) * PW, HW = p.H * p.W;
    const int strips_per_row = (p.W + 3) / 4;
    const int strips = p.H * strips_per_row;
    float* planes = smem_f;                                   // [kConvCi][plane]
    float* red = smem_f + kConvCi * plane;                    // [kConvCo][strips] (pooling only)
    const int co_blocks = (p.co + kConvCo - 1) / kConvCo;
    const long long work = p.B * co_blocks;
    for (long long wi = blockIdx.x; wi < work; wi += gridDim.x) {
        const long long b = wi / co_blocks;
        const int c0 = static_cast<int>(wi - b * co_blocks) * kConvCo;
        const int nco = min(kConvCo, p.co - c0);
        const int items = nco * strips;
        float acc[kConvItems][4];
#pragma unroll
        for (int it = 0; it < kConvItems; ++it) {
            const int item = threadIdx.x + it * 256;
            const float bv = (item < items) ? p.bias[c0 + item / strips] : 0.0f;
            acc[it][0] = acc[it][1] = acc[it][2] = acc[it][3] = bv;
        }
        for (int k0 = 0; k0 < p.ci; k0 += kConvCi) {
            const int nk = min(kConvCi, p.ci - k0);
            __syncthreads();
            for (int i = threadIdx.x; i < nk * plane; i += 256) {
                const int k = i / plane, r = i - k * plane, yy = r / PW, xx = r - yy * PW;
                float v = 0.0f;
                if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) {
                    const int ch = k0 + k;
                    const float* src = (ch < p.ci0) ? p.in0 + (b * p.ci0 + ch) * HW : p.in1 + (b * (p.ci - p.ci0) + (ch - p.ci0)) * HW;
                    v = src[(yy - 1) * p.W + (xx - 1)];
                }
                planes[i] = v;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kConvItems; ++it) {
                const int item = threadIdx.x + it * 256;
                if (item >= items) continue;
                const int c = item / strips, sidx = item - c * strips;
                const int y = sidx / strips_per_row, x0 = (sidx - y * strips_per_row) * 4;
                const float* wc = p.w + (static_cast<long long>(c0 + c) * p.ci + k0) * 9;
                float a0 = acc[it][0], a1 = acc[it][1], a2 = acc[it][2], a3 = acc[it][3];
                for (int k = 0; k < nk; ++k) {
                    const float* ip = planes + k * plane + y * PW + x0;   // top-left of the 3 x 6 window
                    const float* wk = wc + k * 9;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const float i0 = ip[ky * PW + 0], i1 = ip[ky * PW + 1], i2 = ip[ky * PW + 2];
                        const float i3 = (x0 + 3 < PW) ? ip[ky * PW + 3] : 0.0f;
                        const float i4 = (x0 + 4 < PW) ? ip[ky * PW + 4] : 0.0f;
                        const float i5 = (x0 + 5 < PW) ? ip[ky * PW + 5] : 0.0f;
                        const float w0 = __ldg(wk + ky * 3), w1 = __ldg(wk + ky * 3 + 1), w2 = __ldg(wk + ky * 3 + 2);
                        a0 = fmaf(i0, w0, fmaf(i1, w1, fmaf(i2, w2, a0)));
                        a1 = fmaf(i1, w0, fmaf(i2, w1, fmaf(i3, w2, a1)));
                        a2 = fmaf(i2, w0, fmaf(i3, w1, fmaf(i4, w2, a2)));
                        a3 = fmaf(i3, w0, fmaf(i4, w1, fmaf(i5, w2, a3)));
                    }
                }
                acc[it][0] = a0; acc[it][1] = a1; acc[it][2] = a2; acc[it][3] = a3;
            }
        }
#pragma unroll
        for (int it = 0; it < kConvItems; ++it) {
            const int item = threadIdx.x + it * 256;
            if (item >= items) continue;
            const int c = item / strips, sidx = item - c * strips;
            const int y = sidx / strips_per_row, x0 = (sidx - y * strips_per_row) * 4;
            float part = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (x0 + j < p.W) {
                    const float v = silu_exact(acc[it][j]);
                    if (p.out) p.out[(b * p.co + c0 + c) * HW + y * p.W + x0 + j] = v;
                    part += v;
                }
            }
            if (p.pooled) red[c * strips + sidx] = part;
        }
        if (p.pooled) {
            __syncthreads();
            for (int c = threadIdx.x >> 5; c < nco; c += 8) {       // one warp per channel, fixed order
                float sum = 0.0f;
                for (int i = threadIdx.x & 31; i < strips; i += 32) sum += red[c * strips + i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if ((threadIdx.x & 31) == 0) p.pooled[b * p.co + c0 + c] = sum / static_cast<float>(HW);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sinusoid tables: out[r, :dim] = [sin(a_r * f_i), cos(a_r * f_i)], f_i = exp(-ln(1e4) * i / half)
//   mode 0: a_r = r / max(1, rows - 1)   (continuous_time_embedding of idx/(T-1); _positional_embedding)
//   mode 1: a_r = args[r]                (timestep_embedding of t)
// ------------------------------------------------------------------------------------------------
__global__ void sinusoid_kernel(const float* __restrict__ args, int rows, int dim, int mode, float* __restrict__ out) {
    const int half = dim / 2;
    const long long n = static_cast<long long>(rows) * dim;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / dim), c = static_cast<int>(i - static_cast<long long>(r) * dim);
        float v = 0.0f;
        if (c < 2 * half) {
            const int fi = (c < half) ? c : c - half;
            // freqs = exp(-log(10000) * arange(half) / half): fp32 ops in the reference's order
            const float freq = expf(__fdiv_rn(__fmul_rn(-logf(10000.0f), static_cast<float>(fi)), static_cast<float>(half)));
            const float a = (mode == 0) ? __fdiv_rn(static_cast<float>(r), fmaxf(1.0f, static_cast<float>(rows - 1))) : args[r];
            const float x = __fmul_rn(a, freq);
            v = (c < half) ? sinf(x) : cosf(x);
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// token embedding.  h[m, :] = sum_j feat[m, j] * Wf[j, :] + tab[tab_row(m), :] + row_a[b or 0, :] + row_b[b, :]
//   feat: F small features per token (z_t | known_mask | kp_feat  or  x_s | mask channels), gathered
//   from up to 3 sources; Wf is the matching column slice of in_proj.weight, transposed to [F, d].
//   tab: [T, d] table (Stage 1: in_proj(pos sinusoid)[idx]; Stage 2: fixed positional sinusoid[t]).
//   row_a: timestep / level embedding (row stride 0 when batch-constant), row_b: cond_proj(cond_vec)+bias.
// ------------------------------------------------------------------------------------------------
struct EmbedParams {
    const float* src0; int n0;            // fp32 [M, n0]
    const float* src1; int n1;            // fp32 [M, n1] or nullptr
    const unsigned char* src2; int n2;    // uint8 [M, n2] (bool mask) or nullptr
    const float* Wf;                      // [n0 + n1 + n2, d]
    const float* tab;                     // [T, d]
    const long long* tab_idx;             // int64 [M] (Stage 1 idx) or nullptr -> row = m % L
    const float* row_a; long long row_a_stride;
    const float* row_b;                   // [B, d]
    float* h;                             // [M, d]
    long long M; int L, d;
};

__global__ void __launch_bounds__(256) embed_kernel(const EmbedParams p) {
    // one warp per token; lanes stride over d
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int F = p.n0 + p.n1 + p.n2;
    for (long long m = warp; m < p.M; m += nwarps) {
        const long long b = m / p.L;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = 0.0f;
            if (j < p.n0) v = p.src0[m * p.n0 + j];
            else if (j < p.n0 + p.n1) v = p.src1[m * p.n1 + (j - p.n0)];
            else if (j < F) v = p.src2[m * p.n2 + (j - p.n0 - p.n1)] ? 1.0f : 0.0f;
            f[j] = v;
        }
        const long long trow = p.tab_idx ? p.tab_idx[m] : (m - b * p.L);
        const float* tab = p.tab + trow * p.d;
        const float* ra = p.row_a + b * p.row_a_stride;
        const float* rb = p.row_b + b * p.d;
        for (int c = lane; c < p.d; c += 32) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < F) acc = fmaf(f[j], __ldg(p.Wf + j * p.d + c), acc);
            p.h[m * p.d + c] = acc + tab[c] + ra[c] + rb[c];
        }
    }
}

// d % 128 == 0, 16-byte aligned rows: a warp handles the L tokens of one trajectory for one 128-column slice (a lane owns
// one float4 column group), keeps the trajectory's row_a + row_b and its slice of Wf in registers (F <= kF) and streams
// the tokens four at a time (all loads of a group of tokens in flight): per token one table-row read (L1/L2 resident) and
// one coalesced float4 store of h.  Same arithmetic order as embed_kernel: fma chain over the features, + tab + row_a + row_b.
template <int kF, int kTok>
__global__ void __launch_bounds__(256, 3) embed_traj_kernel(const EmbedParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int F = p.n0 + p.n1 + p.n2;
    const long long B = p.M / p.L;
    const int d4 = p.d >> 2, ns = d4 >> 5;                              // column slices of 32 float4 per trajectory
    for (long long u = warp; u < B * ns; u += nwarps) {
        const long long b = u / ns;
        const int c4 = static_cast<int>(u - b * ns) * 32 + lane;
        float4 w[kF];
#pragma unroll
        for (int j = 0; j < kF; ++j) w[j] = (j < F) ? __ldg(reinterpret_cast<const float4*>(p.Wf + j * p.d) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ra = __ldg(reinterpret_cast<const float4*>(p.row_a + b * p.row_a_stride) + c4);
        const float4 rb = __ldg(reinterpret_cast<const float4*>(p.row_b + b * p.d) + c4);
        for (int t0 = 0; t0 < p.L; t0 += kTok) {
            float f[kTok][kF];
            float4 tb[kTok];
#pragma unroll
            for (int i = 0; i < kTok; ++i) {
                const int t = t0 + i;
                if (t < p.L) {
                    const long long m = b * p.L + t;
#pragma unroll
                    for (int j = 0; j < kF; ++j) {
                        float v = 0.0f;
                        if (j < p.n0) v = p.src0[m * p.n0 + j];
                        else if (j < p.n0 + p.n1) v = p.src1[m * p.n1 + (j - p.n0)];
                        else if (j < F) v = p.src2[m * p.n2 + (j - p.n0 - p.n1)] ? 1.0f : 0.0f;
                        f[i][j] = v;
                    }
                    const long long trow = p.tab_idx ? p.tab_idx[m] : t;
                    tb[i] = __ldg(reinterpret_cast<const float4*>(p.tab + trow * p.d) + c4);
                }
            }
#pragma unroll
            for (int i = 0; i < kTok; ++i) {
                const int t = t0 + i;
                if (t < p.L) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < kF; ++j) {
                        acc.x = fmaf(f[i][j], w[j].x, acc.x);
                        acc.y = fmaf(f[i][j], w[j].y, acc.y);
                        acc.z = fmaf(f[i][j], w[j].z, acc.z);
                        acc.w = fmaf(f[i][j], w[j].w, acc.w);
                    }
                    reinterpret_cast<float4*>(p.h + (b * p.L + t) * p.d)[c4] =
                        make_float4(acc.x + tb[i].x + ra.x + rb.x, acc.y + tb[i].y + ra.y + rb.y, acc.z + tb[i].z + ra.z + rb.z, acc.w + tb[i].w + ra.w + rb.w);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (eps = 1e-5, affine) + FiLM: a = LN(h) * (1 + gamma[b]) + beta[b]   (transformer.py:28-41)
//   one warp per token, row held in registers (d <= 512), two-pass variance like ATen.
//   gb: [B, gb_stride] fp32 with gamma at [0, d) and beta at [d, 2d) of the layer's slice; nullptr = no FiLM.
// ------------------------------------------------------------------------------------------------
template <typename TO, int VPL>
__global__ void __launch_bounds__(256) ln_film_kernel(const float* __restrict__ h, const float* __restrict__ lnw,
                                                      const float* __restrict__ lnb, const float* __restrict__ gb,
                                                      long long gb_stride, TO* __restrict__ out, long long M, int L, int d,
                                                      float* __restrict__ hcopy = nullptr) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long m = warp; m < M; m += nwarps) {
        const float4* row = reinterpret_cast<const float4*>(h + m * d);
        float4 v[VPL];
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 * 4 < d) {
                v[i] = row[c4];
                if (hcopy) reinterpret_cast<float4*>(hcopy + m * d)[c4] = v[i];
                sum += v[i].x + v[i].y + v[i].z + v[i].w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / static_cast<float>(d);
        float sq = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 * 4 < d) {
                const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                sq += a * a + b2 * b2 + c * c + e * e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / static_cast<float>(d) + 1e-5f);
        const float* g = gb ? gb + (m / L) * gb_stride : nullptr;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 * 4 < d) {
                const float4 w4 = reinterpret_cast<const float4*>(lnw)[c4];
                const float4 b4 = reinterpret_cast<const float4*>(lnb)[c4];
                float o0 = (v[i].x - mean) * rstd * w4.x + b4.x;
                float o1 = (v[i].y - mean) * rstd * w4.y + b4.y;
                float o2 = (v[i].z - mean) * rstd * w4.z + b4.z;
                float o3 = (v[i].w - mean) * rstd * w4.w + b4.w;
                if (g) {
                    const float4 ga = reinterpret_cast<const float4*>(g)[c4];
                    const float4 be = reinterpret_cast<const float4*>(g + d)[c4];
                    o0 = o0 * (1.0f + ga.x) + be.x;
                    o1 = o1 * (1.0f + ga.y) + be.y;
                    o2 = o2 * (1.0f + ga.z) + be.z;
                    o3 = o3 * (1.0f + ga.w) + be.w;
                }
                if constexpr (sizeof(TO) == 2) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
                    uint2 pk;
                    pk.x = *reinterpret_cast<unsigned*>(&p0);
                    pk.y = *reinterpret_cast<unsigned*>(&p1);
                    reinterpret_cast<uint2*>(out + m * d)[c4] = pk;
                } else {
                    reinterpret_cast<float4*>(out + m * d)[c4] = make_float4(o0, o1, o2, o3);
                }
            }
        }
    }
}



// TMA-staged form (the default): the residual rows stream through a 3-stage shared-memory ring of 16-row chunks filled by
// cp.async.bulk (one elected thread issues, mbarrier completion), so ~70 KB per CTA are in flight regardless of what the warps
// are doing; the register-resident kernel above keeps one 1.5 KB row per warp in flight only while that warp is in its load
// phase and measured 3.3-3.5 TB/s (two rows per warp: 3.1 TB/s, more registers, fewer warps -- rejected).  Same per-lane
// element assignment and reduction order as ln_film_kernel: bit-identical results.
constexpr int kLnRows = 16;          // rows per chunk: one per warp
constexpr int kLnStages = 3;
constexpr int kLnFilmRows = 3;       // FiLM rows staged per chunk (L >= 8: a 16-row chunk touches at most 3 trajectories)
constexpr int kLnThreads = 512;

__device__ __forceinline__ void ln_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000000;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(parity)
            : "memory");
    }
}

// stage layout: [kLnRows x d] residual rows | [kLnFilmRows x 2d] FiLM rows of the chunk's trajectories
// kExact: d == VPL * 128 (256 / 384 / 512): no per-access column predicates.  The first revision executed 389 warp-instructions
// per row at d = 384 and was ISSUE bound (68 % issue-active, 3.3 TB/s): a 64-bit m / L per row, a predicate per access, separate
// subtract / scale / affine / FiLM steps.  Now the chunk's first trajectory and row phase advance incrementally (warp-uniform
// integers, one division per CTA), the row's trajectory slot is two compares (L >= 8: a 16-row chunk spans at most 3), and the
// per-element arithmetic is three FFMAs (normalise; affine; FiLM).
template <typename TO, int VPL, bool kExact>
__global__ void __launch_bounds__(kLnThreads, 2) ln_film_bulk_kernel(const float* __restrict__ h, const float* __restrict__ lnw,
                                                                  const float* __restrict__ lnb, const float* __restrict__ gb,
                                                                  long long gb_stride, TO* __restrict__ out, long long M, int L, int d,
                                                                  float* __restrict__ hcopy) {
    // hcopy != nullptr (training forward): the staged residual rows also go back out as a copy (the backward's saved h), one
    // cp.async.bulk shared -> global per chunk issued by the thread that fills the ring: no registers, no LSU instructions, and
    // the separate device-to-device copy (a second read of h) is gone.
    extern __shared__ __align__(128) unsigned char ln_smem[];
    const size_t stage_floats = static_cast<size_t>(kLnRows) * d + static_cast<size_t>(kLnFilmRows) * 2 * d;
    float* ring = reinterpret_cast<float*>(ln_smem);
    float* saff = ring + kLnStages * stage_floats;                // LayerNorm affine [w | b] (2d floats)
    uint64_t* full = reinterpret_cast<uint64_t*>(saff + 2 * d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long chunks = (M + kLnRows - 1) / kLnRows;
    for (int i = threadIdx.x; i < d; i += kLnThreads) { saff[i] = lnw[i]; saff[d + i] = lnb[i]; }
    const bool film_smem = gb != nullptr && L >= 8 && (gb_stride % 4 == 0);
    auto issue = [&](long long chunk, int stage) {
        const long long r0 = chunk * kLnRows;
        const long long rows = (M - r0 < kLnRows) ? (M - r0) : kLnRows;
        const uint32_t bytes = static_cast<uint32_t>(rows * d * 4);
        const long long t0 = r0 / L, t1 = (r0 + rows - 1) / L;
        const int nt = film_smem ? static_cast<int>(t1 - t0 + 1) : 0;
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[stage]));
        float* dst = ring + stage * stage_floats;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes + static_cast<uint32_t>(nt) * d * 8u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
                     "l"(reinterpret_cast<uint64_t>(h + r0 * d)), "r"(bytes), "r"(bar)
                     : "memory");
        for (int t = 0; t < nt; ++t)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             static_cast<uint32_t>(__cvta_generic_to_shared(dst + static_cast<size_t>(kLnRows) * d + static_cast<size_t>(t) * 2 * d))),
                         "l"(reinterpret_cast<uint64_t>(gb + (t0 + t) * gb_stride)), "r"(static_cast<uint32_t>(d) * 8u), "r"(bar)
                         : "memory");
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLnStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&full[s]))) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int s = 0; s < kLnStages; ++s) {
            const long long c = blockIdx.x + static_cast<long long>(s) * gridDim.x;
            if (c < chunks) issue(c, s);
        }
    // phase of the chunk's first row inside its trajectory, advanced incrementally (rows per step = kLnRows * gridDim.x)
    const long long step_rows = static_cast<long long>(kLnRows) * gridDim.x;
    const int step_rem = static_cast<int>(step_rows % L);
    int rem = static_cast<int>((static_cast<long long>(blockIdx.x) * kLnRows) % L);
    long long it = 0;
    for (long long chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x, ++it) {
        const int stage = static_cast<int>(it % kLnStages);
        ln_mbar_wait(&full[stage], static_cast<uint32_t>((it / kLnStages) & 1));
        const float* tile = ring + stage * stage_floats;
        if (hcopy != nullptr && threadIdx.x == 0) {
            const long long r0 = chunk * kLnRows;
            const long long rows = (M - r0 < kLnRows) ? (M - r0) : kLnRows;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(hcopy + r0 * d)),
                         "r"(static_cast<uint32_t>(__cvta_generic_to_shared(tile))), "r"(static_cast<uint32_t>(rows * d * 4))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        const float* film = tile + static_cast<size_t>(kLnRows) * d;
        const int r = warp;
        const long long m = chunk * kLnRows + r;
        if (m < M) {
            const float4* row = reinterpret_cast<const float4*>(tile + static_cast<size_t>(r) * d);
            float4 v[VPL];
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c4 = lane + 32 * i;
                if (kExact || c4 * 4 < d) {
                    v[i] = row[c4];
                    sum += v[i].x + v[i].y + v[i].z + v[i].w;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / static_cast<float>(d);
            float sq = 0.0f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c4 = lane + 32 * i;
                if (kExact || c4 * 4 < d) {
                    const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                    sq += a * a + b2 * b2 + c * c + e * e;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq / static_cast<float>(d) + 1e-5f);
            const float* g = nullptr;
            if (gb) {
                if (film_smem) {
                    const int x = rem + r;                       // < L + 16 <= 3 L
                    g = film + static_cast<size_t>((x >= L) + (x >= 2 * L)) * 2 * d;
                } else {
                    g = gb + (m / L) * gb_stride;
                }
            }
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c4 = lane + 32 * i;
                if (kExact || c4 * 4 < d) {
                    const float4 w4 = reinterpret_cast<const float4*>(saff)[c4];
                    const float4 b4 = reinterpret_cast<const float4*>(saff + d)[c4];
                    float o0 = (v[i].x - mean) * rstd * w4.x + b4.x;
                    float o1 = (v[i].y - mean) * rstd * w4.y + b4.y;
                    float o2 = (v[i].z - mean) * rstd * w4.z + b4.z;
                    float o3 = (v[i].w - mean) * rstd * w4.w + b4.w;
                    if (g) {
                        const float4 ga = reinterpret_cast<const float4*>(g)[c4];
                        const float4 be = reinterpret_cast<const float4*>(g + d)[c4];
                        o0 = o0 * (1.0f + ga.x) + be.x;
                        o1 = o1 * (1.0f + ga.y) + be.y;
                        o2 = o2 * (1.0f + ga.z) + be.z;
                        o3 = o3 * (1.0f + ga.w) + be.w;
                    }
                    if constexpr (sizeof(TO) == 2) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
                        uint2 pk;
                        pk.x = *reinterpret_cast<unsigned*>(&p0);
                        pk.y = *reinterpret_cast<unsigned*>(&p1);
                        reinterpret_cast<uint2*>(out + m * d)[c4] = pk;
                    } else {
                        reinterpret_cast<float4*>(out + m * d)[c4] = make_float4(o0, o1, o2, o3);
                    }
                }
            }
        }
        __syncthreads();                                         // every warp is done reading this stage
        const long long nxt = chunk + static_cast<long long>(kLnStages) * gridDim.x;
        if (threadIdx.x == 0) {
            if (hcopy != nullptr) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the copy-out has read this stage
            if (nxt < chunks) issue(nxt, stage);
        }
        rem += step_rem;
        if (rem >= L) rem -= L;
    }
    if (hcopy != nullptr && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// output head: y[m, j] = h[m, :] . W[j, :] + b[j], j < D (D <= 4): one warp per token
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) out_head_kernel(const float* __restrict__ h, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ y, long long M, int d,
                                                       int D) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long m = warp; m < M; m += nwarps) {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int c4 = lane; c4 * 4 < d; c4 += 32) {
            const float4 hv = reinterpret_cast<const float4*>(h + m * d)[c4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < D) {
                    const float4 wv = reinterpret_cast<const float4*>(W + static_cast<long long>(j) * d)[c4];
                    acc[j] = fmaf(hv.x, wv.x, fmaf(hv.y, wv.y, fmaf(hv.z, wv.z, fmaf(hv.w, wv.w, acc[j]))));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (lane < D) y[m * D + lane] = acc[lane] + bias[lane];
    }
}

static inline int warp_grid(long long rows) { return grid_for(rows, 8, 8); }

}  // namespace idb200

using namespace idb200;

extern "C" int idb200_sgemm(const void* A, int a_is_bf16, int64_t lda, const float* W, const float* bias, float* out,
                            int64_t ldo, int64_t M, int N, int K, int act, int accumulate, idb200_stream_t stream) {
    IDB_REQUIRE(M >= 0 && N > 0 && K > 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(M == 0 || (A && W && out), IDB200_EINVAL, "NULL pointer");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE((M + 63) / 64 <= 65535 * 32LL, IDB200_EUNSUPPORTED, "M too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (((N + 63) / 64) * ((M + 63) / 64) < num_sms()) {   // fewer 64 x 64 tiles than SMs: the small-tile form (same results)
        dim3 grid((N + 31) / 32, static_cast<unsigned>((M + 31) / 32));
        if (a_is_bf16)
            sgemm_tn_small_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(A), lda, W, bias, out, ldo, M, N, K, act, accumulate);
        else
            sgemm_tn_small_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(A), lda, W, bias, out, ldo, M, N, K, act, accumulate);
        return check_launch("sgemm_tn_small_kernel");
    }
    // grid.y is limited to 65535: loop over row chunks
    const long long chunk = 65535LL * 64;
    for (long long m0 = 0; m0 < M; m0 += chunk) {
        const long long mm = (M - m0 < chunk) ? M - m0 : chunk;
        dim3 grid((N + 63) / 64, static_cast<unsigned>((mm + 63) / 64));
        if (a_is_bf16)
            sgemm_tn_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(A) + m0 * lda, lda, W, bias,
                                                                 out + m0 * ldo, ldo, mm, N, K, act, accumulate);
        else
            sgemm_tn_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(A) + m0 * lda, lda, W, bias, out + m0 * ldo,
                                                         ldo, mm, N, K, act, accumulate);
    }
    return check_launch("sgemm_tn_kernel");
}

extern "C" int idb200_conv_encoder(const float* occ, const float* sdf, int64_t B, int H, int W, int n_layers,
                                   const int* channels /*host [n_layers+1]*/, const float* const* weights /*host*/,
                                   const float* const* biases /*host*/, float* scratch, float* pooled, idb200_stream_t stream) {
    IDB_REQUIRE(n_layers >= 1 && n_layers <= 8, IDB200_EUNSUPPORTED, "1..8 conv layers supported");
    IDB_REQUIRE(B >= 0 && H >= 1 && W >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(channels[0] == 1 || (channels[0] == 2 && sdf), IDB200_EINVAL, "use_sdf is True but sdf missing from cond");
    if (B == 0) return IDB200_OK;
    IDB_REQUIRE(occ && pooled, IDB200_EINVAL, "NULL pointer");
    const int plane = (H + 2) * (W + 2);
    const int strips = H * ((W + 3) / 4);
    IDB_REQUIRE(kConvCo * strips <= kConvItems * 256, IDB200_EUNSUPPORTED, "maze %dx%d too large for the conv kernel", H, W);
    const size_t smem = (static_cast<size_t>(kConvCi) * plane + static_cast<size_t>(kConvCo) * strips) * sizeof(float);
    IDB_REQUIRE(smem <= 227 * 1024, IDB200_EUNSUPPORTED, "conv planes need %zu bytes of shared memory (> 227 KB)", smem);
    int cmid = 0;
    for (int l = 1; l < n_layers; ++l) cmid = cmid > channels[l] ? cmid : channels[l];
    IDB_REQUIRE(n_layers == 1 || scratch, IDB200_EINVAL, "scratch (2 * B * max mid channels * H * W floats) is NULL");
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_silu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        smem_set = smem;
    }
    const int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
    const long long HW = static_cast<long long>(H) * W;
    float* bufs[2] = {scratch, scratch ? scratch + B * cmid * HW : nullptr};
    for (int l = 0; l < n_layers; ++l) {
        ConvLayerParams p{};
        const bool last = (l == n_layers - 1);
        if (l == 0) { p.in0 = occ; p.in1 = sdf; p.ci0 = 1; }
        else { p.in0 = bufs[(l - 1) & 1]; p.in1 = nullptr; p.ci0 = channels[l]; }
        p.w = weights[l]; p.bias = biases[l];
        p.out = last ? nullptr : bufs[l & 1];
        p.pooled = last ? pooled : nullptr;
        p.ci = channels[l]; p.co = channels[l + 1]; p.H = H; p.W = W; p.B = B;
        const long long work = B * ((p.co + kConvCo - 1) / kConvCo);
        const int grid = grid_for(work, 1, per_sm > 0 ? per_sm : 1);
        conv3x3_silu_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
        int rc = check_launch("conv3x3_silu_kernel");
        if (rc) return rc;
    }
    return IDB200_OK;
}

extern "C" int idb200_sinusoid(const float* args, int rows, int dim, int mode, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(rows >= 0 && dim >= 1 && out, IDB200_EINVAL, "bad arguments");
    IDB_REQUIRE(mode == 0 || args, IDB200_EINVAL, "args missing");
    if (rows == 0) return IDB200_OK;
    sinusoid_kernel<<<grid_for(static_cast<long long>(rows) * dim, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(args, rows, dim, mode, out);
    return check_launch("sinusoid_kernel");
}

extern "C" int idb200_embed_tokens(const float* src0, int n0, const float* src1, int n1, const uint8_t* src2, int n2,
                                   const float* Wf, const float* tab, const int64_t* tab_idx, const float* row_a,
                                   int64_t row_a_stride, const float* row_b, float* h, int64_t M, int L, int d,
                                   idb200_stream_t stream) {
    IDB_REQUIRE(M >= 0 && L >= 1 && d >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(n0 + n1 + n2 <= 16 && n0 >= 0 && n1 >= 0 && n2 >= 0, IDB200_EUNSUPPORTED, "at most 16 token features");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(src0 && Wf && tab && row_a && row_b && h, IDB200_EINVAL, "NULL pointer");
    EmbedParams p{src0, n0, src1, n1, src2, n2, Wf, tab, reinterpret_cast<const long long*>(tab_idx), row_a, row_a_stride, row_b, h, M, L, d};
    const int F = n0 + n1 + n2;
    const bool vec = d % 128 == 0 && F <= 8 && M % L == 0 && aligned(Wf, 16) && aligned(tab, 16) && aligned(row_a, 16) &&
                     aligned(row_b, 16) && aligned(h, 16) && row_a_stride % 4 == 0;
    if (vec) {
        const int grid = grid_for((M / L) * (d / 128), 8, 3);
        if (F <= 4) embed_traj_kernel<4, 2><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
        else embed_traj_kernel<8, 2><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
        return check_launch("embed_traj_kernel");
    }
    embed_kernel<<<warp_grid(M), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("embed_kernel");
}

static int ln_film_impl(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride, void* out,
                        int out_is_bf16, int64_t M, int L, int d, float* h_copy, idb200_stream_t stream) {
    IDB_REQUIRE(M >= 0 && L >= 1 && d >= 4 && d % 4 == 0 && d <= 512, IDB200_EUNSUPPORTED, "d must be a multiple of 4, <= 512");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(h && ln_w && ln_b && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(!h_copy || (aligned(h_copy, 16) && h_copy != h), IDB200_EALIGN, "h_copy must be 16-byte aligned and distinct from h");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool bulk_env = !(getenv("IDB200_LN_BULK") && atoi(getenv("IDB200_LN_BULK")) == 0);
    if (bulk_env && d % 4 == 0 && aligned(h, 16) && (!gamma_beta || aligned(gamma_beta, 16)) && M >= 4096) {
        const size_t smem = (static_cast<size_t>(kLnStages) * (static_cast<size_t>(kLnRows) * d + static_cast<size_t>(kLnFilmRows) * 2 * d) + 2 * static_cast<size_t>(d)) * 4 + 64;
        const long long chunks = (M + kLnRows - 1) / kLnRows;
        int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
        if (per_sm > 2) per_sm = 2;
        const int grid_b = grid_for(chunks, 1, per_sm > 0 ? per_sm : 1);
        auto run = [&](auto kern) -> int {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(ln_film_bulk, smem=%zu): %s", smem, cudaGetErrorString(e));
            return IDB200_OK;
        };
#define IDB_LN_LAUNCH(TO, VPL)                                                                                                     \
    do {                                                                                                                           \
        if (d == VPL * 128) {                                                                                                      \
            int rc_ = run(ln_film_bulk_kernel<TO, VPL, true>);                                                                     \
            if (rc_) return rc_;                                                                                                   \
            ln_film_bulk_kernel<TO, VPL, true><<<grid_b, kLnThreads, smem, st>>>(h, ln_w, ln_b, gamma_beta, gb_stride, static_cast<TO*>(out), M, L, d, h_copy); \
        } else {                                                                                                                   \
            int rc_ = run(ln_film_bulk_kernel<TO, VPL, false>);                                                                    \
            if (rc_) return rc_;                                                                                                   \
            ln_film_bulk_kernel<TO, VPL, false><<<grid_b, kLnThreads, smem, st>>>(h, ln_w, ln_b, gamma_beta, gb_stride, static_cast<TO*>(out), M, L, d, h_copy); \
        }                                                                                                                          \
    } while (0)
        if (out_is_bf16) {
            if (d <= 256) IDB_LN_LAUNCH(__nv_bfloat16, 2); else if (d <= 384) IDB_LN_LAUNCH(__nv_bfloat16, 3); else IDB_LN_LAUNCH(__nv_bfloat16, 4);
        } else {
            if (d <= 256) IDB_LN_LAUNCH(float, 2); else if (d <= 384) IDB_LN_LAUNCH(float, 3); else IDB_LN_LAUNCH(float, 4);
        }
#undef IDB_LN_LAUNCH
        return check_launch("ln_film_bulk_kernel");
    }
    const int grid = warp_grid(M);
    if (out_is_bf16) ln_film_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(h, ln_w, ln_b, gamma_beta, gb_stride, static_cast<__nv_bfloat16*>(out), M, L, d, h_copy);
    else ln_film_kernel<float, 4><<<grid, 256, 0, st>>>(h, ln_w, ln_b, gamma_beta, gb_stride, static_cast<float*>(out), M, L, d, h_copy);
    return check_launch("ln_film_kernel");
}

extern "C" int idb200_ln_film(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta,
                              int64_t gb_stride, void* out, int out_is_bf16, int64_t M, int L, int d, idb200_stream_t stream) {
    return ln_film_impl(h, ln_w, ln_b, gamma_beta, gb_stride, out, out_is_bf16, M, L, d, nullptr, stream);
}

extern "C" int idb200_ln_film_save(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                                   void* out, int out_is_bf16, float* h_copy, int64_t M, int L, int d, idb200_stream_t stream) {
    IDB_REQUIRE(h_copy != nullptr, IDB200_EINVAL, "h_copy is NULL (use idb200_ln_film)");
    return ln_film_impl(h, ln_w, ln_b, gamma_beta, gb_stride, out, out_is_bf16, M, L, d, h_copy, stream);
}

extern "C" int idb200_out_head(const float* h, const float* W, const float* bias, float* y, int64_t M, int d, int D,
                               idb200_stream_t stream) {
    IDB_REQUIRE(M >= 0 && d % 4 == 0 && D >= 1 && D <= 4, IDB200_EUNSUPPORTED, "D <= 4 and d % 4 == 0 supported");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(h && W && bias && y, IDB200_EINVAL, "NULL pointer");
    out_head_kernel<<<warp_grid(M), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, W, bias, y, M, d, D);
    return check_launch("out_head_kernel");
}
