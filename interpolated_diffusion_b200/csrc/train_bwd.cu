// Backward kernels of the training steps (src/train/train_interp_levels.py:1142-1173, src/train/train_keypoints.py:505-556):
// everything between the loss gradient (train_tail.cu) and the optimiser that is not a dense contraction.  The dense
// contractions (dX = dY W, dW = dY^T X) run on the tcgen05 GEMM of gemm.cu: dX with the transposed weight as the
// "W" operand, dW as a split-K GEMM over the token dimension on transposed bf16 copies made by transpose_to_bf16.
// Reductions are two-level with a fixed order: results are deterministic run to run.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace idb200 {
namespace tb {

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(unsigned char v) { return static_cast<float>(v); }

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_fwd(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float silu_grad(float x) {
    const float s = sigmoid_f(x);
    return s * (1.0f + x * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------------------------------
// dst[n, m] = bf16(src[m, n]): 64 x 64 tiles through shared memory, coalesced on both sides.
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const TIn* __restrict__ src, long long M, int N,
                                                                __nv_bfloat16* __restrict__ dst) {
    __shared__ float tile[64][65];
    const long long m0 = static_cast<long long>(blockIdx.x) * 64;
    const int n0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int r = ty; r < 64; r += 4) {
        const long long m = m0 + r;
        const int n = n0 + tx;
        tile[r][tx] = (m < M && n < N) ? to_f(src[m * N + n]) : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 64; r += 4) {
        const int n = n0 + r;
        const long long m = m0 + tx;
        if (n < N && m < M) dst[static_cast<long long>(n) * M + m] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// bf16 -> bf16 with 16-byte accesses on both sides (M % 8 == 0, N % 8 == 0): 64 x 64 tile, each thread moves two 8-element
// vectors in and two out.
__global__ void __launch_bounds__(256) transpose_bf16_vec_kernel(const __nv_bfloat16* __restrict__ src, long long M, int N,
                                                                 __nv_bfloat16* __restrict__ dst) {
    __shared__ __nv_bfloat16 tile[64][66];                 // pitch 33 words
    const long long m0 = static_cast<long long>(blockIdx.x) * 64;
    const int n0 = blockIdx.y * 64;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int v = threadIdx.x + it * 256;              // 512 vectors: row r = v / 8, 8-column group cg = v % 8
        const int r = v >> 3, cg = v & 7;
        const long long m = m0 + r;
        const int n = n0 + cg * 8;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (m < M && n < N) val = *reinterpret_cast<const uint4*>(src + m * N + n);
        uint32_t* t32 = reinterpret_cast<uint32_t*>(&tile[r][cg * 8]);
        t32[0] = val.x; t32[1] = val.y; t32[2] = val.z; t32[3] = val.w;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int v = threadIdx.x + it * 256;              // output row n = v / 8 (source column), m group mg = v % 8
        const int c = v >> 3, mg = v & 7;
        const int n = n0 + c;
        const long long m = m0 + mg * 8;
        if (n < N && m < M) {
            __align__(16) __nv_bfloat16 e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = tile[mg * 8 + j][c];
            *reinterpret_cast<uint4*>(dst + static_cast<long long>(n) * M + m) = *reinterpret_cast<const uint4*>(e);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Batched small operations of a training step, one launch each instead of one per parameter (at B = 512 per GPU a step had ~130
// device-to-device copies of gradient slices and ~150 weight casts / transposes of 2-3 us each).  The pointer tables travel as
// kernel parameters (by value), so the launches are CUDA-graph capturable and need no host-to-device copy.
constexpr int kMaxCopy = 96;
struct CopyBatch {
    const float* src[kMaxCopy];
    float* dst[kMaxCopy];
    int n[kMaxCopy];
};
__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ CopyBatch b) {
    const int seg = blockIdx.y;
    const float* __restrict__ src = b.src[seg];
    float* __restrict__ dst = b.dst[seg];
    const int n = b.n[seg];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// dst[m] = bf16(src[m]) [rows, cols] and dst_t[m] = bf16(src[m]^T) [cols, rows] for a list of fp32 matrices (the per-step bf16 operand
// copies of the token GEMMs' weights: W for the forward, W^T as the weight operand of dX = dY W); 64 x 64 tiles through shared memory.
constexpr int kMaxCast = 64;
struct CastBatch {
    const float* src[kMaxCast];
    __nv_bfloat16* dst[kMaxCast];
    __nv_bfloat16* dst_t[kMaxCast];
    int rows[kMaxCast], cols[kMaxCast];
};
__global__ void __launch_bounds__(256) cast_weights_kernel(const __grid_constant__ CastBatch b) {
    __shared__ float tile[64][65];
    const int m = blockIdx.z;
    const int R = b.rows[m], C = b.cols[m];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    if (r0 >= R || c0 >= C) return;
    const float* __restrict__ src = b.src[m];
    __nv_bfloat16* __restrict__ dst = b.dst[m];
    __nv_bfloat16* __restrict__ dst_t = b.dst_t[m];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int r = ty; r < 64; r += 4) {
        const int rr = r0 + r, cc = c0 + tx;
        float v = 0.0f;
        if (rr < R && cc < C) {
            v = src[static_cast<long long>(rr) * C + cc];
            if (dst) dst[static_cast<long long>(rr) * C + cc] = __float2bfloat16_rn(v);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    if (dst_t)
        for (int c = ty; c < 64; c += 4) {
            const int cc = c0 + c, rr = r0 + tx;
            if (cc < C && rr < R) dst_t[static_cast<long long>(cc) * R + rr] = __float2bfloat16_rn(tile[tx][c]);
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// Column sums of src[M, N] (bias gradients; LayerNorm affine gradients from per-trajectory partials).
// Stage 1: block (32 columns, one row slice) -> partial[slice, N]; stage 2: reduce_rows.
template <typename TIn>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const TIn* __restrict__ src, long long M, int N, long long rows_per_slice,
                                                             float* __restrict__ partial, long long part_ld) {
    // segments (blockIdx.z): src is [segs, M, N], partial rows are part_ld = segs * N wide with segment z at column z * N
    src += static_cast<long long>(blockIdx.z) * M * N;
    partial += static_cast<long long>(blockIdx.z) * N;
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slice;
    long long r1 = r0 + rows_per_slice;
    if (r1 > M) r1 = M;
    float acc = 0.0f;
    if (c < N)
        for (long long r = r0 + ty; r < r1; r += 8) acc += to_f(src[r * N + c]);
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < N) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w][tx];
        partial[static_cast<long long>(blockIdx.y) * part_ld + c] = t;
    }
}

// vector form: a thread owns kV consecutive columns (one 16-byte load per row); block = 32 column groups x 8 row lanes
template <typename TIn, int kV>
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const TIn* __restrict__ src, long long M, int N, long long rows_per_slice,
                                                                 float* __restrict__ partial, long long part_ld) {
    src += static_cast<long long>(blockIdx.z) * M * N;     // segments: see colsum_partial_kernel
    partial += static_cast<long long>(blockIdx.z) * N;
    __shared__ float sh[8][32 * kV + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + tx) * kV;
    const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slice;
    long long r1 = r0 + rows_per_slice;
    if (r1 > M) r1 = M;
    float acc[kV];
#pragma unroll
    for (int j = 0; j < kV; ++j) acc[j] = 0.0f;
    if (c0 < N) {
        auto add = [&](const uint4& v) {
            if constexpr (kV == 8) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            } else {
                acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y); acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
            }
        };
        // four independent 16-byte loads in flight per thread (the one-load-per-iteration form was latency bound: 2.1 TB/s);
        // the summation order per thread is unchanged (rows r, r + 8, r + 16, ...), so the result is bit-identical
        long long r = r0 + ty;
        for (; r + 24 < r1; r += 32) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(src + r * N + c0));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(src + (r + 8) * N + c0));
            const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(src + (r + 16) * N + c0));
            const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(src + (r + 24) * N + c0));
            add(v0); add(v1); add(v2); add(v3);
        }
        for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(src + r * N + c0)));
    }
#pragma unroll
    for (int j = 0; j < kV; ++j) sh[ty][tx * kV + j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * kV; c += 256) {
        const int col = blockIdx.x * 32 * kV + c;
        if (col < N) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += sh[w][c];
            partial[static_cast<long long>(blockIdx.y) * part_ld + col] = t;
        }
    }
}

// out[w] (+)= scale * sum_r partial[r, w]   (fixed order)
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, int R, long long W, float scale,
                                                          int accumulate, float* __restrict__ out) {
    const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= W) return;
    float t = 0.0f;
    int r = 0;
    for (; r + 4 <= R; r += 4) {                           // four loads in flight; same (sequential) summation order
        const float p0 = partial[static_cast<long long>(r) * W + w], p1 = partial[static_cast<long long>(r + 1) * W + w];
        const float p2 = partial[static_cast<long long>(r + 2) * W + w], p3 = partial[static_cast<long long>(r + 3) * W + w];
        t += p0; t += p1; t += p2; t += p3;
    }
    for (; r < R; ++r) t += partial[static_cast<long long>(r) * W + w];
    t *= scale;
    out[w] = accumulate ? out[w] + t : t;
}

// ---------------------------------------------------------------------------------------------------------------------
// SiLU forward / backward on bf16 pairs.  mode 0: y = silu(u);  mode 1: y = g * silu'(u).  One-MUFU forms of common.cuh (the GEMM
// epilogues of gemm.cu that fold these passes in use the same functions: bit-identical results either way).
__global__ void __launch_bounds__(256) silu_kernel(const __nv_bfloat162* __restrict__ u, const __nv_bfloat162* g, long long n2,
                                                   int mode, __nv_bfloat162* y) {   // y may alias g
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const float2 uv = __bfloat1622float2(u[i]);
        float2 o;
        if (mode == 0) {
            o = make_float2(silu16_fwd(uv.x), silu16_fwd(uv.y));
        } else {
            const float2 gv = __bfloat1622float2(g[i]);
            o = make_float2(__fmul_rn(gv.x, silu16_grad(uv.x)), __fmul_rn(gv.y, silu16_grad(uv.y)));
        }
        y[i] = __floats2bfloat162_rn(o.x, o.y);
    }
}

// 16-byte (8 x bf16) form of the same kernel for n % 8 == 0 and 16-byte aligned pointers
__global__ void __launch_bounds__(256) silu_vec_kernel(const uint4* __restrict__ u, const uint4* g, long long n8, int mode, uint4* y) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 uv = u[i];
        uint4 gv = mode ? g[i] : make_uint4(0u, 0u, 0u, 0u);
        __nv_bfloat162* u2 = reinterpret_cast<__nv_bfloat162*>(&uv);
        __nv_bfloat162* g2 = reinterpret_cast<__nv_bfloat162*>(&gv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(u2[j]);
            float2 o;
            if (mode == 0) {
                o = make_float2(silu16_fwd(a.x), silu16_fwd(a.y));
            } else {
                const float2 b = __bfloat1622float2(g2[j]);
                o = make_float2(__fmul_rn(b.x, silu16_grad(a.x)), __fmul_rn(b.y, silu16_grad(a.y)));
            }
            u2[j] = __floats2bfloat162_rn(o.x, o.y);
        }
        y[i] = uv;
    }
}

// fp32 variant for the per-trajectory MLPs (level_proj, sg.mlp, t_embed)
__global__ void __launch_bounds__(256) silu_f32_kernel(const float* __restrict__ u, const float* g, long long n, int mode, float* y) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = mode == 0 ? silu_fwd(u[i]) : g[i] * silu_grad(u[i]);
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of a = LN(h) * (1 + gamma) + beta  (transformer.py:35-46: norm + FiLM), one block per trajectory.
//   n = xhat * w + b;  dn = da * (1 + gamma);  dgamma = sum_t da * n;  dbeta = sum_t da;
//   dw_part = sum_t dn * xhat;  db_part = sum_t dn;  dxhat = dn * w;
//   dh += rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
template <int kPerLane>
__global__ void __launch_bounds__(256) ln_film_bwd_kernel(const float* __restrict__ da, const float* __restrict__ h,
                                                          const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                          const float* __restrict__ gb, long long gb_stride, int L,
                                                          float* __restrict__ dh, __nv_bfloat16* __restrict__ dh16,
                                                          float* __restrict__ dgb, long long dgb_stride, float* __restrict__ dwb_part) {
    constexpr int d = kPerLane * 32;
    __shared__ float sh[8][d];
    const long long b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float w[kPerLane], bb[kPerLane], g1[kPerLane];
    float a_dg[kPerLane], a_db[kPerLane], a_dw[kPerLane], a_dbb[kPerLane];
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) {
        const int c = lane + 32 * j;
        w[j] = ln_w[c];
        bb[j] = ln_b[c];
        g1[j] = gb ? 1.0f + gb[b * gb_stride + c] : 1.0f;
        a_dg[j] = a_db[j] = a_dw[j] = a_dbb[j] = 0.0f;
    }
    for (int t = warp; t < L; t += 8) {
        const long long row = (b * L + t) * d;
        float x[kPerLane], g[kPerLane];
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            x[j] = h[row + lane + 32 * j];
            g[j] = da[row + lane + 32 * j];
            s += x[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / d);
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            x[j] -= mean;
            v += x[j] * x[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const float rstd = rsqrtf(v * (1.0f / d) + 1e-5f);
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            x[j] *= rstd;                                   // xhat
            const float n = fmaf(x[j], w[j], bb[j]);
            const float dn = g[j] * g1[j];
            a_dg[j] = fmaf(g[j], n, a_dg[j]);
            a_db[j] += g[j];
            a_dw[j] = fmaf(dn, x[j], a_dw[j]);
            a_dbb[j] += dn;
            g[j] = dn * w[j];                               // dxhat
            s1 += g[j];
            s2 = fmaf(g[j], x[j], s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 *= (1.0f / d);
        s2 *= (1.0f / d);
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            const long long o = row + lane + 32 * j;
            const float v2 = dh[o] + rstd * (g[j] - s1 - x[j] * s2);
            dh[o] = v2;
            if (dh16) dh16[o] = __float2bfloat16_rn(v2);
        }
    }
    // cross-warp sums, one quantity at a time, fixed order
    float* outs[4] = {dgb ? dgb + b * dgb_stride : nullptr, dgb ? dgb + b * dgb_stride + d : nullptr, dwb_part + b * 2 * d,
                      dwb_part + b * 2 * d + d};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float* acc = q == 0 ? a_dg : q == 1 ? a_db : q == 2 ? a_dw : a_dbb;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) sh[warp][lane + 32 * j] = acc[j];
        __syncthreads();
        if (outs[q])
            for (int c = threadIdx.x; c < d; c += 256) {
                float t = 0.0f;
#pragma unroll
                for (int wv = 0; wv < 8; ++wv) t += sh[wv][c];
                outs[q][c] = t;
            }
    }
}

// Two-pass form of the same backward (the default): the one-block-per-trajectory kernel above keeps 4 x d/32 accumulators per
// lane (128 registers, 25 % occupancy) and ran at 1.7 TB/s.  Pass A: a warp per token reduces the four row scalars
// (mean, rstd, mean(dxhat), mean(dxhat * xhat)) -> stats[M, 4].  Pass B: a thread per COLUMN walks the L tokens of its
// trajectory: fully coalesced, no shuffles, ~32 registers; accumulates the column's dgamma / dbeta / dw / db and updates dh.
// 4 consecutive values of the gradient w.r.t. the LayerNorm output
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// Pass A: a warp per token; lane <-> float4 column groups {lane, lane + 32, ...} (16-byte loads; the first revision used scalar
// 4-byte loads and was request bound).  kV4 = d / 128 float4 groups per lane.
template <int kV4, typename TDa>
__global__ void __launch_bounds__(256) ln_bwd_stats_kernel(const TDa* __restrict__ da, const float* __restrict__ h,
                                                           const float* __restrict__ ln_w, const float* __restrict__ gb, long long gb_stride,
                                                           int L, long long M, float4* __restrict__ stats) {
    constexpr int d = kV4 * 128;
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    float4 w[kV4];
#pragma unroll
    for (int j = 0; j < kV4; ++j) w[j] = load4(ln_w + 4 * (lane + 32 * j));
    for (long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        const long long b = m / L;
        const long long row = m * d;
        float4 x[kV4], g[kV4];
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < kV4; ++j) {
            const int c = 4 * (lane + 32 * j);
            x[j] = load4(h + row + c);
            g[j] = load4(da + row + c);
            if (gb) {
                const float4 ga = load4(gb + b * gb_stride + c);
                g[j].x *= 1.0f + ga.x; g[j].y *= 1.0f + ga.y; g[j].z *= 1.0f + ga.z; g[j].w *= 1.0f + ga.w;
            }
            g[j].x *= w[j].x; g[j].y *= w[j].y; g[j].z *= w[j].z; g[j].w *= w[j].w;      // dxhat
            s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / d);
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < kV4; ++j) {
            x[j].x -= mean; x[j].y -= mean; x[j].z -= mean; x[j].w -= mean;
            v = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, fmaf(x[j].w, x[j].w, v))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const float rstd = rsqrtf(v * (1.0f / d) + 1e-5f);
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < kV4; ++j) {
            s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
            s2 = fmaf(g[j].x, x[j].x * rstd, fmaf(g[j].y, x[j].y * rstd, fmaf(g[j].z, x[j].z * rstd, fmaf(g[j].w, x[j].w * rstd, s2))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) stats[m] = make_float4(mean, rstd, s1 * (1.0f / d), s2 * (1.0f / d));
    }
}

// Pass B: a thread per group of FOUR columns walks the L tokens of its trajectory (16-byte loads / stores, coalesced across the
// block = one trajectory's d / 4 column groups), accumulates the columns' dgamma / dbeta / dw / db and updates dh.
// TDa: fp32 or bf16 gradient w.r.t. the LayerNorm output (the training step keeps it in bf16: the dX GEMM writes half the bytes and
// both passes read half).  kDhSum: also emit dwb_part[b, 2d + c] = sum_t of the UPDATED dh -- the bias gradient of the GEMM that
// accumulated into this point of the residual stream (out_proj / previous layer's ff.2), which otherwise costs a full read of dh.
template <typename TDa, bool kDhSum>
__global__ void __launch_bounds__(128) ln_bwd_apply_kernel(const TDa* __restrict__ da, const float* __restrict__ h,
                                                           const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                           const float* __restrict__ gb, long long gb_stride, int L, int d,
                                                           const float4* __restrict__ stats, float* __restrict__ dh,
                                                           __nv_bfloat16* __restrict__ dh16, float* __restrict__ dgb, long long dgb_stride,
                                                           float* __restrict__ dwb_part) {
    const long long b = blockIdx.x;                        // trajectories on the wide grid dimension
    const int c = 4 * threadIdx.x;                         // blockDim.x = d / 4
    const float4 w = load4(ln_w + c), bb = load4(ln_b + c);
    float4 g1 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (gb) {
        const float4 ga = load4(gb + b * gb_stride + c);
        g1 = make_float4(1.0f + ga.x, 1.0f + ga.y, 1.0f + ga.z, 1.0f + ga.w);
    }
    float4 a_dg = make_float4(0.f, 0.f, 0.f, 0.f), a_db = a_dg, a_dw = a_dg, a_dbb = a_dg, a_dh = a_dg;
#pragma unroll 4
    for (int t = 0; t < L; ++t) {
        const long long m = b * L + t;
        const float4 st = stats[m];                         // broadcast load
        const long long o = m * d + c;
        const float4 hv = load4(h + o), g = load4(da + o), dv = load4(dh + o);
        float4 v2;
#define IDB_LN_BWD_LANE(k)                                                    \
        {                                                                      \
            const float xh = (hv.k - st.x) * st.y;                             \
            const float n = fmaf(xh, w.k, bb.k);                               \
            const float dn = g.k * g1.k;                                       \
            a_dg.k = fmaf(g.k, n, a_dg.k);                                     \
            a_db.k += g.k;                                                     \
            a_dw.k = fmaf(dn, xh, a_dw.k);                                     \
            a_dbb.k += dn;                                                     \
            v2.k = dv.k + st.y * (dn * w.k - st.z - xh * st.w);                \
            if (kDhSum) a_dh.k += v2.k;                                        \
        }
        IDB_LN_BWD_LANE(x) IDB_LN_BWD_LANE(y) IDB_LN_BWD_LANE(z) IDB_LN_BWD_LANE(w)
#undef IDB_LN_BWD_LANE
        *reinterpret_cast<float4*>(dh + o) = v2;
        if (dh16) {
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(v2.x, v2.y), p1 = __floats2bfloat162_rn(v2.z, v2.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned*>(&p0);
            pk.y = *reinterpret_cast<const unsigned*>(&p1);
            *reinterpret_cast<uint2*>(dh16 + o) = pk;
        }
    }
    if (dgb) {
        *reinterpret_cast<float4*>(dgb + b * dgb_stride + c) = a_dg;
        *reinterpret_cast<float4*>(dgb + b * dgb_stride + d + c) = a_db;
    }
    constexpr int kW = kDhSum ? 3 : 2;
    *reinterpret_cast<float4*>(dwb_part + b * kW * d + c) = a_dw;
    *reinterpret_cast<float4*>(dwb_part + b * kW * d + d + c) = a_dbb;
    if (kDhSum) *reinterpret_cast<float4*>(dwb_part + b * kW * d + 2 * d + c) = a_dh;
}

// Pass B for SMALL batches (B below ~8 blocks per SM: at the cfg-4 per-GPU shape B = 512 the form above runs 3.5 blocks of 3 warps
// per SM and is latency bound: 0.050 ms against 0.033 ms of HBM time).  kG row groups of d / 4 threads walk interleaved tokens
// (t = g, g + kG, ...) of the trajectory; the groups' column accumulators are merged through shared memory in a fixed order
// (group 0 + 1 + 2 + 3): deterministic, but a different summation order than the one-group form.
template <typename TDa, bool kDhSum, int kG>
__global__ void __launch_bounds__(512) ln_bwd_apply_groups_kernel(const TDa* __restrict__ da, const float* __restrict__ h,
                                                                  const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                  const float* __restrict__ gb, long long gb_stride, int L, int d,
                                                                  const float4* __restrict__ stats, float* __restrict__ dh,
                                                                  __nv_bfloat16* __restrict__ dh16, float* __restrict__ dgb, long long dgb_stride,
                                                                  float* __restrict__ dwb_part) {
    extern __shared__ __align__(16) float4 ln_apply_sm[];      // [kG - 1][5][d / 4]
    const long long b = blockIdx.x;
    const int kCg = d >> 2;
    const int cg = threadIdx.x % kCg, g = threadIdx.x / kCg;
    const int c = 4 * cg;
    const float4 w = load4(ln_w + c), bb = load4(ln_b + c);
    float4 g1 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (gb) {
        const float4 ga = load4(gb + b * gb_stride + c);
        g1 = make_float4(1.0f + ga.x, 1.0f + ga.y, 1.0f + ga.z, 1.0f + ga.w);
    }
    float4 a_dg = make_float4(0.f, 0.f, 0.f, 0.f), a_db = a_dg, a_dw = a_dg, a_dbb = a_dg, a_dh = a_dg;
#pragma unroll 2
    for (int t = g; t < L; t += kG) {
        const long long m = b * L + t;
        const float4 st = stats[m];
        const long long o = m * d + c;
        const float4 hv = load4(h + o), gg = load4(da + o), dv = load4(dh + o);
        float4 v2;
#define IDB_LN_BWD_LANE(k)                                                    \
        {                                                                      \
            const float xh = (hv.k - st.x) * st.y;                             \
            const float n = fmaf(xh, w.k, bb.k);                               \
            const float dn = gg.k * g1.k;                                      \
            a_dg.k = fmaf(gg.k, n, a_dg.k);                                    \
            a_db.k += gg.k;                                                    \
            a_dw.k = fmaf(dn, xh, a_dw.k);                                     \
            a_dbb.k += dn;                                                     \
            v2.k = dv.k + st.y * (dn * w.k - st.z - xh * st.w);                \
            if (kDhSum) a_dh.k += v2.k;                                        \
        }
        IDB_LN_BWD_LANE(x) IDB_LN_BWD_LANE(y) IDB_LN_BWD_LANE(z) IDB_LN_BWD_LANE(w)
#undef IDB_LN_BWD_LANE
        *reinterpret_cast<float4*>(dh + o) = v2;
        if (dh16) {
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(v2.x, v2.y), p1 = __floats2bfloat162_rn(v2.z, v2.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned*>(&p0);
            pk.y = *reinterpret_cast<const unsigned*>(&p1);
            *reinterpret_cast<uint2*>(dh16 + o) = pk;
        }
    }
    if (g > 0) {
        float4* dst = ln_apply_sm + static_cast<size_t>(g - 1) * 5 * kCg + cg;
        dst[0] = a_dg; dst[kCg] = a_db; dst[2 * kCg] = a_dw; dst[3 * kCg] = a_dbb; dst[4 * kCg] = a_dh;
    }
    __syncthreads();
    if (g == 0) {
        auto add4 = [](float4& a, const float4& o) { a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; };
#pragma unroll
        for (int q = 0; q < kG - 1; ++q) {
            const float4* src = ln_apply_sm + static_cast<size_t>(q) * 5 * kCg + cg;
            add4(a_dg, src[0]); add4(a_db, src[kCg]); add4(a_dw, src[2 * kCg]); add4(a_dbb, src[3 * kCg]); add4(a_dh, src[4 * kCg]);
        }
        if (dgb) {
            *reinterpret_cast<float4*>(dgb + b * dgb_stride + c) = a_dg;
            *reinterpret_cast<float4*>(dgb + b * dgb_stride + d + c) = a_db;
        }
        constexpr int kW = kDhSum ? 3 : 2;
        *reinterpret_cast<float4*>(dwb_part + b * kW * d + c) = a_dw;
        *reinterpret_cast<float4*>(dwb_part + b * kW * d + d + c) = a_dbb;
        if (kDhSum) *reinterpret_cast<float4*>(dwb_part + b * kW * d + 2 * d + c) = a_dh;
    }
}

// One-launch form of the two passes: a block per trajectory runs pass A (a warp per token -> the four row scalars, kept in shared
// memory) and then pass B (a thread per group of four columns; the two halves of the block walk the two halves of the trajectory's
// tokens and are merged in a fixed order).  Pass B's second read of da / h (147 KB per trajectory at d = 384, L = 64) comes out of
// L2 instead of HBM: DRAM traffic 22 -> 16 bytes per element.  Same per-element arithmetic as the two kernels above.
template <int kV4, typename TDa, bool kDhSum>
__global__ void __launch_bounds__(64 * kV4) ln_bwd_fused_kernel(const TDa* __restrict__ da, const float* __restrict__ h,
                                                                const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                const float* __restrict__ gb, long long gb_stride, int L,
                                                                float* __restrict__ dh, __nv_bfloat16* __restrict__ dh16,
                                                                float* __restrict__ dgb, long long dgb_stride, float* __restrict__ dwb_part) {
    constexpr int d = kV4 * 128, kThreads = 64 * kV4, kWarps = kThreads / 32, kCg = d / 4;
    extern __shared__ __align__(16) float4 ln_bwd_sm[];
    float4* sstats = ln_bwd_sm;                            // [L]
    float4* smerge = ln_bwd_sm + L;                        // [5][kCg]: the upper half's column accumulators
    const long long b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // ---- pass A
        float4 w[kV4];
#pragma unroll
        for (int j = 0; j < kV4; ++j) w[j] = load4(ln_w + 4 * (lane + 32 * j));
        for (int t = warp; t < L; t += kWarps) {
            const long long row = (b * L + t) * d;
            float4 x[kV4], g[kV4];
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < kV4; ++j) {
                const int c = 4 * (lane + 32 * j);
                x[j] = load4(h + row + c);
                g[j] = load4(da + row + c);
                if (gb) {
                    const float4 ga = load4(gb + b * gb_stride + c);
                    g[j].x *= 1.0f + ga.x; g[j].y *= 1.0f + ga.y; g[j].z *= 1.0f + ga.z; g[j].w *= 1.0f + ga.w;
                }
                g[j].x *= w[j].x; g[j].y *= w[j].y; g[j].z *= w[j].z; g[j].w *= w[j].w;      // dxhat
                s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.0f / d);
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < kV4; ++j) {
                x[j].x -= mean; x[j].y -= mean; x[j].z -= mean; x[j].w -= mean;
                v = fmaf(x[j].x, x[j].x, fmaf(x[j].y, x[j].y, fmaf(x[j].z, x[j].z, fmaf(x[j].w, x[j].w, v))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const float rstd = rsqrtf(v * (1.0f / d) + 1e-5f);
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
            for (int j = 0; j < kV4; ++j) {
                s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
                s2 = fmaf(g[j].x, x[j].x * rstd, fmaf(g[j].y, x[j].y * rstd, fmaf(g[j].z, x[j].z * rstd, fmaf(g[j].w, x[j].w * rstd, s2))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) sstats[t] = make_float4(mean, rstd, s1 * (1.0f / d), s2 * (1.0f / d));
        }
    }
    __syncthreads();
    // ---- pass B
    const int cg = threadIdx.x % kCg, rh = threadIdx.x / kCg;
    const int c = 4 * cg;
    const int half = (L + 1) / 2;
    const int t0 = rh * half, t1 = rh ? L : half;
    const float4 w = load4(ln_w + c), bb = load4(ln_b + c);
    float4 g1 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (gb) {
        const float4 ga = load4(gb + b * gb_stride + c);
        g1 = make_float4(1.0f + ga.x, 1.0f + ga.y, 1.0f + ga.z, 1.0f + ga.w);
    }
    float4 a_dg = make_float4(0.f, 0.f, 0.f, 0.f), a_db = a_dg, a_dw = a_dg, a_dbb = a_dg, a_dh = a_dg;
#pragma unroll 4
    for (int t = t0; t < t1; ++t) {
        const float4 st = sstats[t];
        const long long o = (b * L + t) * d + c;
        const float4 hv = load4(h + o), g = load4(da + o), dv = load4(dh + o);
        float4 v2;
#define IDB_LN_BWD_LANE(k)                                                    \
        {                                                                      \
            const float xh = (hv.k - st.x) * st.y;                             \
            const float n = fmaf(xh, w.k, bb.k);                               \
            const float dn = g.k * g1.k;                                       \
            a_dg.k = fmaf(g.k, n, a_dg.k);                                     \
            a_db.k += g.k;                                                     \
            a_dw.k = fmaf(dn, xh, a_dw.k);                                     \
            a_dbb.k += dn;                                                     \
            v2.k = dv.k + st.y * (dn * w.k - st.z - xh * st.w);                \
            if (kDhSum) a_dh.k += v2.k;                                        \
        }
        IDB_LN_BWD_LANE(x) IDB_LN_BWD_LANE(y) IDB_LN_BWD_LANE(z) IDB_LN_BWD_LANE(w)
#undef IDB_LN_BWD_LANE
        *reinterpret_cast<float4*>(dh + o) = v2;
        if (dh16) {
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(v2.x, v2.y), p1 = __floats2bfloat162_rn(v2.z, v2.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned*>(&p0);
            pk.y = *reinterpret_cast<const unsigned*>(&p1);
            *reinterpret_cast<uint2*>(dh16 + o) = pk;
        }
    }
    if (rh == 1) {
        smerge[0 * kCg + cg] = a_dg; smerge[1 * kCg + cg] = a_db; smerge[2 * kCg + cg] = a_dw; smerge[3 * kCg + cg] = a_dbb;
        smerge[4 * kCg + cg] = a_dh;
    }
    __syncthreads();
    if (rh == 0) {
        auto add4 = [](float4& a, const float4& o) { a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; };
        add4(a_dg, smerge[0 * kCg + cg]); add4(a_db, smerge[1 * kCg + cg]); add4(a_dw, smerge[2 * kCg + cg]); add4(a_dbb, smerge[3 * kCg + cg]);
        add4(a_dh, smerge[4 * kCg + cg]);
        if (dgb) {
            *reinterpret_cast<float4*>(dgb + b * dgb_stride + c) = a_dg;
            *reinterpret_cast<float4*>(dgb + b * dgb_stride + d + c) = a_db;
        }
        constexpr int kW = kDhSum ? 3 : 2;
        *reinterpret_cast<float4*>(dwb_part + b * kW * d + c) = a_dw;
        *reinterpret_cast<float4*>(dwb_part + b * kW * d + d + c) = a_dbb;
        if (kDhSum) *reinterpret_cast<float4*>(dwb_part + b * kW * d + 2 * d + c) = a_dh;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of softmax(q k^T / sqrt(32)) v for one (trajectory, head) per block (nn.MultiheadAttention inside
// transformer.py:39, head_dim 32, L <= 64): recompute P, then dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(P o dP)),
// dQ = dS K / sqrt(32), dK = dS^T Q / sqrt(32).  fp32 in shared memory; every product is an accumulation of outer
// products over a reduction index r with both operands stored r-major.
template <int kL, int kT>
struct AttnBwdCfg {
    static constexpr int kNb = kL / kT;                 // threads per tile edge
    static constexpr int kThreads = kNb * kNb;
    static constexpr int kC = 32 * kT / kL;             // columns per thread for the [kL, 32] outputs
    // shared floats: Qt Kt Vt dOt [32][kL], Q K dO [kL][32], P [kL][kL], dSt [kL][kL]
    static constexpr int kSmemFloats = 4 * 32 * kL + 3 * kL * 32 + 2 * kL * kL;
};

// C[a0 + i][b0 + j] += sum_r A[r][a0 + i] * B[r][b0 + j]
template <int kI, int kJ>
__device__ __forceinline__ void outer_acc(const float* __restrict__ A, int pa, const float* __restrict__ B, int pb, int R, int a0, int b0,
                                          float (&acc)[kI][kJ]) {
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
        float av[kI], bv[kJ];
        if constexpr (kI == 4) {
            const float4 t = *reinterpret_cast<const float4*>(A + r * pa + a0);
            av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else if constexpr (kI == 2) {
            const float2 t = *reinterpret_cast<const float2*>(A + r * pa + a0);
            av[0] = t.x; av[1] = t.y;
        } else {
#pragma unroll
            for (int i = 0; i < kI; ++i) av[i] = A[r * pa + a0 + i];
        }
        if constexpr (kJ == 4) {
            const float4 t = *reinterpret_cast<const float4*>(B + r * pb + b0);
            bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else if constexpr (kJ == 2) {
            const float2 t = *reinterpret_cast<const float2*>(B + r * pb + b0);
            bv[0] = t.x; bv[1] = t.y;
        } else {
#pragma unroll
            for (int j = 0; j < kJ; ++j) bv[j] = B[r * pb + b0 + j];
        }
#pragma unroll
        for (int i = 0; i < kI; ++i)
#pragma unroll
            for (int j = 0; j < kJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

template <int kL, int kT>
__global__ void __launch_bounds__(AttnBwdCfg<kL, kT>::kThreads) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                                     const __nv_bfloat16* __restrict__ dO,
                                                                                     __nv_bfloat16* __restrict__ dqkv, int L, int H, int causal) {
    using Cfg = AttnBwdCfg<kL, kT>;
    constexpr int kNb = Cfg::kNb, kThreads = Cfg::kThreads, kC = Cfg::kC;
    extern __shared__ __align__(16) float sm[];
    float* Qt = sm;                      // [32][kL]
    float* Kt = Qt + 32 * kL;
    float* Vt = Kt + 32 * kL;
    float* dOt = Vt + 32 * kL;
    float* Q = dOt + 32 * kL;            // [kL][32]
    float* K = Q + kL * 32;
    float* dOr = K + kL * 32;
    float* P = dOr + kL * 32;            // [kL][kL]  (P, then dS)
    float* dSt = P + kL * kL;            // [kL][kL]  dS transposed
    const int d = H * 32;
    const long long b = blockIdx.x / H;
    const int hh = blockIdx.x % H;
    const int t = threadIdx.x;
    const float scale = 0.17677669529663687f;   // 1 / sqrt(32)

    // load q, k, v, dO of this head: thread -> (row r fastest, bf16 pair kp)
    for (int e = t; e < kL * 16; e += kThreads) {
        const int r = e % kL, kp = e / kL;
        float2 q = make_float2(0.f, 0.f), k = q, v = q, g = q;
        if (r < L) {
            const long long row = b * L + r;
            const __nv_bfloat16* base = qkv + row * 3 * d + hh * 32 + 2 * kp;
            q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base));
            k = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + d));
            v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + 2 * d));
            g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dO + row * d + hh * 32 + 2 * kp));
        }
        Qt[(2 * kp) * kL + r] = q.x;  Qt[(2 * kp + 1) * kL + r] = q.y;
        Kt[(2 * kp) * kL + r] = k.x;  Kt[(2 * kp + 1) * kL + r] = k.y;
        Vt[(2 * kp) * kL + r] = v.x;  Vt[(2 * kp + 1) * kL + r] = v.y;
        dOt[(2 * kp) * kL + r] = g.x; dOt[(2 * kp + 1) * kL + r] = g.y;
        Q[r * 32 + 2 * kp] = q.x;     Q[r * 32 + 2 * kp + 1] = q.y;
        K[r * 32 + 2 * kp] = k.x;     K[r * 32 + 2 * kp + 1] = k.y;
        dOr[r * 32 + 2 * kp] = g.x;   dOr[r * 32 + 2 * kp + 1] = g.y;
    }
    __syncthreads();

    const int ti = t / kNb, tj = t % kNb;
    const int i0 = ti * kT, j0 = tj * kT;
    // S = scale * Q K^T
    {
        float acc[kT][kT];
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kT; ++j) acc[i][j] = 0.0f;
        outer_acc<kT, kT>(Qt, kL, Kt, kL, 32, i0, j0, acc);
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kT; ++j) P[(i0 + i) * kL + j0 + j] = acc[i][j] * scale;
    }
    __syncthreads();
    // row softmax (one warp per row; kL <= 64: two entries per lane)
    for (int i = t >> 5; i < kL; i += kThreads / 32) {
        const int lane = t & 31;
        float v0 = -INFINITY, v1 = -INFINITY;
        const int jmax = causal ? (i + 1 < L ? i + 1 : L) : L;
        if (lane < jmax) v0 = P[i * kL + lane];
        if (kL > 32 && lane + 32 < jmax) v1 = P[i * kL + lane + 32];
        float m = fmaxf(v0, v1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const bool live = i < L;
        const float e0 = (live && lane < jmax) ? __expf(v0 - m) : 0.0f;
        const float e1 = (live && kL > 32 && lane + 32 < jmax) ? __expf(v1 - m) : 0.0f;
        float s = e0 + e1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = live ? 1.0f / s : 0.0f;
        if (lane < kL) P[i * kL + lane] = e0 * inv;
        if (kL > 32) P[i * kL + lane + 32] = e1 * inv;
    }
    __syncthreads();

    const int ra = (t / (32 / kC)) * kT;          // rows of the [kL, 32] outputs owned by this thread
    const int cb = (t % (32 / kC)) * kC;          // columns
    // dV[j][k] = sum_i P[i][j] dO[i][k]
    {
        float acc[kT][kC];
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kC; ++j) acc[i][j] = 0.0f;
        outer_acc<kT, kC>(P, kL, dOr, 32, kL, ra, cb, acc);
#pragma unroll
        for (int i = 0; i < kT; ++i)
            if (ra + i < L)
#pragma unroll
                for (int j = 0; j < kC; ++j)
                    dqkv[(b * L + ra + i) * 3 * d + 2 * d + hh * 32 + cb + j] = __float2bfloat16_rn(acc[i][j]);
    }
    // dP = dO V^T ; dS = P o (dP - delta)
    {
        float acc[kT][kT];
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kT; ++j) acc[i][j] = 0.0f;
        outer_acc<kT, kT>(dOt, kL, Vt, kL, 32, i0, j0, acc);
        float pv[kT][kT];
#pragma unroll
        for (int i = 0; i < kT; ++i) {
            float part = 0.0f;
#pragma unroll
            for (int j = 0; j < kT; ++j) {
                pv[i][j] = P[(i0 + i) * kL + j0 + j];
                part = fmaf(pv[i][j], acc[i][j], part);
            }
            // the kNb threads of one row group are consecutive lanes
#pragma unroll
            for (int o = kNb / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
#pragma unroll
            for (int j = 0; j < kT; ++j) acc[i][j] = pv[i][j] * (acc[i][j] - part) * scale;      // scale folded into dS
        }
        __syncthreads();                          // every thread has finished reading P (dV product above)
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kT; ++j) {
                P[(i0 + i) * kL + j0 + j] = acc[i][j];
                dSt[(j0 + j) * kL + i0 + i] = acc[i][j];
            }
    }
    __syncthreads();
    // dQ[i][k] = sum_j dS[i][j] K[j][k]   (A = dSt: r = j, contiguous in i)
    {
        float acc[kT][kC];
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kC; ++j) acc[i][j] = 0.0f;
        outer_acc<kT, kC>(dSt, kL, K, 32, kL, ra, cb, acc);
#pragma unroll
        for (int i = 0; i < kT; ++i)
            if (ra + i < L)
#pragma unroll
                for (int j = 0; j < kC; ++j) dqkv[(b * L + ra + i) * 3 * d + hh * 32 + cb + j] = __float2bfloat16_rn(acc[i][j]);
    }
    // dK[j][k] = sum_i dS[i][j] Q[i][k]   (A = dS: r = i, contiguous in j)
    {
        float acc[kT][kC];
#pragma unroll
        for (int i = 0; i < kT; ++i)
#pragma unroll
            for (int j = 0; j < kC; ++j) acc[i][j] = 0.0f;
        outer_acc<kT, kC>(P, kL, Q, 32, kL, ra, cb, acc);
#pragma unroll
        for (int i = 0; i < kT; ++i)
            if (ra + i < L)
#pragma unroll
                for (int j = 0; j < kC; ++j) dqkv[(b * L + ra + i) * 3 * d + d + hh * 32 + cb + j] = __float2bfloat16_rn(acc[i][j]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core attention backward for 32 < L <= 64 (mma.sync m16n8k16 bf16 + ldmatrix; the products are 64 x 64 x 32 per head,
// far below a tcgen05 tile).  One block of 4 warps per (trajectory, head).
//   phase 1, warp w owns query rows 16w..16w+15: S = Q K^T, row softmax in registers, dP = dO V^T, delta = rowsum(P o dP),
//            dS = P o (dP - delta) / sqrt(32), dQ = dS K (register A fragments) -> global; P and dS (bf16) -> shared memory;
//   phase 2, warp w owns key rows 16w..16w+15: dV = P^T dO, dK = dS^T Q with the transposed A fragments read by
//            ldmatrix.trans.  No cross-warp reduction: deterministic.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm4(unsigned (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr_u32(p)));
}
__device__ __forceinline__ void ldsm4t(unsigned (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}

//   Output: the three [64 x 32] gradient tiles are staged in shared memory (dQ over the K tile, dV over the V tile -- both dead after
//   phase 1 --, dK in its own tile) and leave as 16-byte vectors, 64 contiguous bytes per row and part (the first revision stored
//   4-byte fragments straight from the accumulators: 16-byte pieces per row and instruction).  colsum != nullptr: the staged tiles
//   also yield colsum[b, part * d + head * 32 + c] = sum over the trajectory's tokens of the bf16-rounded gradient -- summed over
//   b this is the in_proj bias gradient, which otherwise costs a full read of dqkv.
__global__ void __launch_bounds__(128) attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dO,
                                                                __nv_bfloat16* __restrict__ dqkv, int L, int H, int causal,
                                                                float* __restrict__ colsum) {
    constexpr int PQ = 40;                  // bf16 pitch of the [64][32] operand tiles (80 B rows: conflict-free ldmatrix)
    constexpr int PP = 72;                  // bf16 pitch of the [64][64] P / dS tiles (144 B rows)
    __shared__ __align__(16) __nv_bfloat16 sQ[64 * PQ], sK[64 * PQ], sV[64 * PQ], sG[64 * PQ], sP[64 * PP], sS[64 * PP], sO[64 * PQ];
    const int d = H * 32;
    const long long b = blockIdx.x / H;
    const int hh = blockIdx.x % H;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    // load: 64 rows x 4 chunks of 8 bf16 per operand
    for (int e = t; e < 64 * 4; e += 128) {
        const int r = e >> 2, ch = e & 3;
        uint4 q = make_uint4(0u, 0u, 0u, 0u), k = q, v = q, g = q;
        if (r < L) {
            const long long row = b * L + r;
            const __nv_bfloat16* base = qkv + row * 3 * d + hh * 32 + ch * 8;
            q = *reinterpret_cast<const uint4*>(base);
            k = *reinterpret_cast<const uint4*>(base + d);
            v = *reinterpret_cast<const uint4*>(base + 2 * d);
            g = *reinterpret_cast<const uint4*>(dO + row * d + hh * 32 + ch * 8);
        }
        *reinterpret_cast<uint4*>(&sQ[r * PQ + ch * 8]) = q;
        *reinterpret_cast<uint4*>(&sK[r * PQ + ch * 8]) = k;
        *reinterpret_cast<uint4*>(&sV[r * PQ + ch * 8]) = v;
        *reinterpret_cast<uint4*>(&sG[r * PQ + ch * 8]) = g;
    }
    __syncthreads();
    const int g8 = lane >> 2, tq = lane & 3;
    constexpr float kScale = 0.17677669529663687f;
    constexpr float kScaleLog2 = kScale * 1.4426950408889634f;
    unsigned dq_pk[4][2];                     // this warp's dQ fragments (bf16 pairs), staged after the barrier
    {   // ---- phase 1: query block w
        unsigned qa[2][4], ga[2][4];
        const int arow = w * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldsm4(qa[ks], &sQ[arow * PQ + ks * 16 + (lane >> 4) * 8]);
            ldsm4(ga[ks], &sG[arow * PQ + ks * 16 + (lane >> 4) * 8]);
        }
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) s[nt][c] = dp[nt][c] = 0.0f;
            unsigned kb[4], vb[4];
            ldsm4(kb, &sK[(nt * 8 + (lane & 7)) * PQ + (lane >> 3) * 8]);
            mma16816(s[nt], qa[0], kb[0], kb[1]);
            mma16816(s[nt], qa[1], kb[2], kb[3]);
            ldsm4(vb, &sV[(nt * 8 + (lane & 7)) * PQ + (lane >> 3) * 8]);
            mma16816(dp[nt], ga[0], vb[0], vb[1]);
            mma16816(dp[nt], ga[1], vb[2], vb[3]);
        }
        const int i0 = w * 16 + g8, i1 = i0 + 8;
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = nt * 8 + tq * 2 + (c & 1);
                const int i = (c < 2) ? i0 : i1;
                const bool ok = j < L && (!causal || j <= i);
                s[nt][c] = ok ? s[nt][c] * kScaleLog2 : -INFINITY;
                if (c < 2) m0 = fmaxf(m0, s[nt][c]); else m1 = fmaxf(m1, s[nt][c]);
            }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s[nt][c] = exp2f(s[nt][c] - ((c < 2) ? m0 : m1));          // key 0 is always visible: the max is finite
                if (c < 2) l0 += s[nt][c]; else l1 += s[nt][c];
            }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        float dl0 = 0.0f, dl1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] *= inv0; s[nt][1] *= inv0; s[nt][2] *= inv1; s[nt][3] *= inv1;
            dl0 = fmaf(s[nt][0], dp[nt][0], fmaf(s[nt][1], dp[nt][1], dl0));
            dl1 = fmaf(s[nt][2], dp[nt][2], fmaf(s[nt][3], dp[nt][3], dl1));
        }
        dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1);
        dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
        dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1);
        dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
        unsigned dsa[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const unsigned p01 = pack_bf16x2(s[nt][0], s[nt][1]), p23 = pack_bf16x2(s[nt][2], s[nt][3]);
            const float e0 = s[nt][0] * (dp[nt][0] - dl0) * kScale, e1 = s[nt][1] * (dp[nt][1] - dl0) * kScale;
            const float e2 = s[nt][2] * (dp[nt][2] - dl1) * kScale, e3 = s[nt][3] * (dp[nt][3] - dl1) * kScale;
            const unsigned d01 = pack_bf16x2(e0, e1), d23 = pack_bf16x2(e2, e3);
            const int col = nt * 8 + tq * 2;
            *reinterpret_cast<unsigned*>(&sP[i0 * PP + col]) = p01;
            *reinterpret_cast<unsigned*>(&sP[i1 * PP + col]) = p23;
            *reinterpret_cast<unsigned*>(&sS[i0 * PP + col]) = d01;
            *reinterpret_cast<unsigned*>(&sS[i1 * PP + col]) = d23;
            dsa[nt >> 1][(nt & 1) * 2 + 0] = d01;
            dsa[nt >> 1][(nt & 1) * 2 + 1] = d23;
        }
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[nt][c] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned kb[4];
                ldsm4t(kb, &sK[(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PQ + np * 16 + (lane >> 4) * 8]);
                mma16816(o[np * 2 + 0], dsa[ks], kb[0], kb[1]);
                mma16816(o[np * 2 + 1], dsa[ks], kb[2], kb[3]);
            }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            dq_pk[nt][0] = pack_bf16x2(o[nt][0], o[nt][1]);
            dq_pk[nt][1] = pack_bf16x2(o[nt][2], o[nt][3]);
        }
    }
    __syncthreads();
    {   // the K tile is dead (phase 2 reads P, dS, dO and Q): stage dQ over it
        const int i0 = w * 16 + g8, i1 = i0 + 8;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<unsigned*>(&sK[i0 * PQ + nt * 8 + tq * 2]) = dq_pk[nt][0];
            *reinterpret_cast<unsigned*>(&sK[i1 * PQ + nt * 8 + tq * 2]) = dq_pk[nt][1];
        }
    }
    {   // ---- phase 2: key block w
        float dv[4][4], dk[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) dv[nt][c] = dk[nt][c] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            unsigned pa[4], sa[4];
            const int off = (ks * 16 + ((lane >> 4) & 1) * 8 + (lane & 7)) * PP + w * 16 + ((lane >> 3) & 1) * 8;
            ldsm4t(pa, &sP[off]);
            ldsm4t(sa, &sS[off]);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned gb[4], qb[4];
                const int boff = (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PQ + np * 16 + (lane >> 4) * 8;
                ldsm4t(gb, &sG[boff]);
                mma16816(dv[np * 2 + 0], pa, gb[0], gb[1]);
                mma16816(dv[np * 2 + 1], pa, gb[2], gb[3]);
                ldsm4t(qb, &sQ[boff]);
                mma16816(dk[np * 2 + 0], sa, qb[0], qb[1]);
                mma16816(dk[np * 2 + 1], sa, qb[2], qb[3]);
            }
        }
        const int j0 = w * 16 + g8, j1 = j0 + 8;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int col = nt * 8 + tq * 2;
            *reinterpret_cast<unsigned*>(&sO[j0 * PQ + col]) = pack_bf16x2(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<unsigned*>(&sO[j1 * PQ + col]) = pack_bf16x2(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<unsigned*>(&sV[j0 * PQ + col]) = pack_bf16x2(dv[nt][0], dv[nt][1]);     // the V tile died with phase 1
            *reinterpret_cast<unsigned*>(&sV[j1 * PQ + col]) = pack_bf16x2(dv[nt][2], dv[nt][3]);
        }
    }
    __syncthreads();
    // write-out: 64 rows x 3 parts (dQ | dK | dV) x 4 pieces of 16 bytes; 4 consecutive lanes cover one row's 64 bytes of a part
    for (int e = t; e < 64 * 12; e += 128) {
        const int part = e >> 8, r = (e >> 2) & 63, ch = e & 3;
        if (r < L) {
            const __nv_bfloat16* src = (part == 0 ? sK : part == 1 ? sO : sV) + r * PQ + ch * 8;
            *reinterpret_cast<uint4*>(&dqkv[(b * L + r) * 3 * d + part * d + hh * 32 + ch * 8]) = *reinterpret_cast<const uint4*>(src);
        }
    }
    if (colsum != nullptr && t < 48) {                     // a thread sums a PAIR of columns (the kernel is LSU-wavefront bound: ncu 80 %)
        const int part = t >> 4, c = (t & 15) * 2;
        const __nv_bfloat16* src = (part == 0 ? sK : part == 1 ? sO : sV) + c;
        float a0 = 0.0f, a1 = 0.0f;
        for (int r = 0; r < L; ++r) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + r * PQ));
            a0 += f.x;
            a1 += f.y;
        }
        *reinterpret_cast<float2*>(&colsum[b * 3 * d + part * d + hh * 32 + c]) = make_float2(a0, a1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Out head backward (denoiser out = Linear(d, D), D <= 4): dh[m, :] = sum_j dy[m, j] W[j, :]  (fp32 + bf16 copies)
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ W, long long M, int d, int D,
                                                       float* __restrict__ dh, __nv_bfloat16* __restrict__ dh16) {
    const long long n = M * d;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long m = i / d;
        const int c = static_cast<int>(i - m * d);
        float v = 0.0f;
        for (int j = 0; j < D; ++j) v = fmaf(dy[m * D + j], W[j * d + c], v);
        dh[i] = v;
        if (dh16) dh16[i] = __float2bfloat16_rn(v);
    }
}

// partial[slice][j][c] = sum over the slice's rows of A[m, j] * X[m, c]   (n <= 8 narrow columns of A)
// out head:  A = dy [M, D], X = h_final  -> dW_out [D, d];  in_proj:  A = features [M, D + C], X = dh0 -> dWf [D + C, d]
__global__ void __launch_bounds__(128) narrow_outer_partial_kernel(const float* __restrict__ A, int n, const float* __restrict__ X, long long M,
                                                                   int K, long long rows_per_slice, float* __restrict__ partial) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slice;
    long long r1 = r0 + rows_per_slice;
    if (r1 > M) r1 = M;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    if (c < K) {
        for (long long m = r0; m < r1; ++m) {
            const float x = X[m * K + c];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < n) acc[j] = fmaf(A[m * n + j], x, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < n) partial[(static_cast<long long>(blockIdx.y) * n + j) * K + c] = acc[j];
    }
}

// out[b, c] = sum_t src[b, t, c]   (gradient of the per-trajectory rows added to every token: level vector, cond row)
__global__ void __launch_bounds__(128) token_sum_kernel(const float* __restrict__ src, int L, int d, float* __restrict__ out) {
    const long long b = blockIdx.x;
    for (int c = threadIdx.x; c < d; c += 128) {
        float t = 0.0f;
        for (int l = 0; l < L; ++l) t += src[(b * L + l) * d + c];
        out[b * d + c] = t;
    }
}

// out[i, j] (+)= sum_k A[i * sa0 + k * sa1] * B[j * sb0 + k * sb1]: strided fp32 GEMM for the per-trajectory linears'
// backward (level_proj, cond_proj, maze.fc, sg.mlp, t_embed: M = batch, all dims <= 512)
// 32 x 32 output tile per block, 2 x 2 outputs per thread, k-steps of 32; the tile loads walk whichever index is contiguous in memory
// (sa1 == 1: k, else the row index), so both the plain and the transposed operand forms are coalesced.  (The first revision -- 16 x 16
// tiles, one output per thread, k always on the fast thread index -- read transposed operands with a stride of a whole row: 0.3-0.4 ms
// per call at B = 4096 for 1.2 GFLOP.)  Per-thread accumulation order over k is fixed: deterministic.
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, long long sa0, long long sa1, const float* __restrict__ B,
                                                            long long sb0, long long sb1, float* __restrict__ out, long long ldo, int M, int N,
                                                            int K, int accumulate) {
    __shared__ float As[32][34], Bs[32][34];                // [k][row]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    float acc[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
    for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            {
                const int r = (sa1 == 1) ? (e >> 5) : (e & 31), k = (sa1 == 1) ? (e & 31) : (e >> 5);
                const int ia = i0 + r, ka = k0 + k;
                As[k][r] = (ia < M && ka < K) ? A[ia * sa0 + ka * sa1] : 0.0f;
            }
            {
                const int r = (sb1 == 1) ? (e >> 5) : (e & 31), k = (sb1 == 1) ? (e & 31) : (e >> 5);
                const int jb = j0 + r, kb = k0 + k;
                Bs[k][r] = (jb < N && kb < K) ? B[jb * sb0 + kb * sb1] : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(&As[k][ty * 2]);
            const float2 b = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]);
            acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]);
            acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj) {
            const int i = i0 + ty * 2 + di, j = j0 + tx * 2 + dj;
            if (i < M && j < N) out[i * ldo + j] = accumulate ? out[i * ldo + j] + acc[di][dj] : acc[di][dj];
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// Conv encoder pieces (encoders.py:8-25): activations NHWC bf16 [B, H*W, C] holding the PRE-activation u of each layer.
// col[b*P + p, (ky*3 + kx)*C + c] = act(src[b, (y+ky-1, x+kx-1), c]) (zero outside the plane, zero in the K padding)
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ src, long long B, int Hh, int Ww, int C, int Kpad,
                                                        int act, __nv_bfloat16* __restrict__ col) {
    const int P = Hh * Ww;
    const long long total = B * P * static_cast<long long>(Kpad);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int kk = static_cast<int>(i % Kpad);
        const long long bp = i / Kpad;
        float v = 0.0f;
        if (kk < 9 * C) {
            const int tap = kk / C, c = kk - tap * C;
            const int p = static_cast<int>(bp % P);
            const long long b = bp / P;
            const int y = p / Ww + tap / 3 - 1, x = p % Ww + tap % 3 - 1;
            if (y >= 0 && y < Hh && x >= 0 && x < Ww) {
                v = __bfloat162float(src[(b * P + y * Ww + x) * C + c]);
                if (act) v = silu_fwd(v);
            }
        }
        col[i] = __float2bfloat16_rn(v);
    }
}

// C % 8 == 0: one thread moves 8 channels (16 bytes) of one tap
__global__ void __launch_bounds__(256) im2col3x3_vec_kernel(const __nv_bfloat16* __restrict__ src, long long B, int Hh, int Ww, int C, int Kpad,
                                                            int act, __nv_bfloat16* __restrict__ col) {
    const int P = Hh * Ww;
    const int kv = Kpad / 8;
    const long long total = B * P * static_cast<long long>(kv);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int kk = static_cast<int>(i % kv) * 8;
        const long long bp = i / kv;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (kk < 9 * C) {
            const int tap = kk / C, c = kk - tap * C;
            const int p = static_cast<int>(bp % P);
            const long long b = bp / P;
            const int y = p / Ww + tap / 3 - 1, x = p % Ww + tap % 3 - 1;
            if (y >= 0 && y < Hh && x >= 0 && x < Ww) {
                val = *reinterpret_cast<const uint4*>(src + (b * P + y * Ww + x) * C + c);
                if (act) {
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(h2[j]);
                        h2[j] = __floats2bfloat162_rn(silu_fwd(f.x), silu_fwd(f.y));
                    }
                }
            }
        }
        *reinterpret_cast<uint4*>(col + i * 8) = val;
    }
}

// row-blocked form: a block takes kRows consecutive patch rows; a thread keeps its (tap, channel-group) fixed across them, so the
// integer divisions happen once per thread and once per row instead of once per 16-byte element
constexpr int kIm2colRows = 16;
__global__ void __launch_bounds__(256) im2col3x3_rows_kernel(const __nv_bfloat16* __restrict__ src, long long rows_total, int Hh, int Ww, int C,
                                                             int Kpad, int act, __nv_bfloat16* __restrict__ col) {
    __shared__ long long s_base[kIm2colRows];              // element offset of (b, y, x, 0) in src
    __shared__ int s_y[kIm2colRows], s_x[kIm2colRows];
    const int P = Hh * Ww;
    const int kv = Kpad / 8;
    const long long row0 = static_cast<long long>(blockIdx.x) * kIm2colRows;
    if (threadIdx.x < kIm2colRows) {
        const long long row = row0 + threadIdx.x;
        if (row < rows_total) {
            const long long b = row / P;
            const int p = static_cast<int>(row - b * P);
            s_y[threadIdx.x] = p / Ww;
            s_x[threadIdx.x] = p % Ww;
            s_base[threadIdx.x] = (b * P + p) * C;
        }
    }
    __syncthreads();
    const int kvs = kv < 256 ? kv : 256;                   // threads along the patch row; the rest of the block strides over rows
    const int lanes_r = 256 / kvs;
    const int rl = threadIdx.x / kvs;
    if (rl >= lanes_r) return;
    for (int k8 = threadIdx.x % kvs; k8 < kv; k8 += kvs) {
        const int kk = k8 * 8;
        const bool live = kk < 9 * C;
        const int tap = live ? kk / C : 0;
        const int c = kk - tap * C;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const long long doff = static_cast<long long>(dy * Ww + dx) * C + c;
#pragma unroll 4
        for (int r = rl; r < kIm2colRows; r += lanes_r) {
            const long long row = row0 + r;
            if (row >= rows_total) break;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            const int y = s_y[r] + dy, x = s_x[r] + dx;
            if (live && y >= 0 && y < Hh && x >= 0 && x < Ww) {
                val = *reinterpret_cast<const uint4*>(src + s_base[r] + doff);
                if (act) {
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(h2[j]);
                        h2[j] = __floats2bfloat162_rn(silu_fwd(f.x), silu_fwd(f.y));
                    }
                }
            }
            *reinterpret_cast<uint4*>(col + row * Kpad + kk) = val;
        }
    }
}

// scatter form (the default for C % 8 == 0): a thread owns ONE 16-byte channel group of ONE input pixel: one load, the activation
// applied once (the gather forms above apply it once per tap: 9x the MUFU work -- the 128-channel layer was MUFU-bound at 2.1 TB/s),
// then up to nine 16-byte stores: col[(y - dy, x - dx), tap(dy, dx)] = a[(y, x)].  The zero entries of the patch matrix (taps of a
// border pixel that fall outside the plane) are written by the thread of that pixel; the K padding by the pixel's first thread.
// Consecutive threads hold consecutive channel groups of a pixel, so every store instruction covers C * 2 contiguous bytes per pixel.
__global__ void __launch_bounds__(256) im2col3x3_scatter_kernel(const __nv_bfloat16* __restrict__ src, long long B, int Hh, int Ww, int C,
                                                                int Kpad, int act, __nv_bfloat16* __restrict__ col) {
    const int P = Hh * Ww;
    const int cg = C / 8;
    const long long total = B * P * static_cast<long long>(cg);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long pix = i / cg;                       // b * P + p
        const int c = static_cast<int>(i - pix * cg) * 8;
        const int p = static_cast<int>(pix % P);
        const int y = p / Ww, x = p - y * Ww;
        uint4 val = __ldg(reinterpret_cast<const uint4*>(src + pix * C + c));
        if (act) {
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                h2[j] = __floats2bfloat162_rn(silu_fwd(f.x), silu_fwd(f.y));
            }
        }
        __nv_bfloat16* self = col + pix * Kpad + c;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            // this pixel as the (dy, dx) neighbour of output pixel (y - dy, x - dx)
            if (y - dy >= 0 && y - dy < Hh && x - dx >= 0 && x - dx < Ww)
                *reinterpret_cast<uint4*>(self - static_cast<long long>(dy * Ww + dx) * Kpad + tap * C) = val;
            // this pixel's own tap (dy, dx) falls outside the plane: zero
            if (y + dy < 0 || y + dy >= Hh || x + dx < 0 || x + dx >= Ww) *reinterpret_cast<uint4*>(self + tap * C) = zero;
        }
        if (c == 0)
            for (int k = 9 * C; k < Kpad; k += 8) *reinterpret_cast<uint4*>(col + pix * Kpad + k) = zero;
    }
}

// pooled[b, c] = mean_p silu(u[b, p, c])
// block per image: 64 channels x 4 position groups at a time (the first revision walked the P positions serially per channel with 128
// threads: latency bound, 0.23 ms at B = 512); fixed summation order.
__global__ void __launch_bounds__(256) pool_silu_kernel(const __nv_bfloat16* __restrict__ u, int P, int C, float* __restrict__ pooled) {
    __shared__ float sh[256];
    const long long b = blockIdx.x;
    const int t = threadIdx.x, g = t >> 6;
    for (int c0 = 0; c0 < C; c0 += 64) {
        const int c = c0 + (t & 63);
        float acc = 0.0f;
        if (c < C)
            for (int p = g; p < P; p += 4) acc += silu_fwd(__bfloat162float(u[(b * P + p) * C + c]));
        sh[t] = acc;
        __syncthreads();
        if (g == 0 && c < C) pooled[b * C + c] = ((sh[t] + sh[t + 64]) + (sh[t + 128] + sh[t + 192])) / static_cast<float>(P);
        __syncthreads();
    }
}

// du[b, p, c] = dpooled[b, c] / P * silu'(u[b, p, c])
__global__ void __launch_bounds__(256) pool_silu_bwd_kernel(const __nv_bfloat16* __restrict__ u, const float* __restrict__ dpooled, long long B,
                                                            int P, int C, __nv_bfloat16* __restrict__ du) {
    const long long total = B * P * static_cast<long long>(C);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const float invP = 1.0f / static_cast<float>(P);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = static_cast<int>(i % C);
        const long long b = i / (static_cast<long long>(P) * C);
        du[i] = __float2bfloat16_rn(dpooled[b * C + c] * invP * silu_grad(__bfloat162float(u[i])));
    }
}

template <int kL, int kT>
int launch_attn_bwd(const void* qkv, const void* dO, void* dqkv, long long B, int L, int H, int causal, cudaStream_t st) {
    using Cfg = AttnBwdCfg<kL, kT>;
    constexpr int smem = Cfg::kSmemFloats * 4;
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<kL, kT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(attention_bwd): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    attention_bwd_kernel<kL, kT><<<static_cast<unsigned>(B * H), Cfg::kThreads, smem, st>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dO), static_cast<__nv_bfloat16*>(dqkv), L, H, causal);
    return check_launch("attention_bwd_kernel");
}

template <typename TDa, bool kDhSum>
int ln_bwd_two_pass(const TDa* da, const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, long long gb_stride,
                    long long B, int L, int d, float* dh, __nv_bfloat16* d16, float* dgb, long long dgb_stride, float* dwb_part,
                    float4* stats, cudaStream_t st) {
    IDB_REQUIRE(aligned(da, 16) && aligned(h, 16) && aligned(dh, 16) && aligned(ln_w, 16) && aligned(ln_b, 16) && aligned(dwb_part, 16) &&
                    (!d16 || aligned(d16, 8)) && (!gamma_beta || (aligned(gamma_beta, 16) && gb_stride % 4 == 0)) &&
                    (!dgb || (aligned(dgb, 16) && dgb_stride % 4 == 0)),
                IDB200_EALIGN, "LayerNorm backward (two-pass form) needs 16-byte aligned rows");
    const long long M = B * L;
    // MEASURED AND REJECTED as the default (tools/bench_ln_bwd.py, d = 384, L = 64): the one-launch form takes 0.430 ms at B = 4096 and
    // 0.0795 ms at B = 512 against 0.409 / 0.0753 ms of the two passes -- those already stream their 22 bytes per element at 5.4 TB/s
    // (82 % of the measured copy bandwidth); the block-per-trajectory form moves 16 but is latency bound.  IDB200_LN_BWD_FUSED=1 opts in.
    static const int fused_env = getenv("IDB200_LN_BWD_FUSED") ? atoi(getenv("IDB200_LN_BWD_FUSED")) : 0;
    if (fused_env && L >= 2 && L <= 1024) {
        // pad_kb (dev knob) caps the resident blocks per SM: the re-read of pass B has to find the trajectory's rows in L2
        static const int pad_kb = getenv("IDB200_LN_BWD_PAD_KB") ? atoi(getenv("IDB200_LN_BWD_PAD_KB")) : 0;
        const size_t smem = (static_cast<size_t>(L) + 5 * (d / 4)) * sizeof(float4) + static_cast<size_t>(pad_kb) * 1024;
        auto go = [&](auto kern, int threads) -> int {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
                if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(ln_bwd_fused): %s", cudaGetErrorString(e));
            }
            kern<<<static_cast<unsigned>(B), threads, smem, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, dh, d16, dgb, dgb_stride, dwb_part);
            return check_launch("ln_bwd_fused_kernel");
        };
        switch (d / 128) {
            case 1: return go(ln_bwd_fused_kernel<1, TDa, kDhSum>, 64);
            case 2: return go(ln_bwd_fused_kernel<2, TDa, kDhSum>, 128);
            case 3: return go(ln_bwd_fused_kernel<3, TDa, kDhSum>, 192);
            default: return go(ln_bwd_fused_kernel<4, TDa, kDhSum>, 256);
        }
    }
    const int g1 = grid_for(M * 32, 256, 8);
    switch (d / 128) {
        case 1: ln_bwd_stats_kernel<1, TDa><<<g1, 256, 0, st>>>(da, h, ln_w, gamma_beta, gb_stride, L, M, stats); break;
        case 2: ln_bwd_stats_kernel<2, TDa><<<g1, 256, 0, st>>>(da, h, ln_w, gamma_beta, gb_stride, L, M, stats); break;
        case 3: ln_bwd_stats_kernel<3, TDa><<<g1, 256, 0, st>>>(da, h, ln_w, gamma_beta, gb_stride, L, M, stats); break;
        default: ln_bwd_stats_kernel<4, TDa><<<g1, 256, 0, st>>>(da, h, ln_w, gamma_beta, gb_stride, L, M, stats); break;
    }
    int rc = check_launch("ln_bwd_stats_kernel");
    if (rc) return rc;
    static const int groups_env = getenv("IDB200_LN_BWD_GROUPS") ? atoi(getenv("IDB200_LN_BWD_GROUPS")) : -1;
    const bool small = groups_env >= 0 ? groups_env == 4 : (B < 8ll * num_sms() && L >= 8);
    if (small && d <= 512) {                               // 4 row groups x d / 4 threads (<= 512 threads)
        const size_t smem = static_cast<size_t>(3) * 5 * (d / 4) * sizeof(float4);
        ln_bwd_apply_groups_kernel<TDa, kDhSum, 4><<<static_cast<unsigned>(B), d, smem, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, d, stats,
                                                                                           dh, d16, dgb, dgb_stride, dwb_part);
        return check_launch("ln_bwd_apply_groups_kernel");
    }
    ln_bwd_apply_kernel<TDa, kDhSum><<<static_cast<unsigned>(B), d / 4, 0, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, d, stats, dh, d16,
                                                                                 dgb, dgb_stride, dwb_part);
    return check_launch("ln_bwd_apply_kernel");
}

inline int slices_for(long long M, int col_blocks) {
    long long s = (148 * 8 + col_blocks - 1) / col_blocks;
    const long long max_s = (M + 63) / 64;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return static_cast<int>(s);
}

}  // namespace tb
}  // namespace idb200

using namespace idb200;

extern "C" int idb200_transpose_bf16(const void* src, int src_is_f32, int64_t M, int N, void* dst, idb200_stream_t stream) {
    IDB_REQUIRE(src && dst, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M >= 0 && N > 0, IDB200_EINVAL, "bad shape");
    if (M == 0) return IDB200_OK;
    const dim3 grid(static_cast<unsigned>((M + 63) / 64), static_cast<unsigned>((N + 63) / 64));
    if (src_is_f32)
        tb::transpose_to_bf16_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(src), M, N,
                                                                                                   static_cast<__nv_bfloat16*>(dst));
    else if (M % 8 == 0 && N % 8 == 0 && aligned(src, 16) && aligned(dst, 16))
        tb::transpose_bf16_vec_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(src), M, N,
                                                                                            static_cast<__nv_bfloat16*>(dst));
    else
        tb::transpose_to_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(src), M, N, static_cast<__nv_bfloat16*>(dst));
    return check_launch("transpose_to_bf16_kernel");
}

extern "C" int idb200_colsum_scratch_floats(int64_t M, int N) {
    // the scalar path (32 columns per block) or the vector path (32 x 4 fp32 / 32 x 8 bf16 columns per block), whichever slices more
    const int a = tb::slices_for(M, (N + 31) / 32), b = tb::slices_for(M, (N + 127) / 128), c = tb::slices_for(M, (N + 255) / 256);
    return (a > b ? (a > c ? a : c) : (b > c ? b : c)) * N;
}

static int colsum_impl(const void* src, int src_kind, int segs, int64_t M, int N, float* scratch, float scale, int accumulate, float* out,
                       idb200_stream_t stream) {
    IDB_REQUIRE(src && out && scratch, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M > 0 && N > 0 && segs > 0 && segs <= 65535, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(src_kind >= 0 && src_kind <= 1, IDB200_EINVAL, "src_kind: 0 fp32, 1 bf16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cb = (N + 31) / 32;
    const int S = tb::slices_for(M, cb);                   // (the scratch is sized for this, the larger, slice count)
    const int kV = src_kind == 0 ? 4 : 8;
    const long long ld = static_cast<long long>(segs) * N;
    int S_used = S;
    if (N % kV == 0 && aligned(src, 16)) {
        const int cbv = (N + 32 * kV - 1) / (32 * kV);
        S_used = tb::slices_for(M, cbv);                    // (idb200_colsum_scratch_floats covers the larger of the two)
        const long long rps = (M + S_used - 1) / S_used;
        const dim3 grid(cbv, S_used, segs);
        if (src_kind == 0)
            tb::colsum_partial_vec_kernel<float, 4><<<grid, 256, 0, st>>>(static_cast<const float*>(src), M, N, rps, scratch, ld);
        else
            tb::colsum_partial_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), M, N, rps, scratch, ld);
    } else {
        const long long rps = (M + S - 1) / S;
        const dim3 grid(cb, S, segs);
        if (src_kind == 0)
            tb::colsum_partial_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), M, N, rps, scratch, ld);
        else
            tb::colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), M, N, rps, scratch, ld);
    }
    int rc = check_launch("colsum_partial_kernel");
    if (rc) return rc;
    tb::reduce_rows_kernel<<<static_cast<unsigned>((ld + 255) / 256), 256, 0, st>>>(scratch, S_used, ld, scale, accumulate, out);
    return check_launch("reduce_rows_kernel");
}

extern "C" int idb200_colsum(const void* src, int src_kind, int64_t M, int N, float* scratch, float scale, int accumulate, float* out,
                             idb200_stream_t stream) {
    return colsum_impl(src, src_kind, 1, M, N, scratch, scale, accumulate, out, stream);
}

extern "C" int idb200_colsum_segments(const void* src, int src_kind, int segs, int64_t M, int N, float* scratch, float scale, int accumulate,
                                      float* out, idb200_stream_t stream) {
    return colsum_impl(src, src_kind, segs, M, N, scratch, scale, accumulate, out, stream);
}

extern "C" int idb200_multi_copy_f32(const float* const* srcs, float* const* dsts, const int64_t* counts, int n, idb200_stream_t stream) {
    IDB_REQUIRE(n >= 0 && (n == 0 || (srcs && dsts && counts)), IDB200_EINVAL, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i0 = 0; i0 < n; i0 += tb::kMaxCopy) {
        tb::CopyBatch b{};
        const int m = n - i0 < tb::kMaxCopy ? n - i0 : tb::kMaxCopy;
        int64_t mx = 1;
        for (int i = 0; i < m; ++i) {
            IDB_REQUIRE(srcs[i0 + i] && dsts[i0 + i] && counts[i0 + i] >= 0 && counts[i0 + i] < (1ll << 31), IDB200_EINVAL, "bad segment %d", i0 + i);
            b.src[i] = srcs[i0 + i];
            b.dst[i] = dsts[i0 + i];
            b.n[i] = static_cast<int>(counts[i0 + i]);
            if (counts[i0 + i] > mx) mx = counts[i0 + i];
        }
        const int gx = static_cast<int>((mx + 1023) / 1024 < 64 ? (mx + 1023) / 1024 : 64);
        tb::multi_copy_kernel<<<dim3(gx, m), 256, 0, st>>>(b);
        int rc = check_launch("multi_copy_kernel");
        if (rc) return rc;
    }
    return IDB200_OK;
}

extern "C" int idb200_cast_weights_bf16(const float* const* srcs, void* const* dsts, void* const* dsts_t, const int* rows, const int* cols, int n,
                                        idb200_stream_t stream) {
    IDB_REQUIRE(n >= 0 && (n == 0 || (srcs && dsts && dsts_t && rows && cols)), IDB200_EINVAL, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i0 = 0; i0 < n; i0 += tb::kMaxCast) {
        tb::CastBatch b{};
        const int m = n - i0 < tb::kMaxCast ? n - i0 : tb::kMaxCast;
        int mr = 1, mc = 1;
        for (int i = 0; i < m; ++i) {
            IDB_REQUIRE(srcs[i0 + i] && rows[i0 + i] > 0 && cols[i0 + i] > 0 && (dsts[i0 + i] || dsts_t[i0 + i]), IDB200_EINVAL, "bad matrix %d", i0 + i);
            b.src[i] = srcs[i0 + i];
            b.dst[i] = static_cast<__nv_bfloat16*>(dsts[i0 + i]);
            b.dst_t[i] = static_cast<__nv_bfloat16*>(dsts_t[i0 + i]);
            b.rows[i] = rows[i0 + i];
            b.cols[i] = cols[i0 + i];
            if (rows[i0 + i] > mr) mr = rows[i0 + i];
            if (cols[i0 + i] > mc) mc = cols[i0 + i];
        }
        tb::cast_weights_kernel<<<dim3((mc + 63) / 64, (mr + 63) / 64, m), 256, 0, st>>>(b);
        int rc = check_launch("cast_weights_kernel");
        if (rc) return rc;
    }
    return IDB200_OK;
}

extern "C" int idb200_reduce_rows(const float* partial, int R, int64_t W, float scale, int accumulate, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(partial && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(R > 0 && W > 0, IDB200_EINVAL, "bad shape");
    tb::reduce_rows_kernel<<<static_cast<unsigned>((W + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(partial, R, W, scale,
                                                                                                                  accumulate, out);
    return check_launch("reduce_rows_kernel");
}

extern "C" int idb200_silu_bf16(const void* u, const void* g, int64_t n, int mode, void* y, idb200_stream_t stream) {
    IDB_REQUIRE(u && y && (mode == 0 || g), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(n >= 0 && n % 2 == 0, IDB200_EINVAL, "element count must be even");
    IDB_REQUIRE(mode == 0 || mode == 1, IDB200_EINVAL, "mode: 0 forward, 1 backward");
    if (n == 0) return IDB200_OK;
    if (n % 8 == 0 && aligned(u, 16) && aligned(y, 16) && (mode == 0 || aligned(g, 16))) {
        tb::silu_vec_kernel<<<grid_for(n / 8, 256 * 2, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const uint4*>(u), static_cast<const uint4*>(g), n / 8, mode, static_cast<uint4*>(y));
        return check_launch("silu_vec_kernel");
    }
    tb::silu_kernel<<<grid_for(n / 2, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat162*>(u), static_cast<const __nv_bfloat162*>(g), n / 2, mode, static_cast<__nv_bfloat162*>(y));
    return check_launch("silu_kernel");
}

extern "C" int idb200_silu_f32(const float* u, const float* g, int64_t n, int mode, float* y, idb200_stream_t stream) {
    IDB_REQUIRE(u && y && (mode == 0 || g), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(n >= 0 && (mode == 0 || mode == 1), IDB200_EINVAL, "bad arguments");
    if (n == 0) return IDB200_OK;
    tb::silu_f32_kernel<<<grid_for(n, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(u, g, n, mode, y);
    return check_launch("silu_f32_kernel");
}

extern "C" int idb200_ln_film_bwd(const float* da, const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta,
                                  int64_t gb_stride, int64_t B, int L, int d, float* dh, void* dh_bf16, float* dgb, int64_t dgb_stride,
                                  float* dwb_part, float* stats_scratch, idb200_stream_t stream) {
    IDB_REQUIRE(da && h && ln_w && ln_b && dh && dwb_part, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((gamma_beta == nullptr) == (dgb == nullptr), IDB200_EINVAL, "gamma_beta and dgb must be given together");
    IDB_REQUIRE(B > 0 && L > 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(d == 256 || d == 384 || d == 128 || d == 512, IDB200_EUNSUPPORTED, "d_model must be 128, 256, 384 or 512 (got %d)", d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* d16 = static_cast<__nv_bfloat16*>(dh_bf16);
    const unsigned grid = static_cast<unsigned>(B);
    if (stats_scratch) {                                    // two-pass form: stats_scratch fp32 [B*L, 4], 16-byte aligned
        IDB_REQUIRE(aligned(stats_scratch, 16), IDB200_EALIGN, "stats_scratch must be 16-byte aligned");
        return tb::ln_bwd_two_pass<float, false>(da, h, ln_w, ln_b, gamma_beta, gb_stride, B, L, d, dh, d16, dgb, dgb_stride, dwb_part,
                                                 reinterpret_cast<float4*>(stats_scratch), st);
    }
    switch (d / 32) {
        case 4: tb::ln_film_bwd_kernel<4><<<grid, 256, 0, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, dh, d16, dgb, dgb_stride, dwb_part); break;
        case 8: tb::ln_film_bwd_kernel<8><<<grid, 256, 0, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, dh, d16, dgb, dgb_stride, dwb_part); break;
        case 12: tb::ln_film_bwd_kernel<12><<<grid, 256, 0, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, dh, d16, dgb, dgb_stride, dwb_part); break;
        default: tb::ln_film_bwd_kernel<16><<<grid, 256, 0, st>>>(da, h, ln_w, ln_b, gamma_beta, gb_stride, L, dh, d16, dgb, dgb_stride, dwb_part); break;
    }
    return check_launch("ln_film_bwd_kernel");
}

extern "C" int idb200_ln_film_bwd2(const void* da, int da_is_bf16, const float* h, const float* ln_w, const float* ln_b,
                                   const float* gamma_beta, int64_t gb_stride, int64_t B, int L, int d, float* dh, void* dh_bf16, float* dgb,
                                   int64_t dgb_stride, float* dwb_part, int with_dh_sum, float* stats_scratch, idb200_stream_t stream) {
    IDB_REQUIRE(da && h && ln_w && ln_b && dh && dwb_part && stats_scratch, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((gamma_beta == nullptr) == (dgb == nullptr), IDB200_EINVAL, "gamma_beta and dgb must be given together");
    IDB_REQUIRE(B > 0 && L > 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(d == 256 || d == 384 || d == 128 || d == 512, IDB200_EUNSUPPORTED, "d_model must be 128, 256, 384 or 512 (got %d)", d);
    IDB_REQUIRE(aligned(stats_scratch, 16), IDB200_EALIGN, "stats_scratch must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* d16 = static_cast<__nv_bfloat16*>(dh_bf16);
    float4* stats = reinterpret_cast<float4*>(stats_scratch);
    if (da_is_bf16) {
        const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(da);
        return with_dh_sum ? tb::ln_bwd_two_pass<__nv_bfloat16, true>(g, h, ln_w, ln_b, gamma_beta, gb_stride, B, L, d, dh, d16, dgb, dgb_stride, dwb_part, stats, st)
                           : tb::ln_bwd_two_pass<__nv_bfloat16, false>(g, h, ln_w, ln_b, gamma_beta, gb_stride, B, L, d, dh, d16, dgb, dgb_stride, dwb_part, stats, st);
    }
    const float* g = static_cast<const float*>(da);
    return with_dh_sum ? tb::ln_bwd_two_pass<float, true>(g, h, ln_w, ln_b, gamma_beta, gb_stride, B, L, d, dh, d16, dgb, dgb_stride, dwb_part, stats, st)
                       : tb::ln_bwd_two_pass<float, false>(g, h, ln_w, ln_b, gamma_beta, gb_stride, B, L, d, dh, d16, dgb, dgb_stride, dwb_part, stats, st);
}

static int attention_bwd_impl(const void* qkv, const void* dO, void* dqkv, float* traj_colsum, int64_t B, int L, int H, int causal,
                              int force_simt, idb200_stream_t stream) {
    IDB_REQUIRE(qkv && dO && dqkv, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(!traj_colsum || (L > 32 && !force_simt), IDB200_EUNSUPPORTED,
                "per-trajectory column sums come from the tensor-core kernel only (32 < L <= 64; got L = %d)", L);
    IDB_REQUIRE(B > 0 && H > 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(L >= 1 && L <= 64, IDB200_EUNSUPPORTED, "attention backward supports L <= 64 (got %d)", L);
    IDB_REQUIRE(aligned(qkv, 16) && aligned(dO, 16) && aligned(dqkv, 16), IDB200_EALIGN, "qkv / dO / dqkv must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (L <= 8) return tb::launch_attn_bwd<8, 1>(qkv, dO, dqkv, B, L, H, causal, st);
    if (L <= 16) return tb::launch_attn_bwd<16, 1>(qkv, dO, dqkv, B, L, H, causal, st);
    if (L <= 32) return tb::launch_attn_bwd<32, 2>(qkv, dO, dqkv, B, L, H, causal, st);
    if (force_simt) return tb::launch_attn_bwd<64, 4>(qkv, dO, dqkv, B, L, H, causal, st);
    tb::attention_bwd_mma_kernel<<<static_cast<unsigned>(B * H), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                                               static_cast<const __nv_bfloat16*>(dO),
                                                                               static_cast<__nv_bfloat16*>(dqkv), L, H, causal, traj_colsum);
    return check_launch("attention_bwd_mma_kernel");
}

extern "C" int idb200_attention_bwd(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int H, int causal, int force_simt,
                                    idb200_stream_t stream) {
    return attention_bwd_impl(qkv, dO, dqkv, nullptr, B, L, H, causal, force_simt, stream);
}

extern "C" int idb200_attention_bwd_sums(const void* qkv, const void* dO, void* dqkv, float* traj_colsum, int64_t B, int L, int H, int causal,
                                         idb200_stream_t stream) {
    IDB_REQUIRE(traj_colsum != nullptr, IDB200_EINVAL, "traj_colsum is NULL (use idb200_attention_bwd)");
    return attention_bwd_impl(qkv, dO, dqkv, traj_colsum, B, L, H, causal, 0, stream);
}

extern "C" int idb200_head_bwd(const float* dy, const float* W, int64_t M, int d, int D, float* dh, void* dh_bf16, idb200_stream_t stream) {
    IDB_REQUIRE(dy && W && dh, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M > 0 && d > 0 && D > 0 && D <= 8, IDB200_EINVAL, "bad shape");
    tb::head_bwd_kernel<<<grid_for(M * d, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, W, M, d, D, dh,
                                                                                                    static_cast<__nv_bfloat16*>(dh_bf16));
    return check_launch("head_bwd_kernel");
}

extern "C" int idb200_narrow_outer_scratch_floats(int64_t M, int n, int K) {
    return tb::slices_for(M, (K + 127) / 128) * n * K;
}

extern "C" int idb200_narrow_outer(const float* A, int n, const float* X, int64_t M, int K, float* scratch, int accumulate, float* out,
                                   idb200_stream_t stream) {
    IDB_REQUIRE(A && X && scratch && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M > 0 && K > 0 && n >= 1 && n <= 8, IDB200_EINVAL, "bad shape (n must be 1..8)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cb = (K + 127) / 128;
    const int S = tb::slices_for(M, cb);
    const long long rps = (M + S - 1) / S;
    tb::narrow_outer_partial_kernel<<<dim3(cb, S), 128, 0, st>>>(A, n, X, M, K, rps, scratch);
    int rc = check_launch("narrow_outer_partial_kernel");
    if (rc) return rc;
    const long long W = static_cast<long long>(n) * K;
    tb::reduce_rows_kernel<<<static_cast<unsigned>((W + 255) / 256), 256, 0, st>>>(scratch, S, W, 1.0f, accumulate, out);
    return check_launch("reduce_rows_kernel");
}

extern "C" int idb200_token_sum(const float* src, int64_t B, int L, int d, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(src && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B > 0 && L > 0 && d > 0, IDB200_EINVAL, "bad shape");
    tb::token_sum_kernel<<<static_cast<unsigned>(B), 128, 0, static_cast<cudaStream_t>(stream)>>>(src, L, d, out);
    return check_launch("token_sum_kernel");
}

extern "C" int idb200_sgemm_strided(const float* A, int64_t sa0, int64_t sa1, const float* Bm, int64_t sb0, int64_t sb1, float* out,
                                    int64_t ldo, int M, int N, int K, int accumulate, idb200_stream_t stream) {
    IDB_REQUIRE(A && Bm && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M > 0 && N > 0 && K > 0, IDB200_EINVAL, "bad shape");
    tb::sgemm_strided_kernel<<<dim3((N + 31) / 32, (M + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, sa0, sa1, Bm, sb0, sb1, out,
                                                                                                               ldo, M, N, K, accumulate);
    return check_launch("sgemm_strided_kernel");
}

extern "C" int idb200_im2col3x3(const void* src, int64_t B, int H, int W, int C, int Kpad, int act, void* col, idb200_stream_t stream) {
    IDB_REQUIRE(src && col, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Kpad >= 9 * C, IDB200_EINVAL, "bad shape");
    if (C % 8 == 0 && Kpad % 8 == 0 && aligned(src, 16) && aligned(col, 16)) {
        const long long rows = B * H * W;
        const long long blocks = (rows + tb::kIm2colRows - 1) / tb::kIm2colRows;
        static const bool gather = getenv("IDB200_IM2COL_GATHER") != nullptr;        // dev A/B
        if (!gather) {
            tb::im2col3x3_scatter_kernel<<<grid_for(rows * (C / 8), 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                static_cast<const __nv_bfloat16*>(src), B, H, W, C, Kpad, act, static_cast<__nv_bfloat16*>(col));
            return check_launch("im2col3x3_scatter_kernel");
        }
        if (Kpad / 8 >= 64 && blocks < (1ll << 31)) {
            tb::im2col3x3_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                static_cast<const __nv_bfloat16*>(src), rows, H, W, C, Kpad, act, static_cast<__nv_bfloat16*>(col));
            return check_launch("im2col3x3_rows_kernel");
        }
        tb::im2col3x3_vec_kernel<<<grid_for(B * H * W * (Kpad / 8), 256 * 2, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(src), B, H, W, C, Kpad, act, static_cast<__nv_bfloat16*>(col));
        return check_launch("im2col3x3_vec_kernel");
    }
    tb::im2col3x3_kernel<<<grid_for(B * H * W * Kpad, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(src), B, H, W, C, Kpad, act, static_cast<__nv_bfloat16*>(col));
    return check_launch("im2col3x3_kernel");
}

extern "C" int idb200_pool_silu(const void* u, int64_t B, int P, int C, float* pooled, idb200_stream_t stream) {
    IDB_REQUIRE(u && pooled, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B > 0 && P > 0 && C > 0, IDB200_EINVAL, "bad shape");
    tb::pool_silu_kernel<<<static_cast<unsigned>(B), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(u), P, C, pooled);
    return check_launch("pool_silu_kernel");
}

extern "C" int idb200_pool_silu_bwd(const void* u, const float* dpooled, int64_t B, int P, int C, void* du, idb200_stream_t stream) {
    IDB_REQUIRE(u && dpooled && du, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B > 0 && P > 0 && C > 0, IDB200_EINVAL, "bad shape");
    tb::pool_silu_bwd_kernel<<<grid_for(B * P * C, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(u), dpooled, B, P, C, static_cast<__nv_bfloat16*>(du));
    return check_launch("pool_silu_bwd_kernel");
}
