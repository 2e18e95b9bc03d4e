// K4: two-layer maze conv encoder on tensor cores (implicit GEMM), one launch:
//   occ (+sdf) [B,Cin,H,W] -> conv3x3(Cin->C1)+SiLU -> conv3x3(C1->C2)+SiLU -> mean over H x W -> pooled [B,C2]
//
// Reference: MazeEncoder.forward, src/models/encoders.py:22-25 with the default maze_channels=(32, 64); the second
// conv is 97 % of the encoder's FLOPs (M = H*W pixels, N = C2, K = 9*C1 = 288 per trajectory).
//
// One CTA per trajectory (grid-stride).  The first conv (Cin = 1 or 2: 0.25 MFLOP) runs on CUDA cores straight into a
// zero-bordered channels-last bf16 activation tile [(H+2)*(W+2)][C1+8] in shared memory.  The second conv never
// materialises im2col: for tap (ky,kx) the A fragment of output pixel (y,x) is the C1-vector of padded pixel
// (y+ky, x+kx), fetched with ldmatrix from per-lane row addresses; B fragments come from the bf16 weight matrix
// Wp[C2][9*C1] (+8 pad) resident in shared memory for the CTA's lifetime; mma.sync.m16n8k16 bf16, fp32 accumulate.
// Epilogue: bias + SiLU, rows >= H*W masked, per-channel sums reduced in a fixed order (deterministic pooled mean).
#include <cuda_bf16.h>

#include "common.cuh"

namespace idb200 {
namespace convtc {

__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(static_cast<unsigned>(__cvta_generic_to_shared(p))));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float silu_e(float x) { return x / (1.0f + __expf(-x)); }

struct Params {
    const float* occ;              // [B,1,H,W]
    const float* sdf;              // [B,1,H,W] or nullptr
    const float* w0;               // [C1, Cin, 3, 3] fp32
    const float* b0;               // [C1]
    const __nv_bfloat16* w1;       // [C2, 9*C1] bf16, k = (ky*3+kx)*C1 + c
    const float* b1;               // [C2]
    float* pooled;                 // [B, C2]
    long long B;
    int cin, c1, c2, H, W;
};

constexpr int kMaxNT = 4;          // n-tiles (8 channels) per warp: C2 <= 64

__global__ void __launch_bounds__(256) conv2l_tc_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_c[];
    const int PW = p.W + 2, PP = (p.H + 2) * PW, HW = p.H * p.W;
    const int P1 = p.c1 + 8, K1 = 9 * p.c1, KP = K1 + 8;
    // carve-up (all offsets multiples of 16 bytes)
    __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(smem_c);                           // [PP][P1]
    __nv_bfloat16* w1s = act + ((static_cast<size_t>(PP) * P1 + 7) & ~static_cast<size_t>(7)); // [C2][KP]
    float* inp = reinterpret_cast<float*>(w1s + static_cast<size_t>(p.c2) * KP);             // [cin][PP]
    float* w0s = inp + ((p.cin * PP + 3) & ~3);                                               // [C1][cin][9]
    float* b0s = w0s + p.c1 * p.cin * 9;
    float* b1s = b0s + p.c1;
    float* part = b1s + p.c2;                                                                 // [4][C2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p.c2 * K1; i += 256) w1s[(i / K1) * KP + (i % K1)] = p.w1[i];
    for (int i = tid; i < p.c1 * p.cin * 9; i += 256) w0s[i] = p.w0[i];
    for (int i = tid; i < p.c1; i += 256) b0s[i] = p.b0[i];
    for (int i = tid; i < p.c2; i += 256) b1s[i] = p.b1[i];
    for (int i = tid; i < p.cin * PP; i += 256) inp[i] = 0.0f;                               // zero border, written once

    const int MT = (HW + 15) / 16;
    const int NT = p.c2 / 16;                     // n-tiles per warp (two warps split the channels)
    const int nhalf = warp & 1, mgrp = warp >> 1;
    const int g = lane >> 2, tq = lane & 3;

    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < p.cin * HW; i += 256) {
            const int c = i / HW, r = i - c * HW, y = r / p.W, x = r - y * p.W;
            const float* src = (c == 0) ? p.occ : p.sdf;
            inp[c * PP + (y + 1) * PW + x + 1] = src[b * HW + r];
        }
        __syncthreads();
        // ---- conv0 + SiLU -> act (bf16, channels-last, zero border) ----
        const int cpairs = p.c1 >> 1;
        for (int item = tid; item < PP * cpairs; item += 256) {
            const int pix = item / cpairs, cp = item - pix * cpairs;
            const int yy = pix / PW, xx = pix - yy * PW;
            float v0 = 0.0f, v1 = 0.0f;
            if (yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) {
                float a0 = b0s[2 * cp], a1 = b0s[2 * cp + 1];
                for (int ci = 0; ci < p.cin; ++ci) {
                    const float* ip = inp + ci * PP + (yy - 1) * PW + (xx - 1);
                    const float* wa = w0s + ((2 * cp) * p.cin + ci) * 9;
                    const float* wb = w0s + ((2 * cp + 1) * p.cin + ci) * 9;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float iv = ip[ky * PW + kx];
                            a0 = fmaf(iv, wa[ky * 3 + kx], a0);
                            a1 = fmaf(iv, wb[ky * 3 + kx], a1);
                        }
                }
                v0 = a0 / (1.0f + expf(-a0));
                v1 = a1 / (1.0f + expf(-a1));
            }
            *reinterpret_cast<__nv_bfloat162*>(act + pix * P1 + 2 * cp) = __floats2bfloat162_rn(v0, v1);
        }
        __syncthreads();
        // ---- conv1 as implicit GEMM ----
        float psum[kMaxNT][2];
#pragma unroll
        for (int nt = 0; nt < kMaxNT; ++nt) psum[nt][0] = psum[nt][1] = 0.0f;
        const int nbase = nhalf * (p.c2 >> 1);
        for (int mt = mgrp; mt < MT; mt += 4) {
            float acc[kMaxNT][4];
#pragma unroll
            for (int nt = 0; nt < kMaxNT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
            const int prow = min(mt * 16 + (lane & 15), HW - 1);
            const int py = prow / p.W, px = prow - py * p.W;
            const __nv_bfloat16* abase = act + (py * PW + px) * P1 + (lane >> 4) * 8;
            const __nv_bfloat16* bbase = w1s + (nbase + ((lane >> 4) & 1) * 8 + (lane & 7)) * KP + ((lane >> 3) & 1) * 8;
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                const __nv_bfloat16* arow = abase + (ky * PW + kx) * P1;
                for (int kc = 0; kc < p.c1; kc += 16) {
                    unsigned af[4];
                    ldsm_x4(af, arow + kc);
#pragma unroll
                    for (int j = 0; j < kMaxNT / 2; ++j) {
                        if (2 * j < NT) {
                            unsigned bf[4];
                            ldsm_x4(bf, bbase + (j * 16) * KP + tap * p.c1 + kc);
                            mma16816(acc[2 * j], af, bf[0], bf[1]);
                            mma16816(acc[2 * j + 1], af, bf[2], bf[3]);
                        }
                    }
                }
            }
            const bool ok0 = (mt * 16 + g) < HW, ok1 = (mt * 16 + g + 8) < HW;
#pragma unroll
            for (int nt = 0; nt < kMaxNT; ++nt) {
                if (nt < NT) {
                    const float bb0 = b1s[nbase + nt * 8 + tq * 2], bb1 = b1s[nbase + nt * 8 + tq * 2 + 1];
                    psum[nt][0] += (ok0 ? silu_e(acc[nt][0] + bb0) : 0.0f) + (ok1 ? silu_e(acc[nt][2] + bb0) : 0.0f);
                    psum[nt][1] += (ok0 ? silu_e(acc[nt][1] + bb1) : 0.0f) + (ok1 ? silu_e(acc[nt][3] + bb1) : 0.0f);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < kMaxNT; ++nt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v = psum[nt][j];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0 && nt < NT) part[mgrp * p.c2 + nbase + nt * 8 + tq * 2 + j] = v;
            }
        }
        __syncthreads();
        for (int c = tid; c < p.c2; c += 256)
            p.pooled[b * p.c2 + c] = (((part[c] + part[p.c2 + c]) + part[2 * p.c2 + c]) + part[3 * p.c2 + c]) / static_cast<float>(HW);
    }
}

}  // namespace convtc
}  // namespace idb200

using namespace idb200;

extern "C" int idb200_conv_encoder_tc(const float* occ, const float* sdf, int64_t B, int H, int W, int cin, int c1, int c2,
                                      const float* w0, const float* b0, const void* w1_packed_bf16, const float* b1,
                                      float* pooled, idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && H >= 1 && W >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(cin == 1 || (cin == 2 && sdf), IDB200_EINVAL, "use_sdf is True but sdf missing from cond");
    IDB_REQUIRE(c1 % 16 == 0 && c1 >= 16 && c1 <= 64, IDB200_EUNSUPPORTED, "C1 must be a multiple of 16 in [16, 64]");
    IDB_REQUIRE(c2 % 32 == 0 && c2 >= 32 && c2 <= 16 * convtc::kMaxNT, IDB200_EUNSUPPORTED, "C2 must be 32 or 64");
    if (B == 0) return IDB200_OK;
    IDB_REQUIRE(occ && w0 && b0 && w1_packed_bf16 && b1 && pooled, IDB200_EINVAL, "NULL pointer");
    const size_t PP = static_cast<size_t>(H + 2) * (W + 2);
    const size_t act = ((PP * (c1 + 8) + 7) & ~static_cast<size_t>(7)) * 2;
    const size_t w1s = static_cast<size_t>(c2) * (9 * c1 + 8) * 2;
    const size_t f32s = (((cin * PP + 3) & ~static_cast<size_t>(3)) + static_cast<size_t>(c1) * cin * 9 + c1 + c2 + 4 * c2) * 4;
    const size_t smem = act + w1s + f32s;
    IDB_REQUIRE(smem <= 227 * 1024, IDB200_EUNSUPPORTED, "maze %dx%d with channels (%d,%d) needs %zu bytes of shared memory", H, W, c1, c2, smem);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(convtc::conv2l_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        smem_set = smem;
    }
    const int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
    convtc::Params p{occ, sdf, w0, b0, static_cast<const __nv_bfloat16*>(w1_packed_bf16), b1, pooled, B, cin, c1, c2, H, W};
    const int grid = grid_for(B, 1, per_sm > 0 ? per_sm : 1);
    convtc::conv2l_tc_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("conv2l_tc_kernel");
}
