// First layer and pooling of the implicit-GEMM conv stack (the tap-shifted tcgen05 GEMM itself is idb200_conv3x3_gemm in gemm.cu).
//
// Reference: MazeEncoder.forward, src/models/encoders.py:15-25 (conv3x3 + SiLU stack, mean over H x W, fc).  Deep / wide stacks
// (trainer default maze_channels = 32,64,128,128, src/train/train_interp_levels.py:62) used to run as im2col + GEMM: the patch
// matrix is 9x the activation (30 GB per 16384 mazes) and its write + re-read was 12 % of a large-model generation.  Here the
// activations live in a zero-bordered NHWC layout [B, (H+2)*(W+2), C] so that the input pixel of tap (ky, kx) is a constant ROW
// SHIFT of the output position, which a TMA box coordinate expresses directly.
#include <cuda_bf16.h>

#include "common.cuh"

namespace idb200 {
namespace ci {

// conv3x3(C_in <= 2 -> C1 <= 64) + bias + SiLU on CUDA cores (9 * C_in MACs per output: 0.1 % of the stack's FLOPs), output in the
// zero-bordered NHWC bf16 layout.  One thread per padded position, all C1 channels (weights / bias in shared memory).
__global__ void __launch_bounds__(256) conv_first_nhwc_kernel(const float* __restrict__ x, long long B, int Cin, int H, int W,
                                                              const float* __restrict__ w, const float* __restrict__ bias, int C1,
                                                              __nv_bfloat16* __restrict__ out) {
    extern __shared__ float sw[];                                  // [C1][Cin * 9] | bias [C1]
    const int nw = C1 * Cin * 9;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sw[nw + i] = bias[i];
    __syncthreads();
    const int PW = W + 2, PH = H + 2, P2 = PW * PH;
    const long long total = B * P2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / P2;
        const int pos = static_cast<int>(i - b * P2);
        const int y = pos / PW - 1, xx = pos % PW - 1;
        uint4* dst = reinterpret_cast<uint4*>(out + i * C1);
        if (y < 0 || y >= H || xx < 0 || xx >= W) {
            for (int c = 0; c < C1 / 8; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        float in[18];
        for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = y + t / 3 - 1, xs = xx + t % 3 - 1;
                in[ci * 9 + t] = (yy >= 0 && yy < H && xs >= 0 && xs < W) ? x[((b * Cin + ci) * H + yy) * W + xs] : 0.0f;
            }
        for (int c8 = 0; c8 < C1 / 8; ++c8) {
            uint32_t pk[4];
#pragma unroll
            for (int h2 = 0; h2 < 4; ++h2) {
                float o[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c = c8 * 8 + h2 * 2 + u;
                    float acc = sw[nw + c];
                    for (int k = 0; k < Cin * 9; ++k) acc = fmaf(in[k], sw[c * Cin * 9 + k], acc);
                    o[u] = acc / (1.0f + __expf(-acc));            // SiLU
                }
                __nv_bfloat162 b2 = __floats2bfloat162_rn(o[0], o[1]);
                pk[h2] = *reinterpret_cast<uint32_t*>(&b2);
            }
            dst[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// pooled[b, c] = (1 / (H W)) * sum over the padded positions of act[b, :, c] (border rows are zero): one warp per (b, 64-channel slab)
__global__ void __launch_bounds__(256) pool_bordered_kernel(const __nv_bfloat16* __restrict__ act, long long B, int P2, int C, float inv_hw,
                                                            float* __restrict__ pooled) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int slabs = C / 64;
    for (long long u = warp; u < B * slabs; u += nwarps) {
        const long long b = u / slabs;
        const int c0 = static_cast<int>(u - b * slabs) * 64 + 2 * lane;
        const __nv_bfloat16* src = act + b * P2 * C + c0;
        float s0 = 0.0f, s1 = 0.0f;
        for (int r = 0; r < P2; ++r) {                              // fixed order: deterministic
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + static_cast<long long>(r) * C);
            s0 += __bfloat162float(v.x);
            s1 += __bfloat162float(v.y);
        }
        pooled[b * C + c0] = s0 * inv_hw;
        pooled[b * C + c0 + 1] = s1 * inv_hw;
    }
}

}  // namespace ci
}  // namespace idb200

using namespace idb200;

extern "C" int idb200_conv_first_nhwc(const float* x, int64_t B, int Cin, int H, int W, const float* w, const float* bias, int C1, void* out,
                                      idb200_stream_t stream) {
    IDB_REQUIRE(x && w && bias && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && Cin >= 1 && Cin <= 2 && C1 >= 8 && C1 % 8 == 0 && C1 <= 64, IDB200_EUNSUPPORTED, "first conv: C_in <= 2, C_out a multiple of 8 up to 64");
    if (B == 0) return IDB200_OK;
    const size_t smem = static_cast<size_t>(C1) * (Cin * 9 + 1) * sizeof(float);
    const long long total = B * (H + 2) * (W + 2);
    ci::conv_first_nhwc_kernel<<<grid_for(total, 256, 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(x, B, Cin, H, W, w, bias, C1,
                                                                                                          static_cast<__nv_bfloat16*>(out));
    return check_launch("conv_first_nhwc_kernel");
}

extern "C" int idb200_pool_bordered(const void* act, int64_t B, int H, int W, int C, float* pooled, idb200_stream_t stream) {
    IDB_REQUIRE(act && pooled, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && C % 64 == 0, IDB200_EUNSUPPORTED, "pooling needs C %% 64 == 0");
    if (B == 0) return IDB200_OK;
    const int P2 = (H + 2) * (W + 2);
    ci::pool_bordered_kernel<<<grid_for(B * (C / 64), 8, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), B, P2, C, 1.0f / static_cast<float>(H * W), pooled);
    return check_launch("pool_bordered_kernel");
}
