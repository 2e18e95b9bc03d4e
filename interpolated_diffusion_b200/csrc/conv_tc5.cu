// K4': two-layer maze conv encoder with the second conv on the 5th-generation tensor cores (tcgen05, TMEM accumulators):
//   occ (+sdf) [B,Cin,H,W] -> conv3x3(Cin->32)+SiLU -> conv3x3(32->64)+SiLU -> mean over H x W -> pooled [B,64]
//
// Reference: MazeEncoder.forward, src/models/encoders.py:22-25 with the default maze_channels=(32, 64).
//
// Implicit GEMM without im2col copies.  Output pixels are indexed over PADDED rows, m = y * PW + x with PW = W + 2
// rounded up to a multiple of 8 (24 for W = 21; columns x >= W are junk and masked in the epilogue), so that the A tile
// of output tile t and tap (ky, kx) is a CONTIGUOUS run of 128 activation rows:
//     A[m, :] = act_pad[m + ky * PW + kx, 0:32]          (act_pad = zero-bordered channels-last first-layer output)
// A UMMA descriptor cannot start at an arbitrary row, only at a multiple of the 8-row swizzle atom, so the first conv
// writes its output three times, pre-shifted by kx = 0, 1, 2 (copy_kx[r] = act_pad[r + kx]); the remaining shift
// ky * PW is a multiple of 8 rows by construction of PW.  Rows are 64 bytes (32 bf16 channels) in the SWIZZLE_64B
// K-major layout; the nine weight taps [64 x 32] stay resident in the same layout.  Per maze: 4 tiles x 9 taps x 2
// tcgen05.mma (M=128, N=64, K=16), accumulators double-buffered in TMEM (2 x 4 x 64 columns) so the epilogue of maze
// i (bias + SiLU + masked column sums -> mean) overlaps the first conv and the MMAs of maze i+1.
// The single activation buffer is handed over PER TILE: the first conv walks the pixels in order, signals tile t as soon
// as every activation row that tile reads is written (act_full[t]) and overwrites a band of rows as soon as the last MMA
// tile of the previous maze that reads it has completed (tile_done[t]) -- so the first conv of maze i+1 runs under the
// MMAs of maze i instead of after them (a per-maze hand-over serialised the two: 15.3 k cycles per maze).
// Warps: 0..3 epilogue (TMEM lane quadrants), 4 MMA issuer + TMEM allocator, 5..12 first conv (CUDA cores).
// One persistent CTA per SM.  Requires C1 = 32, C2 = 64, H * PW <= 512 (other shapes: idb200_conv_encoder_tc).
#include <cuda_bf16.h>

#include "tc_common.cuh"

#ifndef IDB200_CONV_IPT
#define IDB200_CONV_IPT 2
#endif

namespace idb200 {
using namespace tc;

namespace conv5 {
constexpr int kThreads = 13 * 32;
constexpr int kC1 = 32, kC2 = 64;
constexpr int kTiles = 4;                                  // M tiles of 128 output positions
constexpr int kMaxPlane = 2 * 26 * 26;                     // floats of one zero-bordered input buffer (cin <= 2, H, W <= 24)

struct Params {
    const float* occ;              // [B,1,H,W]
    const float* sdf;              // [B,1,H,W] or nullptr
    const float* w0;               // [32, Cin, 3, 3] fp32
    const float* b0;               // [32]
    const __nv_bfloat16* w1;       // [64, 9*32] bf16, k = (ky*3+kx)*32 + c
    const float* b1;               // [64]
    float* pooled;                 // [B, 64]
    long long B;
    int cin, H, W, PW, act_rows;   // act_rows: rows of one shifted copy (multiple of 8, >= 512 + 2 * PW)
};

// Shared-memory matrix descriptor, K-major, SWIZZLE_64B: rows of 32 bf16 (64 bytes), 8-row atoms of 512 bytes.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(512 >> 4) << 32) | (1ull << 46) |
           (4ull << 61);
}
// byte offset of the 16-byte chunk `ch` (8 channels) of row `row` in a SWIZZLE_64B matrix of 64-byte rows:
// address bits [4,6) ^= bits [7,9)
__device__ __forceinline__ uint32_t sw64_offset(int row, int ch) { return static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ float silu_t(float x) {
    const float hx = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}
__device__ __forceinline__ unsigned pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1) conv2l_tc5_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));
    const int act_bytes = p.act_rows * 64;
    uint8_t* act = smem;                                              // [3][act_rows][64 B]
    uint8_t* wB = act + 3 * act_bytes;                                // [9][64 x 64 B]  (act_bytes is a multiple of 512)
    float* plane = reinterpret_cast<float*>(wB + 9 * 4096);           // [2][cin][(H+2)*(W+2)]
    float* w0s = plane + 2 * kMaxPlane;                               // [cin*9][32]
    float* b0s = w0s + 2 * 9 * kC1;
    float* b1s = b0s + kC1;
    float* red = b1s + kC2;                                           // [4][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4 * kC2);
    uint64_t* act_full = bars + 0;                                    // [4] conv warps -> MMA: the rows tile t reads are written (8 arrivals)
    uint64_t* tile_done = bars + 4;                                   // [4] MMA commit -> conv warps: tile t has read its rows
    uint64_t* acc_full = bars + 8;                                    // [2] MMA commit -> epilogue
    uint64_t* acc_empty = bars + 10;                                  // [2] epilogue -> MMA         (4 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = p.H * p.W, PWi = p.W + 2, PP = (p.H + 2) * PWi;

    // ---- one-time setup ----
    for (int i = tid; i < 3 * act_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * kMaxPlane; i += kThreads) plane[i] = 0.0f;
    for (int i = tid; i < kC1 * p.cin * 9; i += kThreads) {             // w0 [ch][c][tap] -> w0s[(c*9+tap)*32 + ch]
        const int ch = i / (p.cin * 9), r = i - ch * p.cin * 9;
        w0s[r * kC1 + ch] = p.w0[i];
    }
    for (int i = tid; i < kC1; i += kThreads) b0s[i] = p.b0[i];
    for (int i = tid; i < kC2; i += kThreads) b1s[i] = p.b1[i];
    for (int i = tid; i < 9 * kC2 * 4; i += kThreads) {                // weight taps -> SWIZZLE_64B [64 x 32] tiles, 16-byte chunks
        const int tap = i / (kC2 * 4), r = i - tap * kC2 * 4, n = r >> 2, ch = r & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(p.w1 + n * (9 * kC1) + tap * kC1 + ch * 8);
        *reinterpret_cast<uint4*>(wB + tap * 4096 + sw64_offset(n, ch)) = v;
    }
    if (warp == 4 && lane == 0) {
        for (int i = 0; i < kTiles; ++i) { mbar_init(&act_full[i], 8); mbar_init(&tile_done[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();                                           // weight tiles / zeroed activations -> async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 5) {
        // ===================== first conv (8 warps, CUDA cores) =====================
        const int ct = tid - 5 * 32;                                    // 0..255
        auto load_plane = [&](long long b, int buf) {                   // cp.async: next maze's input lands while this one is computed
            float* dst = plane + buf * kMaxPlane;
            for (int i = ct; i < p.cin * HW; i += 256) {
                const int c = i / HW, r = i - c * HW, y = r / p.W, x = r - y * p.W;
                const float* src = ((c == 0) ? p.occ : p.sdf) + b * HW + r;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + c * PP + (y + 1) * PWi + x + 1)), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int o = ct & 3;                                            // this thread's channel octet (256 % 4 == 0: fixed)
        constexpr int kIPT = IDB200_CONV_IPT;                                          // items per thread and pass: two independent FMA / SiLU chains in flight
        constexpr int kPassPix = 64 * kIPT;                              // a pass = kPassPix consecutive pixels x 4 octets
        const int n_items = HW * 4, n_iters = (n_items + 256 * kIPT - 1) / (256 * kIPT);     // <= 5 (H * W <= 576)
        // Per pass, packed 4 bits each: low 2 bits = the last MMA tile of the PREVIOUS maze that reads a row this pass writes (tiles
        // complete in order; tile t reads padded positions t * 128 .. t * 128 + 129 + 2 PW), and one nibble of tile bits = the tiles
        // whose rows are complete after this pass.
        uint32_t need_bits = 0, sig_bits = 0;
        for (int i = 0; i < n_iters; ++i) {
            const int pmax = min(i * kPassPix + kPassPix - 1, HW - 1);
            need_bits |= static_cast<uint32_t>(min(kTiles - 1, ((pmax / p.W + 1) * p.PW + p.W) >> 7)) << (4 * i);
        }
#pragma unroll
        for (int t = 0; t < kTiles; ++t) {
            const int ylast = min(p.H - 1, (t * 128 + 128 + 2 * p.PW) / p.PW - 1);
            const int last_i = ylast < 0 ? 0 : (ylast * p.W + p.W - 1) / kPassPix;
            sig_bits |= 1u << (4 * last_i + t);
        }
        const float inv_w = 1.0f / static_cast<float>(p.W);             // pix / W for pix < 1024: (pix + 0.5) * (1 / W) truncates exactly
        // cin == 1 (no SDF channel): this thread's 9 x 8 first-layer weights live in registers (18 of the 27 shared-memory loads per item)
        const bool kRegW = p.cin == 1;
        float wr[9][8], br[8];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[tap][j] = kRegW ? w0s[tap * kC1 + o * 8 + j] : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) br[j] = b0s[o * 8 + j];
        long long it = 0;
        if (static_cast<long long>(blockIdx.x) < p.B) load_plane(blockIdx.x, 0);
        for (long long b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
            const int buf = static_cast<int>(it & 1);
            asm volatile("cp.async.wait_group 0;" ::: "memory");        // this thread's part of this maze's plane has landed
            named_barrier_sync(1, 256);                                 // ... and everybody's; every warp is done reading the other buffer
            if (b + gridDim.x < p.B) load_plane(b + gridDim.x, buf ^ 1);  // prefetch the next maze while this one is computed
            const float* pl = plane + buf * kMaxPlane;
            const uint32_t prev = static_cast<uint32_t>((it & 1) ^ 1);    // parity of the previous maze's phases
            // item = (pixel, channel octet): 8 channels of one pixel -> one 16-byte chunk, written to the 3 shifted copies
#pragma unroll 1
            for (int iter = 0; iter < n_iters; ++iter) {
                mbar_wait(&tile_done[(need_bits >> (4 * iter)) & 3u], prev, 70);
                float a[kIPT][8];
                int q[kIPT];
                bool valid[kIPT];
#pragma unroll
                for (int u = 0; u < kIPT; ++u) {
                    const int item = ct + ((iter * kIPT + u) << 8);
                    valid[u] = item < n_items;
                    const int pix = min(item >> 2, HW - 1);
                    const int y = static_cast<int>((static_cast<float>(pix) + 0.5f) * inv_w), x = pix - y * p.W;
                    q[u] = (y + 1) * p.PW + (x + 1);                    // padded linear position of this pixel
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[u][j] = br[j];
                    if (kRegW) {
                        const float* ip = pl + y * PWi + x;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const float iv = ip[(tap / 3) * PWi + (tap % 3)];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[u][j] = fmaf(iv, wr[tap][j], a[u][j]);
                        }
                    } else {
                        for (int c = 0; c < p.cin; ++c) {
                            const float* ip = pl + c * PP + y * PWi + x;
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const float iv = ip[(tap / 3) * PWi + (tap % 3)];
                                const float4 wa = *reinterpret_cast<const float4*>(w0s + (c * 9 + tap) * kC1 + o * 8);
                                const float4 wb = *reinterpret_cast<const float4*>(w0s + (c * 9 + tap) * kC1 + o * 8 + 4);
                                a[u][0] = fmaf(iv, wa.x, a[u][0]); a[u][1] = fmaf(iv, wa.y, a[u][1]); a[u][2] = fmaf(iv, wa.z, a[u][2]); a[u][3] = fmaf(iv, wa.w, a[u][3]);
                                a[u][4] = fmaf(iv, wb.x, a[u][4]); a[u][5] = fmaf(iv, wb.y, a[u][5]); a[u][6] = fmaf(iv, wb.z, a[u][6]); a[u][7] = fmaf(iv, wb.w, a[u][7]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kIPT; ++u) {
                    uint4 pk;
                    pk.x = pack2(silu_t(a[u][0]), silu_t(a[u][1]));
                    pk.y = pack2(silu_t(a[u][2]), silu_t(a[u][3]));
                    pk.z = pack2(silu_t(a[u][4]), silu_t(a[u][5]));
                    pk.w = pack2(silu_t(a[u][6]), silu_t(a[u][7]));
                    if (valid[u]) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) *reinterpret_cast<uint4*>(act + kx * act_bytes + sw64_offset(q[u] - kx, o)) = pk;
                    }
                }
                const uint32_t sig = (sig_bits >> (4 * iter)) & 15u;     // (warp-uniform)
                if (sig) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int t = 0; t < kTiles; ++t) {
                            if ((sig >> t) & 1u) {
                                mbar_wait(&tile_done[t], prev, 74);      // the MMA warp has consumed this barrier's previous phase
                                mbar_arrive(&act_full[t]);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer: warp-uniform schedule, one elected lane issues =====================
        {
            constexpr uint32_t idesc = umma_idesc_bf16(128, kC2);
            const uint32_t sA = smem_u32(act), sW = smem_u32(wB);
            long long it = 0;
            for (long long b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
                const int buf = static_cast<int>(it & 1);
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1, 71);    // the epilogue drained this accumulator buffer
#pragma unroll 1
                for (int t = 0; t < kTiles; ++t) {
                    mbar_wait(&act_full[t], it & 1, 72);                 // the rows this tile reads are written
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * (kTiles * kC2) + t * kC2;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint64_t ad = umma_desc_sw64(sA + kx * act_bytes + (t * 128 + ky * p.PW) * 64);
                        const uint64_t bd = umma_desc_sw64(sW + tap * 4096);
                        if (elect_one_sync()) {
                            umma_bf16(d, ad, bd, idesc, tap ? 1u : 0u);
                            umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                        }
                        __syncwarp();
                    }
                    if (elect_one_sync()) umma_commit(&tile_done[t]);
                    __syncwarp();
                }
                if (elect_one_sync()) umma_commit(&acc_full[buf]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (4 warps: thread <-> accumulator row) =====================
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        long long it = 0;
        for (long long b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
            const int buf = static_cast<int>(it & 1);
            mbar_wait(&acc_full[buf], (it >> 1) & 1, 73);
            tc_fence_after();
            float sum[kC2];
#pragma unroll
            for (int c = 0; c < kC2; ++c) sum[c] = 0.0f;
#pragma unroll 1
            for (int t = 0; t < kTiles; ++t) {
                const int m = t * 128 + warp * 32 + lane;
                const bool valid = (m < p.H * p.PW) && ((m % p.PW) < p.W);
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(tmem_base + lane_base + buf * (kTiles * kC2) + t * kC2, r0);
                tmem_ld_32x32(tmem_base + lane_base + buf * (kTiles * kC2) + t * kC2 + 32, r1);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        sum[c] += silu_t(__uint_as_float(r0[c]) + b1s[c]);
                        sum[32 + c] += silu_t(__uint_as_float(r1[c]) + b1s[32 + c]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            // column sums over the warp's 32 rows: transpose-reduce (lane bit `off` keeps the upper half of the columns)
            int n = kC2;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const int half = n >> 1;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const float send = up ? sum[i] : sum[i + half];
                    const float keep = up ? sum[i + half] : sum[i];
                    sum[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
                n = half;
            }
            const int col = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
            named_barrier_sync(2, 128);                                 // the previous maze's partial sums have been consumed
            red[warp * kC2 + col] = sum[0];
            red[warp * kC2 + col + 1] = sum[1];
            named_barrier_sync(2, 128);
            if (tid < kC2)
                p.pooled[b * kC2 + tid] = (((red[tid] + red[kC2 + tid]) + red[2 * kC2 + tid]) + red[3 * kC2 + tid]) / static_cast<float>(HW);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace conv5
}  // namespace idb200

using namespace idb200;

extern "C" int idb200_conv_encoder_tc5(const float* occ, const float* sdf, int64_t B, int H, int W, int cin, int c1, int c2,
                                       const float* w0, const float* b0, const void* w1_packed_bf16, const float* b1,
                                       float* pooled, idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && H >= 1 && W >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(cin == 1 || (cin == 2 && sdf), IDB200_EINVAL, "use_sdf is True but sdf missing from cond");
    IDB_REQUIRE(c1 == conv5::kC1 && c2 == conv5::kC2, IDB200_EUNSUPPORTED, "the tcgen05 conv encoder is specialised for maze_channels = (32, 64)");
    const int PW = (W + 2 + 7) & ~7;
    IDB_REQUIRE(H * PW <= conv5::kTiles * 128 && cin * (H + 2) * (W + 2) <= conv5::kMaxPlane, IDB200_EUNSUPPORTED,
                "maze %dx%d does not fit four 128-position tiles", H, W);
    if (B == 0) return IDB200_OK;
    IDB_REQUIRE(occ && w0 && b0 && w1_packed_bf16 && b1 && pooled, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(aligned(w1_packed_bf16, 16), IDB200_EALIGN, "packed weights must be 16-byte aligned");
    const int act_rows = conv5::kTiles * 128 + 2 * PW + 8;              // last tile + largest tap shift (a multiple of 8)
    const size_t smem = 3 * static_cast<size_t>(act_rows) * 64 + 9 * 4096 +
                        (2 * conv5::kMaxPlane + 2 * 9 * conv5::kC1 + conv5::kC1 + conv5::kC2 + 4 * conv5::kC2) * 4 + 128 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv5::conv2l_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    IDB_REQUIRE(smem <= 200 * 1024, IDB200_EUNSUPPORTED, "shared memory");
    conv5::Params p{occ, sdf, w0, b0, static_cast<const __nv_bfloat16*>(w1_packed_bf16), b1, pooled, B, cin, H, W, PW, act_rows};
    const int grid = static_cast<int>(B < num_sms() ? B : num_sms());
    conv5::conv2l_tc5_kernel<<<grid, conv5::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("conv2l_tc5_kernel");
}
