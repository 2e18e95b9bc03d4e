// K3d: fused transformer MLP  h[M,256] += W2 * SiLU(W1 * a + b1) + b2  in ONE kernel (d_model = 256).
//
// Reference: the `ff` branch of TransformerBlock.forward, src/models/transformer.py:43-45
// (`x = x + ff(film2(norm2(x)))`, ff = Linear(d, d_ff) -> SiLU -> Linear(d_ff, d)); `a` is the LN+FiLM output.
// The hidden activation [M, d_ff] never leaves the SM: it is 2/3 of the unfused path's HBM traffic.
//
// One CTA per SM, persistent over 128-row tiles.  Per tile the hidden dimension is processed in chunks of 128:
//   FF1(c): acc1[c&1] (TMEM, 128 cols)  = A[128x256] . W1[c*128.., :]^T          16 x tcgen05.mma N=128
//   EPI1(c): acc1 -> +b1 -> SiLU -> bf16 -> H[c&1] in shared memory (SWIZZLE_128B K-major, an MMA A operand)
//   FF2(c): acc2 (TMEM, 256 cols)      += H[c&1][128x128] . W2[:, c*128..]^T     16 x tcgen05.mma N=128
//   final:  h += acc2 + b2   (fp32 tile staged in the idle H buffers, TMA reduce-add into the residual stream)
// The MMA warp runs FF1(c+2) while the 8 epilogue warps do EPI1(c) (acc1 and H are double-buffered), weights stream
// through a 5-slot TMA ring of [128 x 64] bf16 tiles in exactly the order the MMA warp consumes them.
// TMEM: acc2 cols [0,256), acc1 cols [256,384) / [384,512).
#include "fused_common.cuh"

namespace idb200 {
using namespace tc;
using fused::kD;

namespace mlp {
constexpr int kCH = 128;                // hidden chunk
constexpr int kThreads = 384;           // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int kSlots = 5;
constexpr int kTile = 128 * 64 * 2;     // one [128 x 64] bf16 tile = 16 KB
constexpr int kOffA = 0;
constexpr int kOffH = 4 * kTile;
constexpr int kOffRing = kOffH + 4 * kTile;
constexpr int kOffBar = kOffRing + kSlots * kTile;
constexpr int kOffBias = kOffBar + 256;
constexpr int kMaxFF = 2048;
constexpr int kSmem = kOffBias + (kMaxFF + 3 * kD) * 4 + 1024;      // b1 | b2 | ln_w | ln_b

struct Params {
    const float* b1;     // [ff]
    const float* b2;     // [256]
    float* h;            // [M, 256] fp32 residual stream (in/out)
    long long M;
    int ff;
    // LayerNorm + FiLM prologue (kLN kernels): A = LN(h) * (1 + gamma) + beta is produced in shared memory by the compute warps
    const float* lnw;
    const float* lnb;
    const float* gb;
    long long gb_stride;
    int L;
};

__device__ __forceinline__ float silu_fast(float x) {
    const float hx = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}

template <bool kLN>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_h, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));             // keep LDS / STS (the integer round trip hides the address space)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* slot_full = bars + 2;                 // [kSlots]
    uint64_t* slot_empty = slot_full + kSlots;      // [kSlots]
    uint64_t* acc1_full = slot_empty + kSlots;      // [2]
    uint64_t* acc1_empty = acc1_full + 2;           // [2]
    uint64_t* h_full = acc1_empty + 2;              // [2]
    uint64_t* h_empty = h_full + 2;                 // [2]
    uint64_t* acc2_full = h_empty + 2;
    uint64_t* acc2_empty = acc2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
    float* sb1 = reinterpret_cast<float*>(smem + kOffBias);
    float* sb2 = sb1 + kMaxFF;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = p.ff / kCH;
    const long long tiles = (p.M + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w1);
        tma_prefetch_desc(&tmap_w2);
        tma_prefetch_desc(&tmap_h);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, kLN ? 8 : 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&acc1_empty[i], 8);
            mbar_init(&h_full[i], 8);
            mbar_init(&h_empty[i], 1);
        }
        mbar_init(acc2_full, 1);
        mbar_init(acc2_empty, 8);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.ff; i += kThreads) sb1[i] = p.b1[i];
    for (int i = threadIdx.x; i < kD; i += kThreads) sb2[i] = p.b2[i];
    float* slnw = sb2 + kD;
    float* slnb = slnw + kD;
    if (kLN)
        for (int i = threadIdx.x; i < kD; i += kThreads) { slnw[i] = p.lnw[i]; slnb[i] = p.lnb[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t sphase = 0;
            uint32_t tile_n = 0;
            auto load_w = [&](const CUtensorMap* m, int c0, int c1) {
                mbar_wait(&slot_empty[slot], sphase ^ 1, 10);
                mbar_arrive_expect_tx(&slot_full[slot], kTile);
                tma_load_2d(smem + kOffRing + slot * kTile, m, &slot_full[slot], c0, c1);
                if (++slot == kSlots) { slot = 0; sphase ^= 1; }
            };
            auto ff1 = [&](int c) { for (int kb = 0; kb < 4; ++kb) load_w(&tmap_w1, kb * 64, c * kCH); };
            auto ff2 = [&](int c) {
                for (int nh = 0; nh < 2; ++nh)
                    for (int kb = 0; kb < 2; ++kb) load_w(&tmap_w2, c * kCH + kb * 64, nh * 128);
            };
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
                const int m0 = static_cast<int>(tile) * 128;
                if (!kLN) {
                    mbar_wait(a_empty, (tile_n & 1) ^ 1, 11);
                    mbar_arrive_expect_tx(a_full, 4 * kTile);
                    for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + kOffA + kb * kTile, &tmap_a, a_full, kb * 64, m0);
                }
                ff1(0);
                if (nc > 1) ff1(1);
                for (int c = 0; c < nc; ++c) {
                    ff2(c);
                    if (c + 2 < nc) ff1(c + 2);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            int slot = 0;
            uint32_t sphase = 0;
            uint32_t tile_n = 0;
            uint32_t use1[2] = {0, 0};      // completed uses of acc1[b]
            uint32_t useh[2] = {0, 0};      // completed uses of H[b]
            const uint32_t sA = smem_u32(smem + kOffA), sH = smem_u32(smem + kOffH), sR = smem_u32(smem + kOffRing);
            auto ff1 = [&](int c, bool last) {
                const int b = c & 1;
                mbar_wait(&acc1_empty[b], (use1[b] & 1) ^ 1, 20);     // EPI1 drained acc1[b]
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + 256 + b * 128;
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(&slot_full[slot], sphase, 21);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(sA + kb * kTile);
                    const uint64_t bd = umma_desc_sw128(sR + slot * kTile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&slot_empty[slot]);
                    if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                }
                umma_commit(&acc1_full[b]);
                if (last) umma_commit(a_empty);                         // A tile no longer needed
                ++use1[b];
            };
            auto ff2 = [&](int c, bool last) {
                const int b = c & 1;
                mbar_wait(&h_full[b], useh[b] & 1, 22);                 // EPI1 wrote H[b]
                tc_fence_after();
                for (int nh = 0; nh < 2; ++nh) {
                    const uint32_t d_tmem = tmem_base + nh * 128;
                    for (int kb = 0; kb < 2; ++kb) {
                        mbar_wait(&slot_full[slot], sphase, 23);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(sH + (b * 2 + kb) * kTile);
                        const uint64_t bd = umma_desc_sw128(sR + slot * kTile);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (c | kb | k) ? 1u : 0u);
                        umma_commit(&slot_empty[slot]);
                        if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                    }
                }
                umma_commit(&h_empty[b]);
                if (last) umma_commit(acc2_full);
                ++useh[b];
            };
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
                mbar_wait(a_full, tile_n & 1, 24);
                tc_fence_after();
                ff1(0, nc == 1);
                if (nc > 1) ff1(1, nc == 2);
                for (int c = 0; c < nc; ++c) {
                    if (c == 0) {
                        mbar_wait(acc2_empty, (tile_n & 1) ^ 1, 25);    // previous tile's final epilogue drained acc2
                        tc_fence_after();
                    }
                    ff2(c, c == nc - 1);
                    if (c + 2 < nc) ff1(c + 2, c + 3 == nc);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int ew = warp - 4;
        const int q = ew & 3, half = ew >> 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        uint32_t use1[2] = {0, 0}, useh[2] = {0, 0};
        uint32_t tile_n = 0;
        if (kLN && static_cast<long long>(blockIdx.x) < tiles) {
            fused::ln_film_tile(p.h, static_cast<long long>(blockIdx.x) * 128, p.M, p.L, p.gb, p.gb_stride, slnw, slnb, smem + kOffA, ew, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
            for (int c = 0; c < nc; ++c) {
                const int b = c & 1;
                mbar_wait(&acc1_full[b], use1[b] & 1, 30);
                mbar_wait(&h_empty[b], (useh[b] & 1) ^ 1, 31);          // FF2 of the previous use finished reading H[b]
                tc_fence_after();
                uint8_t* hb = smem + kOffH + (b * 2 + half) * kTile;     // this half's 64 columns = one k-block
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + lane_base + 256 + b * 128 + half * 64 + cc * 32, r);
                    tmem_ld_wait();
                    const float* bb = sb1 + c * kCH + half * 64 + cc * 32;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                        // 8 columns -> one 16-byte swizzle chunk
                        const float4 b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                        const float4 b1v = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                        const float v0 = silu_fast(__uint_as_float(r[8 * j + 0]) + b0.x);
                        const float v1 = silu_fast(__uint_as_float(r[8 * j + 1]) + b0.y);
                        const float v2 = silu_fast(__uint_as_float(r[8 * j + 2]) + b0.z);
                        const float v3 = silu_fast(__uint_as_float(r[8 * j + 3]) + b0.w);
                        const float v4 = silu_fast(__uint_as_float(r[8 * j + 4]) + b1v.x);
                        const float v5 = silu_fast(__uint_as_float(r[8 * j + 5]) + b1v.y);
                        const float v6 = silu_fast(__uint_as_float(r[8 * j + 6]) + b1v.z);
                        const float v7 = silu_fast(__uint_as_float(r[8 * j + 7]) + b1v.w);
                        uint4 pk;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v4, v5), h3 = __floats2bfloat162_rn(v6, v7);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0);
                        pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2);
                        pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(hb + sw128_offset(row_in_tile, cc * 32 + j * 8)) = pk;
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();                                // H writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&acc1_empty[b]);
                    mbar_arrive(&h_full[b]);
                }
                ++use1[b];
                ++useh[b];
            }
            if (kLN) {
                // LayerNorm of the next tile before this tile's final epilogue (the last FF1 has been consumed by EPI1, so the
                // A buffer is free): FF1 of the next tile overlaps the epilogue
                const long long next = tile + gridDim.x;
                if (next < tiles) {
                    mbar_wait(a_empty, tile_n & 1, 33);
                    fused::ln_film_tile(p.h, next * 128, p.M, p.L, p.gb, p.gb_stride, slnw, slnb, smem + kOffA, ew, lane);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full);
                }
            }
            // final epilogue: h += acc2 + b2 (fp32 tile staged in the idle H buffers, TMA reduce-add: fused_common.cuh)
            mbar_wait(acc2_full, tile_n & 1, 32);
            tc_fence_after();
            fused::residual_epilogue(tmem_base, sb2, smem + kOffH, &tmap_h, static_cast<int>(tile) * 128, ew, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);
        }
        if (ew == 0 && lane == 0) tma_store_wait_all();                  // all residual updates landed before exit
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mlp

int mlp_fused(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, float* h, long long M, int d, int ff,
              const float* lnw, const float* lnb, const float* gb, long long gb_stride, int L, cudaStream_t st) {
    const bool ln = (A == nullptr);
    IDB_REQUIRE(d == kD, IDB200_EUNSUPPORTED, "fused MLP is specialised for d_model = 256 (got %d)", d);
    IDB_REQUIRE(ff % mlp::kCH == 0 && ff >= mlp::kCH && ff <= mlp::kMaxFF, IDB200_EUNSUPPORTED, "d_ff must be a multiple of 128, <= 2048");
    IDB_REQUIRE(M >= 0, IDB200_EINVAL, "bad shape");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(W1 && b1 && W2 && b2 && h, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(aligned(h, 16), IDB200_EALIGN, "h must be 16-byte aligned");
    if (ln) {
        IDB_REQUIRE(lnw && lnb, IDB200_EINVAL, "NULL LayerNorm parameters");
        IDB_REQUIRE(L >= 1 && M % L == 0, IDB200_EINVAL, "M must be a multiple of L");
        IDB_REQUIRE(L >= 8 ? (L % 8 == 0) : (8 % L == 0), IDB200_EUNSUPPORTED, "fused MLP block needs L | 8 or 8 | L (got %d)", L);
        IDB_REQUIRE(!gb || (aligned(gb, 16) && gb_stride % 4 == 0), IDB200_EALIGN, "gamma_beta must be 16-byte aligned");
    }
    CUtensorMap ta, t1, t2, th;
    int rc = make_tmap_bf16_2d(&ta, ln ? W1 : A, ln ? static_cast<uint64_t>(ff) : static_cast<uint64_t>(M), d, 128, 64);   // unused when ln
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&t1, W1, ff, d, 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&t2, W2, d, ff, 128, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&th, h, 4, static_cast<uint64_t>(M), d, 128, 32);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(mlp::mlp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::kSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp::mlp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", mlp::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    const long long tiles = (M + 127) / 128;
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    mlp::Params p{b1, b2, h, M, ff, lnw, lnb, gb, gb_stride, L};
    if (ln) mlp::mlp_fused_kernel<true><<<grid, mlp::kThreads, mlp::kSmem, st>>>(ta, t1, t2, th, p);
    else mlp::mlp_fused_kernel<false><<<grid, mlp::kThreads, mlp::kSmem, st>>>(ta, t1, t2, th, p);
    return check_launch("mlp_fused_kernel");
}

}  // namespace idb200

extern "C" int idb200_mlp_fused(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, float* h,
                                int64_t M, int d, int ff, idb200_stream_t stream) {
    IDB_REQUIRE(A != nullptr, IDB200_EINVAL, "NULL pointer");
    return idb200::mlp_fused(A, W1, b1, W2, b2, h, M, d, ff, nullptr, nullptr, nullptr, 0, 1, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_mlp_block(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                                const void* W1, const float* b1, const void* W2, const float* b2, int64_t M, int L, int d, int ff,
                                idb200_stream_t stream) {
    return idb200::mlp_fused(nullptr, W1, b1, W2, b2, h, M, d, ff, ln_w, ln_b, gamma_beta, gb_stride, L, static_cast<cudaStream_t>(stream));
}
