// K3e: fused attention half of a transformer block (d_model = 256, 8 heads of 32) in ONE kernel:
//
//     h[M,256] += out_proj( MHA( LN1(h) * (1 + gamma) + beta ) )
//
// Reference: src/models/transformer.py:35-41 (`x = x + attn(film1(norm1(x)))`, nn.MultiheadAttention with packed
// in_proj, optional causal mask of :68-71).  The unfused path is 4 launches (ln_film, QKV GEMM, attention, out_proj
// GEMM) that move 6 KB per token through HBM; here a token costs one fp32 read and one fp32 reduce of h (2 KB).
//
// One CTA per SM, persistent over 128-token tiles (a tile holds 128 / L whole trajectories, so attention never
// leaves the tile).  Warp roles: 0 TMA producer (weights), 1 tcgen05.mma issuer, 2 TMEM allocator, 4..11 compute.
// Per tile:
//   LN      compute warps: LayerNorm + FiLM of the h rows -> X, bf16 K-major SWIZZLE_128B A operand (64 KB)
//   for head group g = 0..3 (2 heads = 64 q + 64 k + 64 v output features, weights pre-packed group-major):
//     GEMM_g  acc[128 x 192] (TMEM cols 0..191) = X . Wqkv_g^T           16 x tcgen05.mma N=192
//     EPI_g   compute warps: acc + bias -> bf16 q|k|v rows in shared memory (padded pitch, ldmatrix-friendly)
//     ATT_g   compute warps: softmax(q k^T / sqrt(32)) v per (16-row block, head) with mma.sync + ldmatrix
//             (block-diagonal over the trajectories of the tile; causal optional) -> O_g, bf16 SWIZZLE_128B [128 x 64]
//     OUT_g   out[128 x 256] (TMEM cols 256..511) += O_g . Wo[:, 64g..64g+63]^T     4 x (N=192 + N=64) tcgen05.mma
//   final   h += out + b_o  (fp32 tile staged in shared memory, TMA reduce-add)
// GEMM_{g+1} runs on the tensor pipe while the compute warps do ATT_g; the LayerNorm of the next tile is done before
// the final epilogue of the current one so the next GEMM_0 overlaps it.  Weight tiles stream through a 3-slot TMA
// ring ([192 x 64] bf16) in exactly the order the MMA warp consumes them.
#include <cstdlib>

#include "fused_common.cuh"

namespace idb200 {
using namespace tc;
using namespace fused;

namespace ab {
constexpr int kThreads = 384;
constexpr int kSlots = 3;
constexpr int kSlotBytes = 192 * 64 * 2;            // 24 KB
constexpr int kOffX = 0;                            // 4 x [128 x 64] bf16
constexpr int kOffQkv = 4 * kTile;                  // 128 x 400 B = 51200
constexpr int kOffO = kOffQkv + 128 * kPitch * 2;   // [128 x 64] bf16 SW128 (16 KB); qkv + O = 66 KB: epilogue staging
constexpr int kOffRing = kOffO + kTile;
constexpr int kOffBar = kOffRing + kSlots * kSlotBytes;
constexpr int kOffBias = kOffBar + 256;             // bqkv 768 | bo 256 | ln_w 256 | ln_b 256 floats
constexpr int kSmem = kOffBias + (768 + 3 * 256) * 4 + 1024;
static_assert(kOffO % 1024 == 0 && kOffRing % 1024 == 0 && kOffQkv % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(kSmem <= 232448, "shared memory budget");

struct Params {
    float* h;               // [M, 256] fp32 residual stream (in/out)
    const float* lnw;
    const float* lnb;
    const float* gb;        // FiLM [B, >= 512] rows = [gamma | beta], or nullptr
    long long gb_stride;
    const float* bqkv;      // [768] group-major (same order as the packed weight rows)
    const float* bo;        // [256]
    long long M;
    int L;
    int causal;
    int dbg;                // dev ablation flags (IDB200_DBG): 1 skip LN, 2 skip attention core, 4 skip EPI, 8 skip residual epilogue
};

__global__ void __launch_bounds__(kThreads, 1)
attn_block_kernel(const __grid_constant__ CUtensorMap tmap_wqkv, const __grid_constant__ CUtensorMap tmap_wo_a,
                  const __grid_constant__ CUtensorMap tmap_wo_b, const __grid_constant__ CUtensorMap tmap_h, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));             // keep LDS / STS (the integer round trip hides the address space)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* x_full = bars + 0;
    uint64_t* x_empty = bars + 1;
    uint64_t* slot_full = bars + 2;                 // [kSlots]
    uint64_t* slot_empty = slot_full + kSlots;      // [kSlots]
    uint64_t* acc_full = slot_empty + kSlots;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* o_full = acc_empty + 1;
    uint64_t* o_empty = o_full + 1;
    uint64_t* out_full = o_empty + 1;
    uint64_t* out_empty = out_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + 1);
    float* sbqkv = reinterpret_cast<float*>(smem + kOffBias);
    float* sbo = sbqkv + 768;
    float* slnw = sbo + 256;
    float* slnb = slnw + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles = (p.M + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_wqkv);
        tma_prefetch_desc(&tmap_wo_a);
        tma_prefetch_desc(&tmap_wo_b);
        tma_prefetch_desc(&tmap_h);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, 8);
        mbar_init(x_empty, 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        mbar_init(o_full, 8);
        mbar_init(o_empty, 1);
        mbar_init(out_full, 1);
        mbar_init(out_empty, 8);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 768; i += kThreads) sbqkv[i] = p.bqkv[i];
    for (int i = threadIdx.x; i < 256; i += kThreads) { sbo[i] = p.bo[i]; slnw[i] = p.lnw[i]; slnb[i] = p.lnb[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc = tmem_base;            // q|k|v of one head group: 192 columns
    const uint32_t tmem_out = tmem_base + 256;      // out_proj accumulator: 256 columns

    if (warp == 0) {
        // ===================== TMA producer (weights) =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t sphase = 0;
            auto load = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes) {
                mbar_wait(&slot_empty[slot], sphase ^ 1, 10);
                mbar_arrive_expect_tx(&slot_full[slot], bytes);
                tma_load_2d(smem + kOffRing + slot * kSlotBytes, m, &slot_full[slot], c0, c1);
                if (++slot == kSlots) { slot = 0; sphase ^= 1; }
            };
            auto qkv = [&](int g) { for (int kb = 0; kb < 4; ++kb) load(&tmap_wqkv, kb * 64, g * 192, kSlotBytes); };
            auto wo = [&](int g) {
                load(&tmap_wo_a, g * 64, 0, 192 * 64 * 2);
                load(&tmap_wo_b, g * 64, 192, 64 * 64 * 2);
            };
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                qkv(0); qkv(1); wo(0); qkv(2); wo(1); qkv(3); wo(2); wo(3);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc192 = umma_idesc_bf16(128, 192);
            constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64);
            int slot = 0;
            uint32_t sphase = 0, n_acc = 0, n_o = 0, tile_n = 0;
            const uint32_t sX = smem_u32(smem + kOffX), sO = smem_u32(smem + kOffO), sR = smem_u32(smem + kOffRing);
            auto gemm = [&](int g) {
                mbar_wait(acc_empty, (n_acc & 1) ^ 1, 20);              // EPI of the previous group drained the accumulator
                tc_fence_after();
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(&slot_full[slot], sphase, 21);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(sX + kb * kTile);
                    const uint64_t bd = umma_desc_sw128(sR + slot * kSlotBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_acc, ad + 2 * k, bd + 2 * k, idesc192, (kb | k) ? 1u : 0u);
                    umma_commit(&slot_empty[slot]);
                    if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                }
                umma_commit(acc_full);
                if (g == 3) umma_commit(x_empty);                       // X may be overwritten by the next tile's LayerNorm
                ++n_acc;
            };
            auto outp = [&](int g) {
                mbar_wait(o_full, n_o & 1, 22);                          // ATT_g wrote O_g
                if (g == 0) mbar_wait(out_empty, (tile_n & 1) ^ 1, 23);  // previous tile's epilogue drained the out accumulator
                tc_fence_after();
                const uint64_t ad = umma_desc_sw128(sO);
                {
                    mbar_wait(&slot_full[slot], sphase, 24);
                    tc_fence_after();
                    const uint64_t bd = umma_desc_sw128(sR + slot * kSlotBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_out, ad + 2 * k, bd + 2 * k, idesc192, (g | k) ? 1u : 0u);
                    umma_commit(&slot_empty[slot]);
                    if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                }
                {
                    mbar_wait(&slot_full[slot], sphase, 25);
                    tc_fence_after();
                    const uint64_t bd = umma_desc_sw128(sR + slot * kSlotBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_out + 192, ad + 2 * k, bd + 2 * k, idesc64, (g | k) ? 1u : 0u);
                    umma_commit(&slot_empty[slot]);
                    if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                }
                umma_commit(o_empty);
                if (g == 3) umma_commit(out_full);
                ++n_o;
            };
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
                mbar_wait(x_full, tile_n & 1, 26);
                tc_fence_after();
                gemm(0); gemm(1); outp(0); gemm(2); outp(1); gemm(3); outp(2); outp(3);
            }
        }
    } else if (warp >= 4) {
        // ===================== compute warps =====================
        const int ew = warp - 4;
        const int q = ew & 3, half = ew >> 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const __nv_bfloat16* sq = reinterpret_cast<const __nv_bfloat16*>(smem + kOffQkv);
        uint8_t* sqb = smem + kOffQkv;
        uint8_t* so = smem + kOffO;
        const int L = p.L;
        uint32_t n_acc = 0, n_o = 0, tile_n = 0;
        if (static_cast<long long>(blockIdx.x) < tiles) {
            if (!(p.dbg & 1)) ln_film_tile(p.h, static_cast<long long>(blockIdx.x) * 128, p.M, L, p.gb, p.gb_stride, slnw, slnb, smem + kOffX, ew, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full);
        }
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                mbar_wait(acc_full, n_acc & 1, 30);
                tc_fence_after();
                named_barrier_sync(1, 256);                              // every warp is done reading the previous q|k|v
                // ---- EPI_g: acc + bias -> bf16 q|k|v rows ----
                if (!(p.dbg & 4))
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const int col = half * 96 + cc * 32;
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_acc + lane_base + col, r);
                    tmem_ld_wait();
                    const float* bb = sbqkv + g * 192 + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                        const float4 b1 = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                        uint4 pk;
                        pk.x = pack2_bf16(__uint_as_float(r[8 * j + 0]) + b0.x, __uint_as_float(r[8 * j + 1]) + b0.y);
                        pk.y = pack2_bf16(__uint_as_float(r[8 * j + 2]) + b0.z, __uint_as_float(r[8 * j + 3]) + b0.w);
                        pk.z = pack2_bf16(__uint_as_float(r[8 * j + 4]) + b1.x, __uint_as_float(r[8 * j + 5]) + b1.y);
                        pk.w = pack2_bf16(__uint_as_float(r[8 * j + 6]) + b1.z, __uint_as_float(r[8 * j + 7]) + b1.w);
                        *reinterpret_cast<uint4*>(sqb + row_in_tile * (kPitch * 2) + (col + 8 * j) * 2) = pk;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
                ++n_acc;
                named_barrier_sync(2, 256);                              // q|k|v of the whole tile are in shared memory
                // ---- ATT_g: warp ew owns rows ew*16 .. +15, both heads of the group ----
                float o[2][4][4] = {};
                if (!(p.dbg & 2)) {
                    const int rb = ew;
                    int kbeg, kend, blk = -1;
                    if (L < 16) { kbeg = rb * 16; kend = kbeg + 16; blk = 31 - __clz(L); }
                    else {
                        kbeg = (rb * 16 / L) * L;
                        kend = p.causal ? rb * 16 + 16 : kbeg + L;
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (L <= 16) attn_unit<2>(sq + hh * 32, sq + 64 + hh * 32, sq + 128 + hh * 32, rb, kbeg, kend, blk, p.causal, lane, o[hh]);
                        else attn_unit<8>(sq + hh * 32, sq + 64 + hh * 32, sq + 128 + hh * 32, rb, kbeg, kend, blk, p.causal, lane, o[hh]);
                    }
                }
                mbar_wait(o_empty, (n_o & 1) ^ 1, 31);                   // OUT_{g-1} finished reading O
                {
                    const int gq = lane >> 2, tq = lane & 3;
                    const int r0 = ew * 16 + gq, r1 = r0 + 8;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const int c = hh * 32 + nt * 8 + tq * 2;
                            *reinterpret_cast<unsigned*>(so + sw128_offset(r0, c)) = pack2_bf16(o[hh][nt][0], o[hh][nt][1]);
                            *reinterpret_cast<unsigned*>(so + sw128_offset(r1, c)) = pack2_bf16(o[hh][nt][2], o[hh][nt][3]);
                        }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_full);
                ++n_o;
            }
            // LayerNorm of the next tile before this tile's final epilogue: GEMM_0 of the next tile overlaps the epilogue
            const long long next = tile + gridDim.x;
            if (next < tiles) {
                mbar_wait(x_empty, tile_n & 1, 32);                      // GEMM_3 finished reading X
                if (!(p.dbg & 1)) ln_film_tile(p.h, next * 128, p.M, L, p.gb, p.gb_stride, slnw, slnb, smem + kOffX, ew, lane);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_full);
            }
            mbar_wait(out_full, tile_n & 1, 33);
            tc_fence_after();
            named_barrier_sync(1, 256);                                  // q|k|v region is reused as fp32 staging
            if (!(p.dbg & 8)) residual_epilogue(tmem_out, sbo, smem + kOffQkv, &tmap_h, static_cast<int>(tile) * 128, ew, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(out_empty);
        }
        if (ew == 0 && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ab

int attn_block(float* h, const float* lnw, const float* lnb, const float* gb, long long gb_stride, const void* wqkv,
               const float* bqkv, const void* wo, const float* bo, long long M, int L, int d, int H, int causal, cudaStream_t st) {
    IDB_REQUIRE(d == kD && H == 8, IDB200_EUNSUPPORTED, "fused attention block is specialised for d_model = 256, 8 heads (got %d, %d)", d, H);
    IDB_REQUIRE(L >= 1 && L <= 128 && (128 % L) == 0, IDB200_EUNSUPPORTED, "fused attention block needs L | 128 (got %d)", L);
    IDB_REQUIRE(M >= 0 && M % L == 0, IDB200_EINVAL, "M must be a multiple of L");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(h && lnw && lnb && wqkv && bqkv && wo && bo, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(aligned(h, 16) && (!gb || (aligned(gb, 16) && gb_stride % 4 == 0)), IDB200_EALIGN, "h / gamma_beta must be 16-byte aligned");
    CUtensorMap tq, ta, tb, th;
    int rc = make_tmap_bf16_2d(&tq, wqkv, 768, 256, 192, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&ta, wo, 256, 256, 192, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tb, wo, 256, 256, 64, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&th, h, 4, static_cast<uint64_t>(M), 256, 128, 32);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ab::attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ab::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", ab::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    const long long tiles = (M + 127) / 128;
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    static const int dbg = getenv("IDB200_DBG") ? atoi(getenv("IDB200_DBG")) : 0;
    ab::Params p{h, lnw, lnb, gb, gb_stride, bqkv, bo, M, L, causal, dbg};
    ab::attn_block_kernel<<<grid, ab::kThreads, ab::kSmem, st>>>(tq, ta, tb, th, p);
    return check_launch("attn_block_kernel");
}

}  // namespace idb200

extern "C" int idb200_attn_block(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                                 const void* wqkv_packed, const float* bqkv_packed, const void* wo, const float* bo, int64_t M,
                                 int L, int d, int H, int causal, idb200_stream_t stream) {
    return idb200::attn_block(h, ln_w, ln_b, gamma_beta, gb_stride, wqkv_packed, bqkv_packed, wo, bo, M, L, d, H, causal,
                              static_cast<cudaStream_t>(stream));
}
