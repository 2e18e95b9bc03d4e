// K3h: the MLP of one TransformerBlock in ONE pair-mode kernel for d_model = 384 (and 256):
//
//     h[M, d] += W2 . SiLU(W1 . a + b1) + b2          a = LayerNorm + FiLM output (bf16)
//
// Reference: the `ff` branch of TransformerBlock.forward, src/models/transformer.py:43-45, for the trainer-default models
// (d_model 384, d_ff 1536: src/train/train_interp_levels.py:57-62).  The per-op path wrote the hidden activation [M, d_ff] (403 MB
// per Stage-1 evaluation of the large model) and read it back: 49 us per 128-token tile-SM for the two GEMMs.  mlp_fused.cu keeps it
// on the SM for d_model = 256 with single-CTA tiles; at d = 384 a single CTA would be bound by the L2 -> shared-memory weight
// stream (2.4 MB per tile), so this kernel runs as CTA pairs (tcgen05 cta_group::2: one M = 256 MMA per k-step for the pair's two
// tiles, each CTA stages half of every weight tile) like the MLP half of encoder_fused.cu.
//
// Tensor memory (512 columns): acc2 = d columns (the FF2 accumulator of the whole tile), acc1 = 128 columns (one FF1 chunk).  The
// hidden dimension is processed in chunks of 128:
//   FF1_c   acc1  = A[128 x d] . W1[128c.., :]^T                          d/64 x 4 tcgen05.mma N = 128
//   EPI1_c  acc1 + b1 -> SiLU -> bf16 H (SWIZZLE_128B A operand, 2 k-blocks, 32 KB)   16 compute warps, thread <-> row x 32 columns
//   FF2_c   acc2 += H . W2[:, 128c..]^T                                   2 k-blocks x 4 x (N = 256 [+ N = 128 at d = 384])
//   final   h += acc2 + b2: fp32 [128 x 32] boxes staged in the idle H tile, TMA reduce-add into the residual stream
// acc1 and H are SINGLE buffers (TMEM: 384 + 128 = 512; shared memory: A 96 KB + H 32 KB + a 3 x 24 KB weight ring), and the tensor
// pipe still runs back to back: issue order FF1_{c+1}, FF2_c; EPI1 releases acc1 right after its TMEM load (so FF1_{c+1} runs under
// the SiLU math), computes the chunk into registers, and only then waits for FF2_{c-1} to have released H.  (The first revision used
// 64-column chunks with double-buffered acc1 / H: an N = 64 pair MMA costs the same ~64 cycles as N = 128 -- 2.3 k cycles per 64
// hidden columns with the SiLU and the weight loads ablated, against 1.5 k nominal.)  A ring slot = half of the d/64 k-blocks of a
// W1 chunk (this CTA's 64 rows, 8 KB each) or one k-block of W2 (this CTA's d/2 permuted output rows).  W2 is row-permuted on the
// host (idb200_mlp_pair_w2_order) so that each CTA's rows are one TMA box: a pair MMA takes the first N/2 rows of its B operand from
// the even CTA and the rest from the odd one.
#include <cstdlib>

#include "fused_common.cuh"

namespace idb200 {
using namespace tc;
using namespace fused;

namespace mp {
constexpr int kThreads = 640;
constexpr int kCW = 16;
constexpr int kCT = kCW * 32;
constexpr int kSlots = 3;
constexpr int kSlotBytes = 192 * 128;               // 24 KB
constexpr int kMaxFF = 2048;
constexpr int kRegsAux = 32, kRegsCompute = 112;

template <int NK>
struct Cfg {
    static constexpr int kD = NK * 64;
    static constexpr int kOffX = 0;                                  // NK x [128 x 64] bf16 SWIZZLE_128B
    static constexpr int kOffH = NK * kTile;                         // 2 k-blocks [128 x 64] bf16 | 2 x [128 x 32] fp32 (final epilogue)
    static constexpr int kOffRing = kOffH + 2 * kTile;
    static constexpr int kOffBar = kOffRing + kSlots * kSlotBytes;
    static constexpr int kOffBias = kOffBar + 256;                   // b1 (kMaxFF) | b2 (kD) | ln_w (kD) | ln_b (kD) (kLN kernels)
    static constexpr int kSmem = kOffBias + (kMaxFF + 3 * kD) * 4 + 1024;
    static constexpr int kN1 = kD < 256 ? kD : 256;                  // FF2: first MMA's N, second's (0 or 128)
    static constexpr int kN2 = kD - kN1;
    static_assert(kD == 256 || kD == 384, "d_model 256 or 384");
    static_assert(kOffRing % 1024 == 0 && kSlotBytes % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
    static_assert(kSmem <= 232448, "shared memory budget");
};

struct Params {
    const float* b1;            // [ff]
    const float* b2;            // [d]
    long long M;
    int ff;
    // kLN kernels: A = LayerNorm(h) * (1 + gamma) + beta is produced in shared memory by the compute warps (no [M, d] operand in HBM)
    const float* h;             // [M, d] fp32 residual stream (LayerNorm input; also updated through tm_h)
    const float* lnw;
    const float* lnb;
    const float* gb;            // FiLM rows [gamma | beta] per trajectory, or nullptr
    long long gb_stride;
    int L;
    int dbg;                    // dev (IDB200_MLP_DBG): 1 = EPI1 without the SiLU math, 2 = the weight producer signals slots without loading
};

__device__ __forceinline__ float silu_t(float x) {
    const float hx = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}

template <int NK, bool kLN>
__global__ void __launch_bounds__(kThreads, 1)
mlp_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                const __grid_constant__ CUtensorMap tm_h, const Params p) {
    using C = Cfg<NK>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* x_full = bars + 0;                    // TMA -> MMA (leader; tx bytes of both CTAs' A tiles)
    uint64_t* x_empty = bars + 1;                   // MMA -> TMA: the last FF1 of the tile has read A (multicast commit)
    uint64_t* slot_full = bars + 2;                 // [kSlots] (leader)
    uint64_t* slot_empty = slot_full + kSlots;      // [kSlots] (multicast commit)
    uint64_t* acc1_full = slot_empty + kSlots;      // MMA -> compute
    uint64_t* acc1_empty = acc1_full + 1;           // compute -> MMA (leader, 2 * kCW)
    uint64_t* hb_full = acc1_empty + 1;             // compute -> MMA (leader, 2 * kCW)
    uint64_t* hb_empty = hb_full + 1;               // MMA -> compute
    uint64_t* acc2_full = hb_empty + 1;             // MMA -> compute
    uint64_t* acc2_empty = acc2_full + 1;           // compute -> MMA (leader, 2 * kCW)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
    float* sb1 = reinterpret_cast<float*>(smem + C::kOffBias);
    float* sb2 = sb1 + kMaxFF;
    float* slnw = sb2 + C::kD;
    float* slnb = slnw + C::kD;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int nc = p.ff / 128;
    const int tiles = static_cast<int>((p.M + 127) / 128);
    const int trips = (tiles + 1) / 2;
    const int trip0 = static_cast<int>(blockIdx.x / 2), trip_stride = static_cast<int>(gridDim.x / 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_w1);
        tma_prefetch_desc(&tm_w2);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, kLN ? 2 * kCW : 1);
        mbar_init(x_empty, 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 1); }
        mbar_init(acc1_full, 1);
        mbar_init(acc1_empty, 2 * kCW);
        mbar_init(hb_full, 2 * kCW);
        mbar_init(hb_empty, 1);
        mbar_init(acc2_full, 1);
        mbar_init(acc2_empty, 2 * kCW);
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    for (int i = threadIdx.x; i < p.ff; i += kThreads) sb1[i] = p.b1[i];
    for (int i = threadIdx.x; i < C::kD; i += kThreads) sb2[i] = p.b2[i];
    if (kLN)
        for (int i = threadIdx.x; i < C::kD; i += kThreads) { slnw[i] = p.lnw[i]; slnb[i] = p.lnb[i]; }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base + C::kD;

    if (warp < 4) {
        setmaxnreg_dec<kRegsAux>();
        if (warp == 0) {
            // ===================== weight producer: in exactly the order the MMA warp consumes the slots =====================
            if (lane == 0) {
                int slot = 0;
                uint32_t sphase = 0;
                const bool noload = (p.dbg & 2) != 0;
                auto begin = [&](uint32_t bytes) -> uint8_t* {
                    mbar_wait(&slot_empty[slot], sphase ^ 1, 10);
                    if (rank == 0) { if (noload) mbar_arrive(&slot_full[slot]); else mbar_arrive_expect_tx(&slot_full[slot], 2 * bytes); }
                    return smem + C::kOffRing + slot * kSlotBytes;
                };
                auto end = [&]() { if (++slot == kSlots) { slot = 0; sphase ^= 1; } };
                auto ff1 = [&](int c) {                                  // this CTA's 64 rows of the chunk: 2 slots x NK/2 k-blocks of 8 KB
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        uint8_t* dst = begin((NK / 2) * 8192);
                        if (!noload)
                            for (int i = 0; i < NK / 2; ++i)
                                tma_load_2d_2sm(dst + i * 8192, &tm_w1, &slot_full[slot], (half * (NK / 2) + i) * 64, c * 128 + static_cast<int>(rank) * 64);
                        end();
                    }
                };
                auto ff2 = [&](int c) {                                  // this CTA's d/2 (permuted) output rows, 2 k-blocks = 2 slots
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb) {
                        uint8_t* dst = begin((C::kD / 2) * 128);
                        if (!noload) tma_load_2d_2sm(dst, &tm_w2, &slot_full[slot], c * 128 + kb * 64, static_cast<int>(rank) * (C::kD / 2));
                        end();
                    }
                };
                for (int trip = trip0; trip < trips; trip += trip_stride) {
#pragma unroll 1
                    for (int c = -1; c < nc; ++c) {
                        if (c + 1 < nc) ff1(c + 1);
                        if (c >= 0) ff2(c);
                    }
                }
            }
        } else if (warp == 3) {
            // ===================== A-tile loader (kLN: the compute warps produce A) =====================
            if (!kLN && lane == 0) {
                uint32_t n = 0;
                for (int trip = trip0; trip < trips; trip += trip_stride, ++n) {
                    const int tile = 2 * trip + static_cast<int>(rank);
                    mbar_wait(x_empty, (n & 1) ^ 1, 11);
                    if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * NK * kTile);
#pragma unroll 1
                    for (int kb = 0; kb < NK; ++kb) tma_load_2d_2sm(smem + C::kOffX + kb * kTile, &tm_a, x_full, kb * 64, tile * 128);
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (the even CTA of the pair) =====================
            if (rank == 0) {
                constexpr uint32_t idesc128 = umma_idesc_bf16(256, 128);
                constexpr uint32_t idescN1 = umma_idesc_bf16(256, C::kN1);
                constexpr uint32_t idescN2 = umma_idesc_bf16(256, C::kN2 > 0 ? C::kN2 : 64);
                int slot = 0;
                uint32_t sphase = 0, n = 0;
                const uint32_t sX = smem_u32(smem + C::kOffX), sH = smem_u32(smem + C::kOffH), sR = smem_u32(smem + C::kOffRing);
                auto wait = [&](uint64_t* bar, uint32_t parity, int tag) {
                    mbar_wait(bar, parity, tag);
                    tc_fence_after();
                };
                auto commit = [&](uint64_t* bar) {
                    if (elect_one_sync()) umma_commit_2sm(bar);
                    __syncwarp();
                };
                auto mma4 = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool first_zero) {
                    const uint64_t ad = umma_desc_sw128(a_addr);
                    const uint64_t bd = umma_desc_sw128(b_addr);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
                    }
                    __syncwarp();
                };
                auto ff1 = [&](int c) {
                    const uint32_t u = (n * static_cast<uint32_t>(nc) + static_cast<uint32_t>(c)) & 1u;      // chunks so far
                    wait(acc1_empty, u ^ 1u, 20);                        // EPI1 of the previous chunk has the accumulator in registers
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        wait(&slot_full[slot], sphase, 21);
                        const uint32_t bs = sR + slot * kSlotBytes;
#pragma unroll 1
                        for (int i = 0; i < NK / 2; ++i) mma4(tmem_acc1, sX + (half * (NK / 2) + i) * kTile, bs + i * 8192, idesc128, half == 0 && i == 0);
                        commit(&slot_empty[slot]);
                        if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                    }
                    commit(acc1_full);
                    if (c == nc - 1) commit(x_empty);                    // A no longer needed: the next tile's load may start
                };
                auto ff2 = [&](int c) {
                    const uint32_t u = (n * static_cast<uint32_t>(nc) + static_cast<uint32_t>(c)) & 1u;
                    wait(hb_full, u, 22);                                // EPI1 wrote H
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb) {
                        wait(&slot_full[slot], sphase, 23);
                        const uint32_t bs = sR + slot * kSlotBytes;
                        mma4(tmem_base, sH + kb * kTile, bs, idescN1, c == 0 && kb == 0);
                        if (C::kN2 > 0) mma4(tmem_base + C::kN1, sH + kb * kTile, bs + (C::kN1 / 2) * 128, idescN2, c == 0 && kb == 0);
                        commit(&slot_empty[slot]);
                        if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                    }
                    commit(hb_empty);
                    if (c == nc - 1) commit(acc2_full);
                };
                for (int trip = trip0; trip < trips; trip += trip_stride, ++n) {
                    wait(x_full, n & 1, 24);
#pragma unroll 1
                    for (int c = -1; c < nc; ++c) {
                        if (c + 1 < nc) ff1(c + 1);
                        if (c == 0) wait(acc2_empty, (n & 1) ^ 1, 25);   // the previous tile's final epilogue drained acc2
                        if (c >= 0) ff2(c);
                    }
                }
            }
        }
    } else {
        setmaxnreg_inc<kRegsCompute>();
        // ===================== compute warps =====================
        const int ew = warp - 4;
        const int q = ew & 3, part = ew >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        uint32_t n = 0;
        // kLN: LayerNorm + FiLM of a tile's rows -> A.  The next tile is normalised between the last EPI1 and the final epilogue of
        // the current one (A is free once the last FF1 has run), so its first FF1s run under that epilogue.
        auto ln_tile = [&](int tile_, uint32_t n_) {
            mbar_wait(x_empty, (n_ & 1) ^ 1, 33);
            ln_film_rows<NK / 2, kCW>(p.h, static_cast<long long>(tile_) * 128, p.M, p.L, p.gb, p.gb_stride, slnw, slnb, smem + C::kOffX, ew, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(x_full);
        };
        if (kLN && trip0 < trips) ln_tile(2 * trip0 + static_cast<int>(rank), 0u);
        for (int trip = trip0; trip < trips; trip += trip_stride, ++n) {
            const int tile = 2 * trip + static_cast<int>(rank);
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                const uint32_t u = (n * static_cast<uint32_t>(nc) + static_cast<uint32_t>(c)) & 1u;
                mbar_wait(acc1_full, u, 30);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld_32x32(tmem_acc1 + lane_base + part * 32, r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(acc1_empty);           // in registers: FF1_{c+1} may overwrite the accumulator
                const float* bb = sb1 + c * 128 + part * 32;
                uint4 pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {                            // 8 columns -> one 16-byte swizzle chunk
                    const float4 b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                    const float4 b1v = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                    if (p.dbg & 1) { pk[j] = make_uint4(r[8 * j], r[8 * j + 1], r[8 * j + 2], r[8 * j + 3]); continue; }
                    pk[j].x = pack2_bf16(silu_t(__uint_as_float(r[8 * j + 0]) + b0.x), silu_t(__uint_as_float(r[8 * j + 1]) + b0.y));
                    pk[j].y = pack2_bf16(silu_t(__uint_as_float(r[8 * j + 2]) + b0.z), silu_t(__uint_as_float(r[8 * j + 3]) + b0.w));
                    pk[j].z = pack2_bf16(silu_t(__uint_as_float(r[8 * j + 4]) + b1v.x), silu_t(__uint_as_float(r[8 * j + 5]) + b1v.y));
                    pk[j].w = pack2_bf16(silu_t(__uint_as_float(r[8 * j + 6]) + b1v.z), silu_t(__uint_as_float(r[8 * j + 7]) + b1v.w));
                }
                mbar_wait(hb_empty, u ^ 1u, 31);                         // FF2 of the previous chunk finished reading H (the math is done: only
                uint8_t* hb = smem + C::kOffH + (part >> 1) * kTile;     // the stores wait)
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(hb + sw128_offset(row, (part & 1) * 32 + j * 8)) = pk[j];
                fence_proxy_async_smem();                                // H writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(hb_full);
            }
            if (kLN && trip + trip_stride < trips) ln_tile(2 * (trip + trip_stride) + static_cast<int>(rank), n + 1);
            // ---- final epilogue: h += acc2 + b2, rounds of 32 columns through two fp32 [128 x 32] boxes in the (idle) H buffers, TMA
            // reduce-add into the residual stream.  (red.global.add.v4.f32 straight from registers -- no staging, no barriers -- was
            // tried: 24 us per tile instead of 7, the L2 reduction units do not keep up with 16-byte requests from 512 threads.) ----
            mbar_wait(acc2_full, n & 1, 32);                             // every FF2 of the tile has completed (so H is idle, too)
            tc_fence_after();
#pragma unroll 1
            for (int rnd = 0; rnd < C::kD / 32; ++rnd) {
                uint32_t r[8];
                tmem_ld_32x8(tmem_base + lane_base + rnd * 32 + part * 8, r);
                tmem_ld_wait();
                uint8_t* box = smem + C::kOffH + (rnd & 1) * kTile;
                const float* bv = sb2 + rnd * 32 + part * 8;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bv + 4 * j);
                    const float4 v = make_float4(__uint_as_float(r[4 * j + 0]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y,
                                                 __uint_as_float(r[4 * j + 2]) + b4.z, __uint_as_float(r[4 * j + 3]) + b4.w);
                    *reinterpret_cast<float4*>(box + row * 128 + ((((part * 2 + j) ^ (row & 7)) & 7) << 4)) = v;
                }
                fence_proxy_async_smem();
                named_barrier_sync(1, kCT);
                if (ew == 0 && lane == 0) {
                    tma_reduce_add_2d(&tm_h, box, rnd * 32, tile * 128);  // rows >= M are clipped by the tensor map
                    tma_store_commit();
                    tma_store_wait_read<1>();                            // the other box (round rnd - 1) has been read
                }
                named_barrier_sync(2, kCT);
            }
            if (ew == 0 && lane == 0) tma_store_wait_read<0>();          // H is reused by EPI1 of the next tile
            named_barrier_sync(1, kCT);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(acc2_empty);
        }
        if (ew == 0 && lane == 0) tma_store_wait_all();                  // all residual updates landed before exit
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

template <int NK, bool kLN>
int launch(const CUtensorMap& ta, const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& th, const Params& p, long long tiles,
           cudaStream_t st) {
    using C = Cfg<NK>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(mlp_pair_kernel<NK, kLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(mlp_pair, smem=%d): %s", C::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    const long long trips = (tiles + 1) / 2;
    const long long pairs = trips < num_sms() / 2 ? trips : num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_pair_kernel<NK, kLN>, ta, t1, t2, th, p);
    if (e != cudaSuccess) return fail(IDB200_ECUDA, "mlp_pair_kernel: %s", cudaGetErrorString(e));
    return check_launch("mlp_pair_kernel");
}

}  // namespace mp
}  // namespace idb200

using namespace idb200;

// order[i] = the row of W2 [d, ff] that goes to row i of the packed matrix: each CTA's rows contiguous
extern "C" int idb200_mlp_pair_w2_order(int d, int* order) {
    IDB_REQUIRE(order && (d == 256 || d == 384), IDB200_EINVAL, "d_model 256 or 384");
    const int n1 = d < 256 ? d : 256, n2 = d - n1;
    int i = 0;
    for (int r = 0; r < 2; ++r) {
        for (int j = 0; j < n1 / 2; ++j) order[i++] = r * (n1 / 2) + j;
        for (int j = 0; j < n2 / 2; ++j) order[i++] = n1 + r * (n2 / 2) + j;
    }
    return IDB200_OK;
}

static int mlp_pair_impl(const void* a, const float* ln_w, const float* ln_b, const float* gamma_beta, long long gb_stride, int L, const void* W1,
                         const float* b1, const void* W2_packed, const float* b2, float* h, long long M, int d, int ff, cudaStream_t st) {
    const bool ln = (a == nullptr);
    IDB_REQUIRE(d == 256 || d == 384, IDB200_EUNSUPPORTED, "pair-mode fused MLP supports d_model 256 or 384 (got %d)", d);
    IDB_REQUIRE(ff % 128 == 0 && ff >= 128 && ff <= mp::kMaxFF, IDB200_EUNSUPPORTED, "d_ff must be a multiple of 128 in [128, 2048] (got %d)", ff);
    IDB_REQUIRE(M >= 0 && M < (1ll << 37), IDB200_EINVAL, "bad shape");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(W1 && b1 && W2_packed && b2 && h, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(aligned(h, 16), IDB200_EALIGN, "h must be 16-byte aligned");
    if (ln) {
        IDB_REQUIRE(ln_w && ln_b && L >= 1 && M % L == 0, IDB200_EINVAL, "LayerNorm parameters / M must be a multiple of L");
        IDB_REQUIRE(!gamma_beta || (aligned(gamma_beta, 16) && gb_stride % 4 == 0), IDB200_EALIGN, "gamma_beta must be 16-byte aligned");
    }
    CUtensorMap ta, t1, t2, th;
    int rc = make_tmap_bf16_2d(&t1, W1, static_cast<uint64_t>(ff), static_cast<uint64_t>(d), 64, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&t2, W2_packed, static_cast<uint64_t>(d), static_cast<uint64_t>(ff), static_cast<uint32_t>(d / 2), 64);
    if (rc) return rc;
    rc = make_tmap_2d(&th, h, 4, static_cast<uint64_t>(M), static_cast<uint64_t>(d), 128, 32);
    if (rc) return rc;
    if (ln) ta = t1;                                                     // unused by the kLN kernels
    else {
        rc = make_tmap_bf16_2d(&ta, a, static_cast<uint64_t>(M), static_cast<uint64_t>(d), 128, 64);
        if (rc) return rc;
    }
    static const int dbg = getenv("IDB200_MLP_DBG") ? atoi(getenv("IDB200_MLP_DBG")) : 0;
    mp::Params p{b1, b2, M, ff, h, ln_w, ln_b, gamma_beta, gb_stride, L, dbg};
    const long long tiles = (M + 127) / 128;
    if (ln) return d == 256 ? mp::launch<4, true>(ta, t1, t2, th, p, tiles, st) : mp::launch<6, true>(ta, t1, t2, th, p, tiles, st);
    return d == 256 ? mp::launch<4, false>(ta, t1, t2, th, p, tiles, st) : mp::launch<6, false>(ta, t1, t2, th, p, tiles, st);
}

extern "C" int idb200_mlp_pair(const void* a, const void* W1, const float* b1, const void* W2_packed, const float* b2, float* h, int64_t M, int d,
                               int ff, idb200_stream_t stream) {
    IDB_REQUIRE(a != nullptr, IDB200_EINVAL, "NULL pointer");
    return mlp_pair_impl(a, nullptr, nullptr, nullptr, 0, 1, W1, b1, W2_packed, b2, h, M, d, ff, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_ln_mlp_pair(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride, int L, const void* W1,
                                  const float* b1, const void* W2_packed, const float* b2, int64_t M, int d, int ff, idb200_stream_t stream) {
    return mlp_pair_impl(nullptr, ln_w, ln_b, gamma_beta, gb_stride, L, W1, b1, W2_packed, b2, h, M, d, ff, static_cast<cudaStream_t>(stream));
}
