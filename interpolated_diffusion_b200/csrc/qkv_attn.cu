// K3g: in_proj + multi-head self attention of one TransformerBlock in ONE kernel, for the models the whole-encoder kernel does
// not take (d_model = 384, 12 heads: the trainer defaults of src/train/train_interp_levels.py:57-62; also d_model = 256):
//
//     o[M, d] = MHA(a)            a = LayerNorm + FiLM output (bf16),  q|k|v = a . Wqkv^T + b never leave the SM
//
// Reference: nn.MultiheadAttention inside src/models/transformer.py:11,39 (in_proj + scaled-dot-product attention; the
// out_proj stays a separate residual GEMM).  The per-op path wrote the packed projections qkv [M, 3d] (302 MB per Stage-1
// evaluation of the large model) and read them back in the attention kernel: two memory-bound launches, 36 us per 128-token
// tile-SM; here a tile's q|k|v go from the accumulator (tensor memory) through a bf16 staging tile in shared memory straight
// into the in-tile attention core.
//
// Same machinery as the attention half of encoder_fused.cu: CTA pairs (tcgen05 cta_group::2: one M = 256 MMA per k-step for the
// pair's two 128-token tiles, each CTA stages half of every weight tile), head groups of 2 heads (q|k|v = 192 accumulator
// columns, double-buffered so GEMM_{g+1} runs under EPI_g / ATT_g), 16 compute warps (thread <-> tile row x column quarter in
// the epilogue; warp <-> (16-row block, head) in the attention core: mma.sync m16n8k16 + ldmatrix, block-diagonal over the
// tile's trajectories, causal optional), O_g leaves through a SWIZZLE_128B tile and a TMA store.  A (the tile's rows of `a`)
// arrives by TMA; rows beyond M are zero-filled on load and clipped on store.
#include <cstdlib>

#include "fused_common.cuh"

namespace idb200 {
using namespace tc;
using namespace fused;

namespace qa {
constexpr int kThreads = 640;
constexpr int kCW = 16;
constexpr int kCT = kCW * 32;
constexpr int kSlots = 4;
constexpr int kSlotBytes = 96 * 128;                // pair mode: this CTA's 96 of the group's 192 weight rows, one 64-wide k-block
constexpr int kRegsAux = 32, kRegsCompute = 112;    // 128 * 32 + 512 * 112 = 640 * 96

template <int NG>
struct Cfg {
    static constexpr int kOffX = 0;                                 // NG x [128 x 64] bf16 SWIZZLE_128B (A operand)
    static constexpr int kOffQkv = NG * kTile;                      // staged q|k|v rows (pitch kPitch)
    static constexpr int kOffO = kOffQkv + 128 * kPitch * 2;        // [128 x 64] bf16 SWIZZLE_128B (TMA store source)
    static constexpr int kOffRing = kOffO + kTile;
    static constexpr int kOffBar = kOffRing + kSlots * kSlotBytes;
    static constexpr int kOffBias = kOffBar + 256;                  // bqkv (head-group-major) 3 d floats | ln_w d | ln_b d (kLN kernels)
    static constexpr int kSmem = kOffBias + (NG * 192 + 2 * NG * 64) * 4 + 1024;
    static_assert(kOffO % 1024 == 0 && kOffRing % 1024 == 0 && kSlotBytes % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
    static_assert(kSmem <= 232448, "shared memory budget");
};

struct Params {
    const float* bqkv;          // [3 d] head-group-major
    long long M;
    int L, causal;
    // kLN kernels: A = LayerNorm(h) * (1 + gamma) + beta is produced in shared memory by the compute warps (no [M, d] operand in HBM)
    const float* h;             // [M, d] fp32 residual stream (read only)
    const float* lnw;
    const float* lnb;
    const float* gb;            // FiLM rows [gamma | beta] per trajectory, or nullptr
    long long gb_stride;
};

template <int NG, bool kLN>
__global__ void __launch_bounds__(kThreads, 1)
qkv_attn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_o,
                const Params p) {
    using C = Cfg<NG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* x_full = bars + 0;                    // TMA -> MMA: the pair's two A tiles have landed (leader, tx bytes of both CTAs)
    uint64_t* x_empty = bars + 1;                   // MMA -> TMA: every GEMM of the tile has read A (multicast commit)
    uint64_t* slot_full = bars + 2;                 // [kSlots] (leader)
    uint64_t* slot_empty = slot_full + kSlots;      // [kSlots] (multicast commit)
    uint64_t* acc_full = slot_empty + kSlots;       // [2] MMA -> compute (multicast commit)
    uint64_t* acc_empty = acc_full + 2;             // [2] compute -> MMA (leader, 2 * kCW arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sbias = reinterpret_cast<float*>(smem + C::kOffBias);
    float* slnw = sbias + NG * 192;
    float* slnb = slnw + NG * 64;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int tiles = static_cast<int>((p.M + 127) / 128);
    const int trips = (tiles + 1) / 2;
    const int trip0 = static_cast<int>(blockIdx.x / 2), trip_stride = static_cast<int>(gridDim.x / 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_o);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, kLN ? 2 * kCW : 1);
        mbar_init(x_empty, 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kCW); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    for (int i = threadIdx.x; i < NG * 192; i += kThreads) sbias[i] = p.bqkv[i];
    if (kLN)
        for (int i = threadIdx.x; i < NG * 64; i += kThreads) { slnw[i] = p.lnw[i]; slnb[i] = p.lnb[i]; }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        setmaxnreg_dec<kRegsAux>();
        if (warp == 0) {
            // ===================== weight producer (free-running across groups and tiles) =====================
            if (lane == 0) {
                int slot = 0;
                uint32_t sphase = 0;
                for (int trip = trip0; trip < trips; trip += trip_stride) {
#pragma unroll 1
                    for (int g = 0; g < NG; ++g) {
#pragma unroll 1
                        for (int kb = 0; kb < NG; ++kb) {
                            mbar_wait(&slot_empty[slot], sphase ^ 1, 10);
                            if (rank == 0) mbar_arrive_expect_tx(&slot_full[slot], 2 * kSlotBytes);
                            tma_load_2d_2sm(smem + C::kOffRing + slot * kSlotBytes, &tm_w, &slot_full[slot], kb * 64, g * 192 + static_cast<int>(rank) * 96);
                            if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                        }
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
                    if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * NG * kTile);
#pragma unroll 1
                    for (int kb = 0; kb < NG; ++kb) tma_load_2d_2sm(smem + C::kOffX + kb * kTile, &tm_a, x_full, kb * 64, tile * 128);
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (the even CTA of the pair) =====================
            if (rank == 0) {
                constexpr uint32_t idesc192 = umma_idesc_bf16(256, 192);
                int slot = 0;
                uint32_t sphase = 0, n = 0;
                const uint32_t sX = smem_u32(smem + C::kOffX), sR = smem_u32(smem + C::kOffRing);
                for (int trip = trip0; trip < trips; trip += trip_stride, ++n) {
                    mbar_wait(x_full, n & 1, 20);
                    tc_fence_after();
#pragma unroll 1
                    for (int g = 0; g < NG; ++g) {
                        const int b = g & 1;
                        const uint32_t use = (n * (NG / 2) + static_cast<uint32_t>(g >> 1)) & 1u;       // uses of accumulator b so far
                        mbar_wait(&acc_empty[b], use ^ 1u, 21);
                        tc_fence_after();
#pragma unroll 1
                        for (int kb = 0; kb < NG; ++kb) {
                            mbar_wait(&slot_full[slot], sphase, 22);
                            tc_fence_after();
                            const uint64_t ad = umma_desc_sw128(sX + kb * kTile);
                            const uint64_t bd = umma_desc_sw128(sR + slot * kSlotBytes);
                            if (elect_one_sync()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem_base + b * 256, ad + 2 * k, bd + 2 * k, idesc192, (kb == 0 && k == 0) ? 0u : 1u);
                            }
                            __syncwarp();
                            if (elect_one_sync()) umma_commit_2sm(&slot_empty[slot]);
                            __syncwarp();
                            if (++slot == kSlots) { slot = 0; sphase ^= 1; }
                        }
                        if (elect_one_sync()) umma_commit_2sm(&acc_full[b]);
                        __syncwarp();
                    }
                    if (elect_one_sync()) umma_commit_2sm(x_empty);
                    __syncwarp();
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
        const __nv_bfloat16* sq = reinterpret_cast<const __nv_bfloat16*>(smem + C::kOffQkv);
        uint8_t* sqb = smem + C::kOffQkv;
        uint8_t* so = smem + C::kOffO;
        const int L = p.L;
        const int rb = ew & 7, hh = ew >> 3;            // attention unit: 16-row block, head of the group
        uint32_t okbits = 0;                            // L < 16: block-diagonal mask of the 16 x 16 score block
        if (L < 16) {
            const int lg = 31 - __clz(L);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int key = nt * 8 + (lane & 3) * 2 + (c & 1), qr = (lane >> 2) + (c < 2 ? 0 : 8);
                    okbits |= ((((key ^ qr) >> lg) == 0) ? 1u : 0u) << (nt * 4 + c);
                }
        }
        uint32_t n = 0;
        // kLN: LayerNorm + FiLM of a tile's rows -> A (bf16 SWIZZLE_128B k-blocks); the tile after the current one is normalised right
        // after the last group's EPI (its GEMMs then run under that group's attention), the first one before the loop
        auto ln_tile = [&](int tile_, uint32_t n_) {
            mbar_wait(x_empty, (n_ & 1) ^ 1, 33);                        // every GEMM of the previous tile has read A
            ln_film_rows<NG / 2, kCW>(p.h, static_cast<long long>(tile_) * 128, p.M, p.L, p.gb, p.gb_stride, slnw, slnb, smem + C::kOffX, ew, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(x_full);
        };
        if (kLN && trip0 < trips) ln_tile(2 * trip0 + static_cast<int>(rank), 0u);
        for (int trip = trip0; trip < trips; trip += trip_stride, ++n) {
            const int tile = 2 * trip + static_cast<int>(rank);
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                const int b = g & 1;
                const uint32_t use = (n * (NG / 2) + static_cast<uint32_t>(g >> 1)) & 1u;
                mbar_wait(&acc_full[b], use, 30);
                tc_fence_after();
                // ---- EPI_g: accumulator + bias -> bf16 q|k|v rows (this thread: its row, 48 of the 192 columns) ----
                {
                    uint32_t ra[32], rc[16];
                    const uint32_t ta = tmem_base + b * 256 + lane_base + part * 48;
                    tmem_ld_32x32(ta, ra);
                    tmem_ld_32x16(ta + 32, rc);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&acc_empty[b]);    // the accumulator is in registers: GEMM_{g+2} may overwrite it
                    const float* bb = sbias + g * 192 + part * 48;
                    uint8_t* dst = sqb + row * (kPitch * 2) + part * 96;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                        const float4 b1 = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                        const uint32_t* v = (j < 4) ? &ra[8 * j] : &rc[8 * (j - 4)];
                        uint4 pk;
                        pk.x = pack2_bf16(__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y);
                        pk.y = pack2_bf16(__uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w);
                        pk.z = pack2_bf16(__uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y);
                        pk.w = pack2_bf16(__uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w);
                        *reinterpret_cast<uint4*>(dst + 16 * j) = pk;
                    }
                }
                named_barrier_sync(1, kCT);                              // q|k|v of the whole tile are staged
                if (kLN && g == NG - 1 && trip + trip_stride < trips) ln_tile(2 * (trip + trip_stride) + static_cast<int>(rank), n + 1);
                // ---- ATT_g: this warp's (16-row block, head) ----
                float o[4][4];
                {
                    const __nv_bfloat16* qh = sq + hh * 32;
                    if (L < 16) attn_unit_fast<2, 1>(qh, qh + 64, qh + 128, rb, rb * 16, rb * 16 + 16, okbits, lane, o);
                    else if (p.causal) {
                        const int kbeg = (rb * 16 / L) * L, kend = rb * 16 + 16;
                        if (L == 16) attn_unit_fast<2, 2>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                        else attn_unit_fast<4, 2>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                    } else {
                        const int kbeg = (rb * 16 / L) * L, kend = kbeg + L;
                        if (L == 16) attn_unit_fast<2, 0>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                        else attn_unit_fast<4, 0>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                    }
                }
                if (ew == 0 && lane == 0) tma_store_wait_read<0>();      // the previous O tile has been read by the TMA engine
                named_barrier_sync(2, kCT);                              // ... and every warp is done reading the staged q|k|v
                {
                    const int gq = lane >> 2, tq = lane & 3;
                    const int r0 = rb * 16 + gq;
                    uint8_t* o0 = so + r0 * 128 + tq * 4;                // sw128_offset(r0, c): chunk (c >> 3) ^ (r0 & 7); r0 + 8: + 1024 bytes
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int ch = ((hh * 4 + nt) ^ (r0 & 7)) << 4;
                        *reinterpret_cast<unsigned*>(o0 + ch) = pack2_bf16(o[nt][0], o[nt][1]);
                        *reinterpret_cast<unsigned*>(o0 + ch + 1024) = pack2_bf16(o[nt][2], o[nt][3]);
                    }
                }
                fence_proxy_async_smem();
                named_barrier_sync(3, kCT);
                if (ew == 0 && lane == 0) {
                    tma_store_2d(&tm_o, so, g * 64, tile * 128);         // rows >= M are clipped
                    tma_store_commit();
                }
            }
        }
        if (ew == 0 && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

template <int NG, bool kLN>
int launch(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const Params& p, long long tiles, cudaStream_t st) {
    using C = Cfg<NG>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(qkv_attn_kernel<NG, kLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(qkv_attn, smem=%d): %s", C::kSmem, cudaGetErrorString(e));
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, qkv_attn_kernel<NG, kLN>, ta, tw, to, p);
    if (e != cudaSuccess) return fail(IDB200_ECUDA, "qkv_attn_kernel: %s", cudaGetErrorString(e));
    return check_launch("qkv_attn_kernel");
}

}  // namespace qa
}  // namespace idb200

using namespace idb200;

static int qkv_attention_impl(const void* a, const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, long long gb_stride,
                              const void* wqkv_packed, const float* bqkv_packed, void* o, long long M, int L, int d, int H, int causal,
                              cudaStream_t st) {
    const bool ln = (a == nullptr);
    IDB_REQUIRE(d == 256 || d == 384, IDB200_EUNSUPPORTED, "fused in_proj + attention supports d_model 256 or 384 (got %d)", d);
    IDB_REQUIRE(H * 32 == d, IDB200_EUNSUPPORTED, "head_dim must be 32 (d = %d, H = %d)", d, H);
    IDB_REQUIRE(L >= 1 && L <= 128 && (128 % L) == 0, IDB200_EUNSUPPORTED, "fused in_proj + attention needs L | 128 (got %d)", L);
    IDB_REQUIRE(M >= 0 && M % L == 0 && M < (1ll << 37), IDB200_EINVAL, "M must be a multiple of L (and below 2^37)");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE((a || h) && wqkv_packed && bqkv_packed && o, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(aligned(bqkv_packed, 16), IDB200_EALIGN, "bias must be 16-byte aligned");
    if (ln) {
        IDB_REQUIRE(ln_w && ln_b, IDB200_EINVAL, "NULL LayerNorm parameters");
        IDB_REQUIRE(aligned(h, 16) && (!gamma_beta || (aligned(gamma_beta, 16) && gb_stride % 4 == 0)), IDB200_EALIGN,
                    "h / gamma_beta must be 16-byte aligned");
    }
    CUtensorMap ta, tw, to;
    int rc = make_tmap_bf16_2d(&tw, wqkv_packed, static_cast<uint64_t>(3 * d), static_cast<uint64_t>(d), 96, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&to, o, static_cast<uint64_t>(M), static_cast<uint64_t>(d), 128, 64);
    if (rc) return rc;
    if (ln) ta = to;                                                     // unused by the kLN kernels
    else {
        rc = make_tmap_bf16_2d(&ta, a, static_cast<uint64_t>(M), static_cast<uint64_t>(d), 128, 64);
        if (rc) return rc;
    }
    qa::Params p{bqkv_packed, M, L, causal, h, ln_w, ln_b, gamma_beta, gb_stride};
    const long long tiles = (M + 127) / 128;
    if (ln) return d == 256 ? qa::launch<4, true>(ta, tw, to, p, tiles, st) : qa::launch<6, true>(ta, tw, to, p, tiles, st);
    return d == 256 ? qa::launch<4, false>(ta, tw, to, p, tiles, st) : qa::launch<6, false>(ta, tw, to, p, tiles, st);
}

extern "C" int idb200_qkv_attention(const void* a, const void* wqkv_packed, const float* bqkv_packed, void* o, int64_t M, int L, int d, int H,
                                    int causal, idb200_stream_t stream) {
    IDB_REQUIRE(a != nullptr, IDB200_EINVAL, "NULL pointer");
    return qkv_attention_impl(a, nullptr, nullptr, nullptr, nullptr, 0, wqkv_packed, bqkv_packed, o, M, L, d, H, causal, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_ln_qkv_attention(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                                       const void* wqkv_packed, const float* bqkv_packed, void* o, int64_t M, int L, int d, int H, int causal,
                                       idb200_stream_t stream) {
    IDB_REQUIRE(h != nullptr, IDB200_EINVAL, "NULL pointer");
    return qkv_attention_impl(nullptr, h, ln_w, ln_b, gamma_beta, gb_stride, wqkv_packed, bqkv_packed, o, M, L, d, H, causal,
                              static_cast<cudaStream_t>(stream));
}
