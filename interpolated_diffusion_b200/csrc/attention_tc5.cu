// K3b-tc5: multi-head self attention (bidirectional or causal, head_dim 32, any L <= 256) on the 5th-generation tensor
// cores: S = Q K^T and O = P V are tcgen05.mma with the score tile and the output accumulator in tensor memory; the softmax
// runs on CUDA cores between the two, thread <-> query row (the accumulator's natural layout: row statistics are in-thread,
// no shuffles), P goes back to the tensor core as a bf16 SWIZZLE_128B A operand in shared memory.
//
// Reference: nn.MultiheadAttention inside src/models/transformer.py:11,39 on the packed projections
// qkv [M, 3d] (columns [q | k | v], head h at columns 32h of each third): softmax(Q K^T / sqrt(32) [+ causal mask]) V
// per (trajectory, head); the -inf upper-triangular mask of transformer.py:68-71 is the `causal` flag.
//
// Work unit = (128-token tile, head PAIR).  A head pair is 64 columns = one 128-byte swizzle row, so one TMA box
// [128 tokens x 64 columns] per operand feeds both heads: head j of the pair is k-steps 2j, 2j+1 of the Q / K tiles
// (K-major operands) and columns 32j.. of the V tile, which is read in place as an MN-major B operand (keys are the
// reduction dimension of P V) -- no transposes, no per-head staging.
//   L <= 128 (kNK = 128): a tile holds G = floor(128 / L) whole trajectories; scores between different trajectories of
//                         the tile are masked (block diagonal), keys == the tile's tokens.  Two load stages.
//   L  > 128 (kNK = 256): unit = (trajectory, head pair): K / V (256 key rows, masked beyond L) are loaded once and
//                         serve the ceil(L / 128) query tiles; a causal first query tile only multiplies 128 keys.
// 384 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11 softmax (two groups of four
// warps; group g <-> head g of the pair, warp w <-> TMEM lanes 32 (w % 4)..).  Per group the chain is
//   S_g = Q_g K_g^T (tensor)  ->  softmax_g (CUDA cores)  ->  O_g = P_g V_g (tensor, accumulator aliased onto the first
//   columns of S_g)  ->  O_g / rowsum -> bf16 -> global
// and the two groups' chains interleave, so the tensor pipe, the MUFU / FMA pipes and TMA overlap.
#include <cstdlib>

#include "tc_common.cuh"

namespace idb200 {
using namespace tc;

namespace at5 {
#ifdef IDB200_ATT5_DEBUG
#define AT5_DBG(...) do { if (blockIdx.x == 0 && lane == 0) printf(__VA_ARGS__); } while (0)
#else
#define AT5_DBG(...) do { } while (0)
#endif
constexpr int kThreads = 384;
constexpr int kTile = 128 * 64 * 2;                 // [128 x 64] bf16 SWIZZLE_128B block
constexpr float kScaleLog2 = 0.17677669529663687f * 1.4426950408889634f;     // log2(e) / sqrt(32)

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kCtas = CTAs per SM.  kNK = 128: two co-resident CTAs (one load stage each, 256 TMEM columns each) overlap each other's
// MMA <-> softmax hand-offs -- best when the softmax dominates (L >= 32: 0.61 -> 0.57 ms at L = 64, H = 12, B = 8192); one CTA
// with two load stages is better when the unit is load-bound (L = 8: 48 KB per unit for 8 keys per row; 0.43 vs 0.55 ms).
template <int kNK, int kCtas>
struct Cfg {
    static constexpr int kStages = (kNK == 128 && kCtas == 1) ? 2 : 1;
    static constexpr int kCtasPerSm = kCtas;
    static constexpr int kTmemCols = 2 * kNK;                                 // S_g at column g * kNK (O_g aliased onto its first columns)
    static constexpr int kQTiles = kNK == 128 ? 1 : 2;
    static constexpr int kQBytes = kQTiles * kTile;
    static constexpr int kKBytes = kNK * 128;
    static constexpr int kStageBytes = kQBytes + 2 * kKBytes;
    static constexpr int kPBytes = 128 * kNK * 2;                            // one group's P tile: kNK / 64 k-blocks
    static constexpr int kOffP = kStages * kStageBytes;
    static constexpr int kOffBar = kOffP + 2 * kPBytes;
    static constexpr int kSmem = kOffBar + 256 + 1024;
    static_assert(kSmem <= 232448, "shared memory budget");
};

struct Params {
    __nv_bfloat16* out;         // [M, d]
    long long B;                // trajectories
    int L, H, causal;
    int G;                      // kNK = 128: trajectories per tile
    long long tiles;            // kNK = 128: ceil(B / G); kNK = 256: B
    int pv_wide;                // 1: P V with N = 64 (both heads' V columns; each group keeps its 32): fallback form
};

template <int kNK, int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) attn_tc5_kernel(const __grid_constant__ CUtensorMap tm_qkv, const Params p) {
    using C = Cfg<kNK, kCtas>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* qk_full = bars;                       // [2] TMA -> MMA
    uint64_t* v_full = bars + 2;                    // [2]
    uint64_t* qk_empty = bars + 4;                  // [2] MMA -> TMA (commit after the unit's last S MMA)
    uint64_t* v_empty = bars + 6;                   // [2] (commit after the unit's last P V MMA)
    uint64_t* s_full = bars + 8;                    // [2 groups] MMA -> softmax
    uint64_t* p_full = bars + 10;                   // [2] softmax -> MMA (4 warp arrivals)
    uint64_t* o_full = bars + 12;                   // [2] MMA -> softmax
    uint64_t* s_free = bars + 14;                   // [2] softmax -> MMA: O read out, S / O columns reusable (4 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.H * 32;
    const int HP = p.H / 2;
    const long long units = p.tiles * HP;
    const int nq = kNK == 128 ? 1 : (p.L + 127) / 128;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&qk_empty[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&s_free[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, C::kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (first token row of the Q / K tiles, head pair)
    auto unit_rows = [&](long long u, long long& row0, int& hp) {
        const long long t = u / HP;
        hp = static_cast<int>(u - t * HP);
        row0 = kNK == 128 ? t * p.G * p.L : t * p.L;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            long long it = 0;
            for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
                const int st = static_cast<int>(it % C::kStages);
                const uint32_t ph = static_cast<uint32_t>((it / C::kStages) & 1);
                long long row0;
                int hp;
                unit_rows(u, row0, hp);
                uint8_t* base = smem + st * C::kStageBytes;
                mbar_wait(&qk_empty[st], ph ^ 1, 10);
                mbar_arrive_expect_tx(&qk_full[st], nq * kTile + C::kKBytes);
                for (int q = 0; q < nq; ++q) tma_load_2d(base + q * kTile, &tm_qkv, &qk_full[st], hp * 64, static_cast<int>(row0 + q * 128));
                for (int k = 0; k < kNK / 128; ++k)
                    tma_load_2d(base + C::kQBytes + k * kTile, &tm_qkv, &qk_full[st], d + hp * 64, static_cast<int>(row0 + k * 128));
                mbar_wait(&v_empty[st], ph ^ 1, 11);
                mbar_arrive_expect_tx(&v_full[st], C::kKBytes);
                for (int k = 0; k < kNK / 128; ++k)
                    tma_load_2d(base + C::kQBytes + C::kKBytes + k * kTile, &tm_qkv, &v_full[st], 2 * d + hp * 64, static_cast<int>(row0 + k * 128));
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the schedule; one elected lane issues) =====================
        uint32_t n[2] = {0, 0};
        long long it = 0;
        for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            const int st = static_cast<int>(it % C::kStages);
            const uint32_t ph = static_cast<uint32_t>((it / C::kStages) & 1);
            const uint32_t sQ = smem_u32(smem + st * C::kStageBytes), sK = sQ + C::kQBytes, sV = sK + C::kKBytes;
            mbar_wait(&qk_full[st], ph, 20);
            tc_fence_after();
            for (int qt = 0; qt < nq; ++qt) {
                const int keff = (kNK == 256 && p.causal && qt == 0) ? 128 : kNK;
                const uint32_t idesc_s = umma_idesc_bf16(128, keff);
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&s_free[g], (n[g] & 1) ^ 1, 21);           // O of this group's previous problem has been read out
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(sQ + qt * kTile) + 4 * g;
                    const uint64_t bd = umma_desc_sw128(sK) + 4 * g;
                    if (elect_one_sync()) {
                        umma_bf16(tmem_base + g * kNK, ad, bd, idesc_s, 0u);
                        umma_bf16(tmem_base + g * kNK, ad + 2, bd + 2, idesc_s, 1u);
                        umma_commit(&s_full[g]);
                    }
                    __syncwarp();
                    AT5_DBG("mma issued S g=%d keff=%d\n", g, keff);
                }
                if (qt == nq - 1) {
                    if (elect_one_sync()) umma_commit(&qk_empty[st]);   // Q / K of this stage are free once the S MMAs complete
                    __syncwarp();
                }
                if (qt == 0) {
                    mbar_wait(&v_full[st], ph, 22);
                    tc_fence_after();
                }
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(&p_full[g], n[g] & 1, 23);                 // P_g is in shared memory (and S_g has been consumed)
                    tc_fence_after();
                    const uint32_t sP = smem_u32(smem + C::kOffP + g * C::kPBytes);
                    const uint32_t idesc_o = (p.pv_wide ? umma_idesc_bf16(128, 64) : umma_idesc_bf16(128, 32)) | (1u << 16);   // B MN-major
                    const uint64_t vd = umma_desc_sw128_mn(sV) + (p.pv_wide ? 0 : 4 * g);
                    if (elect_one_sync()) {
                        for (int ks = 0; ks < keff / 16; ++ks) {
                            const uint64_t pd = umma_desc_sw128(sP + (ks >> 2) * kTile) + 2 * (ks & 3);
                            umma_bf16(tmem_base + g * kNK, pd, vd + 128 * ks, idesc_o, ks > 0 ? 1u : 0u);
                        }
                        umma_commit(&o_full[g]);
                    }
                    __syncwarp();
                    ++n[g];
                }
            }
            if (elect_one_sync()) umma_commit(&v_empty[st]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== softmax warps =====================
        const int g = (warp - 4) >> 2;                                   // group <-> head of the pair
        const int q = warp & 3;                                          // TMEM lane quadrant
        const int row = q * 32 + lane;
        const uint32_t tS = tmem_base + g * kNK + (static_cast<uint32_t>(q * 32) << 16);
        // (explicit st.shared below: a __builtin_assume(__isShared()) on this derived pointer made the compiler treat the whole
        // branch as unreachable and delete it)
        const uint32_t sPw = smem_u32(smem + C::kOffP + g * C::kPBytes);
        const int L = p.L;
        uint32_t n = 0;
        for (long long u = blockIdx.x; u < units; u += gridDim.x) {
            long long row0;
            int hp;
            unit_rows(u, row0, hp);
            const long long tile = u / HP;
            for (int qt = 0; qt < nq; ++qt, ++n) {
                const int keff = (kNK == 256 && p.causal && qt == 0) ? 128 : kNK;
                // this thread's query row: token, validity, and the (contiguous) range of key columns it attends to
                int lo = 0, hi = 0;
                long long token = -1;
                if (kNK == 128) {
                    const int j = row / L;
                    if (j < p.G && tile * p.G + j < p.B) {
                        lo = j * L;
                        hi = p.causal ? row + 1 : lo + L;
                        token = row0 + row;
                    }
                } else {
                    const int qpos = qt * 128 + row;
                    if (qpos < L) {
                        hi = p.causal ? qpos + 1 : L;
                        token = row0 + qpos;
                    }
                }
                const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? (lo & ~31) : 0x7fffffff);
                const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
                AT5_DBG("sm w%d wait s_full n=%u wlo=%d whi=%d\n", warp, n, wlo, whi);
                mbar_wait(&s_full[g], n & 1, 30);
                tc_fence_after();
                AT5_DBG("sm w%d got s_full\n", warp);
                // ---- pass 1: row maximum over the valid columns ----
                float mx = -INFINITY;
                for (int c0 = wlo; c0 < whi; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tS + c0, r);
                    tmem_ld_wait();
                    if (__all_sync(0xffffffffu, c0 >= lo && c0 + 32 <= hi)) {      // every row of the warp attends to the whole chunk
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int c = c0 + j;
                            if (c >= lo && c < hi) mx = fmaxf(mx, __uint_as_float(r[j]));
                        }
                    }
                }
                const float mneg = (mx == -INFINITY) ? 0.0f : mx * kScaleLog2;
                AT5_DBG("sm w%d pass1 done mx=%f\n", warp, mx);
                // ---- pass 2: p = exp2((s - max) * log2e / sqrt(32)); row sum; bf16 P tile (zeros where masked) ----
                float sum = 0.0f;
                for (int c0 = 0; c0 < keff; c0 += 32) {
                    uint32_t pk[16];
                    if (c0 >= wlo && c0 < whi) {
                        uint32_t r[32];
                        tmem_ld_32x32(tS + c0, r);
                        tmem_ld_wait();
                        const bool full = __all_sync(0xffffffffu, c0 >= lo && c0 + 32 <= hi);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = c0 + 2 * j;
                            float e0 = ex2(fmaf(__uint_as_float(r[2 * j]), kScaleLog2, -mneg));
                            float e1 = ex2(fmaf(__uint_as_float(r[2 * j + 1]), kScaleLog2, -mneg));
                            if (!full) {
                                e0 = (c >= lo && c < hi) ? e0 : 0.0f;
                                e1 = (c + 1 >= lo && c + 1 < hi) ? e1 : 0.0f;
                            }
                            sum += e0 + e1;
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                            pk[j] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = 0u;
                    }
                    const uint32_t blk = sPw + (c0 >> 6) * kTile;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + sw128_offset(row, (c0 & 63) + 8 * i)), "r"(pk[4 * i]),
                                     "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                                     : "memory");
                }
                tc_fence_before();                                       // the TMEM reads of S are done before the MMA overwrites it with O
                fence_proxy_async_smem();                                // P (generic-proxy writes) -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);
                AT5_DBG("sm w%d arrived p_full sum=%f\n", warp, sum);
                // ---- O_g / rowsum -> bf16 -> global ----
                mbar_wait(&o_full[g], n & 1, 31);
                tc_fence_after();
                {
                    uint32_t r[32];
                    tmem_ld_32x32(tS + (p.pv_wide ? g * 32 : 0), r);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[g]);              // registers hold O: the group's TMEM columns are free
                    if (token >= 0) {
                        const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
                        uint4* dst = reinterpret_cast<uint4*>(p.out + token * d + hp * 64 + g * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t w[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2 * k]) * inv, __uint_as_float(r[8 * i + 2 * k + 1]) * inv);
                                w[k] = *reinterpret_cast<uint32_t*>(&b2);
                            }
                            dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

template <int kNK, int kCtas>
static int launch(const CUtensorMap& tm, const Params& p, long long units, cudaStream_t st) {
    using C = Cfg<kNK, kCtas>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc5_kernel<kNK, kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(attn_tc5, smem=%d): %s", C::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    const long long cap = static_cast<long long>(num_sms()) * C::kCtasPerSm;
    const int grid = static_cast<int>(units < cap ? units : cap);
    attn_tc5_kernel<kNK, kCtas><<<grid, kThreads, C::kSmem, st>>>(tm, p);
    return check_launch("attn_tc5_kernel");
}


// ------------------------------------------------------------------------------------------------
// Double-buffered form for L <= 128 (one CTA per SM, 640 threads).  In the kernel above a group's chain
//   S MMA -> softmax -> P V MMA -> O read-out -> (next) S MMA
// is serial: four MMA <-> CUDA-core hand-offs per problem, and at L = 8 (8 keys per row) the hand-offs ARE the time
// (3.4 us per unit measured against 1.2 us of HBM time).  Here every group has TWO score / output buffers in tensor memory
// (4 x 128 columns = all 512) and two P tiles, the O read-out belongs to separate epilogue warps, and the MMA warp issues
// S(u+1) before P V(u): the softmax warps of a group run back to back, the read-out of O(u-1) and the MMAs of u+1 hide under
// softmax(u).  The row sums travel from the softmax to the epilogue warps through a dead score column of the same buffer.
// Warps: 0 TMA, 1 MMA, 2 TMEM alloc, 4..11 softmax (group = (w-4)/4), 12..19 epilogue (group = (w-12)/4).
// ------------------------------------------------------------------------------------------------
constexpr int kDbThreads = 640;
struct DbCfg {
    static constexpr int kStageBytes = 3 * kTile;                            // Q | K | V tiles [128 x 64]
    static constexpr int kPBytes = 2 * kTile;                                // [128 x 128] bf16
    static constexpr int kOffP = 2 * kStageBytes;
    static constexpr int kOffBar = kOffP + 4 * kPBytes;
    static constexpr int kSmem = kOffBar + 256 + 1024;
    static_assert(kSmem <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}

__global__ void __launch_bounds__(kDbThreads, 1) attn_tc5_db_kernel(const __grid_constant__ CUtensorMap tm_qkv, const Params p) {
    using C = DbCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* qk_full = bars;                       // [2 stages]
    uint64_t* v_full = bars + 2;
    uint64_t* qk_empty = bars + 4;
    uint64_t* v_empty = bars + 6;
    uint64_t* s_full = bars + 8;                    // [group * 2 + buffer]
    uint64_t* p_full = bars + 12;
    uint64_t* o_full = bars + 16;
    uint64_t* s_free = bars + 20;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.H * 32;
    const int HP = p.H / 2;
    const long long units = p.tiles * HP;
    const int L = p.L;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&qk_empty[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&s_free[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto unit_rows = [&](long long u, long long& row0, int& hp) {
        const long long t = u / HP;
        hp = static_cast<int>(u - t * HP);
        row0 = t * p.G * L;
    };

    if (warp == 0) {
        if (lane == 0) {
            long long it = 0;
            for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
                const int st = static_cast<int>(it & 1);
                const uint32_t ph = static_cast<uint32_t>((it >> 1) & 1);
                long long row0;
                int hp;
                unit_rows(u, row0, hp);
                uint8_t* base = smem + st * C::kStageBytes;
                mbar_wait(&qk_empty[st], ph ^ 1, 10);
                mbar_arrive_expect_tx(&qk_full[st], 2 * kTile);
                tma_load_2d(base, &tm_qkv, &qk_full[st], hp * 64, static_cast<int>(row0));
                tma_load_2d(base + kTile, &tm_qkv, &qk_full[st], d + hp * 64, static_cast<int>(row0));
                mbar_wait(&v_empty[st], ph ^ 1, 11);
                mbar_arrive_expect_tx(&v_full[st], kTile);
                tma_load_2d(base + 2 * kTile, &tm_qkv, &v_full[st], 2 * d + hp * 64, static_cast<int>(row0));
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);          // B (V) MN-major
        long long n_units = 0;
        for (long long u = blockIdx.x; u < units; u += gridDim.x) ++n_units;
        auto pv_phase = [&](long long it) {                                       // P V of the unit with local index `it`
            const int st = static_cast<int>(it & 1);
            const uint32_t ph = static_cast<uint32_t>((it >> 1) & 1);
            const int b = static_cast<int>(it & 1);
            const uint32_t sV = smem_u32(smem + st * C::kStageBytes + 2 * kTile);
            mbar_wait(&v_full[st], ph, 22);
            tc_fence_after();
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&p_full[g * 2 + b], ph, 23);
                tc_fence_after();
                const uint32_t sP = smem_u32(smem + C::kOffP + (g * 2 + b) * C::kPBytes);
                const uint64_t vd = umma_desc_sw128_mn(sV) + 4 * g;
                const uint32_t dcol = tmem_base + (g * 2 + b) * 128;
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(dcol, umma_desc_sw128(sP + (ks >> 2) * kTile) + 2 * (ks & 3), vd + 128 * ks, idesc_o, ks > 0 ? 1u : 0u);
                    umma_commit(&o_full[g * 2 + b]);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(&v_empty[st]);
            __syncwarp();
        };
        for (long long it = 0; it < n_units; ++it) {
            const int st = static_cast<int>(it & 1);
            const uint32_t ph = static_cast<uint32_t>((it >> 1) & 1);
            const int b = static_cast<int>(it & 1);
            const uint32_t sQ = smem_u32(smem + st * C::kStageBytes), sK = sQ + kTile;
            mbar_wait(&qk_full[st], ph, 20);
            tc_fence_after();
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&s_free[g * 2 + b], ph ^ 1, 21);                        // O of problem it - 2 has been read out of this buffer
                tc_fence_after();
                const uint64_t ad = umma_desc_sw128(sQ) + 4 * g, bd = umma_desc_sw128(sK) + 4 * g;
                const uint32_t dcol = tmem_base + (g * 2 + b) * 128;
                if (elect_one_sync()) {
                    umma_bf16(dcol, ad, bd, idesc_s, 0u);
                    umma_bf16(dcol, ad + 2, bd + 2, idesc_s, 1u);
                    umma_commit(&s_full[g * 2 + b]);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(&qk_empty[st]);
            __syncwarp();
            if (it > 0) pv_phase(it - 1);
        }
        if (n_units > 0) pv_phase(n_units - 1);
    } else if (warp >= 4 && warp < 12) {
        // ===================== softmax warps =====================
        const int g = (warp - 4) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        long long it = 0;
        for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            const int b = static_cast<int>(it & 1);
            const uint32_t ph = static_cast<uint32_t>((it >> 1) & 1);
            const uint32_t tS = tmem_base + (g * 2 + b) * 128 + (static_cast<uint32_t>(q * 32) << 16);
            const uint32_t sPw = smem_u32(smem + C::kOffP + (g * 2 + b) * C::kPBytes);
            const long long tile = u / HP;
            int lo = 0, hi = 0;
            {
                const int j = row / L;
                if (j < p.G && tile * p.G + j < p.B) {
                    lo = j * L;
                    hi = p.causal ? row + 1 : lo + L;
                }
            }
            const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? (lo & ~31) : 0x7fffffff);
            const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
            mbar_wait(&s_full[g * 2 + b], ph, 30);
            tc_fence_after();
            float mx = -INFINITY;
            for (int c0 = wlo; c0 < whi; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tS + c0, r);
                tmem_ld_wait();
                if (__all_sync(0xffffffffu, c0 >= lo && c0 + 32 <= hi)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = c0 + j;
                        if (c >= lo && c < hi) mx = fmaxf(mx, __uint_as_float(r[j]));
                    }
                }
            }
            const float mneg = (mx == -INFINITY) ? 0.0f : mx * kScaleLog2;
            float sum = 0.0f;
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t pk[16];
                if (c0 >= wlo && c0 < whi) {
                    uint32_t r[32];
                    tmem_ld_32x32(tS + c0, r);
                    tmem_ld_wait();
                    const bool full = __all_sync(0xffffffffu, c0 >= lo && c0 + 32 <= hi);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + 2 * j;
                        float e0 = ex2(fmaf(__uint_as_float(r[2 * j]), kScaleLog2, -mneg));
                        float e1 = ex2(fmaf(__uint_as_float(r[2 * j + 1]), kScaleLog2, -mneg));
                        if (!full) {
                            e0 = (c >= lo && c < hi) ? e0 : 0.0f;
                            e1 = (c + 1 >= lo && c + 1 < hi) ? e1 : 0.0f;
                        }
                        sum += e0 + e1;
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
                        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = 0u;
                }
                const uint32_t blk = sPw + (c0 >> 6) * kTile;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + sw128_offset(row, (c0 & 63) + 8 * i)), "r"(pk[4 * i]),
                                 "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                                 : "memory");
            }
            tmem_st_32x1(tS + 32, __float_as_uint(sum));                   // row sum -> a dead score column (O only takes columns 0..31)
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g * 2 + b]);
        }
    } else if (warp >= 12) {
        // ===================== epilogue warps: O / rowsum -> bf16 -> global =====================
        const int g = (warp - 12) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        long long it = 0;
        for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            const int b = static_cast<int>(it & 1);
            const uint32_t ph = static_cast<uint32_t>((it >> 1) & 1);
            const uint32_t tS = tmem_base + (g * 2 + b) * 128 + (static_cast<uint32_t>(q * 32) << 16);
            long long row0;
            int hp;
            unit_rows(u, row0, hp);
            const long long tile = u / HP;
            const int j = row / L;
            const bool live = j < p.G && tile * p.G + j < p.B;
            mbar_wait(&o_full[g * 2 + b], ph, 31);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tS, r);
            const float sum = __uint_as_float(tmem_ld_32x1(tS + 32));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g * 2 + b]);
            if (live) {
                const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
                uint4* dst = reinterpret_cast<uint4*>(p.out + (row0 + row) * d + hp * 64 + g * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2 * k]) * inv, __uint_as_float(r[8 * i + 2 * k + 1]) * inv);
                        w[k] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int launch_db(const CUtensorMap& tm, const Params& p, long long units, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc5_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DbCfg::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(attn_tc5_db, smem=%d): %s", DbCfg::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    const int grid = static_cast<int>(units < num_sms() ? units : num_sms());
    attn_tc5_db_kernel<<<grid, kDbThreads, DbCfg::kSmem, st>>>(tm, p);
    return check_launch("attn_tc5_db_kernel");
}

}  // namespace at5

// bf16 qkv [B*L, 3d] -> out bf16 [B*L, d]; H even, L <= 256.  Returns IDB200_EUNSUPPORTED (without setting an error the
// caller would surface) only through the explicit checks below: the dispatcher in attention.cu tests `attention_tc5_supported`.
bool attention_tc5_supported(const void* qkv, const void* out, long long B, int L, int H) {
    return (H % 2 == 0) && L >= 1 && L <= 256 && aligned(qkv, 16) && aligned(out, 16) && B * L < (1ll << 31);
}

int attention_tc5(const void* qkv, void* out, long long B, int L, int H, int causal, cudaStream_t st) {
    IDB_REQUIRE(attention_tc5_supported(qkv, out, B, L, H), IDB200_EUNSUPPORTED, "tcgen05 attention needs an even head count, L <= 256, 16-byte aligned buffers");
    const int d = H * 32;
    CUtensorMap tm;
    int rc = make_tmap_bf16_2d(&tm, qkv, static_cast<uint64_t>(B) * L, static_cast<uint64_t>(3) * d, 128, 64);
    if (rc) return rc;
    static const int wide = getenv("IDB200_ATT5_WIDE") ? atoi(getenv("IDB200_ATT5_WIDE")) : 0;
    at5::Params p{};
    p.out = static_cast<__nv_bfloat16*>(out);
    p.B = B; p.L = L; p.H = H; p.causal = causal; p.pv_wide = wide;
    if (L <= 128) {
        p.G = 128 / L;
        p.tiles = (B + p.G - 1) / p.G;
        static const int db = getenv("IDB200_ATT5_DB") ? atoi(getenv("IDB200_ATT5_DB")) : 1;
        if (db) return at5::launch_db(tm, p, p.tiles * (H / 2), st);
        if (L < 32) return at5::launch<128, 1>(tm, p, p.tiles * (H / 2), st);
        return at5::launch<128, 2>(tm, p, p.tiles * (H / 2), st);
    }
    p.G = 1;
    p.tiles = B;
    return at5::launch<256, 1>(tm, p, p.tiles * (H / 2), st);
}

}  // namespace idb200
