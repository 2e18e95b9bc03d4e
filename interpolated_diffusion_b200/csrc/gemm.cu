// K3a: token GEMM  out[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)  on tcgen05 / TMEM fed by TMA.
//
// This is the dense building block of the two denoisers (reference: the nn.Linear / packed-QKV /
// out_proj / ff GEMMs inside src/models/transformer.py:35-46, src/models/denoiser_keypoints.py:82-113,
// src/models/denoiser_interp_levels.py:64-84, which today run as cuBLAS calls from PyTorch eager).
//
// Structure (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: A tile [128 x 64] and W tile [BN x 64] bf16 per k-block, SWIZZLE_128B,
//               kStages-deep mbarrier ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per k-block,
//               accumulating in TMEM; tcgen05.commit frees the smem slot / publishes the accumulator
//   warp 2      TMEM allocator (512 columns = two BN-wide accumulator stages)
//   warps 4..11 epilogue: tcgen05.ld (thread <-> output row), bias / SiLU / residual, global stores;
//               overlaps the next tile's MMAs through the second accumulator stage
// M is arbitrary (TMA zero-fills out-of-range rows, stores are row-guarded); N % BN == 0; K % 64 == 0.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace idb200 {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint32_t box_rows,
                 uint32_t box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(IDB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if (!aligned(base, 16) || (cols * elem_bytes) % 16 != 0) return fail(IDB200_EALIGN, "TMA needs 16-byte aligned base and row pitch");
    if (box_cols * elem_bytes != 128) return fail(IDB200_EINVAL, "SWIZZLE_128B boxes are 128 bytes wide");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = (elem_bytes == 2) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IDB200_ECUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return IDB200_OK;
}

// bf16 [rows, cols] view with an explicit row pitch (bytes, multiple of 16); pitch < cols * 2 gives OVERLAPPING rows (row r =
// elements [r * pitch / 2, r * pitch / 2 + cols)): legal for TMA loads, used by the implicit conv to read two adjacent
// 32-channel pixels as one 64-wide k-block
int make_tmap_bf16_2d_pitch(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                            uint32_t box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(IDB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if (!aligned(base, 16) || pitch_bytes % 16 != 0) return fail(IDB200_EALIGN, "TMA needs 16-byte aligned base and row pitch");
    if (box_cols * 2 != 128) return fail(IDB200_EINVAL, "SWIZZLE_128B boxes are 128 bytes wide");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IDB200_ECUDA, "cuTensorMapEncodeTiled (pitched) failed (%d)", static_cast<int>(r));
    return IDB200_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, base, 2, rows, cols, box_rows, box_cols);
}

using namespace tc;

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int kBiasFloats = 2048;       // bias vector staged in shared memory (N <= 2048)

// 4 / 5: training forms with a second bf16 [M,N] tensor `aux` (idb200_gemm_bf16_aux):
//   EPI_BF16_SILU_DUAL  out = bf16(acc + bias) (the pre-activation u, kept for the backward), aux = bf16(SiLU(out)) -- ff.0 forward
//   EPI_BF16_DSILU      out = bf16(bf16(acc) * SiLU'(aux)), aux = u read through TMA -- dU = (dY W2) . SiLU'(u) of the backward
// Both reproduce the separate SiLU kernels of csrc/train_bwd.cu bit for bit (same roundings, the one-MUFU silu16_* of common.cuh).
enum { EPI_BF16 = 0, EPI_SILU_BF16 = 1, EPI_RESID_F32 = 2, EPI_F32 = 3, EPI_BF16_SILU_DUAL = 4, EPI_BF16_DSILU = 5 };
constexpr int kConvMaxKB = 36;         // k-blocks of the implicit conv: 9 taps x C / 64 (C <= 256)

// kPair: two CTAs of a cluster work on one [256 x BN] tile with tcgen05 cta_group::2: each CTA stages its own 128 rows of A but
// only BN/2 rows of W, so the L2 -> shared-memory fill per flop drops by (128 + BN) / (128 + BN/2) (the unpaired kernel is bound
// by that fill: 40 KB per k-block at BN = 192 against 72 B/ns per SM measured -> at most ~49 % of the tensor peak).
template <int BN, bool kPair = false>
struct GemmCfg {
    static constexpr int kBRows = kPair ? BN / 2 : BN;                       // rows of W staged by one CTA
    static constexpr int kStageBytes = kBM * kBK * 2 + kBRows * kBK * 2;
    // shared memory: [ring | barriers 256 | bias 8 KB | pad | 4 x 16 KB output staging slabs (TMA-store epilogue)]
    static constexpr int kStagingBytes = 4 * 16384;
    static constexpr int kFixedBytes = 1024 /*align*/ + 256 /*barriers*/ + kBiasFloats * 4 + 1024 /*align*/ + kStagingBytes;
    static constexpr int kFit = (227 * 1024 - kFixedBytes) / kStageBytes;
    static constexpr int kStages = kFit < 8 ? kFit : 8;
    static constexpr int kAccStages = (2 * BN <= 512) ? 2 : 1;
    static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
};

struct GemmParams {
    const float* bias;      // [N] or nullptr
    void* out;              // bf16 [M,N] (EPI_BF16 / EPI_SILU_BF16), fp32 [M,N] (EPI_F32), fp32 residual in/out (EPI_RESID_F32)
    long long M;
    int N, K, epilogue;
    int splits;             // split-K: tile = (split, m, n); split s covers K/splits of the reduction, out += s * M * N (EPI_F32)
    int tma_out;            // 1: the epilogue stages 16 KB slabs in shared memory and writes them with TMA (tmap_out)
    // implicit 3x3 convolution ("tap-shifted" GEMM, conv_P2 > 0): A is a zero-bordered NHWC activation [B * P2, C] (P2 = (H+2) *
    // (W+2) padded positions per image); k-block kb reads the A box at row m0 + conv_shift[kb], column conv_col[kb] -- the input
    // pixel of tap (ky, kx) for output position m is row m + (ky-1) * PW + (kx-1) -- so no im2col matrix exists.  The epilogue
    // writes zeros to border positions (the next layer's zero padding).
    int conv_P2, conv_PW, conv_PH;
    int conv_shift[kConvMaxKB];
    int conv_col[kConvMaxKB];
    // EPI_BF16_DSILU only (nullptr otherwise): colpart [8 * ceil(M / 256), N] fp32 receives, per 128-row block and epilogue warp, the
    // column sums of the bf16-rounded output over the warp's 32 rows -- summed over its rows this is the bias gradient of ff.0
    // (column sums of dU), which otherwise costs a full extra read of dU.
    float* colpart;
    int out_f16;            // EPI_BF16 only: the 16-bit output is IEEE half instead of bf16 (epilogue | 0x100 at the entry point)
};


// eight fp32 -> one 16-byte chunk of bf16 (or IEEE half) pairs
__device__ __forceinline__ uint4 pack8_16(const float* v, bool f16) {
    uint4 pk;
    if (f16) {
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]), h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    return pk;
}

// SiLU(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2: one MUFU (tanh.approx, rel. error ~2^-11, below the bf16
// rounding of the stored result) instead of ex2 + rcp.
__device__ __forceinline__ float silu_f(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// kMN = false: A [M,K], W [N,K] row-major (K-major operands): out = A W^T.
// kMN = true : both operands stored reduction-major, A [K,M], W [K,N] row-major (MN-major UMMA operands): out = A^T W.  This is
// the weight-gradient form dW = dY^T X with the tokens as K: TMA boxes are [64 tokens x 64 features] (one SW128 atom column),
// an operand tile is a row of such boxes 8 KB apart (descriptor LBO), 8-token groups 1 KB apart (SBO), and one UMMA_K = 16
// tokens advances the start address by 2 KB.
template <int BN, bool kMN, bool kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux, const GemmParams p) {
    using Cfg = GemmCfg<BN, kPair>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));             // keep LDS / STS (the integer round trip hides the address space)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                              // [kStages]  TMA -> MMA
    uint64_t* empty_bar = bars + Cfg::kStages;              // [kStages]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * Cfg::kStages;           // [kAccStages] MMA -> epilogue
    uint64_t* acc_empty = acc_full + Cfg::kAccStages;       // [kAccStages] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::kAccStages);
    uint64_t* aux_bar = acc_empty + Cfg::kAccStages + 1;   // [2 halves x 2 buffers] EPI_BF16_DSILU: a u slab has landed
    float* sbias = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);
    uint8_t* staging = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem + Cfg::kStages * Cfg::kStageBytes + 256 + kBiasFloats * 4) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(staging));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = p.N / BN;
    // pair mode: a "tile" is a pair of 128-row M tiles; CTA `rank` of the cluster owns rows (2 * pm + rank) * 128
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const long long m_tiles = kPair ? (p.M + 2 * kBM - 1) / (2 * kBM) : (p.M + kBM - 1) / kBM;
    const long long mn_tiles = m_tiles * n_tiles;
    const long long tiles = mn_tiles * p.splits;
    const int k_blocks = p.K / kBK / p.splits;             // per split
    const long long tile0 = kPair ? blockIdx.x / 2 : blockIdx.x;
    const long long tile_stride = kPair ? gridDim.x / 2 : gridDim.x;
    constexpr int kMRows = kPair ? 2 * kBM : kBM;          // rows of one scheduling tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < Cfg::kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kPair ? 16 : 8); }
        for (int i = 0; i < 4; ++i) mbar_init(&aux_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    const bool bias_smem = p.bias != nullptr && p.N <= kBiasFloats;
    if (bias_smem)
        for (int i = threadIdx.x; i < p.N; i += kGemmThreads) sbias[i] = p.bias[i];
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();               // the peer's barriers must exist before remote arrives / TMA credits
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = tile0; tile < tiles; tile += tile_stride) {
                const long long mn = tile % mn_tiles;
                const int kb0 = static_cast<int>(tile / mn_tiles) * k_blocks;
                const int m0 = static_cast<int>(mn / n_tiles) * kMRows + static_cast<int>(rank) * kBM;
                const int n0 = static_cast<int>(mn % n_tiles) * BN + static_cast<int>(rank) * Cfg::kBRows * (kPair ? 1 : 0);
                for (int kb = kb0; kb < kb0 + k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + kBM * kBK * 2;
                    if constexpr (kPair) {
                        // both CTAs' bytes are credited to the even CTA's barrier (it issues the pair MMA)
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        if constexpr (kMN) {
#pragma unroll
                            for (int hb = 0; hb < kBM / 64; ++hb) tma_load_2d_2sm(sa + hb * 8192, &tmap_a, &full_bar[stage], m0 + hb * 64, kb * kBK);
#pragma unroll
                            for (int hb = 0; hb < Cfg::kBRows / 64; ++hb) tma_load_2d_2sm(sb + hb * 8192, &tmap_w, &full_bar[stage], n0 + hb * 64, kb * kBK);
                        } else {
                            tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], p.conv_P2 ? p.conv_col[kb] : kb * kBK, p.conv_P2 ? m0 + p.conv_shift[kb] : m0);
                            tma_load_2d_2sm(sb, &tmap_w, &full_bar[stage], kb * kBK, n0);
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    if constexpr (kMN) {
#pragma unroll
                        for (int hb = 0; hb < kBM / 64; ++hb) tma_load_2d(sa + hb * 8192, &tmap_a, &full_bar[stage], m0 + hb * 64, kb * kBK);
#pragma unroll
                        for (int hb = 0; hb < BN / 64; ++hb) tma_load_2d(sb + hb * 8192, &tmap_w, &full_bar[stage], n0 + hb * 64, kb * kBK);
                    } else {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], p.conv_P2 ? p.conv_col[kb] : kb * kBK, p.conv_P2 ? m0 + p.conv_shift[kb] : m0);
                        tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * kBK, n0);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====  (the whole warp walks the schedule; one elected lane issues: tcgen05 instructions inside a
        // divergent `if (lane == 0)` compile to ELECT / BRA.U.ANY waterfall loops of ~100 cycles each, which made the issue of a
        // k-block (4 MMAs + commit) as long as its execution)
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kMRows, BN) | (kMN ? ((1u << 15) | (1u << 16)) : 0u);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile0; tile < tiles; tile += tile_stride) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1, 2);      // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);         // TMA bytes landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = kMN ? umma_desc_sw128_mn(sa) : umma_desc_sw128(sa);
                    const uint64_t bdesc = kMN ? umma_desc_sw128_mn(sa + kBM * kBK * 2) : umma_desc_sw128(sa + kBM * kBK * 2);
                    constexpr int kStep = kMN ? 128 : 2;            // 16-byte units per UMMA_K
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            if constexpr (kPair) umma_bf16_2sm(d_tmem, adesc + kStep * k, bdesc + kStep * k, idesc, (kb | k) ? 1u : 0u);
                            else umma_bf16(d_tmem, adesc + kStep * k, bdesc + kStep * k, idesc, (kb | k) ? 1u : 0u);
                        }
                        if constexpr (kPair) umma_commit_2sm(&empty_bar[stage]);   // both CTAs' slots free once these MMAs retire
                        else umma_commit(&empty_bar[stage]);       // smem slot free once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) {
                    if constexpr (kPair) umma_commit_2sm(&acc_full[acc]);
                    else umma_commit(&acc_full[acc]);              // accumulator complete
                }
                __syncwarp();
                if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warps (4 + q) and (8 + q) share TMEM lanes [32q, 32q+32) and alternate
        // over the 32-column chunks, so every SM sub-partition has two epilogue warps to hide TMEM / MUFU latency
        const int ew = warp - 4;
        const int q = ew & 3, half = ew >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t slab_it = 0;
        const bool leader = q == 0 && lane == 0;               // issues this half's TMA stores / aux loads
        // EPI_BF16_DSILU: the u slab of (tile t, slab sl) of this half lands in staging buffer (it & 1) of the half
        auto aux_issue = [&](long long t, int sl, uint32_t it) {
            const long long mn_ = t % mn_tiles;
            const int m0_ = static_cast<int>(mn_ / n_tiles) * kMRows + static_cast<int>(rank) * kBM;
            const int n0_ = static_cast<int>(mn_ % n_tiles) * BN;
            uint64_t* bar = &aux_bar[half * 2 + (it & 1)];
            mbar_arrive_expect_tx(bar, 16384);
            tma_load_2d(staging + half * 2 * 16384 + (it & 1) * 16384, &tmap_aux, bar, n0_ + sl * 64, m0_);
        };
        if (p.epilogue == EPI_BF16_DSILU && leader && tile0 < tiles && half < BN / 64) aux_issue(tile0, half, 0);
        for (long long tile = tile0; tile < tiles; tile += tile_stride) {
            const long long mn = tile % mn_tiles;
            const long long m0 = (mn / n_tiles) * kMRows + static_cast<long long>(rank) * kBM;
            const int n0 = static_cast<int>(mn % n_tiles) * BN;
            mbar_wait(&acc_full[acc], acc_phase, 4);
            tc_fence_after();
            const long long row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            bool border = false;                                    // implicit conv: this row is a zero-padding position
            if (p.conv_P2) {
                const int pos = static_cast<int>(row % p.conv_P2);
                const int y = pos / p.conv_PW, x = pos - y * p.conv_PW;
                border = y == 0 || y == p.conv_PH - 1 || x == 0 || x == p.conv_PW - 1;
            }
            if (p.tma_out && p.epilogue >= EPI_BF16_SILU_DUAL) {
                // Training forms with the MLP's SiLU pass folded in (one TMEM read per element, one-MUFU SiLU / SiLU').  First
                // revision: two staging passes per slab with exp + IEEE division and, for SiLU', a TMA load of the u slab issued
                // and awaited inside the slab: the step got 10 % SLOWER than with the separate HBM-bound passes (ALU-issue bound
                // epilogue + one exposed HBM round trip per slab).  Now:
                //   4: both outputs of a slab are staged at once (u -> buffer 0, SiLU(u) -> buffer 1 of the half), two TMA stores;
                //   5: the u slab of the NEXT slab (possibly of the next tile) is requested half a slab ahead into the other
                //      buffer, SiLU' is applied in place over the landed u slab.
                // The buffers are released by cp.async.bulk.wait_group.read after the first half of a slab's math, when the
                // previous slab's stores have long read them.
                const int n_slabs = BN / 64;
                const int r_in = q * 32 + lane;
                uint8_t* stage_h = staging + half * 2 * 16384;
                for (int sl = half; sl < n_slabs; sl += 2) {
                    const bool dual = p.epilogue == EPI_BF16_SILU_DUAL;
                    uint8_t* buf0 = dual ? stage_h : stage_h + (slab_it & 1) * 16384;
                    uint8_t* buf1 = stage_h + 16384;                      // dual only: SiLU(u)
                    if (!dual) mbar_wait(&aux_bar[half * 2 + (slab_it & 1)], (slab_it >> 1) & 1, 6);
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = sl * 64 + cc * 32;
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c, r);
                        tmem_ld_wait();
                        uint4 pk0[4], pk1[4];
                        if (dual) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float v[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[8 * j + k]);
                                if (bias_smem) {
                                    const float4 b0 = *reinterpret_cast<const float4*>(&sbias[n0 + c + 8 * j]);
                                    const float4 b1 = *reinterpret_cast<const float4*>(&sbias[n0 + c + 8 * j + 4]);
                                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                                    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                                } else if (p.bias) {
#pragma unroll
                                    for (int k = 0; k < 8; ++k) v[k] += __ldg(p.bias + n0 + c + 8 * j + k);
                                }
                                uint32_t* w0 = reinterpret_cast<uint32_t*>(&pk0[j]);
                                uint32_t* w1 = reinterpret_cast<uint32_t*>(&pk1[j]);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 u2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                                    const float2 uf = __bfloat1622float2(u2);
                                    const __nv_bfloat162 f2 = __floats2bfloat162_rn(silu16_fwd(uf.x), silu16_fwd(uf.y));
                                    w0[k] = *reinterpret_cast<const uint32_t*>(&u2);
                                    w1[k] = *reinterpret_cast<const uint32_t*>(&f2);
                                }
                            }
                            if (cc == 0) {
                                if (leader) tma_store_wait_read<0>();           // the previous slab's two stores are done with the buffers
                                named_barrier_sync(1 + half, 128);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int off = r_in * 128 + (((cc * 4 + j) ^ (r_in & 7)) << 4);
                                *reinterpret_cast<uint4*>(buf0 + off) = pk0[j];
                                *reinterpret_cast<uint4*>(buf1 + off) = pk1[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int off = r_in * 128 + (((cc * 4 + j) ^ (r_in & 7)) << 4);
                                const uint4 uu = *reinterpret_cast<const uint4*>(buf0 + off);
                                const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&uu);
                                uint32_t* w0 = reinterpret_cast<uint32_t*>(&pk0[j]);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 uf = __bfloat1622float2(u2[k]);
                                    const float g0 = bf16_round(__uint_as_float(r[8 * j + 2 * k]));
                                    const float g1 = bf16_round(__uint_as_float(r[8 * j + 2 * k + 1]));
                                    const __nv_bfloat162 o2 = __floats2bfloat162_rn(__fmul_rn(g0, silu16_grad(uf.x)), __fmul_rn(g1, silu16_grad(uf.y)));
                                    w0[k] = *reinterpret_cast<const uint32_t*>(&o2);
                                }
                                *reinterpret_cast<uint4*>(buf0 + off) = pk0[j];
                            }
                            if (cc == 0 && leader) {                            // request the next slab's u into the other buffer
                                int nsl = sl + 2;
                                long long nt = tile;
                                if (nsl >= n_slabs) { nsl = half; nt = tile + tile_stride; }
                                if (nt < tiles) {
                                    tma_store_wait_read<0>();                   // the previous slab's store is done reading it
                                    aux_issue(nt, nsl, slab_it + 1);
                                }
                            }
                        }
                    }
                    if (!dual && p.colpart != nullptr) {
                        // column sums of the warp's OWN 32 rows of the staged slab (only __syncwarp needed: no other warp's rows are
                        // read, and the buffer is not handed to TMA yet): lane <-> (16-byte piece = 8 columns, row group of 4);
                        // 8 conflict-free LDS.128 per thread, then two shuffle steps over the 4 row groups
                        __syncwarp();
                        const int piece = lane & 7, rg = q * 32 + (lane >> 3);
                        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rg + 4 * i;
                            const uint4 v = *reinterpret_cast<const uint4*>(buf0 + r * 128 + ((piece ^ (r & 7)) << 4));
                            const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f = __bfloat1622float2(v2[k]);
                                cs[2 * k] += f.x;
                                cs[2 * k + 1] += f.y;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
                            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
                        }
                        if (lane < 8) {
                            float* dst = p.colpart + ((m0 / kBM) * 4 + q) * p.N + n0 + sl * 64 + piece * 8;
                            *reinterpret_cast<float4*>(dst) = make_float4(cs[0], cs[1], cs[2], cs[3]);
                            *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
                        }
                    }
                    fence_proxy_async_smem();
                    named_barrier_sync(3 + half, 128);
                    if (leader) {
                        tma_store_2d(&tmap_out, buf0, n0 + sl * 64, static_cast<int>(m0));
                        if (dual) tma_store_2d(&tmap_aux, buf1, n0 + sl * 64, static_cast<int>(m0));
                        tma_store_commit();
                    }
                    ++slab_it;
                }
            } else if (p.tma_out) {
                // Coalesced epilogue.  A thread owns one ROW of the accumulator (tcgen05.ld 32x32b), so direct stores touch 32
                // different lines per instruction: measured, they (not the MMAs) bounded the kernel at ~45 % of its store-less
                // speed.  Instead each half (4 warps = 128 rows) fills a 16 KB slab (64 bf16 / 32 fp32 columns, SWIZZLE_128B
                // rows) in shared memory and one thread hands it to TMA (store, or reduce-add for the residual epilogue); two
                // slabs per half ping-pong so the conversion of slab i+1 overlaps the store of slab i.
                const bool out16 = p.epilogue == EPI_BF16 || p.epilogue == EPI_SILU_BF16;
                const int slab_cols = out16 ? 64 : 32;
                const int n_slabs = BN / slab_cols;
                const int r_in = q * 32 + lane;
                uint8_t* stage_h = staging + half * 2 * 16384;
                for (int sl = half; sl < n_slabs; sl += 2) {
                    uint8_t* buf = stage_h + (slab_it & 1) * 16384;
                    if (q == 0 && lane == 0) tma_store_wait_read<1>();          // the store that last read `buf` is done with it
                    named_barrier_sync(1 + half, 128);
                    const int nch = out16 ? 2 : 1;
                    for (int cc = 0; cc < nch; ++cc) {
                        const int c = sl * slab_cols + cc * 32;
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c, r);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (bias_smem) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 bv = *reinterpret_cast<const float4*>(&sbias[n0 + c + 4 * j]);
                                v[4 * j + 0] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
                            }
                        } else if (p.bias) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n0 + c + j);
                        }
                        if (out16) {
                            if (p.epilogue == EPI_SILU_BF16) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                            }
                            if (border) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 pk = pack8_16(&v[8 * j], p.out_f16 != 0);
                                const int piece = cc * 4 + j;
                                *reinterpret_cast<uint4*>(buf + r_in * 128 + ((piece ^ (r_in & 7)) << 4)) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4)) =
                                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                    fence_proxy_async_smem();
                    named_barrier_sync(3 + half, 128);
                    if (q == 0 && lane == 0) {
                        if (p.epilogue == EPI_RESID_F32) tma_reduce_add_2d(&tmap_out, buf, n0 + sl * slab_cols, static_cast<int>(m0));
                        else tma_store_2d(&tmap_out, buf, n0 + sl * slab_cols, static_cast<int>(m0));
                        tma_store_commit();
                    }
                    ++slab_it;
                }
            } else {
#pragma unroll 1
            for (int c = half * 32; c < BN; c += 64) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c, r);
                tmem_ld_wait();
                if (row < p.M) {
                    float v[32];
                    if (bias_smem) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bv = *reinterpret_cast<const float4*>(&sbias[n0 + c + 4 * j]);
                            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bv.x;
                            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bv.y;
                            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bv.z;
                            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bv.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            v[j] = __uint_as_float(r[j]);
                            if (p.bias) v[j] += __ldg(p.bias + n0 + c + j);
                        }
                    }
                    const long long o = (tile / mn_tiles) * p.M * p.N + row * p.N + n0 + c;
                    if (p.epilogue == EPI_BF16 || p.epilogue == EPI_SILU_BF16) {
                        if (p.epilogue == EPI_SILU_BF16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dst[j] = pack8_16(&v[8 * j], p.out_f16 != 0);
                        }
                    } else {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
                        if (p.epilogue == EPI_RESID_F32) {
                            float4 hv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = dst[j];            // all loads first (MLP)
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                dst[j] = make_float4(v[4 * j] + hv[j].x, v[4 * j + 1] + hv[j].y, v[4 * j + 2] + hv[j].z, v[4 * j + 3] + hv[j].w);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kPair) mbar_arrive_leader(&acc_empty[acc]);      // the even CTA's issuer waits for both epilogues
                else mbar_arrive(&acc_empty[acc]);
            }
            if (++acc == Cfg::kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_out && q == 0 && lane == 0) tma_store_wait_all();           // stores complete before the CTA exits
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

static bool pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("IDB200_GEMM_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

template <int BN, bool kMN, bool kPair>
static int launch_gemm_impl(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const GemmParams& p, cudaStream_t st,
                            const CUtensorMap* taux = nullptr) {
    const CUtensorMap& tx = taux ? *taux : to;
    using Cfg = GemmCfg<BN, kPair>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, kMN, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
        attr_set = true;
    }
    const long long rows = kPair ? 2 * kBM : kBM;
    const long long tiles = ((p.M + rows - 1) / rows) * (p.N / BN) * p.splits;
    if constexpr (kPair) {
        const int clusters = static_cast<int>(tiles < num_sms() / 2 ? tiles : num_sms() / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<BN, kMN, true>, ta, tw, to, tx, p);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaLaunchKernelEx(gemm pair): %s", cudaGetErrorString(e));
        return check_launch("gemm_bf16_tn_kernel<pair>");
    } else {
        const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
        gemm_bf16_tn_kernel<BN, kMN, false><<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ta, tw, to, tx, p);
        return check_launch("gemm_bf16_tn_kernel");
    }
}

// pair mode needs BN / 2 rows of W per CTA to be whole swizzle atoms (K-major: 8 rows; MN-major: 64-feature boxes) and at least
// two M tiles of work
template <int BN, bool kMN = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tw_half, const CUtensorMap& to, const GemmParams& p,
                       cudaStream_t st, const CUtensorMap* taux = nullptr) {
    constexpr bool can_pair = kMN ? (BN % 128 == 0) : (BN % 64 == 0);
    if constexpr (can_pair) {
        if (pair_enabled() && p.M > kBM) return launch_gemm_impl<BN, kMN, true>(ta, tw_half, to, p, st, taux);
    }
    return launch_gemm_impl<BN, kMN, false>(ta, tw, to, p, st, taux);
}

// out[s] [M,N] fp32 = A[K_s, M]^T W[K_s, N] for the s-th slice of the K rows (MN-major operands, see the kernel comment)
int gemm_bf16_nn_splitk(const void* A, const void* W, float* partial, long long M, int N, long long K, int splits, cudaStream_t st) {
    IDB_REQUIRE(A && W && partial, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(M > 0 && N > 0 && K > 0 && M % 8 == 0 && K < (1ll << 31), IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(K % kBK == 0, IDB200_EUNSUPPORTED, "K (rows of both operands) must be a multiple of %d (got %lld)", kBK, K);
    IDB_REQUIRE(N % 64 == 0, IDB200_EUNSUPPORTED, "N must be a multiple of 64 (got %d)", N);
    IDB_REQUIRE(splits >= 1 && (K / kBK) % splits == 0, IDB200_EINVAL, "splits (%d) must divide K / %d", splits, kBK);
    IDB_REQUIRE(aligned(partial, 16), IDB200_EALIGN, "out must be 16-byte aligned");
    int BN = 0;
    static const bool prefer_pair = !(getenv("IDB200_GEMM_NN_PREFER_PAIR") && getenv("IDB200_GEMM_NN_PREFER_PAIR")[0] == '0');
    if (prefer_pair) {                                    // pair mode needs BN % 128 == 0 here: 128 before 192
        for (int cand : {256, 128, 192, 64})
            if (N % cand == 0) { BN = cand; break; }
    } else {
        for (int cand : {256, 192, 128, 64})
            if (N % cand == 0) { BN = cand; break; }
    }
    CUtensorMap ta, tw;
    int rc = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), 64, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), 64, 64);
    if (rc) return rc;
    GemmParams p{nullptr, partial, M, N, static_cast<int>(K), EPI_F32, splits, 0};
    switch (BN) {                                           // MN-major boxes are [64 x 64] in both modes: same tensor map
        case 256: return launch_gemm<256, true>(ta, tw, tw, ta, p, st);
        case 192: return launch_gemm<192, true>(ta, tw, tw, ta, p, st);
        case 128: return launch_gemm<128, true>(ta, tw, tw, ta, p, st);
        default: return launch_gemm<64, true>(ta, tw, tw, ta, p, st);
    }
}

int gemm_bf16_tn(const void* A, const void* W, const float* bias, void* out, long long M, int N, int K, int epilogue,
                 cudaStream_t st, int splits, void* aux = nullptr, float* colpart = nullptr) {
    const int out_f16 = (epilogue & 0x100) ? 1 : 0;                     // epilogue 0 | 0x100: IEEE half output
    epilogue &= 0xff;
    IDB_REQUIRE(!out_f16 || epilogue == EPI_BF16, IDB200_EINVAL, "the half-precision output flag goes with epilogue 0");
    IDB_REQUIRE(!colpart || (epilogue == EPI_BF16_DSILU && aligned(colpart, 16)), IDB200_EINVAL, "colpart goes with epilogue 5 (16-byte aligned)");
    IDB_REQUIRE(A && W && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((epilogue >= EPI_BF16_SILU_DUAL) == (aux != nullptr), IDB200_EINVAL, "epilogues 4 / 5 need the aux tensor (and only they take one)");
    IDB_REQUIRE(M >= 0 && N > 0 && K > 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(K % kBK == 0, IDB200_EUNSUPPORTED, "K must be a multiple of %d (got %d)", kBK, K);
    IDB_REQUIRE(N % 32 == 0, IDB200_EUNSUPPORTED, "N must be a multiple of 32 (got %d)", N);
    IDB_REQUIRE(epilogue >= 0 && epilogue <= 5, IDB200_EINVAL, "unknown epilogue %d", epilogue);
    IDB_REQUIRE(splits >= 1 && (K / kBK) % splits == 0, IDB200_EINVAL, "splits (%d) must divide K / %d", splits, kBK);
    IDB_REQUIRE(splits == 1 || (epilogue == EPI_F32 && bias == nullptr), IDB200_EINVAL, "split-K needs the fp32 epilogue without bias");
    IDB_REQUIRE(aligned(out, 16) && (!aux || aligned(aux, 16)), IDB200_EALIGN, "out / aux must be 16-byte aligned");
    IDB_REQUIRE(!aux || (N % 64 == 0 && M < (1ll << 31)), IDB200_EUNSUPPORTED, "the aux epilogues need N %% 64 == 0");
    if (M == 0) return IDB200_OK;
    int BN = 0;
    for (int cand : {256, 192, 128, 96, 64, 32})
        if (N % cand == 0) { BN = cand; break; }
    CUtensorMap ta, tw;
    int rc = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), kBM, kBK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint32_t>(BN), kBK);
    if (rc) return rc;
    CUtensorMap tw_half = tw;                               // pair mode: each CTA stages BN / 2 rows of W
    if (BN % 64 == 0) {
        rc = make_tmap_bf16_2d(&tw_half, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint32_t>(BN / 2), kBK);
        if (rc) return rc;
    }
    // TMA-store epilogue: 16 KB slabs of 64 bf16 / 32 fp32 columns (N % 64 == 0 covers both); split-K partials keep the direct stores
    static const bool tma_epi = !(getenv("IDB200_GEMM_TMA_EPI") && getenv("IDB200_GEMM_TMA_EPI")[0] == '0');
    const bool out16 = epilogue == EPI_BF16 || epilogue == EPI_SILU_BF16 || aux != nullptr;
    const int use_tma = ((tma_epi || aux) && splits == 1 && BN % 64 == 0 && M < (1ll << 31)) ? 1 : 0;
    CUtensorMap to = ta, tx = ta;
    if (use_tma) {
        rc = make_tmap_2d(&to, out, out16 ? 2 : 4, static_cast<uint64_t>(M), static_cast<uint64_t>(N), kBM, out16 ? 64 : 32);
        if (rc) return rc;
        if (aux) {
            rc = make_tmap_2d(&tx, aux, 2, static_cast<uint64_t>(M), static_cast<uint64_t>(N), kBM, 64);
            if (rc) return rc;
        }
    }
    GemmParams p{bias, out, M, N, K, epilogue, splits, use_tma};
    p.colpart = colpart;
    p.out_f16 = out_f16;
    const CUtensorMap* px = aux ? &tx : nullptr;
    switch (BN) {
        case 256: return launch_gemm<256>(ta, tw, tw_half, to, p, st, px);
        case 192: return launch_gemm<192>(ta, tw, tw_half, to, p, st, px);
        case 128: return launch_gemm<128>(ta, tw, tw_half, to, p, st, px);
        case 96: return launch_gemm<96>(ta, tw, tw_half, to, p, st, px);
        case 64: return launch_gemm<64>(ta, tw, tw_half, to, p, st, px);
        default: return launch_gemm<32>(ta, tw, tw_half, to, p, st, px);
    }
}


// 3x3 convolution (padding 1) as a tap-shifted GEMM on zero-bordered NHWC activations (encoders.py:15-24, one conv layer):
//   act [B * P2, C] bf16, P2 = (H + 2) * (W + 2); out [B * P2, N] bf16 = epi(conv(act) + bias), border positions written as zeros.
//   Wm: bf16 [N, nkb * 64], k-block layout (host: models/_engine.py::conv_tap_weight):
//     C % 64 == 0: k = tap * C + c (tap = ky * 3 + kx), nkb = 9 C / 64;
//     C == 32:     k-block (ky, j) = [tap (ky, 2j) channels | tap (ky, 2j + 1) channels (zeros for the missing 4th tap)], nkb = 6 -- the A box
//                  is two ADJACENT pixels read through an overlapping-row view of the activation tensor.
int conv3x3_gemm(const void* act, int C, const void* Wm, const float* bias, void* out, int N, long long B, int H, int Wd, int epilogue,
                 cudaStream_t st) {
    IDB_REQUIRE(act && Wm && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && H >= 1 && Wd >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(C == 32 || (C % 64 == 0 && 9 * C / 64 <= kConvMaxKB), IDB200_EUNSUPPORTED, "implicit conv needs C_in == 32 or a multiple of 64 up to 256 (got %d)", C);
    IDB_REQUIRE(N % 64 == 0, IDB200_EUNSUPPORTED, "implicit conv needs C_out %% 64 == 0 (got %d)", N);
    IDB_REQUIRE(epilogue == EPI_BF16 || epilogue == EPI_SILU_BF16, IDB200_EINVAL, "implicit conv writes bf16 (epilogue 0 or 1)");
    if (B == 0) return IDB200_OK;
    const int PW = Wd + 2, PH = H + 2, P2 = PW * PH;
    const long long M = B * P2;
    IDB_REQUIRE(M < (1ll << 31), IDB200_EUNSUPPORTED, "batch too large for one call");
    GemmParams p{bias, out, M, N, 0, epilogue, 1, 1};
    p.conv_P2 = P2; p.conv_PW = PW; p.conv_PH = PH;
    int nkb = 0;
    CUtensorMap ta;
    int rc;
    if (C == 32) {
        for (int ky = 0; ky < 3; ++ky)
            for (int j = 0; j < 2; ++j) { p.conv_shift[nkb] = (ky - 1) * PW + (j == 0 ? -1 : 1); p.conv_col[nkb] = 0; ++nkb; }
        rc = make_tmap_bf16_2d_pitch(&ta, act, static_cast<uint64_t>(M - 1), 64, 64, kBM, kBK);    // row r = pixels r, r + 1
    } else {
        for (int tap = 0; tap < 9; ++tap)
            for (int cb = 0; cb < C / 64; ++cb) { p.conv_shift[nkb] = (tap / 3 - 1) * PW + (tap % 3 - 1); p.conv_col[nkb] = cb * 64; ++nkb; }
        rc = make_tmap_bf16_2d(&ta, act, static_cast<uint64_t>(M), static_cast<uint64_t>(C), kBM, kBK);
    }
    if (rc) return rc;
    p.K = nkb * kBK;
    int BN = 0;
    for (int cand : {256, 192, 128, 64})
        if (N % cand == 0) { BN = cand; break; }
    CUtensorMap tw, tw_half, to;
    rc = make_tmap_bf16_2d(&tw, Wm, static_cast<uint64_t>(N), static_cast<uint64_t>(p.K), static_cast<uint32_t>(BN), kBK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tw_half, Wm, static_cast<uint64_t>(N), static_cast<uint64_t>(p.K), static_cast<uint32_t>(BN / 2), kBK);
    if (rc) return rc;
    rc = make_tmap_2d(&to, out, 2, static_cast<uint64_t>(M), static_cast<uint64_t>(N), kBM, 64);
    if (rc) return rc;
    switch (BN) {
        case 256: return launch_gemm<256>(ta, tw, tw_half, to, p, st);
        case 192: return launch_gemm<192>(ta, tw, tw_half, to, p, st);
        case 128: return launch_gemm<128>(ta, tw, tw_half, to, p, st);
        default: return launch_gemm<64>(ta, tw, tw_half, to, p, st);
    }
}

}  // namespace idb200

extern "C" int idb200_conv3x3_gemm(const void* act, int C, const void* Wm, const float* bias, void* out, int N, int64_t B, int H, int W,
                                   int epilogue, idb200_stream_t stream) {
    return idb200::conv3x3_gemm(act, C, Wm, bias, out, N, B, H, W, epilogue, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_gemm_bf16(const void* A, const void* W, const float* bias, void* out, int64_t M, int N, int K,
                                int epilogue, idb200_stream_t stream) {
    return idb200::gemm_bf16_tn(A, W, bias, out, M, N, K, epilogue, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int idb200_gemm_bf16_aux(const void* A, const void* W, const float* bias, void* out, void* aux, int64_t M, int N, int K,
                                    int epilogue, idb200_stream_t stream) {
    return idb200::gemm_bf16_tn(A, W, bias, out, M, N, K, epilogue, static_cast<cudaStream_t>(stream), 1, aux);
}

extern "C" int idb200_gemm_bf16_dsilu_sums(const void* A, const void* W, void* out, void* aux, float* colpart, int64_t M, int N, int K,
                                           idb200_stream_t stream) {
    IDB_REQUIRE(colpart != nullptr, IDB200_EINVAL, "colpart is NULL (use idb200_gemm_bf16_aux)");
    return idb200::gemm_bf16_tn(A, W, nullptr, out, M, N, K, idb200::EPI_BF16_DSILU, static_cast<cudaStream_t>(stream), 1, aux, colpart);
}

extern "C" int idb200_gemm_bf16_nn_splitk(const void* A, const void* W, float* partial, int64_t M, int N, int64_t K, int splits,
                                          idb200_stream_t stream) {
    return idb200::gemm_bf16_nn_splitk(A, W, partial, M, N, K, splits, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_gemm_bf16_splitk(const void* A, const void* W, float* partial, int64_t M, int N, int K, int splits,
                                       idb200_stream_t stream) {
    return idb200::gemm_bf16_tn(A, W, nullptr, partial, M, N, K, IDB200_EPI_F32, static_cast<cudaStream_t>(stream), splits);
}
