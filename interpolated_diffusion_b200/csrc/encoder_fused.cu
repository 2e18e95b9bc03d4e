// K3f: the WHOLE TransformerEncoder (all layers) of both denoisers in ONE persistent kernel, d_model = 256, 8 heads.
//
//     for each layer:  h += out_proj(MHA(LN1(h) (1+g1) + b1'));   h += ff2(SiLU(ff1(LN2(h) (1+g2) + b2')))
//
// Reference: src/models/transformer.py:35-46 (TransformerBlock.forward) iterated by :73-82 (TransformerEncoder).
//
// A 128-token tile (128 / L whole trajectories) never interacts with another tile, so a CTA carries its tile
// through every layer with the fp32 residual stream RESIDENT IN TENSOR MEMORY: h = TMEM columns [0,256).  The
// out-projection and FF2 MMAs accumulate straight onto h (the residual add is the accumulate flag), LayerNorm reads
// h back with tcgen05.ld in the accumulator's natural thread-per-row layout (row statistics are in-thread sums, no
// shuffles), and HBM sees one fp32 read and one write of h per tile for the whole encoder (the two-kernels-per-layer
// path moved 4 KB per token per layer).  The biases of the accumulating GEMMs (b_o, b_2) are never added to TMEM:
// the host passes their running sums cb (pending bias before each LayerNorm) and the kernel reads h + cb.
//
// One CTA per SM, 640 threads: warp 0 TMA producer (weights, 5-slot ring of [128 x 64] bf16 tiles, in exactly the
// order the MMA warp consumes them, free-running across phases/layers), warp 1 tcgen05.mma issuer, warp 2 TMEM
// allocator, warp 3 per-layer parameter loader (cp.async.bulk 1-D), warps 4..19 compute (thread <-> tile row x
// column quarter: 4 warps per SM sub-partition hide the TMEM / shared-memory / MUFU latencies of each other).
// Per layer:
//   LN1   compute: h(+cb1) -> LayerNorm, FiLM -> X (bf16 K-major SWIZZLE_128B A operand, 64 KB)
//   for head group g = 0..3:
//     GEMM_g  acc (TMEM cols 256..447) = X . [Wq_g;Wk_g;Wv_g]^T                 16 x (N=128 + N=64) tcgen05.mma
//     EPI_g   compute: acc + bias -> bf16 q|k|v rows in shared memory
//     ATT_g   compute: softmax(q k^T / sqrt(32)) v per (16-row block, head), mma.sync + ldmatrix -> O_g (A operand)
//     OUT_g   h (TMEM cols 0..255) += O_g . Wo[:, 64g..64g+63]^T                2 x 4 tcgen05.mma N=128
//   LN2   compute: h(+cb2) -> X
//   for hidden chunk c = 0..ff/128-1:
//     FF1_c   acc1[c&1] (TMEM cols 256.. / 384..) = X . W1[128c.., :]^T
//     EPI1_c  compute: acc1 + b1 -> SiLU -> bf16 H[c&1] (A operand)
//     FF2_c   h += H[c&1] . W2[:, 128c..]^T
// GEMM_{g+1} overlaps ATT_g, FF1_{c+2} overlaps EPI1_c; the shared-memory scratch (q|k|v staging + O, or H) and the
// TMEM scratch columns [256,512) are time-shared by the two phases (ordered by the x_full / h_ready barriers).
#include <cuda_fp16.h>
#include <cstdlib>
#include <type_traits>

#include "fused_common.cuh"

namespace idb200 {
using namespace tc;
using namespace fused;

namespace ef {
constexpr int kThreads = 640;
constexpr int kCW = 16;                             // compute warps
constexpr int kCT = kCW * 32;                       // compute threads (named-barrier width)
constexpr int kSlots = 5;
#ifndef IDB200_EF_REGS_AUX
#define IDB200_EF_REGS_AUX 32
#define IDB200_EF_REGS_COMPUTE 112
#endif
constexpr int kRegsAux = IDB200_EF_REGS_AUX, kRegsCompute = IDB200_EF_REGS_COMPUTE;   // 128 * aux + 512 * compute <= 640 * 96
constexpr int kSlotBytes = kTile;                   // [128 x 64] bf16 (the V tile [64 x 64] uses half a slot)
constexpr int kMaxFF = 1024;
constexpr int kOffX = 0;                            // 4 x [128 x 64] bf16
constexpr int kOffS = 4 * kTile;                    // scratch: q|k|v staging (51200) + O (16384)  |  H[2][2] (65536)
constexpr int kOffQkv = kOffS;
constexpr int kOffO = kOffS + 128 * kPitch * 2;
constexpr int kOffH = kOffS;
constexpr int kOffRing = kOffO + kTile;
constexpr int kOffBar = kOffRing + kSlots * kSlotBytes;
constexpr int kOffStat = kOffO + 12288;             // float2 [4][128] LayerNorm partial statistics: the tail of the O tile, idle
                                                    // while a LayerNorm runs (every MMA reading O / H has completed: h_ready)
constexpr int kOffERb = kOffS + 32768;              // staged prologue operands (p.e_stage; H[1], dead between the last FF2 and EPI1_1):
constexpr int kOffEWf = kOffERb + 16384;            //   row_b rows of the tile's trajectories (<= 16 KB) | Wf (<= 8 KB) | row_a (1 KB)
constexpr int kOffERa = kOffEWf + 8192;
static_assert(kOffERa + 1024 <= kOffO + 8192, "staged prologue operands stay clear of the head's partial sums and the LayerNorm statistics");
constexpr int kOffPA = kOffBar + 256;             // ln1_w 256 | ln1_b 256 | cb1 256 | bqkv 768
constexpr int kPAFloats = 1536;
constexpr int kOffPM = kOffPA + kPAFloats * 4;      // ln2_w 256 | ln2_b 256 | cb2 256 | b1 ff
constexpr int kPMFloatsMax = 768 + kMaxFF;
constexpr int kSmem = kOffPM + kPMFloatsMax * 4 + 1024;
static_assert(kOffO % 1024 == 0 && kOffRing % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(kOffH + 4 * kTile <= kOffRing, "H buffers fit the scratch region");
static_assert(kSmem <= 232448, "shared memory budget");

constexpr uint32_t kColAcc = 256;                   // TMEM scratch columns

struct Params {
    float* h;                   // [M, 256] fp32 residual stream (in/out)
    const float* params;        // per layer: PA (1536 floats) | PM (768 + ff floats)
    const float* cb_total;      // [256] pending bias of the last layer's ff.2 (added when h is written back)
    const float* gb;            // FiLM [B, 2 * n_layers, 512] rows = [gamma | beta] per LayerNorm, or nullptr
    long long gb_stride;        // floats between trajectories (512: the rows of a tile are contiguous -> one bulk copy per LayerNorm)
    long long gb_ln_stride;     // floats between the 2 * n_layers LayerNorm slots
    long long M;
    int L;
    int causal;
    int ff;
    int n_layers;
    // optional fused token assembly (replaces the read of h): idb200_embed_tokens semantics
    const float* e_src0;        // [M, n0] fp32, or nullptr: read h
    const float* e_src1;        // [M, n1] fp32 or nullptr
    const unsigned char* e_src2;  // [M, n2] uint8 or nullptr
    int e_n0, e_n1, e_n2;
    const float* e_wf;          // [n0+n1+n2, 256]
    const float* e_tab;         // [rows, 256]
    const long long* e_tab_idx; // [M] or nullptr (row = position in the trajectory)
    const float* e_row_a;       // [B or 1, 256]
    long long e_row_a_stride;
    const float* e_row_b;       // [B, 256]
    int e_stage;                // 1: tab has <= 64 rows and F <= 8 -> the table is staged in the (idle) X region by TMA, swizzled
    // optional fused output head (replaces the write of h): y[M, D] = (h + bias_last) . W^T + b
    const float* o_w;           // [D, 256] or nullptr: write h
    const float* o_b;           // [D]
    float* o_y;                 // [M, D]
    int o_D;                    // <= 4
    int film_mode;              // kFilmNone / kFilmRaw ([gamma | beta]) / kFilmFolded ([scale | shift], LayerNorm affine folded in)
    unsigned long long* prof;   // dev: [P_N] cycle sums (kProf kernels only)
    int dbg_skip;               // dev (kProf kernels only, IDB200_DBG_SKIP): compute warps only do the barrier handshakes (garbage results)
};

// SiLU(acc + b) = hx * tanh(hx) + hx with hx = 0.5 * (acc + b); hb = 0.5 * b
__device__ __forceinline__ float silu_half(float acc, float hb) {
    const float hx = fmaf(acc, 0.5f, hb);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}

// LayerNorm(eps 1e-5) + FiLM of the tile's rows, read from the TMEM-resident residual stream.
// Thread <-> (row, column quarter `part`).  Pass 1 adds the pending bias of the last accumulating GEMM (which never adds
// its own), writes the row back to TMEM, and sums statistics in-thread over its 64 columns (shifted by the first
// element), merged with the other three quarters' through shared memory (Chan's formula).  Pass 2 normalises and
// writes the bf16 SWIZZLE_128B k-block `part` of the A operand X.
// film: this thread's trajectory row of 512 floats, or nullptr (dead row: plain LayerNorm affine).  mode:
//   kFilmFolded  [scale | shift] with the LayerNorm affine already folded in: y = n * scale + shift
//   kFilmRaw     [gamma | beta]: y = (n * w + b) * (1 + gamma) + beta        kFilmNone  y = n * w + b
// kFilmSmem: the row is in shared memory (staged by bulk copies; wait on film_full first), else in global memory.
enum { kFilmNone = 0, kFilmRaw = 1, kFilmFolded = 2, kFilmFolded16 = 3 };   // Folded16: [scale | shift] as IEEE half (rows of 512 halves, staged in shared memory only)
// explicit shared-space 16-byte load (pointer selects hide the address space from the compiler: generic LD is ~3x slower here)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
template <bool kFilmSmem>
__device__ __forceinline__ void ln_tmem(uint32_t tmem_row, const float* sw, const float* sb, const float* spend, const float* film, int mode,
                                     uint64_t* film_full, uint32_t film_parity, float2* stat, uint8_t* X, int row, int part, long long* tt) {
    // (inlined: the pointers derive from the kernel's shared-memory base, whose address space the caller has already asserted)
    const int c0 = part * 64;
    const uint32_t t0 = tmem_row + c0;
    // ---- pass 1: x = h + pending bias (kept in registers for pass 2, written back to TMEM); statistics ----
    float xs, s1 = 0.0f, s2 = 0.0f;
    uint32_t r[2][32];
    {
        if (tt) tt[4] = clock64();
        tmem_ld_32x32(t0, r[0]);
        tmem_ld_32x32(t0 + 32, r[1]);
        tmem_ld_wait();
        if (tt) tt[3] = clock64();
        const uint32_t pb4 = smem_u32(spend + c0);
        xs = __uint_as_float(r[0][0]) + lds128(pb4).x;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 pb = lds128(pb4 + (u * 8 + j) * 16);
                const float x0 = __uint_as_float(r[u][4 * j + 0]) + pb.x, x1 = __uint_as_float(r[u][4 * j + 1]) + pb.y;
                const float x2 = __uint_as_float(r[u][4 * j + 2]) + pb.z, x3 = __uint_as_float(r[u][4 * j + 3]) + pb.w;
                r[u][4 * j + 0] = __float_as_uint(x0);
                r[u][4 * j + 1] = __float_as_uint(x1);
                r[u][4 * j + 2] = __float_as_uint(x2);
                r[u][4 * j + 3] = __float_as_uint(x3);
                const float d0 = x0 - xs, d1 = x1 - xs, d2 = x2 - xs, d3 = x3 - xs;
                s1 += (d0 + d1) + (d2 + d3);
                s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
            }
            tmem_st_32x32(t0 + u * 32, r[u]);
        }
    }
    stat[part * 128 + row] = make_float2(xs + s1 * (1.0f / 64.0f), s2 - s1 * s1 * (1.0f / 64.0f));
    if (tt) tt[0] = clock64();
    named_barrier_sync(3, kCT);
    if (tt) tt[1] = clock64();
    float mean = 0.0f, m2 = 0.0f;
    float2 st[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { st[i] = stat[i * 128 + row]; mean += st[i].x; m2 += st[i].y; }
    mean *= 0.25f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float dm = st[i].x - mean; m2 = fmaf(dm * dm, 64.0f, m2); }
    const float rstd = rsqrtf(fmaxf(m2 * (1.0f / 256.0f), 0.0f) + 1e-5f);
    const float shift = -mean * rstd;
    if (kFilmSmem) mbar_wait(film_full, film_parity, 58);
    if (tt) tt[2] = clock64();
    // ---- pass 2: normalise, FiLM, pack -- from the registers of pass 1 (no second TMEM read) ----
    uint8_t* xt = X + part * kTile;
    // scale / shift: the folded FiLM row (staged in shared memory, or in global memory when L < 8) or the LayerNorm affine
    const bool folded = (mode == kFilmFolded) && film != nullptr;
    const bool folded16 = kFilmSmem && (mode == kFilmFolded16) && film != nullptr;
    const bool raw = (mode == kFilmRaw) && film != nullptr;
    const bool gfold = !kFilmSmem && folded;
    const uint32_t scs = smem_u32((kFilmSmem && folded) ? film : sw) + c0 * 4;
    const uint32_t shs = smem_u32((kFilmSmem && folded) ? film + 256 : sb) + c0 * 4;
    const uint32_t fls = (kFilmSmem && raw) ? smem_u32(film) + c0 * 4 : 0u;
    const uint32_t f16s = smem_u32(film) + c0 * 2;                       // half table: scale at [0, 256), shift at [256, 512)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int col = c0 + cc * 32;
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {                                 // 8 columns -> one 16-byte swizzle chunk
            float y[8];
            if (folded16) {
                // half the shared-memory bytes per column: the LayerNorm is bound by the LSU return path (a warp-wide LDS.128
                // delivers 512 B in 4 cycles whether or not it is a broadcast), not by issue slots
                const float4 s16 = lds128(f16s + (cc * 4 + j2) * 16), h16 = lds128(f16s + 512 + (cc * 4 + j2) * 16);
                const uint32_t sv[4] = {__float_as_uint(s16.x), __float_as_uint(s16.y), __float_as_uint(s16.z), __float_as_uint(s16.w)};
                const uint32_t hv[4] = {__float_as_uint(h16.x), __float_as_uint(h16.y), __float_as_uint(h16.z), __float_as_uint(h16.w)};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j2 * 8 + 2 * u;                        // column offset inside the 32-column half
                    const float2 sc2 = __half22float2(*reinterpret_cast<const __half2*>(&sv[u]));
                    const float2 sh2 = __half22float2(*reinterpret_cast<const __half2*>(&hv[u]));
                    y[2 * u + 0] = fmaf(fmaf(__uint_as_float(r[cc][j + 0]), rstd, shift), sc2.x, sh2.x);
                    y[2 * u + 1] = fmaf(fmaf(__uint_as_float(r[cc][j + 1]), rstd, shift), sc2.y, sh2.y);
                }
            } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j2 * 2 + u;
                float4 sc, sh;
                if (!gfold) {
                    sc = lds128(scs + (cc * 8 + j) * 16);
                    sh = lds128(shs + (cc * 8 + j) * 16);
                } else {
                    sc = __ldg(reinterpret_cast<const float4*>(film + col) + j);
                    sh = __ldg(reinterpret_cast<const float4*>(film + 256 + col) + j);
                }
                y[4 * u + 0] = fmaf(fmaf(__uint_as_float(r[cc][4 * j + 0]), rstd, shift), sc.x, sh.x);
                y[4 * u + 1] = fmaf(fmaf(__uint_as_float(r[cc][4 * j + 1]), rstd, shift), sc.y, sh.y);
                y[4 * u + 2] = fmaf(fmaf(__uint_as_float(r[cc][4 * j + 2]), rstd, shift), sc.z, sh.z);
                y[4 * u + 3] = fmaf(fmaf(__uint_as_float(r[cc][4 * j + 3]), rstd, shift), sc.w, sh.w);
                if (raw) {
                    const float4* f4 = reinterpret_cast<const float4*>(film + col) + j;
                    const float4 g = kFilmSmem ? lds128(fls + (cc * 8 + j) * 16) : __ldg(f4);
                    const float4 t = kFilmSmem ? lds128(fls + 1024 + (cc * 8 + j) * 16) : __ldg(f4 + 64);
                    y[4 * u + 0] = fmaf(y[4 * u + 0], 1.0f + g.x, t.x);
                    y[4 * u + 1] = fmaf(y[4 * u + 1], 1.0f + g.y, t.y);
                    y[4 * u + 2] = fmaf(y[4 * u + 2], 1.0f + g.z, t.z);
                    y[4 * u + 3] = fmaf(y[4 * u + 3], 1.0f + g.w, t.w);
                }
            }
            }
            uint4 pk;
            pk.x = pack2_bf16(y[0], y[1]);
            pk.y = pack2_bf16(y[2], y[3]);
            pk.z = pack2_bf16(y[4], y[5]);
            pk.w = pack2_bf16(y[6], y[7]);
            *reinterpret_cast<uint4*>(xt + sw128_offset(row, cc * 32 + 8 * j2)) = pk;
        }
    }
    tmem_st_wait();                                                      // the write-back of pass 1 has landed (before x_full is signalled)
}

// kProf (dev, IDB200_PROF=1): compute warp 0 lane 0 accumulates clock64() spans per phase into p.prof[0..15]
enum { P_LOAD, P_LN1, P_WACC, P_EPI, P_ATT, P_WO, P_OWR, P_WH1, P_LN2, P_WACC1, P_EPI1, P_WH2, P_STORE, P_WPA, P_LNP1, P_LNBAR, P_LNFILM, P_LNLD, P_LNENTRY, P_LDPRE, P_LDWAIT, P_MSLOT, P_MCOMP, P_MISSUE, P_N };

// kPair: the kernel runs as clusters of two CTAs (tcgen05 cta_group::2).  Each CTA still owns one 128-token tile (its
// rows of h in its own TMEM, its own X / scratch / parameters), but the even CTA issues ONE M=256 MMA for both tiles
// and each CTA stages only half of every weight tile (N/2 rows), so the L2 -> shared-memory weight stream per SM -- the
// limiter of the single-CTA form at 128-token tiles -- is halved.  compute -> MMA barriers live in the even CTA (both
// CTAs' compute warps arrive there), MMA -> compute / producer barriers are signalled in both CTAs by multicast commits.
// tm_qk / tm_v / tm_w1: in pair mode tm_qk has a 96-row box (half of a head group's q|k|v rows), tm_w1 a 64-row box.
template <bool kProf, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
encoder_fused_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_v,
                     const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_tab, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __builtin_assume(__isShared(smem));             // the integer round trip hides the address space: keep LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* x_full = bars + 0;                    // compute -> MMA: X (LayerNorm output) written          (kCW arrivals)
    uint64_t* h_ready = bars + 1;                   // MMA -> compute: every accumulate into h has completed (commit)
    uint64_t* slot_full = bars + 2;                 // [kSlots]
    uint64_t* slot_empty = slot_full + kSlots;      // [kSlots]
    uint64_t* acc_full = slot_empty + kSlots;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* o_full = acc_empty + 1;
    uint64_t* o_empty = o_full + 1;
    uint64_t* acc1_full = o_empty + 1;              // [2]
    uint64_t* acc1_empty = acc1_full + 2;           // [2]
    uint64_t* hb_full = acc1_empty + 2;             // [2]
    uint64_t* hb_empty = hb_full + 2;               // [2]
    uint64_t* pa_full = hb_empty + 2;
    uint64_t* pa_empty = pa_full + 1;
    uint64_t* pm_full = pa_empty + 1;
    uint64_t* pm_empty = pm_full + 1;
    uint64_t* film_full = pm_empty + 1;             // FiLM rows of the next LayerNorm staged in the scratch region (tx bytes)
    uint64_t* stg_free = film_full + 1;             // every compute warp has read what it needs of the staged q|k|v (kCW arrivals)
    uint64_t* tab_full = stg_free + 1;              // token-assembly table staged in the X region (tx bytes), one phase per tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tab_full + 1);
    float2* stat = reinterpret_cast<float2*>(smem + kOffStat);
    float* sPA = reinterpret_cast<float*>(smem + kOffPA);
    float* sPM = reinterpret_cast<float*>(smem + kOffPM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = p.ff / 128;
    const int NL = p.n_layers;
    const int tiles = static_cast<int>((p.M + 127) / 128);          // (the host checks M < 2^38: 32-bit tile / trip counters, fewer live registers)
    const int pm_floats = 768 + p.ff;
    const int layer_floats = kPAFloats + pm_floats;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;          // 0 = the CTA that issues the pair's MMAs
    constexpr int kArr = kPair ? 2 * kCW : kCW;                      // arrivals on the compute -> MMA barriers
    // persistent loop: trip t handles tile t (single) or tiles 2t, 2t+1 (pair; a trailing odd tile leaves the peer a dead tile)
    const int trips = kPair ? (tiles + 1) / 2 : tiles;
    const int trip0 = static_cast<int>(kPair ? blockIdx.x / 2 : blockIdx.x);
    const int trip_stride = static_cast<int>(kPair ? gridDim.x / 2 : gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qk);
        tma_prefetch_desc(&tm_v);
        tma_prefetch_desc(&tm_wo);
        tma_prefetch_desc(&tm_w1);
        tma_prefetch_desc(&tm_w2);
        if (p.e_stage) tma_prefetch_desc(&tm_tab);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, kArr);
        mbar_init(h_ready, 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kArr);
        mbar_init(o_full, kArr);
        mbar_init(o_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&acc1_empty[i], kArr);
            mbar_init(&hb_full[i], kArr);
            mbar_init(&hb_empty[i], 1);
        }
        mbar_init(pa_full, 1);
        mbar_init(pa_empty, kCW);
        mbar_init(pm_full, 1);
        mbar_init(pm_empty, kCW);
        mbar_init(film_full, 1);
        mbar_init(stg_free, kCW);
        mbar_init(tab_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if (kPair) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_h = tmem_base;
    const uint32_t tmem_acc = tmem_base + kColAcc;

    // Register re-balancing (setmaxnreg, per warpgroup of 4 warps): the kernel is compiled for 96 registers per thread (640 threads);
    // warps 0..3 (producer / MMA issuer / allocator / parameter loader) give theirs back, the 16 compute warps -- whose LayerNorm
    // holds 64 row values per thread across a barrier and spilled at 96 -- take them.
    if (warp < 4) {
    setmaxnreg_dec<kRegsAux>();
    if (warp == 0) {
        // ===================== TMA producer (weights) =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t sphase = 0;
            // one ring slot: `n` tile loads (each `bytes`, `step` apart in the slot); in pair mode this CTA loads its half of
            // the rows and the bytes of both CTAs are credited to the even CTA's slot_full barrier
            auto fill = [&](const CUtensorMap* m, int c0, int c1, uint32_t bytes, int n, int dc0, uint32_t step) {
                mbar_wait(&slot_empty[slot], sphase ^ 1, 10);
                uint8_t* dst = smem + kOffRing + slot * kSlotBytes;
                if (kPair) {
                    if (rank == 0) mbar_arrive_expect_tx(&slot_full[slot], 2 * n * bytes);
                    for (int i = 0; i < n; ++i) tma_load_2d_2sm(dst + i * step, m, &slot_full[slot], c0 + i * dc0, c1);
                } else {
                    mbar_arrive_expect_tx(&slot_full[slot], n * bytes);
                    for (int i = 0; i < n; ++i) tma_load_2d(dst + i * step, m, &slot_full[slot], c0 + i * dc0, c1);
                }
                if (++slot == kSlots) { slot = 0; sphase ^= 1; }
            };
            for (int trip = trip0; trip < trips; trip += trip_stride) {
                for (int l = 0; l < NL; ++l) {
                    auto qkv = [&](int g) {
#pragma unroll 1
                        for (int kb = 0; kb < 4; ++kb) {
                            if (kPair) fill(&tm_qk, kb * 64, l * 768 + g * 192 + rank * 96, 96 * 128, 1, 0, 0);
                            else {
                                fill(&tm_qk, kb * 64, l * 768 + g * 192, kTile, 1, 0, 0);
                                fill(&tm_v, kb * 64, l * 768 + g * 192 + 128, kTile / 2, 1, 0, 0);
                            }
                        }
                    };
                    auto wo = [&](int g) {
                        if (kPair) fill(&tm_wo, g * 64, l * 256 + rank * 128, kTile, 1, 0, 0);
                        else {
                            fill(&tm_wo, g * 64, l * 256, kTile, 1, 0, 0);
                            fill(&tm_wo, g * 64, l * 256 + 128, kTile, 1, 0, 0);
                        }
                    };
                    auto ff1 = [&](int c) {
                        if (kPair) {                                         // 2 slots x (2 k-blocks of this CTA's 64 rows)
#pragma unroll 1
                            for (int j = 0; j < 2; ++j) fill(&tm_w1, j * 128, l * p.ff + c * 128 + rank * 64, kTile / 2, 2, 64, kTile / 2);
                        } else {
#pragma unroll 1
                            for (int kb = 0; kb < 4; ++kb) fill(&tm_w1, kb * 64, l * p.ff + c * 128, kTile, 1, 0, 0);
                        }
                    };
                    auto ff2 = [&](int c) {
                        if (kPair) {
#pragma unroll 1
                            for (int kb = 0; kb < 2; ++kb) fill(&tm_w2, c * 128 + kb * 64, l * 256 + rank * 128, kTile, 1, 0, 0);
                        } else {
#pragma unroll 1
                            for (int i = 0; i < 4; ++i) fill(&tm_w2, c * 128 + (i & 1) * 64, l * 256 + (i >> 1) * 128, kTile, 1, 0, 0);
                        }
                    };
                    // attention half: G0 G1 O0 G2 O1 G3 O2 O3 (rolled: one copy of each body keeps the code small)
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {
                        if ((0x2Bu >> i) & 1u) qkv(i < 2 ? i : (i + 1) >> 1);
                        else wo(i < 7 ? (i >> 1) - 1 : 3);
                    }
#pragma unroll 1
                    for (int c = -2; c < nc; ++c) {                    // FF1_{c+2} first: it only needs acc1 drained (signalled early)
                        if (c + 2 < nc) ff1(c + 2);
                        if (c >= 0) ff2(c);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair mode: the even CTA only) =====================
        // The whole warp runs the schedule (warp-uniform control flow, uniform descriptors); only the tcgen05 instructions are
        // predicated on one lane.  Issuing from inside an `if (lane == 0)` region made ptxas wrap every MMA in an
        // ELECT / R2UR.BROADCAST waterfall loop (~18 instructions and ~100 cycles per MMA: the issue rate, not the tensor
        // pipe, was the limit of the MLP half).
        if (rank == 0) {
            const bool issue = lane == 0;
            constexpr uint32_t kM = kPair ? 256 : 128;
            constexpr uint32_t idesc64 = umma_idesc_bf16(kM, 64);
            constexpr uint32_t idesc128 = umma_idesc_bf16(kM, 128);
            constexpr uint32_t idesc192 = umma_idesc_bf16(kM, 192);
            constexpr uint32_t idesc256 = umma_idesc_bf16(kM, 256);
            int slot = 0;
            uint32_t sphase = 0, n_l = 0;                                // n_l: layers issued so far (kernel lifetime)
            // barrier phases are derived from the position in the layer (see the compute warps): no counters, and above all no
            // dynamically indexed use1[b] / useh[b] arrays -- those lived in local memory (L2 latency) on the issue path
            const uint32_t uses0 = static_cast<uint32_t>((nc + 1) >> 1), uses1 = static_cast<uint32_t>(nc >> 1);
            const uint32_t sX = smem_u32(smem + kOffX), sO = smem_u32(smem + kOffO), sH = smem_u32(smem + kOffH), sR = smem_u32(smem + kOffRing);
            unsigned long long macc[3] = {0, 0, 0};
            long long mprev = kProf ? clock64() : 0;
            auto mstamp = [&](int what) {
                if (kProf) {
                    const long long t = clock64();
                    macc[what] += static_cast<unsigned long long>(t - mprev);
                    mprev = t;
                }
            };
            unsigned long long mtag[12] = {};                            // kProf: wait cycles per barrier tag (20..31)
            auto wait = [&](uint64_t* bar, uint32_t parity, int tag) {   // barriers the peer CTA also arrives on / credits
                mstamp(2);
                const long long tw = kProf ? clock64() : 0;
                mbar_wait(bar, parity, tag);
                tc_fence_after();
                if (kProf) mtag[tag - 20] += static_cast<unsigned long long>(clock64() - tw);
                mstamp(tag == 21 || tag == 22 || tag == 24 || tag == 25 || tag == 27 || tag == 29 ? 0 : 1);
            };
            auto commit = [&](uint64_t* bar) {
                if (elect_one_sync()) { if (kPair) umma_commit_2sm(bar); else umma_commit(bar); }
                __syncwarp();
            };
            // 4 x (K = 16) MMAs of one 64-wide k-block: A tile at a_addr, B tile at b_addr
            auto mma4 = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool first_zero) {
                const uint64_t ad = umma_desc_sw128(a_addr);
                const uint64_t bd = umma_desc_sw128(b_addr);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (kPair) umma_bf16_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
                        else umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first_zero && k == 0) ? 0u : 1u);
                    }
                }
                __syncwarp();
            };
            auto slot_begin = [&](int tag) -> uint32_t {
                wait(&slot_full[slot], sphase, tag);
                return sR + slot * kSlotBytes;
            };
            auto slot_end = [&]() {
                commit(&slot_empty[slot]);
                if (++slot == kSlots) { slot = 0; sphase ^= 1; }
            };
            auto gemm = [&](int g) {
                wait(acc_empty, static_cast<uint32_t>((g & 1) ^ 1), 20);   // EPI of the previous group drained the accumulator
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb) {
                    if (kPair) {
                        const uint32_t b = slot_begin(21);
                        mma4(tmem_acc, sX + kb * kTile, b, idesc192, kb == 0);           // q | k | v (192 columns)
                        slot_end();
                    } else {
                        uint32_t b = slot_begin(21);
                        mma4(tmem_acc, sX + kb * kTile, b, idesc128, kb == 0);           // q | k  (128 columns)
                        slot_end();
                        b = slot_begin(22);
                        mma4(tmem_acc + 128, sX + kb * kTile, b, idesc64, kb == 0);      // v      (64 columns)
                        slot_end();
                    }
                }
                commit(acc_full);
            };
            auto outp = [&](int g) {
                wait(o_full, static_cast<uint32_t>(g & 1), 23);          // ATT_g wrote O_g
                if (kPair) {
                    const uint32_t b = slot_begin(24);
                    mma4(tmem_h, sO, b, idesc256, false);                // h += O_g . Wo[:, 64g..]^T
                    slot_end();
                } else {
                    uint32_t b = slot_begin(24);
                    mma4(tmem_h, sO, b, idesc128, false);                // h[:, 0:128]
                    slot_end();
                    b = slot_begin(25);
                    mma4(tmem_h + 128, sO, b, idesc128, false);          // h[:, 128:256]
                    slot_end();
                }
                commit(o_empty);
                if (g == 3) commit(h_ready);
            };
            auto ff1 = [&](int c) {
                const int b = c & 1;
                const uint32_t ub = (n_l * (b ? uses1 : uses0) + static_cast<uint32_t>(c >> 1)) & 1u;   // uses of buffer b so far
                wait(&acc1_empty[b], ub ^ 1u, 26);                       // EPI1 drained acc1[b]
                if (kPair) {
#pragma unroll 1
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t bs = slot_begin(27);
                        mma4(tmem_acc + b * 128, sX + (2 * j) * kTile, bs, idesc128, j == 0);
                        mma4(tmem_acc + b * 128, sX + (2 * j + 1) * kTile, bs + kTile / 2, idesc128, false);
                        slot_end();
                    }
                } else {
#pragma unroll 1
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t bs = slot_begin(27);
                        mma4(tmem_acc + b * 128, sX + kb * kTile, bs, idesc128, kb == 0);
                        slot_end();
                    }
                }
                commit(&acc1_full[b]);
            };
            auto ff2 = [&](int c) {
                const int b = c & 1;
                const uint32_t ub = (n_l * (b ? uses1 : uses0) + static_cast<uint32_t>(c >> 1)) & 1u;
                wait(&hb_full[b], ub, 28);                               // EPI1 wrote H[b]
                if (kPair) {
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t bs = slot_begin(29);
                        mma4(tmem_h, sH + (b * 2 + kb) * kTile, bs, idesc256, false);
                        slot_end();
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t bs = slot_begin(29);
                        mma4(tmem_h + (i >> 1) * 128, sH + (b * 2 + (i & 1)) * kTile, bs, idesc128, false);
                        slot_end();
                    }
                }
                commit(&hb_empty[b]);
                if (c == nc - 1) commit(h_ready);
            };
            for (int trip = trip0; trip < trips; trip += trip_stride) {
                for (int l = 0; l < NL; ++l, ++n_l) {
                    wait(x_full, 0u, 30);
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {                        // G0 G1 O0 G2 O1 G3 O2 O3
                        if ((0x2Bu >> i) & 1u) gemm(i < 2 ? i : (i + 1) >> 1);
                        else outp(i < 7 ? (i >> 1) - 1 : 3);
                    }
                    wait(x_full, 1u, 31);
#pragma unroll 1
                    for (int c = -2; c < nc; ++c) {                    // FF1_{c+2} first: it only needs acc1 drained (signalled early)
                        if (c + 2 < nc) ff1(c + 2);
                        if (c >= 0) ff2(c);
                    }
                }
            }
            if (kProf && issue) {
                mstamp(2);
                for (int i = 0; i < 3; ++i) atomicAdd(p.prof + P_MSLOT + i, macc[i] * (kPair ? 2 : 1));   // per tile: the pair's issuer serves two
                for (int i = 0; i < 12; ++i) atomicAdd(p.prof + P_N + i, mtag[i] * (kPair ? 2 : 1));
            }
        }
    } else if (warp == 3) {
        // ===================== per-layer parameter loader =====================
        if (lane == 0) {
            uint32_t n = 0;
            for (int trip = trip0; trip < trips; trip += trip_stride) {
                for (int l = 0; l < NL; ++l, ++n) {
                    const float* src = p.params + static_cast<long long>(l) * layer_floats;
                    mbar_wait<true>(pa_empty, (n & 1) ^ 1, 40);
                    mbar_arrive_expect_tx(pa_full, kPAFloats * 4);
                    bulk_load_1d(sPA, src, kPAFloats * 4, pa_full);
                    mbar_wait<true>(pm_empty, (n & 1) ^ 1, 41);
                    mbar_arrive_expect_tx(pm_full, pm_floats * 4);
                    bulk_load_1d(sPM, src + kPAFloats, pm_floats * 4, pm_full);
                }
            }
        }
    }
    } else {
        setmaxnreg_inc<kRegsCompute>();
        // ===================== compute warps =====================
        const int ew = warp - 4;
        const int q = ew & 3, part = ew >> 2;            // TMEM lane quadrant (warp % 4), column quarter
        const int row = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t tmem_row = tmem_h + lane_base;
        const __nv_bfloat16* sq = reinterpret_cast<const __nv_bfloat16*>(smem + kOffQkv);
        uint8_t* sqb = smem + kOffQkv;
        uint8_t* so = smem + kOffO;
        const int L = p.L;
        // Barrier phases are DERIVED, not counted: every per-layer barrier completes an even number of phases per layer (4 head groups,
        // 2 LayerNorms, 2 h_ready commits), so its parity depends on the position inside the layer only; the chunk buffers' use
        // counts follow from the layer counter n_p.  (Ten live counters per thread spilled around the LayerNorms: local memory goes
        // to L2 here -- 226 KB of the SM's 256 KB are shared memory.)
        uint32_t n_p = 0;
        const uint32_t uses0 = static_cast<uint32_t>((nc + 1) >> 1), uses1 = static_cast<uint32_t>(nc >> 1);   // chunk uses of buffer 0 / 1 per layer
        unsigned long long pacc[P_N] = {};
        long long tprev = kProf ? clock64() : 0;
        auto stamp = [&](int what) {
            if (kProf) {
                const long long t = clock64();
                pacc[what] += static_cast<unsigned long long>(t - tprev);
                tprev = t;
            }
        };
        // FiLM staging (L >= 8: at most 16 trajectories per tile): one thread copies the [gamma | beta] rows of the tile's
        // trajectories for LayerNorm `which` (0/1) of layer l into the first 32 KB of the scratch region, which is idle
        // between the last attention read and EPI1_0 / between FF2 of the last even chunk and EPI_0.
        const bool skip = kProf && p.dbg_skip;
        const bool film_smem = (p.gb != nullptr) && L >= 8;
        const int film_elem = p.film_mode == kFilmFolded16 ? 2 : 4;      // bytes per table element; a trajectory's row is 512 elements
        const uint32_t film_row_bytes = 512u * static_cast<uint32_t>(film_elem);
        auto stage_film = [&](int tile_, int l_, int which) {
            if (film_smem && ew == 0 && lane == 0) {
                const long long t0 = static_cast<long long>(tile_) * 128 / L;
                const long long left = p.M / L - t0;
                const int nt = left <= 0 ? 0 : static_cast<int>(left < 128 / L ? left : 128 / L);   // 0: the pair's dead tile
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(film_full, static_cast<uint32_t>(nt) * film_row_bytes);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.gb) + (t0 * p.gb_stride + (2 * l_ + which) * p.gb_ln_stride) * film_elem;
                if (p.gb_stride == 512) {                                // LayerNorm-major table: the tile's rows are contiguous
                    if (nt > 0) bulk_load_1d(smem + kOffS, src, static_cast<uint32_t>(nt) * film_row_bytes, film_full);
                } else {
                    for (int t = 0; t < nt; ++t) bulk_load_1d(smem + kOffS + t * film_row_bytes, src + t * p.gb_stride * film_elem, film_row_bytes, film_full);
                }
            }
        };
        auto tile_of = [&](int trip) { return kPair ? 2 * trip + static_cast<int>(rank) : trip; };
        auto arrive_mma = [&](uint64_t* bar) { if (kPair) mbar_arrive_leader(bar); else mbar_arrive(bar); };
        if (trip0 < trips) stage_film(tile_of(trip0), 0, 0);
        // Token-assembly table (p.e_stage: <= 64 rows of 256 floats): eight [64 x 32] fp32 SWIZZLE_128B boxes into the X region,
        // which is idle from the last FF1 of a tile to the first LayerNorm of the next.  The tile's rows gather their table row
        // with one conflict-free LDS.128 per 4 columns (a row-per-thread gather from global memory costs 32 L1 wavefronts per
        // load instruction: ~8 k cycles per tile, like the row-per-thread read of h it replaces).  The table is the same for every
        // tile; it is re-staged per tile because X is the LayerNorm output in between.
        // tab_full completes TWO phases per tile (the copy, then a plain arrive just before the next copy is issued), so the prologue
        // always waits for parity 0 and no per-tile counter stays live across the layers.
        // Next to it, in the part of the scratch region that is dead at that time (H[1]): the tile's row_b rows (contiguous: one row
        // per trajectory), the feature weights Wf and (when batch-constant) row_a -- every operand of the prologue is then read
        // with shared-memory latency, and its loops have no global-memory round trips.
        auto stage_tab = [&](int tile_, bool first) {
            if (p.e_stage && ew == 1 && lane == 0) {
                const long long t0 = static_cast<long long>(tile_) * 128 / L;
                const long long left = p.M / L - t0;
                const uint32_t nt = left <= 0 ? 0u : static_cast<uint32_t>(left < 128 / L ? left : 128 / L);   // 0: the pair's dead tile
                const uint32_t wf_bytes = static_cast<uint32_t>(p.e_n0 + p.e_n1 + p.e_n2) * 1024u;
                const uint32_t ra_bytes = p.e_row_a_stride == 0 ? 1024u : 0u;
                if (!first) mbar_arrive(tab_full);
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(tab_full, 65536u + nt * 1024u + wf_bytes + ra_bytes);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) tma_load_2d(smem + kOffX + c * 8192, &tm_tab, tab_full, c * 32, 0);
                if (nt) bulk_load_1d(smem + kOffERb, p.e_row_b + t0 * kD, nt * 1024u, tab_full);
                bulk_load_1d(smem + kOffEWf, p.e_wf, wf_bytes, tab_full);
                if (ra_bytes) bulk_load_1d(smem + kOffERa, p.e_row_a, 1024u, tab_full);
            }
        };
        if (trip0 < trips) stage_tab(tile_of(trip0), true);
        // attention work unit of this warp: 16-row block rb, head hh of the group
        const int rb = ew & 7, hh = ew >> 3;
        uint32_t okbits = 0;                             // L < 16: block-diagonal mask of the 16 x 16 score block (per thread, fixed)
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
        for (int trip = trip0; trip < trips; trip += trip_stride) {
            const int tile = tile_of(trip);
            const long long m0 = static_cast<long long>(tile) * 128;
            const long long m = m0 + row;
            const bool live = m < p.M;
            // ---- residual stream tile -> TMEM (this thread: its row, columns part*64 .. +63) ----
            if (p.e_src0 == nullptr) {
                const float4* src = reinterpret_cast<const float4*>(p.h + m * kD + part * 64);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t r[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = live ? src[cc * 8 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                        r[4 * j + 0] = __float_as_uint(v.x);
                        r[4 * j + 1] = __float_as_uint(v.y);
                        r[4 * j + 2] = __float_as_uint(v.z);
                        r[4 * j + 3] = __float_as_uint(v.w);
                    }
                    tmem_st_32x32(tmem_row + part * 64 + cc * 32, r);
                }
                tmem_st_wait();
            } else if (p.e_stage) {
                // fused token assembly + in_proj with every operand staged in shared memory (idb200_embed_tokens; same fp32 operation
                // order: fma chain over the features, + tab + row_a + row_b).  Dead rows compute on slot 0 and are zeroed at the end.
                const int F = p.e_n0 + p.e_n1 + p.e_n2;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.0f;
                int trow = 0;
                if (live) {
                    trow = p.e_tab_idx ? static_cast<int>(p.e_tab_idx[m]) : (row % L);
                    trow = min(max(trow, 0), 63);                        // (an out-of-range index must not leave the staged boxes)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < p.e_n0) f[j] = p.e_src0[m * p.e_n0 + j];
                        else if (j < p.e_n0 + p.e_n1) f[j] = p.e_src1[m * p.e_n1 + (j - p.e_n0)];
                        else if (j < F) f[j] = p.e_src2[m * p.e_n2 + (j - p.e_n0 - p.e_n1)] ? 1.0f : 0.0f;
                    }
                }
                const bool ra_s = p.e_row_a_stride == 0;
                const float4* ra4 = reinterpret_cast<const float4*>(p.e_row_a + (live ? m / L : 0) * p.e_row_a_stride) + part * 16;
                const uint32_t col_s = static_cast<uint32_t>(part) * 256u;                                  // this thread's 64 columns, bytes
                const uint32_t tab_s = smem_u32(smem + kOffX) + static_cast<uint32_t>(part * 2) * 8192u + static_cast<uint32_t>(trow) * 128u;
                const uint32_t tsw = static_cast<uint32_t>(trow & 7);
                const uint32_t rb_s = smem_u32(smem + kOffERb) + static_cast<uint32_t>(live ? row / L : 0) * 1024u + col_s;
                const uint32_t wf_s = smem_u32(smem + kOffEWf) + col_s, ras = smem_u32(smem + kOffERa) + col_s;
                stamp(P_LDPRE);                                          // (dev: tile turnaround up to here; P_LDWAIT = the wait for the staged operands)
                mbar_wait(tab_full, 0u, 62);
                stamp(P_LDWAIT);                                         // (dev: tile turnaround up to the staged operands' arrival)
                auto body = [&](auto kFtag) {
                    constexpr int kF = decltype(kFtag)::value;
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t r[32];
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const uint32_t co = static_cast<uint32_t>(cc * 8 + j4) * 16u;
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < kF; ++j) {
                                if (j >= F) break;                       // (rows >= F of the staged Wf are not written)
                                const float4 w = lds128(wf_s + static_cast<uint32_t>(j) * 1024u + co);
                                acc.x = fmaf(f[j], w.x, acc.x);
                                acc.y = fmaf(f[j], w.y, acc.y);
                                acc.z = fmaf(f[j], w.z, acc.z);
                                acc.w = fmaf(f[j], w.w, acc.w);
                            }
                            const float4 tb = lds128(tab_s + static_cast<uint32_t>(cc) * 8192u + ((static_cast<uint32_t>(j4) ^ tsw) << 4));
                            const float4 a = ra_s ? lds128(ras + co) : __ldg(ra4 + cc * 8 + j4);
                            const float4 b = lds128(rb_s + co);
                            r[4 * j4 + 0] = live ? __float_as_uint(acc.x + tb.x + a.x + b.x) : 0u;
                            r[4 * j4 + 1] = live ? __float_as_uint(acc.y + tb.y + a.y + b.y) : 0u;
                            r[4 * j4 + 2] = live ? __float_as_uint(acc.z + tb.z + a.z + b.z) : 0u;
                            r[4 * j4 + 3] = live ? __float_as_uint(acc.w + tb.w + a.w + b.w) : 0u;
                        }
                        tmem_st_32x32(tmem_row + part * 64 + cc * 32, r);
                    }
                };
                if (F <= 4) body(std::integral_constant<int, 4>{}); else body(std::integral_constant<int, 8>{});
                tmem_st_wait();
            } else {
                // fused token assembly + in_proj (idb200_embed_tokens; same fp32 operation order):
                //   h[m, :] = [src0 | src1 | src2][m, :] . Wf + tab[row] + row_a[b] + row_b[b]
                const int F = p.e_n0 + p.e_n1 + p.e_n2;
                float f[16];
                long long trow = 0, bb = 0;
                if (live) {
                    bb = m / L;
                    trow = p.e_tab_idx ? p.e_tab_idx[m] : (m - bb * L);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = 0.0f;
                        if (j < p.e_n0) v = p.e_src0[m * p.e_n0 + j];
                        else if (j < p.e_n0 + p.e_n1) v = p.e_src1[m * p.e_n1 + (j - p.e_n0)];
                        else if (j < F) v = p.e_src2[m * p.e_n2 + (j - p.e_n0 - p.e_n1)] ? 1.0f : 0.0f;
                        f[j] = v;
                    }
                }
                const float4* tab4 = reinterpret_cast<const float4*>(p.e_tab + trow * kD) + part * 16;
                const float4* ra4 = reinterpret_cast<const float4*>(p.e_row_a + bb * p.e_row_a_stride) + part * 16;
                const float4* rb4 = reinterpret_cast<const float4*>(p.e_row_b + bb * kD) + part * 16;
                const float4* wf4 = reinterpret_cast<const float4*>(p.e_wf) + part * 16;
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t r[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < F) {
                                    const float4 w = __ldg(wf4 + j * (kD / 4) + cc * 8 + j4);
                                    acc.x = fmaf(f[j], w.x, acc.x);
                                    acc.y = fmaf(f[j], w.y, acc.y);
                                    acc.z = fmaf(f[j], w.z, acc.z);
                                    acc.w = fmaf(f[j], w.w, acc.w);
                                }
                            }
                            const float4 tb = __ldg(tab4 + cc * 8 + j4), a = __ldg(ra4 + cc * 8 + j4), b = __ldg(rb4 + cc * 8 + j4);
                            acc = make_float4(acc.x + tb.x + a.x + b.x, acc.y + tb.y + a.y + b.y, acc.z + tb.z + a.z + b.z, acc.w + tb.w + a.w + b.w);
                        }
                        r[4 * j4 + 0] = __float_as_uint(acc.x);
                        r[4 * j4 + 1] = __float_as_uint(acc.y);
                        r[4 * j4 + 2] = __float_as_uint(acc.z);
                        r[4 * j4 + 3] = __float_as_uint(acc.w);
                    }
                    tmem_st_32x32(tmem_row + part * 64 + cc * 32, r);
                }
                tmem_st_wait();
            }
            stamp(P_LOAD);
            // FiLM row of this thread's trajectory: global (L < 8) or staged in the scratch region (L >= 8, slot = row / L)
            const float* gbtraj = (p.gb != nullptr && live) ? p.gb + (m / L) * p.gb_stride : nullptr;
            const float* sfilm = (p.gb != nullptr && live) ? reinterpret_cast<const float*>(smem + kOffS + (row / L) * film_row_bytes) : nullptr;
            for (int l = 0; l < NL; ++l, ++n_p) {
                // ================= attention half =================
                mbar_wait(pa_full, n_p & 1, 50);
                stamp(P_WPA);
                long long tt[5] = {0, 0, 0, 0, 0};
                if (skip) { if (film_smem) { mbar_wait(film_full, 0u, 58); } }
                else if (film_smem) ln_tmem<true>(tmem_row, sPA, sPA + 256, sPA + 512, sfilm, p.film_mode, film_full, 0u, stat, smem + kOffX, row, part, kProf ? tt : nullptr);
                else ln_tmem<false>(tmem_row, sPA, sPA + 256, sPA + 512, gbtraj ? gbtraj + (2 * l) * p.gb_ln_stride : nullptr, p.film_mode, nullptr, 0, stat, smem + kOffX, row, part, kProf ? tt : nullptr);
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) arrive_mma(x_full);
                if (kProf) { pacc[P_LNP1] += tt[0] - tprev; pacc[P_LNBAR] += tt[1] - tt[0]; pacc[P_LNFILM] += tt[2] - tt[1]; pacc[P_LNLD] += tt[3] - tprev; pacc[P_LNENTRY] += tt[4] - tprev; }
                stamp(P_LN1);
                const float* sbqkv = sPA + 768;
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    mbar_wait(acc_full, static_cast<uint32_t>(g & 1), 51);
                    tc_fence_after();
                    if (g > 0) mbar_wait(stg_free, static_cast<uint32_t>((g + 1) & 1), 60);  // every warp is done reading the previous q|k|v
                    stamp(P_WACC);
                    // ---- EPI_g: acc + bias -> bf16 q|k|v rows (this thread: its row, 48 of the 192 columns) ----
                    if (!skip) {
                        uint32_t ra[32], rc[16];
                        tmem_ld_32x32(tmem_acc + lane_base + part * 48, ra);
                        tmem_ld_32x16(tmem_acc + lane_base + part * 48 + 32, rc);
                        tmem_ld_wait();
                        const float* bb = sbqkv + g * 192 + part * 48;
                        uint8_t* dst = sqb + row * (kPitch * 2) + part * 96;
                        // Only the q columns (the first 64 of the group's 192) get their in_proj bias here: a k bias shifts every score
                        // of a query by the same amount (softmax-invariant) and a v bias passes through the attention unchanged
                        // (rows of P sum to 1), so the host drops the former and folds the latter into the pending out_proj bias.
                        // This epilogue is bound by shared-memory return bandwidth (12 of its 18 128-bit accesses were bias loads).
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const uint32_t* v = (j < 4) ? &ra[8 * j] : &rc[8 * (j - 4)];
                            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                            if (part * 48 + 8 * j < 64) {                    // (warp-uniform)
                                b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                                b1 = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                            }
                            uint4 pk;
                            pk.x = pack2_bf16(__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y);
                            pk.y = pack2_bf16(__uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w);
                            pk.z = pack2_bf16(__uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y);
                            pk.w = pack2_bf16(__uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w);
                            *reinterpret_cast<uint4*>(dst + 16 * j) = pk;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_mma(acc_empty);
                    named_barrier_sync(2, kCT);                          // q|k|v of the whole tile are in shared memory
                    stamp(P_EPI);
                    // ---- ATT_g: this warp's (16-row block, head) ----
                    float o[4][4] = {};
                    if (!skip) {
                        const __nv_bfloat16* qh = sq + hh * 32;
                        if (L < 16) attn_unit_fast<2, 1>(qh, qh + 64, qh + 128, rb, rb * 16, rb * 16 + 16, okbits, lane, o);
                        else if (p.causal) {
                            const int kbeg = (rb * 16 / L) * L, kend = rb * 16 + 16;
                            if (L == 16) attn_unit_fast<2, 2>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                            else attn_unit_fast<4, 2>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                        } else {
                            const int kbeg = (rb * 16 / L) * L, kend = kbeg + L;
                            if (L == 16) attn_unit_fast<2, 0>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                            else if (L == 64) attn_unit_fast<8, 0>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);   // all 64 keys in one pass: no online rescale
                            else attn_unit_fast<4, 0>(qh, qh + 64, qh + 128, rb, kbeg, kend, 0u, lane, o);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(stg_free);                // this warp is done reading the staged q|k|v
                    stamp(P_ATT);
                    mbar_wait(o_empty, static_cast<uint32_t>((g & 1) ^ 1), 52);   // OUT_{g-1} finished reading O
                    stamp(P_WO);
                    if (!skip) {
                        const int gq = lane >> 2, tq = lane & 3;
                        const int r0 = rb * 16 + gq;
                        uint8_t* o0 = so + r0 * 128 + tq * 4;            // sw128_offset(r0, c): chunk (c >> 3) ^ (r0 & 7); r0 + 8: + 1024 bytes
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const int ch = ((hh * 4 + nt) ^ (r0 & 7)) << 4;
                            *reinterpret_cast<unsigned*>(o0 + ch) = pack2_bf16(o[nt][0], o[nt][1]);
                            *reinterpret_cast<unsigned*>(o0 + ch + 1024) = pack2_bf16(o[nt][2], o[nt][3]);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) arrive_mma(o_full);
                    stamp(P_OWR);
                }
                if (film_smem && ew == 0 && lane == 0) {                 // once every warp is done with the staged q|k|v: stage LN2's
                    mbar_wait(stg_free, 1u, 61);                         // FiLM rows over them (after o_full: OUT_3 is not held up)
                    stage_film(tile, l, 1);
                }
                if (lane == 0) mbar_arrive(pa_empty);                    // (after the __syncwarp above: the warp is done with PA)
                // ================= MLP half =================
                mbar_wait(pm_full, n_p & 1, 53);
                mbar_wait(h_ready, 0u, 54);                              // OUT_3 has landed in h
                tc_fence_after();
                stamp(P_WH1);
                if (skip) { if (film_smem) { mbar_wait(film_full, 1u, 58); } }
                else if (film_smem) ln_tmem<true>(tmem_row, sPM, sPM + 256, sPM + 512, sfilm, p.film_mode, film_full, 1u, stat, smem + kOffX, row, part, kProf ? tt : nullptr);
                else ln_tmem<false>(tmem_row, sPM, sPM + 256, sPM + 512, gbtraj ? gbtraj + (2 * l + 1) * p.gb_ln_stride : nullptr, p.film_mode, nullptr, 0, stat, smem + kOffX, row, part, kProf ? tt : nullptr);
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) arrive_mma(x_full);
                if (kProf) { pacc[P_LNP1] += tt[0] - tprev; pacc[P_LNBAR] += tt[1] - tt[0]; pacc[P_LNFILM] += tt[2] - tt[1]; pacc[P_LNLD] += tt[3] - tprev; pacc[P_LNENTRY] += tt[4] - tprev; }
                stamp(P_LN2);
                const float* sb1 = sPM + 768;
#pragma unroll 1
                for (int c = 0; c < nc; ++c) {
                    const int b = c & 1;
                    const uint32_t ub = (n_p * (b ? uses1 : uses0) + static_cast<uint32_t>(c >> 1)) & 1u;   // uses of buffer b so far
                    mbar_wait(&acc1_full[b], ub, 55);
                    tc_fence_after();
                    stamp(P_WACC1);
                    if (!skip) {
                        // this thread: its row, columns part*32 .. +31 of the 128-column chunk (k-block part >> 1 of H[b])
                        uint8_t* hb = smem + kOffH + (b * 2 + (part >> 1)) * kTile;
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_acc + lane_base + b * 128 + part * 32, r);
                        tmem_ld_wait();
                        tc_fence_before();                               // the accumulator is in registers: release it at once, so that
                        __syncwarp();                                    // FF1_{c+2} runs under the SiLU / store work of this chunk
                        if (lane == 0) arrive_mma(&acc1_empty[b]);
                        const float* bb = sb1 + c * 128 + part * 32;    // 0.5 * b1 (pre-halved on the host)
                        uint4 pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {                    // 8 columns -> one 16-byte swizzle chunk
                            const float4 b0 = *reinterpret_cast<const float4*>(bb + 8 * j);
                            const float4 b1v = *reinterpret_cast<const float4*>(bb + 8 * j + 4);
                            pk[j].x = pack2_bf16(silu_half(__uint_as_float(r[8 * j + 0]), b0.x), silu_half(__uint_as_float(r[8 * j + 1]), b0.y));
                            pk[j].y = pack2_bf16(silu_half(__uint_as_float(r[8 * j + 2]), b0.z), silu_half(__uint_as_float(r[8 * j + 3]), b0.w));
                            pk[j].z = pack2_bf16(silu_half(__uint_as_float(r[8 * j + 4]), b1v.x), silu_half(__uint_as_float(r[8 * j + 5]), b1v.y));
                            pk[j].w = pack2_bf16(silu_half(__uint_as_float(r[8 * j + 6]), b1v.z), silu_half(__uint_as_float(r[8 * j + 7]), b1v.w));
                        }
                        mbar_wait(&hb_empty[b], ub ^ 1u, 56);           // FF2 of the previous use finished reading H[b] (only the stores wait)
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(hb + sw128_offset(row, (part & 1) * 32 + j * 8)) = pk[j];
                    } else {
                        mbar_wait(&hb_empty[b], ub ^ 1u, 56);
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();                            // H writes -> visible to the tensor core
                    __syncwarp();
                    if (skip && lane == 0) arrive_mma(&acc1_empty[b]);
                    if (lane == 0) arrive_mma(&hb_full[b]);
                    stamp(P_EPI1);
                }
                if (lane == 0) mbar_arrive(pm_empty);
                if (film_smem && ew == 0 && lane == 0) {
                    // H[0] is dead once FF2 of the last even chunk has completed: stage the next LayerNorm's FiLM rows there
                    const int nxt = (l + 1 < NL) ? trip : trip + trip_stride;
                    if (nxt < trips) {
                        mbar_wait(&hb_empty[0], (((n_p + 1u) * uses0) & 1u) ^ 1u, 59);
                        stage_film(tile_of(nxt), (l + 1 < NL) ? l + 1 : 0, 0);
                    }
                }
                mbar_wait(h_ready, 1u, 57);                              // the last FF2 has landed in h
                tc_fence_after();
                stamp(P_WH2);
                if (l == NL - 1 && trip + trip_stride < trips) stage_tab(tile_of(trip + trip_stride), false);   // every MMA reading X / H has completed
            }
            // ---- TMEM -> residual stream (+ the pending bias), or straight through the output head ----
            if (p.o_w != nullptr) {
                const float4* cbt = reinterpret_cast<const float4*>(p.cb_total + part * 64);
                uint32_t r[2][32];
                tmem_ld_32x32(tmem_row + part * 64, r[0]);
                tmem_ld_32x32(tmem_row + part * 64 + 32, r[1]);
                tmem_ld_wait();
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 cb = __ldg(cbt + cc * 8 + j);
                        const float x0 = __uint_as_float(r[cc][4 * j + 0]) + cb.x, x1 = __uint_as_float(r[cc][4 * j + 1]) + cb.y;
                        const float x2 = __uint_as_float(r[cc][4 * j + 2]) + cb.z, x3 = __uint_as_float(r[cc][4 * j + 3]) + cb.w;
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            if (o < p.o_D) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(p.o_w + o * kD + part * 64) + cc * 8 + j);
                                acc[o] = fmaf(x0, w.x, fmaf(x1, w.y, fmaf(x2, w.z, fmaf(x3, w.w, acc[o]))));
                            }
                        }
                    }
                float4* hp = reinterpret_cast<float4*>(smem + kOffO + 8192);     // [4 parts][128 rows] partial dot products
                hp[part * 128 + row] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                named_barrier_sync(3, kCT);
                if (part == 0 && live) {
                    const float4 a0 = hp[row], a1 = hp[128 + row], a2 = hp[256 + row], a3 = hp[384 + row];
                    const float y[4] = {((a0.x + a1.x) + (a2.x + a3.x)), ((a0.y + a1.y) + (a2.y + a3.y)), ((a0.z + a1.z) + (a2.z + a3.z)),
                                        ((a0.w + a1.w) + (a2.w + a3.w))};
#pragma unroll
                    for (int o = 0; o < 4; ++o)
                        if (o < p.o_D) p.o_y[m * p.o_D + o] = y[o] + __ldg(p.o_b + o);
                }
                named_barrier_sync(3, kCT);                              // hp is reused as LayerNorm statistics / O tile by the next tile
            } else {
                float4* dst = reinterpret_cast<float4*>(p.h + m * kD + part * 64);
                const float4* cbt = reinterpret_cast<const float4*>(p.cb_total + part * 64);
                uint32_t r[2][32];
                tmem_ld_32x32(tmem_row + part * 64, r[0]);
                tmem_ld_32x32(tmem_row + part * 64 + 32, r[1]);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 cb = __ldg(cbt + cc * 8 + j);
                            dst[cc * 8 + j] = make_float4(__uint_as_float(r[cc][4 * j + 0]) + cb.x, __uint_as_float(r[cc][4 * j + 1]) + cb.y,
                                                          __uint_as_float(r[cc][4 * j + 2]) + cb.z, __uint_as_float(r[cc][4 * j + 3]) + cb.w);
                        }
                }
            }
            stamp(P_STORE);
        }
        if (kProf && ew == 0 && lane == 0)
            for (int i = 0; i < P_N; ++i) atomicAdd(p.prof + i, pacc[i]);
    }

    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();     // pair: the peer's shared memory / TMEM stay valid until both are done
    if (warp == 2) {
        tc_fence_after();
        if (kPair) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ef

int encoder_fused(float* h, const float* params, const float* cb_total, const float* gb, long long gb_stride, long long gb_ln_stride, int film_folded,
                  const void* wqkv, const void* wo, const void* w1, const void* w2, long long M, int L, int d, int H, int ff, int n_layers,
                  int causal, const idb200_embed_t* emb, const idb200_head_t* head, cudaStream_t st) {
    IDB_REQUIRE(d == kD && H == 8, IDB200_EUNSUPPORTED, "fused encoder is specialised for d_model = 256, 8 heads (got %d, %d)", d, H);
    IDB_REQUIRE(L >= 1 && L <= 128 && (128 % L) == 0, IDB200_EUNSUPPORTED, "fused encoder needs L | 128 (got %d)", L);
    IDB_REQUIRE(ff % 128 == 0 && ff >= 128 && ff <= ef::kMaxFF, IDB200_EUNSUPPORTED, "fused encoder needs d_ff a multiple of 128, <= 1024 (got %d)", ff);
    IDB_REQUIRE(n_layers >= 1, IDB200_EINVAL, "n_layers must be >= 1");
    IDB_REQUIRE(M >= 0 && M % L == 0, IDB200_EINVAL, "M must be a multiple of L");
    IDB_REQUIRE(M < (1ll << 37), IDB200_EUNSUPPORTED, "M must be below 2^37 tokens (32-bit tile counters)");
    if (M == 0) return IDB200_OK;
    IDB_REQUIRE(params && cb_total && wqkv && wo && w1 && w2, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((h || (emb && head)), IDB200_EINVAL, "h may be NULL only when both the token assembly and the output head are fused");
    if (emb) {
        IDB_REQUIRE(emb->src0 && emb->Wf && emb->tab && emb->row_a && emb->row_b, IDB200_EINVAL, "NULL pointer in idb200_embed_t");
        IDB_REQUIRE(emb->n0 >= 1 && emb->n1 >= 0 && emb->n2 >= 0 && emb->n0 + emb->n1 + emb->n2 <= 16, IDB200_EUNSUPPORTED, "at most 16 token features");
        IDB_REQUIRE((emb->n1 == 0 || emb->src1) && (emb->n2 == 0 || emb->src2), IDB200_EINVAL, "NULL feature source");
        IDB_REQUIRE(aligned(emb->Wf, 16) && aligned(emb->tab, 16) && aligned(emb->row_a, 16) && aligned(emb->row_b, 16) && emb->row_a_stride % 4 == 0,
                    IDB200_EALIGN, "idb200_embed_t rows must be 16-byte aligned");
    }
    if (head) {
        IDB_REQUIRE(head->W && head->bias && head->y && head->D >= 1 && head->D <= 4, IDB200_EUNSUPPORTED, "output head needs 1 <= D <= 4");
        IDB_REQUIRE(aligned(head->W, 16), IDB200_EALIGN, "head weights must be 16-byte aligned");
    }
    IDB_REQUIRE(film_folded != 2 || !gb || L >= 8, IDB200_EUNSUPPORTED, "the bf16 [scale | shift] table is staged in shared memory: needs L >= 8");
    IDB_REQUIRE(film_folded != 2 || !gb || (gb_stride % 8 == 0 && gb_ln_stride % 8 == 0), IDB200_EALIGN, "bf16 FiLM rows must be 16-byte aligned");
    IDB_REQUIRE((!h || aligned(h, 16)) && aligned(params, 16) && aligned(cb_total, 16) && (!gb || (aligned(gb, 16) && gb_stride % 4 == 0 && gb_ln_stride % 4 == 0)),
                IDB200_EALIGN, "h / params / gamma_beta must be 16-byte aligned");
    static const bool pair_env = !(getenv("IDB200_ENCODER_PAIR") && atoi(getenv("IDB200_ENCODER_PAIR")) == 0);
    const long long tiles = (M + 127) / 128;
    const bool pair = pair_env && tiles >= 2;
    CUtensorMap tqk, tv, two, t1, t2;
    int rc = make_tmap_bf16_2d(&tqk, wqkv, static_cast<uint64_t>(n_layers) * 768, 256, pair ? 96 : 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tv, wqkv, static_cast<uint64_t>(n_layers) * 768, 256, 64, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&two, wo, static_cast<uint64_t>(n_layers) * 256, 256, 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&t1, w1, static_cast<uint64_t>(n_layers) * ff, 256, pair ? 64 : 128, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&t2, w2, static_cast<uint64_t>(n_layers) * 256, ff, 128, 64);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ef::encoder_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::kSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ef::encoder_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::kSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ef::encoder_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::kSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ef::encoder_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::kSmem);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(smem=%d): %s", ef::kSmem, cudaGetErrorString(e));
        attr = true;
    }
    int grid;
    if (pair) {
        const long long trips = (tiles + 1) / 2;
        static const int max_pairs = getenv("IDB200_EF_MAXPAIRS") ? atoi(getenv("IDB200_EF_MAXPAIRS")) : 1 << 30;   // dev: L2-contention probe
        long long pairs = trips < num_sms() / 2 ? trips : num_sms() / 2;
        if (pairs > max_pairs) pairs = max_pairs;
        grid = static_cast<int>(2 * pairs);
    } else {
        grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    }
    ef::Params p{};
    p.h = h; p.params = params; p.cb_total = cb_total; p.gb = gb; p.gb_stride = gb_stride; p.gb_ln_stride = gb_ln_stride; p.M = M; p.L = L; p.causal = causal;
    p.ff = ff; p.n_layers = n_layers; p.film_mode = gb ? (film_folded == 2 ? ef::kFilmFolded16 : film_folded ? ef::kFilmFolded : ef::kFilmRaw) : ef::kFilmNone; p.prof = nullptr;
    p.dbg_skip = getenv("IDB200_DBG_SKIP") ? atoi(getenv("IDB200_DBG_SKIP")) : 0;
    CUtensorMap ttab = tqk;                                                // (unused unless the table is staged)
    if (emb) {
        const int F = emb->n0 + emb->n1 + emb->n2;
        if (emb->tab_rows >= 1 && emb->tab_rows <= 64 && F <= 8 && L >= 8) {
            rc = make_tmap_2d(&ttab, emb->tab, 4, static_cast<uint64_t>(emb->tab_rows), 256, 64, 32);
            if (rc) return rc;
            p.e_stage = 1;
        }
        p.e_src0 = emb->src0; p.e_src1 = emb->src1; p.e_src2 = emb->src2; p.e_n0 = emb->n0; p.e_n1 = emb->n1; p.e_n2 = emb->n2;
        p.e_wf = emb->Wf; p.e_tab = emb->tab; p.e_tab_idx = reinterpret_cast<const long long*>(emb->tab_idx);
        p.e_row_a = emb->row_a; p.e_row_a_stride = emb->row_a_stride; p.e_row_b = emb->row_b;
    }
    if (head) { p.o_w = head->W; p.o_b = head->bias; p.o_y = head->y; p.o_D = head->D; }
    static const bool prof = getenv("IDB200_PROF") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ef::kThreads);
    cfg.dynamicSmemBytes = ef::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pair ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    auto launch = [&](bool profiled) -> cudaError_t {
        if (pair) return profiled ? cudaLaunchKernelEx(&cfg, ef::encoder_fused_kernel<true, true>, tqk, tv, two, t1, t2, ttab, p)
                                  : cudaLaunchKernelEx(&cfg, ef::encoder_fused_kernel<false, true>, tqk, tv, two, t1, t2, ttab, p);
        return profiled ? cudaLaunchKernelEx(&cfg, ef::encoder_fused_kernel<true, false>, tqk, tv, two, t1, t2, ttab, p)
                        : cudaLaunchKernelEx(&cfg, ef::encoder_fused_kernel<false, false>, tqk, tv, two, t1, t2, ttab, p);
    };
    if (prof) {                                                           // dev only: synchronous, prints the phase breakdown
        static unsigned long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, (ef::P_N + 12) * sizeof(unsigned long long));
        cudaMemsetAsync(dprof, 0, (ef::P_N + 12) * sizeof(unsigned long long), st);
        p.prof = dprof;
        cudaError_t e = launch(true);
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "encoder_fused_kernel<prof>: %s", cudaGetErrorString(e));
        unsigned long long hp[ef::P_N + 12];
        cudaMemcpyAsync(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        static const char* names[ef::P_N] = {"load", "ln1", "wait_acc", "epi", "att", "wait_o", "o_write", "wait_h1", "ln2", "wait_acc1", "epi1",
                                             "wait_h2", "store", "wait_pa", "(ln_pass1", "ln_bar", "ln_film", "ln_tmem_ld", "ln_entry)", "(load_turnaround", "load_wait_staged)", "[mma: wait_slot", "wait_compute", "issue]"};
        const double units = static_cast<double>(tiles) * n_layers;
        fprintf(stderr, "encoder_fused prof (cycles per tile-layer, L=%d, pair=%d):", L, pair ? 1 : 0);
        double tot = 0;
        for (int i = 0; i < ef::P_N; ++i) { fprintf(stderr, " %s=%.0f", names[i], hp[i] / units); if (i < ef::P_LNP1) tot += hp[i] / units; }
        fprintf(stderr, " total=%.0f\n", tot);
        static const char* tags[12] = {"acc_empty", "slot_qkv", "slot_v", "o_full", "slot_wo", "slot_wo2", "acc1_empty", "slot_ff1", "hb_full", "slot_ff2", "x_full_ln1", "x_full_ln2"};
        fprintf(stderr, "  mma issuer waits by barrier:");
        for (int i = 0; i < 12; ++i) fprintf(stderr, " %s=%.0f", tags[i], hp[ef::P_N + i] / units);
        fprintf(stderr, "\n");
        return check_launch("encoder_fused_kernel<prof>");
    }
    cudaError_t e = launch(false);
    if (e != cudaSuccess) return fail(IDB200_ECUDA, "encoder_fused_kernel: %s", cudaGetErrorString(e));
    return check_launch("encoder_fused_kernel");
}

}  // namespace idb200

extern "C" int idb200_encoder_fused(float* h, const float* layer_params, const float* bias_last, const float* film, int64_t film_stride,
                                    int64_t film_ln_stride, int film_folded, const void* wqkv_packed, const void* wo, const void* w1, const void* w2,
                                    int64_t M, int L, int d, int H, int ff, int n_layers, int causal, idb200_stream_t stream) {
    return idb200::encoder_fused(h, layer_params, bias_last, film, film_stride, film_ln_stride, film_folded, wqkv_packed, wo, w1, w2, M, L, d, H,
                                 ff, n_layers, causal, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int idb200_denoiser_fused(const idb200_embed_t* embed, const idb200_head_t* head, float* h, const float* layer_params,
                                     const float* bias_last, const float* film, int64_t film_stride, int64_t film_ln_stride,
                                     int film_folded, const void* wqkv_packed, const void* wo, const void* w1, const void* w2, int64_t M,
                                     int L, int d, int H, int ff, int n_layers, int causal, idb200_stream_t stream) {
    return idb200::encoder_fused(h, layer_params, bias_last, film, film_stride, film_ln_stride, film_folded, wqkv_packed, wo, w1, w2, M, L, d, H,
                                 ff, n_layers, causal, embed, head, static_cast<cudaStream_t>(stream));
}
