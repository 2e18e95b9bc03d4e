// Pieces shared by the fused transformer-block kernels (attn_block.cu, mlp_fused.cu): the LayerNorm + FiLM
// prologue that writes a 128-token bf16 MMA operand tile straight into shared memory, legacy mma.sync helpers
// for the in-tile attention core, and the TMEM -> TMA-reduce residual epilogue.
#pragma once
#include "tc_common.cuh"

namespace idb200 {
namespace fused {

using namespace tc;

constexpr int kD = 256;                 // d_model of both denoisers (8 heads x 32)
constexpr int kTile = 128 * 64 * 2;     // one [128 x 64] bf16 SWIZZLE_128B block = 16 KB
constexpr int kPitch = 200;             // bf16 elements per staged q|k|v row (400 B: conflict-free ldmatrix)

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}

// L2 prefetch of the h rows (and FiLM rows) one compute warp will normalise for tile m0: issued a tile ahead so the
// LayerNorm prologue's global loads hit L2 instead of HBM (the prologue is on the critical path of the compute warps).
__device__ __forceinline__ void ln_prefetch_l2(const float* __restrict__ h, long long m0, long long M, int L,
                                               const float* __restrict__ gb, long long gb_stride, int ew, int lane) {
    const long long r0 = m0 + ew * 16;
    if (r0 >= M) return;
    const long long rows = (M - r0 < 16) ? (M - r0) : 16;
    if (lane == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(h + r0 * kD), "r"(static_cast<uint32_t>(rows * kD * 4)) : "memory");
    if (gb != nullptr) {
        const long long t0 = r0 / L, t1 = (r0 + rows - 1) / L;
        if (t0 + lane <= t1)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gb + (t0 + lane) * gb_stride), "r"(2048u) : "memory");
    }
}

// LayerNorm(eps 1e-5) * (1 + gamma) + beta of the 128 rows m0..m0+127 of h [M, 256] (src/models/transformer.py:28-41),
// written as the bf16 K-major SWIZZLE_128B A operand X (4 k-blocks of [128 x 64]).  Called by the 8 compute warps
// (ew = 0..7, 16 rows each); a warp owns a row: lane <-> float4 columns {lane, lane + 32}, two-pass statistics by
// shuffles (same arithmetic as ln_film_kernel).  Rows are processed in batches of 4 with the next batch's loads in
// flight.  Rows >= M are written as zeros.  L = tokens per trajectory (FiLM parameters are per trajectory:
// gamma_beta row = m / L, [gamma (256) | beta (256)]).
__device__ __forceinline__ void ln_film_tile(const float* __restrict__ h, long long m0, long long M, int L,
                                             const float* __restrict__ gb, long long gb_stride, const float* s_lnw,
                                             const float* s_lnb, uint8_t* X, int ew, int lane) {
    const float4 wA = reinterpret_cast<const float4*>(s_lnw)[lane], wB = reinterpret_cast<const float4*>(s_lnw)[lane + 32];
    const float4 bA = reinterpret_cast<const float4*>(s_lnb)[lane], bB = reinterpret_cast<const float4*>(s_lnb)[lane + 32];
    const uint32_t offA = static_cast<uint32_t>((lane >> 4) * kTile), offB = static_cast<uint32_t>((2 + (lane >> 4)) * kTile);
    const int colin = 4 * (lane & 15);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[2][4], vb[2][4], film[2][4];             // double-buffered: rows, and [gamma A, gamma B, beta A, beta B] of the batch
    auto issue = [&](int batch, int buf) {
        const int rbase = ew * 16 + batch * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long m = m0 + rbase + i;
            if (m < M) {
                const float4* row = reinterpret_cast<const float4*>(h + m * kD);
                va[buf][i] = row[lane];
                vb[buf][i] = row[lane + 32];
            } else {
                va[buf][i] = zero4;
                vb[buf][i] = zero4;
            }
        }
        if (gb != nullptr && L >= 4 && m0 + rbase < M) {        // 4-row batches never straddle a trajectory when L >= 4
            const float4* g = reinterpret_cast<const float4*>(gb + ((m0 + rbase) / L) * gb_stride);
            film[buf][0] = g[lane];
            film[buf][1] = g[lane + 32];
            film[buf][2] = g[64 + lane];
            film[buf][3] = g[96 + lane];
        }
    };
    issue(0, 0);
#pragma unroll
    for (int batch = 0; batch < 4; ++batch) {
        const int buf = batch & 1;
        if (batch + 1 < 4) issue(batch + 1, buf ^ 1);
        const int rbase = ew * 16 + batch * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rbase + i;
            const long long m = m0 + r;
            const float4 xa = va[buf][i], xb = vb[buf][i];
            float sum = (xa.x + xa.y + xa.z + xa.w) + (xb.x + xb.y + xb.z + xb.w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / static_cast<float>(kD);
            const float a0 = xa.x - mean, a1 = xa.y - mean, a2 = xa.z - mean, a3 = xa.w - mean;
            const float b0 = xb.x - mean, b1 = xb.y - mean, b2 = xb.z - mean, b3 = xb.w - mean;
            float sq = (a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3) + (b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq / static_cast<float>(kD) + 1e-5f);
            float y0 = a0 * rstd * wA.x + bA.x, y1 = a1 * rstd * wA.y + bA.y, y2 = a2 * rstd * wA.z + bA.z, y3 = a3 * rstd * wA.w + bA.w;
            float z0 = b0 * rstd * wB.x + bB.x, z1 = b1 * rstd * wB.y + bB.y, z2 = b2 * rstd * wB.z + bB.z, z3 = b3 * rstd * wB.w + bB.w;
            if (gb != nullptr && m < M) {
                float4 gA = film[buf][0], gB = film[buf][1], tA = film[buf][2], tB = film[buf][3];
                if (L < 4) {
                    const float4* g = reinterpret_cast<const float4*>(gb + (m / L) * gb_stride);
                    gA = g[lane];
                    gB = g[lane + 32];
                    tA = g[64 + lane];
                    tB = g[96 + lane];
                }
                y0 = y0 * (1.0f + gA.x) + tA.x; y1 = y1 * (1.0f + gA.y) + tA.y; y2 = y2 * (1.0f + gA.z) + tA.z; y3 = y3 * (1.0f + gA.w) + tA.w;
                z0 = z0 * (1.0f + gB.x) + tB.x; z1 = z1 * (1.0f + gB.y) + tB.y; z2 = z2 * (1.0f + gB.z) + tB.z; z3 = z3 * (1.0f + gB.w) + tB.w;
            }
            if (m >= M) { y0 = y1 = y2 = y3 = z0 = z1 = z2 = z3 = 0.0f; }
            const uint32_t so = sw128_offset(r, colin);
            *reinterpret_cast<uint2*>(X + offA + so) = pack4_bf16(y0, y1, y2, y3);
            *reinterpret_cast<uint2*>(X + offB + so) = pack4_bf16(z0, z1, z2, z3);
        }
    }
}

// Generic-width form of ln_film_tile for the pair-mode kernels (qkv_attn.cu, mlp_pair.cu): d = VPL * 128 (256 / 384), called by
// kWarps compute warps (ew = 0 .. kWarps - 1, 128 / kWarps rows each), a warp owns a row: lane <-> float4 column groups
// {lane + 32 i}; batches of 4 rows with the next batch's loads in flight.  X: VPL * 2 k-blocks of [128 x 64] bf16 SWIZZLE_128B.
// gb: raw FiLM rows [gamma (d) | beta (d)] per trajectory (row m / L) or nullptr.  Same arithmetic as ln_film_kernel.
template <int VPL, int kWarps>
__device__ __forceinline__ void ln_film_rows(const float* __restrict__ h, long long m0, long long M, int L, const float* __restrict__ gb,
                                             long long gb_stride, const float* s_lnw, const float* s_lnb, uint8_t* X, int ew, int lane) {
    constexpr int d = VPL * 128;
    constexpr int kRows = 128 / kWarps;
    constexpr int kB = 2;                                  // rows per batch (two batches of loads in flight: 2 * kB * VPL float4 registers)
    static_assert(kRows % kB == 0, "rows per warp in whole batches");
    const int colin = 4 * (lane & 15);
    float4 va[2][kB][VPL];
    auto issue = [&](int batch, int buf) {
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            const long long m = m0 + ew * kRows + batch * kB + i;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                va[buf][i][v] = (m < M) ? reinterpret_cast<const float4*>(h + m * d)[lane + 32 * v] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    issue(0, 0);
#pragma unroll
    for (int batch = 0; batch < kRows / kB; ++batch) {
        const int buf = batch & 1;
        if (batch + 1 < kRows / kB) issue(batch + 1, buf ^ 1);
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            const int r = ew * kRows + batch * kB + i;
            const long long m = m0 + r;
            float sum = 0.0f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) sum += (va[buf][i][v].x + va[buf][i][v].y) + (va[buf][i][v].z + va[buf][i][v].w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / static_cast<float>(d);
            float sq = 0.0f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const float a0 = va[buf][i][v].x - mean, a1 = va[buf][i][v].y - mean, a2 = va[buf][i][v].z - mean, a3 = va[buf][i][v].w - mean;
                sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = rsqrtf(sq / static_cast<float>(d) + 1e-5f);
            const float* g = (gb != nullptr && m < M) ? gb + (m / L) * gb_stride : nullptr;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = lane + 32 * v;
                const float4 w4 = reinterpret_cast<const float4*>(s_lnw)[c4], b4 = reinterpret_cast<const float4*>(s_lnb)[c4];
                float y0 = (va[buf][i][v].x - mean) * rstd * w4.x + b4.x, y1 = (va[buf][i][v].y - mean) * rstd * w4.y + b4.y;
                float y2 = (va[buf][i][v].z - mean) * rstd * w4.z + b4.z, y3 = (va[buf][i][v].w - mean) * rstd * w4.w + b4.w;
                if (g != nullptr) {
                    const float4 ga = __ldg(reinterpret_cast<const float4*>(g) + c4), be = __ldg(reinterpret_cast<const float4*>(g + d) + c4);
                    y0 = y0 * (1.0f + ga.x) + be.x; y1 = y1 * (1.0f + ga.y) + be.y; y2 = y2 * (1.0f + ga.z) + be.z; y3 = y3 * (1.0f + ga.w) + be.w;
                }
                if (m >= M) { y0 = y1 = y2 = y3 = 0.0f; }
                // column 4 c4 .. 4 c4 + 3 lives in k-block c4 / 16 = 2 v + (lane >> 4), at column 4 (lane & 15) of it
                *reinterpret_cast<uint2*>(X + (2 * v + (lane >> 4)) * kTile + sw128_offset(r, colin)) = pack4_bf16(y0, y1, y2, y3);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// legacy tensor-core helpers (mma.sync m16n8k16 + ldmatrix) for the in-tile attention core
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldsm_x4_trans(unsigned (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void hmma_16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}

// softmax(q k^T * scale) v for one (16-row block, head): rows rb*16..+15 of the tile, keys kbeg..kend-1 (tile rows).
// NT = key tiles of 8 per 64-key step that can be live (2 when the key range is <= 16 rows, else 8).
// lgblk >= 0: trajectories of 2^lgblk < 16 rows share the 16-row block (block-diagonal mask); -1: none.
template <int NT>
__device__ __forceinline__ void attn_unit(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int rb, int kbeg,
                                          int kend, int lgblk, int causal, int lane, float (&o)[4][4]) {
    constexpr float kScaleLog2 = 0.17677669529663687f * 1.4426950408889634f;
    unsigned qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int row = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(qa[ks], Q + row * kPitch + ks * 16 + (lane >> 4) * 8);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[nt][c] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = rb * 16 + g, qrow1 = qrow0 + 8;
    for (int k0 = kbeg; k0 < kend; k0 += 8 * NT) {
        const int kw = min(8 * NT, kend - k0);                          // multiple of 16
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) s[nt][c] = 0.0f;
            if (nt * 8 < kw) {
                unsigned kb[4];
                ldsm_x4(kb, K + (k0 + nt * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);
                hmma_16816(s[nt], qa[0], kb[0], kb[1]);
                hmma_16816(s[nt], qa[1], kb[2], kb[3]);
            }
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int key = k0 + nt * 8 + tq * 2 + (c & 1);
                const int qr = (c < 2) ? qrow0 : qrow1;
                const bool ok = (nt * 8 < kw) && (!causal || key <= qr) && (lgblk < 0 || ((key ^ qr) >> lgblk) == 0);
                s[nt][c] = ok ? s[nt][c] * kScaleLog2 : -INFINITY;
            }
            bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);         // finite: a query always sees its own key
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.0f, rs1 = 0.0f;
        unsigned pa[NT / 2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float p0 = exp2f(s[nt][0] - mn0), p1 = exp2f(s[nt][1] - mn0);
            const float p2 = exp2f(s[nt][2] - mn1), p3 = exp2f(s[nt][3] - mn1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p2, p3);
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        if (k0 > kbeg) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
        }
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 16 < kw) {
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    unsigned vb[4];
                    const int key = k0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    ldsm_x4_trans(vb, V + key * kPitch + np * 16 + (lane >> 4) * 8);
                    hmma_16816(o[np * 2 + 0], pa[ks], vb[0], vb[1]);
                    hmma_16816(o[np * 2 + 1], pa[ks], vb[2], vb[3]);
                }
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= i0; o[nt][1] *= i0; o[nt][2] *= i1; o[nt][3] *= i1; }
}

// pure-register variants (not volatile: the compiler may interleave independent chains)
__device__ __forceinline__ void hmma_16816_nv(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Leaner single-head unit used by the whole-encoder kernel (one unit per warp): softmax scale folded into the exponent
// (p = 2^(s * c - m * c), one FFMA per score), ex2.approx / rcp.approx, non-volatile MMAs, and the mask specialised at
// compile time: kMode 0 none (whole trajectories of >= 16 rows, bidirectional), 1 block-diagonal from 8 precomputed bits
// (trajectories of < 16 rows share the 16-row block; bit nt*4+c of okbits), 2 causal (key <= query row).
template <int NT, int kMode>
__device__ __forceinline__ void attn_unit_fast(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int rb, int kbeg,
                                               int kend, uint32_t okbits, int lane, float (&o)[4][4]) {
    constexpr float kS = 0.17677669529663687f * 1.4426950408889634f;
    unsigned qa[2][4];
    {
        const __nv_bfloat16* qp = Q + (rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 8;
        ldsm_x4(qa[0], qp);
        ldsm_x4(qa[1], qp + 16);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[nt][c] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = rb * 16 + g, qrow1 = qrow0 + 8;
    const __nv_bfloat16* kp = K + (lane & 7) * kPitch + (lane >> 3) * 8;
    const __nv_bfloat16* vp = V + ((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 8;
    for (int k0 = kbeg; k0 < kend; k0 += 8 * NT) {
        const int kw = min(8 * NT, kend - k0);                          // multiple of 16
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) s[nt][c] = 0.0f;
            if (nt * 8 < kw) {
                unsigned kb[4];
                ldsm_x4(kb, kp + (k0 + nt * 8) * kPitch);
                hmma_16816_nv(s[nt], qa[0], kb[0], kb[1]);
                hmma_16816_nv(s[nt], qa[1], kb[2], kb[3]);
            }
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                bool ok = nt * 8 < kw;
                if (kMode == 1) ok = ok && ((okbits >> (nt * 4 + c)) & 1u);
                if (kMode == 2) ok = ok && (k0 + nt * 8 + tq * 2 + (c & 1) <= ((c < 2) ? qrow0 : qrow1));
                if (kMode != 0 || NT * 8 > 16) s[nt][c] = ok ? s[nt][c] : -INFINITY;
            }
            bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);         // finite: a query always sees its own key
        const float c0 = ex2_fast((m0 - mn0) * kS), c1 = ex2_fast((m1 - mn1) * kS);
        m0 = mn0;
        m1 = mn1;
        const float n0 = -mn0 * kS, n1 = -mn1 * kS;
        float rs0 = 0.0f, rs1 = 0.0f;
        unsigned pa[NT / 2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float p0 = ex2_fast(fmaf(s[nt][0], kS, n0)), p1 = ex2_fast(fmaf(s[nt][1], kS, n0));
            const float p2 = ex2_fast(fmaf(s[nt][2], kS, n1)), p3 = ex2_fast(fmaf(s[nt][3], kS, n1));
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p2, p3);
        }
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        if (k0 > kbeg) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
        }
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 16 < kw) {
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    unsigned vb[4];
                    ldsm_x4_trans(vb, vp + (k0 + ks * 16) * kPitch + np * 16);
                    hmma_16816_nv(o[np * 2 + 0], pa[ks], vb[0], vb[1]);
                    hmma_16816_nv(o[np * 2 + 1], pa[ks], vb[2], vb[3]);
                }
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = rcp_fast(l0), i1 = rcp_fast(l1);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= i0; o[nt][1] *= i0; o[nt][2] *= i1; o[nt][3] *= i1; }
}


// Both heads of a head group at once (two independent dependency chains interleaved instruction by instruction):
// softmax(q k^T / sqrt(32)) v for the 16-row block rb, keys kbeg..kend-1 (tile rows), from the staged bf16 rows
// sq[row][q 64 | k 64 | v 64] (pitch kPitch, head hh at column offset 32 hh of each part).  NT key tiles of 8 per step.
template <int NT>
__device__ __forceinline__ void attn_unit2(const __nv_bfloat16* sq, int rb, int kbeg, int kend, int lgblk, int causal, int lane,
                                           float (&o)[2][4][4]) {
    constexpr float kScaleLog2 = 0.17677669529663687f * 1.4426950408889634f;
    unsigned qa[2][2][4];
    {
        const __nv_bfloat16* qp = sq + (rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 8;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) ldsm_x4(qa[hh][ks], qp + hh * 32 + ks * 16);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[hh][nt][c] = 0.0f;
    float mx[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}}, ls[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = rb * 16 + g, qrow1 = qrow0 + 8;
    const __nv_bfloat16* kp = sq + (lane & 7) * kPitch + 64 + (lane >> 3) * 8;
    const __nv_bfloat16* vp = sq + ((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + 128 + (lane >> 4) * 8;
    for (int k0 = kbeg; k0 < kend; k0 += 8 * NT) {
        const int kw = min(8 * NT, kend - k0);                          // multiple of 16
        float s[2][NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[hh][nt][c] = 0.0f;
            if (nt * 8 < kw) {
                unsigned kb[2][4];
                ldsm_x4(kb[0], kp + (k0 + nt * 8) * kPitch);
                ldsm_x4(kb[1], kp + (k0 + nt * 8) * kPitch + 32);
                hmma_16816_nv(s[0][nt], qa[0][0], kb[0][0], kb[0][1]);
                hmma_16816_nv(s[1][nt], qa[1][0], kb[1][0], kb[1][1]);
                hmma_16816_nv(s[0][nt], qa[0][1], kb[0][2], kb[0][3]);
                hmma_16816_nv(s[1][nt], qa[1][1], kb[1][2], kb[1][3]);
            }
        }
        float bm[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int key = k0 + nt * 8 + tq * 2 + (c & 1);
                const int qr = (c < 2) ? qrow0 : qrow1;
                const bool ok = (nt * 8 < kw) && (!causal || key <= qr) && (lgblk < 0 || ((key ^ qr) >> lgblk) == 0);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) s[hh][nt][c] = ok ? s[hh][nt][c] * kScaleLog2 : -INFINITY;
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                bm[hh][0] = fmaxf(bm[hh][0], fmaxf(s[hh][nt][0], s[hh][nt][1]));
                bm[hh][1] = fmaxf(bm[hh][1], fmaxf(s[hh][nt][2], s[hh][nt][3]));
            }
        }
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int i = 0; i < 2; ++i) bm[hh][i] = fmaxf(bm[hh][i], __shfl_xor_sync(0xffffffffu, bm[hh][i], off));
        float corr[2][2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float mn = fmaxf(mx[hh][i], bm[hh][i]);            // finite: a query always sees its own key
                corr[hh][i] = ex2_fast(mx[hh][i] - mn);
                mx[hh][i] = mn;
            }
        unsigned pa[2][NT / 2][4];
        float rs[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float p0 = ex2_fast(s[hh][nt][0] - mx[hh][0]), p1 = ex2_fast(s[hh][nt][1] - mx[hh][0]);
                const float p2 = ex2_fast(s[hh][nt][2] - mx[hh][1]), p3 = ex2_fast(s[hh][nt][3] - mx[hh][1]);
                rs[hh][0] += p0 + p1;
                rs[hh][1] += p2 + p3;
                pa[hh][nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p0, p1);
                pa[hh][nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p2, p3);
            }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            ls[hh][0] = ls[hh][0] * corr[hh][0] + rs[hh][0];
            ls[hh][1] = ls[hh][1] * corr[hh][1] + rs[hh][1];
        }
        if (k0 > kbeg) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    o[hh][nt][0] *= corr[hh][0]; o[hh][nt][1] *= corr[hh][0];
                    o[hh][nt][2] *= corr[hh][1]; o[hh][nt][3] *= corr[hh][1];
                }
        }
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            if (ks * 16 < kw) {
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    unsigned vb[2][4];
                    ldsm_x4_trans(vb[0], vp + (k0 + ks * 16) * kPitch + np * 16);
                    ldsm_x4_trans(vb[1], vp + (k0 + ks * 16) * kPitch + 32 + np * 16);
                    hmma_16816_nv(o[0][np * 2 + 0], pa[0][ks], vb[0][0], vb[0][1]);
                    hmma_16816_nv(o[1][np * 2 + 0], pa[1][ks], vb[1][0], vb[1][1]);
                    hmma_16816_nv(o[0][np * 2 + 1], pa[0][ks], vb[0][2], vb[0][3]);
                    hmma_16816_nv(o[1][np * 2 + 1], pa[1][ks], vb[1][2], vb[1][3]);
                }
            }
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ls[hh][i] += __shfl_xor_sync(0xffffffffu, ls[hh][i], 1);
            ls[hh][i] += __shfl_xor_sync(0xffffffffu, ls[hh][i], 2);
        }
        const float i0 = rcp_fast(ls[hh][0]), i1 = rcp_fast(ls[hh][1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { o[hh][nt][0] *= i0; o[hh][nt][1] *= i0; o[hh][nt][2] *= i1; o[hh][nt][3] *= i1; }
    }
}

// ------------------------------------------------------------------------------------------------
// residual epilogue: h[m0 .. m0+127, 0..255] += acc (TMEM, 256 fp32 columns at tmem_acc) + bias.
// The fp32 tile is staged through `stage` (64 KB, 1024-byte aligned) as [128 x 32] SWIZZLE_128B boxes and added to
// global memory by TMA reduce-add: no read of h, fully coalesced, rows >= M are clipped by the tensor map.  Four
// rounds of 64 columns, double-buffered (2 boxes per round) so the TMA engine drains round r while the warps stage
// round r+1.  Called by the 8 compute warps; uses named barriers 1 and 2 (256 threads).  On return the TMA engine has
// finished READING `stage` (the global writes may still be in flight), so the caller may overwrite it at once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void residual_epilogue(uint32_t tmem_acc, const float* s_bias, uint8_t* stage, const CUtensorMap* tmap_h,
                                                  int m0, int ew, int lane) {
    const int q = ew & 3, half = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
#pragma unroll 1
    for (int rnd = 0; rnd < 4; ++rnd) {
        const int col = rnd * 64 + half * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + lane_base + col, r);
        tmem_ld_wait();
        uint8_t* box = stage + ((rnd & 1) * 2 + half) * kTile;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + col + 4 * j);
            const float4 v = make_float4(__uint_as_float(r[4 * j + 0]) + bv.x, __uint_as_float(r[4 * j + 1]) + bv.y,
                                         __uint_as_float(r[4 * j + 2]) + bv.z, __uint_as_float(r[4 * j + 3]) + bv.w);
            *reinterpret_cast<float4*>(box + row * 128 + ((j ^ (row & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        named_barrier_sync(1, 256);
        if (ew == 0 && lane == 0) {
            tma_reduce_add_2d(tmap_h, stage + ((rnd & 1) * 2 + 0) * kTile, rnd * 64, m0);
            tma_reduce_add_2d(tmap_h, stage + ((rnd & 1) * 2 + 1) * kTile, rnd * 64 + 32, m0);
            tma_store_commit();
            if (rnd < 3) tma_store_wait_read<1>();              // the other buffer (round rnd-1) has been read
            else tma_store_wait_read<0>();                      // last round: `stage` is free for the caller on return
        }
        named_barrier_sync(2, 256);
    }
}

}  // namespace fused
}  // namespace idb200
