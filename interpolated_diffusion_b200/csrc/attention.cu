// K3b: small-sequence multi-head attention (bidirectional or causal), head_dim = 32.
//
// Reference: nn.MultiheadAttention inside src/models/transformer.py:11,39 (packed QKV already projected by
// the token GEMM; this file is softmax(Q K^T / sqrt(32) [+ causal mask]) V per trajectory and head; the
// -inf upper-triangular mask of transformer.py:68-71 is the `causal` flag).
//
//   attn_simt   fp32 CUDA-core path: L = 8 (Stage 1, 0.5 % of the FLOPs) and the fp32 check mode.
//               Lanes map to (problem, query); several (trajectory, head) problems share a warp when L < 32.
//   attn_mma    bf16 mma.sync m16n8k16 path with ldmatrix fragments and an online softmax over 64-key
//               blocks: L = 16..256 (Stage 2, L = 64; long-horizon causal L = 256).  One warp per
//               (head, 16-query block); Q/K/V of the CTA's (trajectory, head group) are staged in shared
//               memory with cp.async.  Attention is <= 14 % of the FLOPs, the tcgen05
//               pipeline is reserved for the projections.
// qkv: [M, 3d] rows = tokens, columns [q | k | v], head h at columns h*32 inside each third.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace idb200 {

// attention_tc5.cu: the tcgen05 / TMEM path (default for bf16)
bool attention_tc5_supported(const void* qkv, const void* out, long long B, int L, int H);
int attention_tc5(const void* qkv, void* out, long long B, int L, int H, int causal, cudaStream_t st);

constexpr int kHD = 32;

template <typename T> __device__ __forceinline__ float ld_f32(const T* p);
template <> __device__ __forceinline__ float ld_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f32(T* p, float v);
template <> __device__ __forceinline__ void st_f32<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// SIMT path
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, long long B, int L, int H,
                                                        int causal) {
    extern __shared__ float smem_kv[];
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = (L < 32) ? 32 / L : 1;          // problems per warp
    const int d = H * kHD;
    const long long problems = B * H;
    const long long groups = (problems + G - 1) / G;
    float* Ks = smem_kv + static_cast<size_t>(warp) * 2 * G * L * (kHD + 1);
    float* Vs = Ks + static_cast<size_t>(G) * L * (kHD + 1);
    const float scale = rsqrtf(static_cast<float>(kHD));
    for (long long grp = static_cast<long long>(blockIdx.x) * warps + warp; grp < groups; grp += static_cast<long long>(gridDim.x) * warps) {
        __syncwarp();
        // stage K and V of the G problems: element e -> (g, j, c)
        for (int e = lane; e < G * L * kHD; e += 32) {
            const int c = e % kHD, j = (e / kHD) % L, g = e / (kHD * L);
            const long long pr = grp * G + g;
            float kv = 0.0f, vv = 0.0f;
            if (pr < problems) {
                const long long b = pr / H;
                const int h = static_cast<int>(pr - b * H);
                const T* row = qkv + (b * L + j) * 3 * d;
                kv = ld_f32<T>(row + d + h * kHD + c);
                vv = ld_f32<T>(row + 2 * d + h * kHD + c);
            }
            Ks[(g * L + j) * (kHD + 1) + c] = kv;
            Vs[(g * L + j) * (kHD + 1) + c] = vv;
        }
        __syncwarp();
        const int reps = (L + 31) / 32;
        for (int r = 0; r < reps; ++r) {
            const int g = (L < 32) ? lane / L : 0;
            const int i = (L < 32) ? lane % L : lane + 32 * r;
            const long long pr = grp * G + g;
            if (g >= G || i >= L || pr >= problems) continue;
            const long long b = pr / H;
            const int h = static_cast<int>(pr - b * H);
            const T* qrow = qkv + (b * L + i) * 3 * d + h * kHD;
            float q[kHD], o[kHD];
#pragma unroll
            for (int c = 0; c < kHD; ++c) { q[c] = ld_f32<T>(qrow + c) * scale; o[c] = 0.0f; }
            float mx = -INFINITY, den = 0.0f;
            const int jend = causal ? i + 1 : L;
            for (int j = 0; j < jend; ++j) {
                const float* kr = Ks + (g * L + j) * (kHD + 1);
                const float* vr = Vs + (g * L + j) * (kHD + 1);
                float s = 0.0f;
#pragma unroll
                for (int c = 0; c < kHD; ++c) s = fmaf(q[c], kr[c], s);
                const float mn = fmaxf(mx, s);
                const float corr = expf(mx - mn);
                const float pj = expf(s - mn);
                den = den * corr + pj;
#pragma unroll
                for (int c = 0; c < kHD; ++c) o[c] = fmaf(pj, vr[c], o[c] * corr);
                mx = mn;
            }
            const float inv = 1.0f / den;
            T* orow = out + (b * L + i) * d + h * kHD;
#pragma unroll
            for (int c = 0; c < kHD; ++c) st_f32<T>(orow + c, o[c] * inv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mma.sync path (bf16)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_ptr))));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* smem_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_ptr))));
}
// D(16x8 f32) += A(16x16 bf16, row) * B(16x8 bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}

constexpr int kPitch = 40;      // bf16 elements per staged row (80 bytes): conflict-free ldmatrix rows

// grid.x = B * (H / HG); CTA stages Q, K, V of HG heads of one trajectory: [3][HG][L][kPitch] bf16
// blk > 0: block-diagonal attention -- the L rows of a "sequence" are L / blk independent trajectories of blk tokens
// (Stage 1: blk = 8 tokens per trajectory, 8 trajectories per 64-row sequence), a query only sees keys of its own block.
// kBlk: block-diagonal (short-trajectory) instantiation; it alone skips the dead key tiles of its 16-key steps (in the long-sequence
// instantiation the extra branches cost 8 % at L = 256, measured)
template <bool kBlk>
__global__ void __launch_bounds__(256) attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                       long long B, int L, int H, int HG, int causal, int blk) {
    extern __shared__ __align__(16) unsigned char smem_attn[];
    __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    const int d = H * kHD;
    const int groups_per_b = H / HG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const size_t part = static_cast<size_t>(HG) * L * kPitch;       // elements per Q / K / V region
    const float scale_log2 = rsqrtf(static_cast<float>(kHD)) * 1.4426950408889634f;
    for (long long cta = blockIdx.x; cta < B * groups_per_b; cta += gridDim.x) {
        const long long b = cta / groups_per_b;
        const int h0 = static_cast<int>(cta - b * groups_per_b) * HG;
        __syncthreads();
        // stage: 16-byte chunks; per token row, per third (q/k/v), HG*32 contiguous bf16 = HG*4 chunks.  A warp takes whole
        // token rows and a lane keeps its chunk slots (third, head, 16-byte piece) fixed across them: the index divisions run
        // once per lane and unit instead of once per chunk (they were a quarter of this kernel's instructions)
        const int chunks_per_tok = 3 * HG * 4;
        if (chunks_per_tok >= 32) {
            const __nv_bfloat16* gbase = qkv + b * L * 3 * d + h0 * kHD;
            for (int r = lane; r < chunks_per_tok; r += 32) {
                const int third = r / (HG * 4), rr = r - third * (HG * 4);
                const int hh = rr >> 2, ck = rr & 3;
                const __nv_bfloat16* src0 = gbase + third * d + hh * kHD + ck * 8;
                __nv_bfloat16* dst0 = sm + third * part + static_cast<size_t>(hh) * L * kPitch + ck * 8;
                for (int tok = warp; tok < L; tok += nwarps) {
                    const __nv_bfloat16* src = src0 + static_cast<size_t>(tok) * 3 * d;
                    __nv_bfloat16* dst = dst0 + tok * kPitch;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
                }
            }
        } else {                                                     // few heads per unit (long sequences): flat chunk loop
            for (int e = threadIdx.x; e < L * chunks_per_tok; e += blockDim.x) {
                const int tok = e / chunks_per_tok, r = e - tok * chunks_per_tok;
                const int third = r / (HG * 4), rr = r - third * (HG * 4);
                const int hh = rr >> 2, ck = rr & 3;
                const __nv_bfloat16* src = qkv + (b * L + tok) * 3 * d + third * d + (h0 + hh) * kHD + ck * 8;
                __nv_bfloat16* dst = sm + third * part + (static_cast<size_t>(hh) * L + tok) * kPitch + ck * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        const int qblocks = L / 16;
        for (int task = warp; task < HG * qblocks; task += nwarps) {
            const int hh = task / qblocks, qb = task - hh * qblocks;
            const __nv_bfloat16* Qs = sm + (static_cast<size_t>(hh) * L + qb * 16) * kPitch;
            const __nv_bfloat16* Ks = sm + part + static_cast<size_t>(hh) * L * kPitch;
            const __nv_bfloat16* Vs = sm + 2 * part + static_cast<size_t>(hh) * L * kPitch;
            // Q fragments: two k-steps (dims 0-15, 16-31); ldmatrix x4: matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7),
            // (rows 0-7, k 8-15), (rows 8-15, k 8-15) -> a0..a3
            unsigned qa[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
                const int col = ks * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qa[ks], Qs + row * kPitch + col);
            }
            float o[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[nt][c] = 0.0f;
            float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;   // rows g and g+8 of the 16-row block
            const int g = lane >> 2, tq = lane & 3;
            const int qrow0 = qb * 16 + g, qrow1 = qrow0 + 8;
            const int lgblk = kBlk ? 31 - __clz(blk) : 0;
            int kbeg = 0, kend = causal ? (qb + 1) * 16 : L;               // keys needed by this query block
            if (kBlk) {                                                    // only the blocks the 16 query rows belong to
                kbeg = (((qb * 16) / blk) * blk) & ~15;
                const int ke = ((qb * 16 + 15) / blk + 1) * blk;
                kend = min(kend, min(L, (ke + 15) & ~15));
            }
            for (int k0 = kbeg; k0 < kend; k0 += 64) {
                const int kw = min(64, kend - k0);                         // multiple of 16
                float s[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[nt][c] = 0.0f;
                    if (nt * 8 < kw) {
                        // B fragments of K for keys k0 + nt*8 .. +7: matrices (dims 0-7), (8-15), (16-23), (24-31)
                        unsigned kb[4];
                        ldmatrix_x4(kb, Ks + (k0 + nt * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);
                        mma_bf16_16816(s[nt], qa[0], kb[0], kb[1]);
                        mma_bf16_16816(s[nt], qa[1], kb[2], kb[3]);
                    }
                }
                // scale, mask, block max
                float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    if (kBlk && nt * 8 >= kw) {                             // dead key tile of a short step: no mask / max work
                        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = -INFINITY;
                        continue;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int key = k0 + nt * 8 + tq * 2 + (c & 1);
                        const int qr = (c < 2) ? qrow0 : qrow1;
                        // blk is a power of two (2, 4, 8): same-trajectory test by shift, not by two integer divisions per element
                        // (measured: the divisions made this kernel ALU-issue bound, 1700 instructions per 16-row task)
                        const bool ok = (nt * 8 < kw) && (!causal || key <= qr) && (!kBlk || ((key ^ qr) >> lgblk) == 0);
                        s[nt][c] = ok ? s[nt][c] * scale_log2 : -INFINITY;
                    }
                    bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
                    bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
                }
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
                bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
                bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
                const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);     // finite: key 0 is always visible
                const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
                m0 = mn0; m1 = mn1;
                float rs0 = 0.0f, rs1 = 0.0f;
                unsigned pa[4][4];                                          // P as A fragments, one per 16-key step
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    if (kBlk && nt * 8 >= kw) {
                        pa[nt >> 1][(nt & 1) * 2 + 0] = 0u;
                        pa[nt >> 1][(nt & 1) * 2 + 1] = 0u;
                        continue;
                    }
                    const float p0 = exp2f(s[nt][0] - mn0), p1 = exp2f(s[nt][1] - mn0);
                    const float p2 = exp2f(s[nt][2] - mn1), p3 = exp2f(s[nt][3] - mn1);
                    rs0 += p0 + p1;
                    rs1 += p2 + p3;
                    // C layout of key tiles (2j, 2j+1) -> A fragment of k-step j: a0 = rows g keys 0-7, a1 = rows g+8 keys 0-7,
                    // a2 = rows g keys 8-15, a3 = rows g+8 keys 8-15
                    pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
                    pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
                }
                l0 = l0 * c0 + rs0;
                l1 = l1 * c1 + rs1;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
                // O += P V: k-steps of 16 keys; B fragments of V via transposed ldmatrix:
                // matrices (keys 0-7, dims n0..n0+7), (keys 8-15, same dims), (keys 0-7, dims n0+8..), (keys 8-15, ...)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks * 16 < kw) {
#pragma unroll
                        for (int np = 0; np < 2; ++np) {
                            unsigned vb[4];
                            const int key = k0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                            const int col = np * 16 + (lane >> 4) * 8;
                            ldmatrix_x4_trans(vb, Vs + key * kPitch + col);
                            mma_bf16_16816(o[np * 2 + 0], pa[ks], vb[0], vb[1]);
                            mma_bf16_16816(o[np * 2 + 1], pa[ks], vb[2], vb[3]);
                        }
                    }
                }
            }
            // the row sums live in the 4 lanes of a quad
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            const float i0 = 1.0f / l0, i1 = 1.0f / l1;
            __nv_bfloat16* o0 = out + (b * L + qrow0) * d + (h0 + hh) * kHD;
            __nv_bfloat16* o1 = out + (b * L + qrow1) * d + (h0 + hh) * kHD;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<unsigned*>(o0 + nt * 8 + tq * 2) = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
                *reinterpret_cast<unsigned*>(o1 + nt * 8 + tq * 2) = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Cross attention (KeypointSelector, src/models/keypoint_selector.py:22-38): Lq query tokens attend La + Lb memory tokens
// (La spatial tokens + Lb extra tokens kept in a second buffer; key order is irrelevant to softmax(QK^T)V).
//   q [B, Lq, d] bf16 (projected queries), kv_a [B, La, 2d], kv_b [B, Lb, 2d] bf16 ([K | V] projections), out [B, Lq, d].
// One CTA per (sample, head): Q, K, V of the head staged in shared memory (keys padded to a multiple of 16 with zeros and
// masked), 16-query blocks per warp, online softmax over 64-key steps, same mma.sync / ldmatrix fragments as attn_mma_kernel.
__global__ void __launch_bounds__(128) cross_attn_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv_a,
                                                             const __nv_bfloat16* __restrict__ kv_b, __nv_bfloat16* __restrict__ out,
                                                             int Lq, int La, int Lb, int H) {
    extern __shared__ __align__(16) unsigned char smem_attn[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    const int d = H * kHD;
    const int Lk = La + Lb, Lkp = (Lk + 15) & ~15;
    __nv_bfloat16* Ks = Qs + static_cast<size_t>(Lq) * kPitch;
    __nv_bfloat16* Vs = Ks + static_cast<size_t>(Lkp) * kPitch;
    const long long b = blockIdx.x / H;
    const int hh = blockIdx.x % H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale_log2 = rsqrtf(static_cast<float>(kHD)) * 1.4426950408889634f;
    for (int e = threadIdx.x; e < Lq * 4; e += blockDim.x) {
        const int tok = e >> 2, ck = e & 3;
        *reinterpret_cast<uint4*>(Qs + tok * kPitch + ck * 8) = *reinterpret_cast<const uint4*>(q + (b * Lq + tok) * d + hh * kHD + ck * 8);
    }
    for (int e = threadIdx.x; e < Lkp * 8; e += blockDim.x) {
        const int tok = e >> 3, r = e & 7, third = r >> 2, ck = r & 3;              // third: 0 = K, 1 = V
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (tok < La) v = *reinterpret_cast<const uint4*>(kv_a + (b * La + tok) * 2 * d + third * d + hh * kHD + ck * 8);
        else if (tok < Lk) v = *reinterpret_cast<const uint4*>(kv_b + (b * Lb + (tok - La)) * 2 * d + third * d + hh * kHD + ck * 8);
        *reinterpret_cast<uint4*>((third ? Vs : Ks) + tok * kPitch + ck * 8) = v;
    }
    __syncthreads();
    const int g = lane >> 2, tq = lane & 3;
    for (int qb = warp; qb < Lq / 16; qb += blockDim.x >> 5) {
        unsigned qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            ldmatrix_x4(qa[ks], Qs + (qb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + ks * 16 + (lane >> 4) * 8);
        float o[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[nt][c] = 0.0f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
        for (int k0 = 0; k0 < Lkp; k0 += 64) {
            const int kw = min(64, Lkp - k0);
            float sc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int c = 0; c < 4; ++c) sc[nt][c] = 0.0f;
                if (nt * 8 < kw) {
                    unsigned kb[4];
                    ldmatrix_x4(kb, Ks + (k0 + nt * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);
                    mma_bf16_16816(sc[nt], qa[0], kb[0], kb[1]);
                    mma_bf16_16816(sc[nt], qa[1], kb[2], kb[3]);
                }
            }
            float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int key = k0 + nt * 8 + tq * 2 + (c & 1);
                    sc[nt][c] = (nt * 8 < kw && key < Lk) ? sc[nt][c] * scale_log2 : -INFINITY;
                }
                bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
                bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
            }
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
            const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);          // finite: key 0 exists (Lk >= 1)
            const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
            m0 = mn0; m1 = mn1;
            float rs0 = 0.0f, rs1 = 0.0f;
            unsigned pa[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = exp2f(sc[nt][0] - mn0), p1 = exp2f(sc[nt][1] - mn0);
                const float p2 = exp2f(sc[nt][2] - mn1), p3 = exp2f(sc[nt][3] - mn1);
                rs0 += p0 + p1;
                rs1 += p2 + p3;
                pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
                pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
            }
            l0 = l0 * c0 + rs0;
            l1 = l1 * c1 + rs1;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (ks * 16 < kw) {
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
                        unsigned vb[4];
                        ldmatrix_x4_trans(vb, Vs + (k0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + np * 16 + (lane >> 4) * 8);
                        mma_bf16_16816(o[np * 2 + 0], pa[ks], vb[0], vb[1]);
                        mma_bf16_16816(o[np * 2 + 1], pa[ks], vb[2], vb[3]);
                    }
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        __nv_bfloat16* o0 = out + (b * Lq + qb * 16 + g) * d + hh * kHD;
        __nv_bfloat16* o1 = o0 + 8 * d;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<unsigned*>(o0 + nt * 8 + tq * 2) = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
            *reinterpret_cast<unsigned*>(o1 + nt * 8 + tq * 2) = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
        }
    }
}

// Start / goal maps of the selector (keypoint_selector.py:113-146): out [B, 2, H, W].  inv_2sigma2 > 0: Gaussian bumps;
// inv_2sigma2 <= 0 (the reference's sg_map_sigma <= 0 branch, :129-139): one-hot at the rounded cell (torch.round = half to even).
__global__ void __launch_bounds__(256) sg_map_kernel(const float* __restrict__ sg, long long B, int Hh, int Ww, float inv_2sigma2,
                                                     float* __restrict__ out) {
    const long long total = B * 2 * Hh * Ww;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = static_cast<int>(i % Ww);
        const int y = static_cast<int>((i / Ww) % Hh);
        const int which = static_cast<int>((i / (static_cast<long long>(Ww) * Hh)) % 2);
        const long long b = i / (2ll * Hh * Ww);
        const float cx = fminf(fmaxf(sg[b * 4 + which * 2 + 0], 0.0f), 1.0f) * static_cast<float>(Ww - 1);
        const float cy = fminf(fmaxf(sg[b * 4 + which * 2 + 1], 0.0f), 1.0f) * static_cast<float>(Hh - 1);
        if (inv_2sigma2 <= 0.0f) {
            const int xi = min(max(static_cast<int>(rintf(cx)), 0), Ww - 1), yi = min(max(static_cast<int>(rintf(cy)), 0), Hh - 1);
            out[i] = (x == xi && y == yi) ? 1.0f : 0.0f;
            continue;
        }
        const float dx = static_cast<float>(x) - cx, dy = static_cast<float>(y) - cy;
        out[i] = expf(-(dx * dx + dy * dy) * inv_2sigma2);
    }
}

}  // namespace idb200

using namespace idb200;

extern "C" int idb200_attention(const void* qkv, void* out, int is_bf16, int64_t B, int L, int H, int causal, int force_simt,
                                idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && L >= 1 && H >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(L <= 256, IDB200_EUNSUPPORTED, "attention supports L <= 256 (got %d)", L);
    if (B == 0) return IDB200_OK;
    IDB_REQUIRE(qkv && out, IDB200_EINVAL, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // force_simt: 0 = auto (bf16: tcgen05 path, any L), 1 = fp32-arithmetic SIMT kernel, 2 = legacy mma.sync path (A/B + cross-check)
    static const bool tc5_env = !(getenv("IDB200_ATTN_TC5") && atoi(getenv("IDB200_ATTN_TC5")) == 0);
    if (is_bf16 && force_simt == 0 && tc5_env && attention_tc5_supported(qkv, out, B, L, H))
        return attention_tc5(qkv, out, B, L, H, causal, st);
    if (force_simt == 2) force_simt = 0;
    // short sequences (Stage 1, L = 8): pack Lp / L trajectories into one block-diagonal "sequence" for the mma path
    int Lp = L, blk = 0;
    long long Bp = B;
    if (is_bf16 && !force_simt && L < 16 && (L == 8 || L == 4 || L == 2)) {
        for (int cand : {64, 32, 16}) {
            if (cand % L == 0 && B % (cand / L) == 0) { Lp = cand; blk = L; Bp = B / (cand / L); break; }
        }
    }
    const bool use_mma = is_bf16 && !force_simt && (Lp % 16 == 0) && aligned(qkv, 16) && aligned(out, 4);
    if (use_mma) {
        int HG = H;
        while (HG > 1 && (3 * static_cast<size_t>(HG) * Lp * kPitch * 2 > 100 * 1024 || H % HG != 0)) --HG;
        const size_t smem = 3 * static_cast<size_t>(HG) * Lp * kPitch * 2;
        static size_t smem_set[2] = {0, 0};
        const int variant = blk > 0 ? 1 : 0;
        if (smem > smem_set[variant]) {
            cudaError_t e = variant ? cudaFuncSetAttribute(attn_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                                    : cudaFuncSetAttribute(attn_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            smem_set[variant] = smem;
        }
        const int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
        const int grid = grid_for(Bp * (H / HG), 1, per_sm > 0 ? per_sm : 1);
        if (variant)
            attn_mma_kernel<true><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), Bp, Lp, H, HG, causal, blk);
        else
            attn_mma_kernel<false><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), Bp, Lp, H, HG, causal, blk);
        return check_launch("attn_mma_kernel");
    }
    const int G = (L < 32) ? 32 / L : 1;
    const int warps = (L <= 64) ? 4 : 1;
    const size_t smem = static_cast<size_t>(warps) * 2 * G * L * (kHD + 1) * sizeof(float);
    const long long groups = (B * H + G - 1) / G;
    const int grid = grid_for(groups, warps, 8);
    if (is_bf16) {
        static size_t set_b = 0;
        if (smem > set_b) { cudaFuncSetAttribute(attn_simt_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); set_b = smem; }
        attn_simt_kernel<__nv_bfloat16><<<grid, warps * 32, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), B, L, H, causal);
    } else {
        static size_t set_f = 0;
        if (smem > set_f) { cudaFuncSetAttribute(attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); set_f = smem; }
        attn_simt_kernel<float><<<grid, warps * 32, smem, st>>>(static_cast<const float*>(qkv), static_cast<float*>(out), B, L, H, causal);
    }
    return check_launch("attn_simt_kernel");
}

extern "C" int idb200_cross_attention(const void* q, const void* kv_a, const void* kv_b, void* out, int64_t B, int Lq, int La, int Lb,
                                      int H, idb200_stream_t stream) {
    IDB_REQUIRE(q && kv_a && out && (Lb == 0 || kv_b), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && H >= 1 && La >= 1 && Lb >= 0, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(Lq >= 16 && Lq % 16 == 0, IDB200_EUNSUPPORTED, "the number of queries must be a multiple of 16 (got %d)", Lq);
    IDB_REQUIRE(aligned(q, 16) && aligned(kv_a, 16) && (Lb == 0 || aligned(kv_b, 16)) && aligned(out, 4), IDB200_EALIGN,
                "q / kv must be 16-byte aligned");
    if (B == 0) return IDB200_OK;
    const int Lkp = (La + Lb + 15) & ~15;
    const size_t smem = (static_cast<size_t>(Lq) + 2 * static_cast<size_t>(Lkp)) * kPitch * 2;
    IDB_REQUIRE(smem <= 200 * 1024, IDB200_EUNSUPPORTED, "too many memory tokens for one CTA (%d)", La + Lb);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(cross_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        smem_set = smem;
    }
    IDB_REQUIRE(B * H < (1ll << 31), IDB200_EUNSUPPORTED, "batch too large");
    cross_attn_mma_kernel<<<static_cast<unsigned>(B * H), 128, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv_a), static_cast<const __nv_bfloat16*>(kv_b),
        static_cast<__nv_bfloat16*>(out), Lq, La, Lb, H);
    return check_launch("cross_attn_mma_kernel");
}

extern "C" int idb200_sg_map(const float* start_goal, int64_t B, int H, int W, float sigma, float* out, idb200_stream_t stream) {
    IDB_REQUIRE(start_goal && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && H >= 1 && W >= 1, IDB200_EINVAL, "bad arguments");
    if (B == 0) return IDB200_OK;
    sg_map_kernel<<<grid_for(B * 2 * H * W, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        start_goal, B, H, W, sigma > 0.0f ? 1.0f / (2.0f * sigma * sigma) : 0.0f, out);
    return check_launch("sg_map_kernel");
}
