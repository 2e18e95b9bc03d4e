// K1: nested anchor masks + piecewise-linear Interp(x0 | M_s), and its relatives.
//
// Reference arithmetic being replaced (paths under the reference root):
//   src/corruptions/keyframes.py:172-209  build_nested_masks_batch  (rand -> argsort -> per level cat/sort/scatter_)
//   src/corruptions/keyframes.py:348-380  interpolate_from_indices  (searchsorted, 5 gathers, lerp, scatter_, velocity)
//   src/train/train_interp_levels.py:444-510  _distance_alpha / _corrupt_from_anchors
//
// Design (HBM-bound, SURVEY.md 8d: 4600 algorithmic bytes per T=64,D=4,S=3 trajectory):
//   one warp per trajectory, grid-stride over B.  The [T,D] row is read once with 8/16-byte
//   vector loads (lane <-> timestep), the T-2 scores are staged in shared memory and ranked by an
//   all-pairs count, level masks are warp ballots, anchor indices are popc prefix sums, the
//   segment endpoints are fetched from a shared-memory copy of the row, and every level is
//   written with one coalesced vector store per lane.  No temporaries ever reach HBM.
//   All fp32 arithmetic uses the explicitly rounded intrinsics in the reference's operation
//   order, so results are bit-identical to the eager PyTorch ops (no FMA contraction).
#include "common.cuh"

namespace idb200 {

constexpr int kMaxLevels = 8;
constexpr unsigned kFull = 0xffffffffu;

struct NestedParams {
    const float* x0;
    const float* scores;
    long long score_stride;
    unsigned char* masks;
    long long* idx_out;
    float* x_levels;
    long long level_stride;
    long long B;
    int T, n, n_levels, s_lo, s_hi, flags;
    float dt;
    int thr[kMaxLevels];          // interior positions taken at level s
    int width[kMaxLevels];        // idx row width W_s
    long long idx_off[kMaxLevels];// sum of W_j, j < s
};

template <int D> struct VecOf;
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };

__device__ __forceinline__ float lerp_rn(float vl, float vr, float w) {
    // left + w * (right - left), each op rounded (keyframes.py:371)
    return __fadd_rn(vl, __fmul_rn(w, __fsub_rn(vr, vl)));
}

// Packed fp32x2 add / sub (FADD2 on sm_100a): two independent round-to-nearest results per
// instruction.  The multiply stays scalar: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even
// with explicit .rn, which would break bit parity with the reference's separately rounded ops.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float2 lerp_rn(float2 a, float2 b, float w) {
    float dx, dy;
    const unsigned long long l = pack2(a.x, a.y);
    unpack2(sub2_rn(pack2(b.x, b.y), l), dx, dy);
    float2 o;
    unpack2(add2_rn(l, pack2(__fmul_rn(w, dx), __fmul_rn(w, dy))), o.x, o.y);
    return o;
}
__device__ __forceinline__ float4 lerp_rn(float4 a, float4 b, float w) {
    const float2 lo = lerp_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), w);
    const float2 hi = lerp_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), w);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// count += [a < m] for two candidates at once: sign bits of the packed difference (a - m)
__device__ __forceinline__ void count_lt2(int& cnt, float a0, float a1, unsigned long long m2) {
    float d0, d1;
    unpack2(sub2_rn(pack2(a0, a1), m2), d0, d1);
    cnt += __float_as_uint(d0) >> 31;
    cnt += __float_as_uint(d1) >> 31;
}

template <int E> struct K1Cfg {
    static constexpr int kWarps = (E <= 4) ? 8 : 4;
};

template <int E, int D>
struct WarpScratch {
    alignas(16) float sc[32 * E + 4];                      // staged scores (padded with +inf)
    alignas(16) float cv[(D ? D : 1) * 32 * E];            // anchor values of the current level, compacted
    alignas(8) int2 seg_lr[32 * E];                        // (left, right) anchor position of segment k
    unsigned mw[kMaxLevels * E];                           // level mask words: bit (t & 31) of word (t >> 5)
};

// E = ceil(T / 32) mask words per level; D in {0 (masks only), 2, 4}.
template <int E, int D>
__global__ void __launch_bounds__(K1Cfg<E>::kWarps * 32, (E >= 8 ? 4 : 1)) nested_masks_interp_kernel(const NestedParams p) {
    constexpr int kWarps = K1Cfg<E>::kWarps;
    __shared__ WarpScratch<E, D> scratch[kWarps];
    __shared__ float rtab[256];            // rtab[g] = RN(1 / max(g, 1))
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch<E, D>& ws = scratch[warp];
    const long long warps_total = static_cast<long long>(gridDim.x) * kWarps;
    const int T = p.T, n = p.n;
    const bool noend = (p.flags & IDB200_F_NO_ENDPOINTS) != 0;
    const bool desc = (p.flags & IDB200_F_DESCENDING) != 0;
    const int n4 = (n + 3) & ~3;
    const float kInf = __int_as_float(0x7f800000);
    const unsigned lane_le = kFull >> (31 - lane);
    using V = typename VecOf<(D ? D : 2)>::type;

    if (D) {
        for (int i = threadIdx.x; i < 256; i += kWarps * 32) rtab[i] = __frcp_rn(static_cast<float>(max(i, 1)));
        __syncthreads();
    }

    for (long long b = static_cast<long long>(blockIdx.x) * kWarps + warp; b < p.B; b += warps_total) {
        // ---- load scores (lane <-> timestep) and the row ----------------------------------------
        const float* srow = p.scores + b * p.score_stride;
        float me[E];
        int jj[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int t = lane + 32 * e;
            const int j = noend ? t : t - 1;
            const bool vj = (j >= 0) && (j < n);
            jj[e] = vj ? j : -1;
            float s = vj ? __ldg(srow + j) : kInf;
            if (desc && vj) s = -s;
            me[e] = s + 0.0f;             // canonicalise -0 -> +0 so the sign-bit count is exact
        }
        V xv[E];
        if (D) {
            const V* xrow = reinterpret_cast<const V*>(p.x0) + b * T;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int t = lane + 32 * e;
                if (t < T) xv[e] = __ldg(xrow + t);
            }
        }
        __syncwarp();                      // previous iteration's readers are done with ws
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (jj[e] >= 0) ws.sc[jj[e]] = me[e];
        if (lane < n4 - n) ws.sc[n + lane] = kInf;
        __syncwarp();

        // ---- which interior positions does each level take? -------------------------------------
        // Level s takes the thr_s lowest-ranked scores.  E <= 2 (T <= 64): stable rank by an all-pairs count -- every lane streams
        // the staged scores with broadcast 16-byte loads and counts sign bits of packed differences; a tie makes both partners
        // miss a count, so sum(rank) < n(n-1)/2 detects ties exactly and those rows take the index-tie-break loop.
        // E >= 4 (T = 128 / 256): the all-pairs count is O(n^2) = 3.1 k instructions per lane at n = 254 and bounded the kernel
        // (0.34 of the HBM roofline).  Only the level CUTS matter, so the warp sorts the scores instead -- a bitonic network over
        // the 32 E register-resident values, ~0.8 k instructions -- and a position is in level s iff its score is below the
        // thr_s-th sorted value.  A tie straddling a cut (sorted[thr-1] == sorted[thr]) routes the row to the exact loop.
        int cnt[E];
        float cut[kMaxLevels];
        bool by_cut = false;
        if constexpr (E >= 4) {
            float v[E];
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = me[e];                 // any initial arrangement; virtual index i = lane * E + e
#pragma unroll
            for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (j < E) {                                       // partner in the same lane
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            if ((e & j) == 0) {
                                const bool up = (((lane * E + e) & k) == 0);
                                const float lo = fminf(v[e], v[e | j]), hi = fmaxf(v[e], v[e | j]);
                                v[e] = up ? lo : hi;
                                v[e | j] = up ? hi : lo;
                            }
                        }
                    } else {                                           // partner in lane ^ (j / E)
                        const int lj = j / E;
                        const bool upper = (lane & lj) != 0;
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const float o = __shfl_xor_sync(kFull, v[e], lj);
                            const bool up = (((lane * E + e) & k) == 0);
                            v[e] = (up != upper) ? fminf(v[e], o) : fmaxf(v[e], o);
                        }
                    }
                }
            }
            float* sorted = reinterpret_cast<float*>(ws.cv);           // idle until the interpolation stage
#pragma unroll
            for (int e = 0; e < E; ++e) sorted[lane * E + e] = v[e];
            __syncwarp();
            by_cut = true;
            for (int s = 0; s < p.n_levels; ++s) {
                const int thr = p.thr[s];
                if (thr <= 0) cut[s] = -kInf;
                else if (thr >= n) cut[s] = kInf;
                else {
                    cut[s] = sorted[thr];
                    if (sorted[thr - 1] == sorted[thr]) by_cut = false;   // tie across the cut: exact path (warp-uniform)
                }
            }
            __syncwarp();
            if (!by_cut) {
#pragma unroll
                for (int e = 0; e < E; ++e) cnt[e] = 0;
                for (int u = 0; u < n; ++u) {
                    const float su = ws.sc[u];
#pragma unroll
                    for (int e = 0; e < E; ++e) cnt[e] += ((su < me[e]) || (su == me[e] && u < jj[e])) ? 1 : 0;
                }
            }
        } else {
            unsigned long long m2[E];
#pragma unroll
            for (int e = 0; e < E; ++e) { cnt[e] = 0; m2[e] = pack2(me[e], me[e]); }
            for (int u = 0; u < n4; u += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&ws.sc[u]);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    count_lt2(cnt[e], v.x, v.y, m2[e]);
                    count_lt2(cnt[e], v.z, v.w, m2[e]);
                }
            }
            int rsum = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) rsum += (jj[e] >= 0) ? cnt[e] : 0;
            rsum = __reduce_add_sync(kFull, rsum);
            if (rsum != (n * (n - 1)) / 2) {   // warp-uniform: exact stable order (lower index first)
#pragma unroll
                for (int e = 0; e < E; ++e) cnt[e] = 0;
                for (int u = 0; u < n; ++u) {
                    const float su = ws.sc[u];
#pragma unroll
                    for (int e = 0; e < E; ++e) cnt[e] += ((su < me[e]) || (su == me[e] && u < jj[e])) ? 1 : 0;
                }
            }
        }

        // ---- level masks (ballots) --------------------------------------------------------------
        for (int s = 0; s < p.n_levels; ++s) {
            const int thr = p.thr[s];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int t = lane + 32 * e;
                const bool in = by_cut ? (me[e] < cut[s]) : (cnt[e] < thr);
                const bool a = (t < T) && ((!noend && (t == 0 || t == T - 1)) || (jj[e] >= 0 && in));
                const unsigned w = __ballot_sync(kFull, a);
                if (lane == 0) ws.mw[s * E + e] = w;
            }
        }
        __syncwarp();

        // ---- bool masks [B, n_levels, T] --------------------------------------------------------
        if (p.masks) {
            unsigned char* mrow = p.masks + b * static_cast<long long>(p.n_levels) * T;
            if ((T & 7) == 0) {
                const int per = T >> 3, chunks = p.n_levels * per;
                for (int c = lane; c < chunks; c += 32) {
                    const int s = c / per, t0 = (c - s * per) << 3;
                    const unsigned bits = (ws.mw[s * E + (t0 >> 5)] >> (t0 & 31)) & 0xffu;
                    // spread 4 bits to 4 bytes: copies at shifts 0,7,14,21 do not overlap
                    const unsigned lo = ((bits & 0xfu) * 0x204081u) & 0x01010101u;
                    const unsigned hi = ((bits >> 4) * 0x204081u) & 0x01010101u;
                    *reinterpret_cast<uint2*>(mrow + (static_cast<long long>(c) << 3)) = make_uint2(lo, hi);
                }
            } else {
                const int total = p.n_levels * T;
                for (int i = lane; i < total; i += 32) {
                    const int s = i / T, t = i - s * T;
                    mrow[i] = (ws.mw[s * E + (t >> 5)] >> (t & 31)) & 1u;
                }
            }
        }

        // ---- ascending anchor indices per level (popc prefix) -----------------------------------
        if (p.idx_out) {
            for (int s = 0; s < p.n_levels; ++s) {
                const int W = p.width[s];
                long long* irow = p.idx_out + p.idx_off[s] * p.B + b * W;
                int base = 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned w = ws.mw[s * E + e];
                    if ((w >> lane) & 1u) {
                        const int k = base + __popc(w & lane_le) - 1;
                        if (k < W) irow[k] = lane + 32 * e;
                    }
                    base += __popc(w);
                }
            }
        }

        // ---- Interp(x0 | M_s) for the requested levels ------------------------------------------
        // Anchor k publishes its value and position into compacted shared lists (segment k spans
        // [pos_k, pos_{k+1}]); timestep t finds its segment as popc(mask bits <= t) - 1 and fetches
        // both endpoints with independent loads.  w = (t-l)/(r-l) is the correctly rounded quotient:
        // q = off*RN(1/gap), one FMA residual and one FMA correction (exhaustively equal to IEEE
        // division for 0 <= off, gap <= 255).
        if (D) {
            for (int s = p.s_lo; s <= p.s_hi; ++s) {
                int seg[E];
                bool anchor[E];
                int base = 0;
                __syncwarp();              // readers of the previous level's lists are done
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned w = ws.mw[s * E + e];
                    const int t = lane + 32 * e;
                    seg[e] = base + __popc(w & lane_le) - 1;
                    anchor[e] = ((w >> lane) & 1u) != 0;
                    if (anchor[e]) {
                        reinterpret_cast<V*>(ws.cv)[seg[e]] = xv[e];
                        ws.seg_lr[seg[e]].x = t;
                        if (seg[e] > 0) ws.seg_lr[seg[e] - 1].y = t;
                    }
                    base += __popc(w);
                }
                if (lane == 0) ws.seg_lr[base - 1].y = T - 1;      // last anchor closes on itself
                __syncwarp();
                const int klast = base - 1;
                V yv[E];
                const bool all_anchors = (base == T);               // K_s = T (e.g. levels 0 / 1 of T = 256, K = 32, S = 4): Interp is the identity
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int t = lane + 32 * e;
                    if (t >= T) continue;
                    if (all_anchors) { yv[e] = xv[e]; continue; }
                    const int k1 = min(seg[e] + 1, klast);
                    const int2 lr = ws.seg_lr[seg[e]];
                    const V vl = reinterpret_cast<const V*>(ws.cv)[seg[e]];
                    const V vr = reinterpret_cast<const V*>(ws.cv)[k1];
                    const int gap = lr.y - lr.x;
                    const float off = static_cast<float>(t - lr.x);
                    const float rg = rtab[gap];
                    const float q = __fmul_rn(off, rg);
                    const float rem = __fmaf_rn(-static_cast<float>(max(gap, 1)), q, off);
                    const float wt = __fmaf_rn(rem, rg, q);
                    const V y = lerp_rn(vl, vr, wt);
                    yv[e] = anchor[e] ? xv[e] : y;      // anchors: exact copy (scatter_ at keyframes.py:372)
                }
                V* orow = reinterpret_cast<V*>(p.x_levels + static_cast<long long>(s - p.s_lo) * p.level_stride) + b * T;
                if (D == 4 && (p.flags & IDB200_F_RECOMPUTE_VELOCITY)) {
                    // v[t] = (pos[t+1] - pos[t]) / dt, v[T-1] = 0   (keyframes.py:373-379)
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int t = lane + 32 * e;
                        float nx = __shfl_down_sync(kFull, yv[e].x, 1);
                        float ny = __shfl_down_sync(kFull, yv[e].y, 1);
                        if (e + 1 < E) {
                            const float fx = __shfl_sync(kFull, yv[(e + 1 < E) ? e + 1 : e].x, 0);
                            const float fy = __shfl_sync(kFull, yv[(e + 1 < E) ? e + 1 : e].y, 0);
                            if (lane == 31) { nx = fx; ny = fy; }
                        }
                        if (t < T) {
                            float4 o;
                            o.x = yv[e].x;
                            o.y = yv[e].y;
                            o.z = (t == T - 1) ? 0.0f : __fdiv_rn(__fsub_rn(nx, yv[e].x), p.dt);
                            o.w = (t == T - 1) ? 0.0f : __fdiv_rn(__fsub_rn(ny, yv[e].y), p.dt);
                            *reinterpret_cast<float4*>(&orow[t]) = o;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int t = lane + 32 * e;
                        if (t < T) orow[t] = yv[e];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path of K1 for the headline shape: T == 64 (two timesteps per lane), NL <= 4 levels,
// endpoint mode.  Same algorithm as the generic kernel above with every shape decision made at
// compile time: level masks live in registers, the level loop is unrolled, no runtime divisions.
// ------------------------------------------------------------------------------------------------
template <int D>
struct FastScratch {
    alignas(16) float xin[2][(D ? D : 1) * 64];   // cp.async landing buffers for the row (double-buffered)
    alignas(16) float sin_[2][64];                // ... and for the 62 scores
    alignas(16) float sc[64 + 4];
    alignas(16) float cv[(D ? D : 1) * 64];
    alignas(8) int2 seg_lr[64];
    alignas(16) unsigned mw[8];
};

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int D, int NL>
__global__ void __launch_bounds__(256) nested_masks_interp_t64_kernel(const NestedParams p) {
    constexpr int kWarps = 8;
    constexpr int T = 64, n = 62;
    __shared__ FastScratch<D> scratch[kWarps];
    __shared__ float rtab[64];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    FastScratch<D>& ws = scratch[warp];
    const long long warps_total = static_cast<long long>(gridDim.x) * kWarps;
    const bool desc = (p.flags & IDB200_F_DESCENDING) != 0;
    const float kInf = __int_as_float(0x7f800000);
    const unsigned lane_le = kFull >> (31 - lane);
    const bool s8 = ((p.score_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.scores) & 7) == 0);  // 8-byte score rows
    using V = typename VecOf<(D ? D : 2)>::type;

    if (D) {
        if (threadIdx.x < 64) rtab[threadIdx.x] = __frcp_rn(static_cast<float>(max(static_cast<int>(threadIdx.x), 1)));
        __syncthreads();
    }
    if (lane < 2) ws.sc[62 + lane] = kInf;          // padding never changes

    // Software pipeline: the next trajectory's row and scores stream into shared memory with cp.async
    // (no registers held) while the current one is ranked and interpolated.
    auto prefetch = [&](long long b, int buf) {
        const float* srow = p.scores + b * p.score_stride;
        if (s8) {
            if (lane < 31) cp_async_8(&ws.sin_[buf][2 * lane], srow + 2 * lane);
        } else {
            cp_async_4(&ws.sin_[buf][lane], srow + lane);
            if (lane < 30) cp_async_4(&ws.sin_[buf][lane + 32], srow + lane + 32);
        }
        if (D) {
            const V* xrow = reinterpret_cast<const V*>(p.x0) + b * T;
            if (D == 4) {
                cp_async_16(&reinterpret_cast<V*>(ws.xin[buf])[lane], xrow + lane);
                cp_async_16(&reinterpret_cast<V*>(ws.xin[buf])[lane + 32], xrow + lane + 32);
            } else {
                cp_async_16(&reinterpret_cast<float4*>(ws.xin[buf])[lane], reinterpret_cast<const float4*>(xrow) + lane);
            }
        }
        cp_async_commit();
    };

    long long b = static_cast<long long>(blockIdx.x) * kWarps + warp;
    int buf = 0;
    if (b < p.B) prefetch(b, 0);
    for (; b < p.B; b += warps_total, buf ^= 1) {
        cp_async_wait_all();
        __syncwarp();
        // lane <-> timesteps t0 = lane, t1 = lane + 32; interior score index j = t - 1
        float s0 = (lane >= 1) ? ws.sin_[buf][lane - 1] : kInf;
        float s1 = (lane < 31) ? ws.sin_[buf][lane + 31] : kInf;
        V x0v, x1v;
        if (D) {
            x0v = reinterpret_cast<const V*>(ws.xin[buf])[lane];
            x1v = reinterpret_cast<const V*>(ws.xin[buf])[lane + 32];
        }
        if (b + warps_total < p.B) prefetch(b + warps_total, buf ^ 1);   // buffer buf^1 was consumed last iteration
        if (desc) { s0 = -s0; s1 = -s1; }           // (+inf sentinels become -inf but are never staged)
        const float me0 = s0 + 0.0f, me1 = s1 + 0.0f;
        __syncwarp();
        if (lane >= 1) ws.sc[lane - 1] = me0;
        if (lane < 31) ws.sc[lane + 31] = me1;
        __syncwarp();

        int c0 = 0, c1 = 0;
        {
            const unsigned long long m0 = pack2(me0, me0), m1 = pack2(me1, me1);
#pragma unroll
            for (int u = 0; u < 64; u += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&ws.sc[u]);
                count_lt2(c0, v.x, v.y, m0);
                count_lt2(c0, v.z, v.w, m0);
                count_lt2(c1, v.x, v.y, m1);
                count_lt2(c1, v.z, v.w, m1);
            }
        }
        const bool v0 = lane >= 1, v1 = lane < 31;  // element holds an interior score
        int rsum = (v0 ? c0 : 0) + (v1 ? c1 : 0);
        rsum = __reduce_add_sync(kFull, rsum);
        if (rsum != (n * (n - 1)) / 2) {            // ties: exact stable order (lower index first)
            c0 = 0; c1 = 0;
            for (int u = 0; u < n; ++u) {
                const float su = ws.sc[u];
                c0 += ((su < me0) || (su == me0 && u < lane - 1)) ? 1 : 0;
                c1 += ((su < me1) || (su == me1 && u < lane + 31)) ? 1 : 0;
            }
        }

        unsigned w0[NL], w1[NL];
#pragma unroll
        for (int s = 0; s < NL; ++s) {
            const int thr = p.thr[s];
            w0[s] = __ballot_sync(kFull, lane == 0 || c0 < thr);          // t = 0 is always an anchor
            w1[s] = __ballot_sync(kFull, lane == 31 || c1 < thr);         // t = 63 is always an anchor
        }

        if (p.masks) {
            // [NL, 64] bytes per trajectory = NL * 8 chunks of 8 bytes, one per lane
            if (lane == 0) {
#pragma unroll
                for (int s = 0; s < NL; ++s) { ws.mw[2 * s] = w0[s]; ws.mw[2 * s + 1] = w1[s]; }
            }
            __syncwarp();
            if (lane < NL * 8) {
                const unsigned word = ws.mw[lane >> 2];
                const unsigned bits = (word >> ((lane & 3) << 3)) & 0xffu;
                const unsigned lo = ((bits & 0xfu) * 0x204081u) & 0x01010101u;
                const unsigned hi = ((bits >> 4) * 0x204081u) & 0x01010101u;
                reinterpret_cast<uint2*>(p.masks + b * (NL * 64))[lane] = make_uint2(lo, hi);
            }
        }

        if (p.idx_out) {
#pragma unroll
            for (int s = 0; s < NL; ++s) {
                const int W = p.width[s];
                long long* irow = p.idx_out + p.idx_off[s] * p.B + b * W;
                const int k0 = __popc(w0[s] & lane_le) - 1;
                const int k1 = __popc(w0[s]) + __popc(w1[s] & lane_le) - 1;
                if (((w0[s] >> lane) & 1u) && k0 < W) irow[k0] = lane;
                if (((w1[s] >> lane) & 1u) && k1 < W) irow[k1] = lane + 32;
            }
        }

        if (D) {
#pragma unroll
            for (int s = 0; s < NL; ++s) {
                if (s < p.s_lo || s > p.s_hi) continue;               // warp-uniform
                const int n0 = __popc(w0[s]);
                const int seg0 = __popc(w0[s] & lane_le) - 1;
                const int seg1 = n0 + __popc(w1[s] & lane_le) - 1;
                const bool a0 = (w0[s] >> lane) & 1u, a1 = (w1[s] >> lane) & 1u;
                const int klast = n0 + __popc(w1[s]) - 1;
                __syncwarp();                                          // previous level's readers are done
                if (a0) {
                    reinterpret_cast<V*>(ws.cv)[seg0] = x0v;
                    ws.seg_lr[seg0].x = lane;
                    if (seg0 > 0) ws.seg_lr[seg0 - 1].y = lane;
                }
                if (a1) {
                    reinterpret_cast<V*>(ws.cv)[seg1] = x1v;
                    ws.seg_lr[seg1].x = lane + 32;
                    ws.seg_lr[seg1 - 1].y = lane + 32;                 // seg1 >= 1: t = 0 is an anchor
                    if (lane == 31) ws.seg_lr[seg1].y = 63;            // last anchor closes on itself
                }
                __syncwarp();
                V* orow = reinterpret_cast<V*>(p.x_levels + static_cast<long long>(s - p.s_lo) * p.level_stride) + b * T;
                V y0 = x0v, y1 = x1v;
                if (!a0) {
                    const int2 lr = ws.seg_lr[seg0];
                    const V vl = reinterpret_cast<const V*>(ws.cv)[seg0];
                    const V vr = reinterpret_cast<const V*>(ws.cv)[seg0 + 1];
                    const int gap = lr.y - lr.x;                       // >= 2 for a non-anchor
                    const float off = static_cast<float>(lane - lr.x);
                    const float rg = rtab[gap];
                    const float q = __fmul_rn(off, rg);
                    const float wt = __fmaf_rn(__fmaf_rn(-static_cast<float>(gap), q, off), rg, q);
                    y0 = lerp_rn(vl, vr, wt);
                }
                if (!a1) {
                    const int2 lr = ws.seg_lr[seg1];
                    const V vl = reinterpret_cast<const V*>(ws.cv)[seg1];
                    const V vr = reinterpret_cast<const V*>(ws.cv)[min(seg1 + 1, klast)];
                    const int gap = lr.y - lr.x;
                    const float off = static_cast<float>(lane + 32 - lr.x);
                    const float rg = rtab[gap];
                    const float q = __fmul_rn(off, rg);
                    const float wt = __fmaf_rn(__fmaf_rn(-static_cast<float>(gap), q, off), rg, q);
                    y1 = lerp_rn(vl, vr, wt);
                }
                if (D == 4 && (p.flags & IDB200_F_RECOMPUTE_VELOCITY)) {
                    // v[t] = (pos[t+1] - pos[t]) / dt, v[T-1] = 0   (keyframes.py:373-379)
                    float nx0 = __shfl_down_sync(kFull, y0.x, 1), ny0 = __shfl_down_sync(kFull, y0.y, 1);
                    const float nx1 = __shfl_down_sync(kFull, y1.x, 1), ny1 = __shfl_down_sync(kFull, y1.y, 1);
                    const float fx = __shfl_sync(kFull, y1.x, 0), fy = __shfl_sync(kFull, y1.y, 0);
                    if (lane == 31) { nx0 = fx; ny0 = fy; }
                    float4 o0, o1;
                    o0.x = y0.x; o0.y = y0.y;
                    o0.z = __fdiv_rn(__fsub_rn(nx0, y0.x), p.dt);
                    o0.w = __fdiv_rn(__fsub_rn(ny0, y0.y), p.dt);
                    o1.x = y1.x; o1.y = y1.y;
                    o1.z = (lane == 31) ? 0.0f : __fdiv_rn(__fsub_rn(nx1, y1.x), p.dt);
                    o1.w = (lane == 31) ? 0.0f : __fdiv_rn(__fsub_rn(ny1, y1.y), p.dt);
                    *reinterpret_cast<float4*>(&orow[lane]) = o0;
                    *reinterpret_cast<float4*>(&orow[lane + 32]) = o1;
                } else {
                    orow[lane] = y0;
                    orow[lane + 32] = y1;
                }
            }
        }
    }
}

template <int D, int NL>
static int launch_nested_t64(const NestedParams& p, cudaStream_t st) {
    const int grid = grid_for(p.B, 8, 8);
    nested_masks_interp_t64_kernel<D, NL><<<grid, 256, 0, st>>>(p);
    return check_launch("nested_masks_interp_t64_kernel");
}

template <int D>
static int dispatch_nested_t64(const NestedParams& p, cudaStream_t st) {
    switch (p.n_levels) {
        case 1: return launch_nested_t64<D, 1>(p, st);
        case 2: return launch_nested_t64<D, 2>(p, st);
        case 3: return launch_nested_t64<D, 3>(p, st);
        default: return launch_nested_t64<D, 4>(p, st);
    }
}

template <int E, int D>
static int launch_nested(const NestedParams& p, cudaStream_t st) {
    constexpr int kWarps = K1Cfg<E>::kWarps;
    const int grid = grid_for(p.B, kWarps, 8);
    nested_masks_interp_kernel<E, D><<<grid, kWarps * 32, 0, st>>>(p);
    return check_launch("nested_masks_interp_kernel");
}

template <int D>
static int dispatch_nested_E(const NestedParams& p, cudaStream_t st) {
    if (p.T <= 32) return launch_nested<1, D>(p, st);
    if (p.T <= 64) return launch_nested<2, D>(p, st);
    if (p.T <= 128) return launch_nested<4, D>(p, st);
    return launch_nested<8, D>(p, st);
}

// ------------------------------------------------------------------------------------------------
// general interpolate_from_indices: one CTA per trajectory, idx staged in shared memory,
// (t, d) flattened over the threads; searchsorted(right=True) is an upper-bound binary search.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound_smem(const int* a, int K, int t) {
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= t) lo = mid + 1; else hi = mid;
    }
    return lo;            // number of idx <= t
}

__device__ __forceinline__ float interp_eval(const int* sidx, const float* v, int K, int D, int t, int d) {
    const int cnt = upper_bound_smem(sidx, K, t);
    if (cnt > 0 && sidx[cnt - 1] == t) return v[static_cast<long long>(cnt - 1) * D + d];  // scatter_: last duplicate wins
    const int seg = min(max(cnt - 1, 0), K - 2);
    const int l = sidx[seg], r = sidx[seg + 1];
    const float w = __fdiv_rn(static_cast<float>(t - l), static_cast<float>(max(r - l, 1)));
    return lerp_rn(v[static_cast<long long>(seg) * D + d], v[static_cast<long long>(seg + 1) * D + d], w);
}

__global__ void __launch_bounds__(128) interp_indices_kernel(const long long* __restrict__ idx, const float* __restrict__ vals,
                                                             long long B, int K, int T, int D, int vel, float dt,
                                                             float* __restrict__ y) {
    extern __shared__ int sidx[];
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) sidx[k] = static_cast<int>(idx[b * K + k]);
        __syncthreads();
        const float* v = vals + b * K * D;
        float* yo = y + b * T * D;
        const int total = T * D;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int t = i / D, d = i - t * D;
            float out;
            if (vel && d >= 2) {
                if (t == T - 1) out = 0.0f;
                else out = __fdiv_rn(__fsub_rn(interp_eval(sidx, v, K, D, t + 1, d - 2), interp_eval(sidx, v, K, D, t, d - 2)), dt);
            } else {
                out = interp_eval(sidx, v, K, D, t, d);
            }
            yo[i] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1c: _corrupt_from_anchors (train_interp_levels.py:458-510), one CTA per row.
// shared: idx (int) [K], noisy anchor values [K*D], final positions [T*2] (for the velocity pass)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) corrupt_from_anchors_kernel(
    const float* __restrict__ source, const long long* __restrict__ idx, const long long* __restrict__ idx_gather,
    const float* __restrict__ anchor_noise, const float* __restrict__ path_noise, const long long* __restrict__ row_index,
    long long B, int K, int T, int D, float sigma, float anchor_sigma, int mode_dist, int clamp_endpoints, int vel,
    float dt, float* __restrict__ out) {
    extern __shared__ int smem_i[];
    int* sidx = smem_i;
    float* svals = reinterpret_cast<float*>(smem_i + K);
    float* spos = svals + K * D;
    for (long long r = blockIdx.x; r < B; r += gridDim.x) {
        const long long row = row_index ? row_index[r] : r;
        const float* src = source + row * T * D;
        float* orow = out + row * T * D;
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) sidx[k] = static_cast<int>(idx[r * K + k]);
        __syncthreads();
        for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
            const int k = i / D, d = i - k * D;
            const int g = idx_gather ? static_cast<int>(idx_gather[r * K + k]) : sidx[k];
            float v = src[static_cast<long long>(g) * D + d];
            if (anchor_noise && d < 2) {               // :486-494
                float nv = __fmul_rn(anchor_noise[(r * K + k) * 2 + d], anchor_sigma);
                if (clamp_endpoints && (sidx[k] == 0 || sidx[k] == T - 1)) nv = 0.0f;
                v = __fadd_rn(v, nv);
            }
            svals[i] = v;
        }
        __syncthreads();
        const int total = T * D;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int t = i / D, d = i - t * D;
            if (vel && d >= 2) continue;
            float x = interp_eval(sidx, svals, K, D, t, d);
            if (path_noise && d < 2) {                 // :496-504
                float alpha = 1.0f;
                if (mode_dist) {                       // _distance_alpha :444-455
                    const int cnt = upper_bound_smem(sidx, K, t);
                    const int seg = min(max(cnt - 1, 0), K - 2);
                    const int l = sidx[seg], rr = sidx[seg + 1];
                    const int gap = max(rr - l, 1);
                    const int dist = min(t - l, rr - t);
                    alpha = __fdiv_rn(__fmul_rn(2.0f, static_cast<float>(dist)), static_cast<float>(gap));
                    alpha = fminf(fmaxf(alpha, 0.0f), 1.0f);
                }
                const float nz = __fmul_rn(path_noise[(r * T + t) * 2 + d], sigma);
                x = __fadd_rn(x, __fmul_rn(nz, alpha));
            }
            if (vel && d < 2) spos[t * 2 + d] = x;
            orow[i] = x;
        }
        if (vel) {                                     // :505-509
            __syncthreads();
            for (int i = threadIdx.x; i < T * 2; i += blockDim.x) {
                const int t = i >> 1, d = i & 1;
                const float v = (t == T - 1) ? 0.0f : __fdiv_rn(__fsub_rn(spos[(t + 1) * 2 + d], spos[t * 2 + d]), dt);
                orow[t * D + 2 + d] = v;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// K1c, single launch: build_interp_adjacent_batch / build_interp_level_batch (train_interp_levels.py:227-383) for the
// WHOLE batch -- each row picks its own level s = s_idx[row], so K_s, sigma_s and the anchor set are per-row quantities and
// there is no per-level python loop, boolean-mask gather or host sync.  One CTA per row; the anchor list of level s (and
// s - 1 for x_prev) is compacted from the row's nested mask by a ballot scan (the masks ARE the index lists: idx_levels[s]
// is the ascending set of M_s), then the body is corrupt_from_anchors_kernel's: noisy anchors, interpolation, tent-weighted
// path noise, velocity.  Noise: either PROVIDED (parity mode: the caller drew it with the reference's generator calls;
// per-row layout anchor_noise [B, 2, Kmax, 2], path_noise [B, 2, T, 2], slot 0 = x_s, slot 1 = x_prev) or drawn in the
// kernel from Philox4x32-10 + Box-Muller keyed by (seed, offset | row, slot, kind, position) (speed mode: same
// distribution, not the reference's stream); the drawn noise can be exported in the provided-noise layout, which makes the
// two modes checkable against each other bit for bit.
// ------------------------------------------------------------------------------------------------
struct AdjParams {
    const float* source;
    const unsigned char* masks;       // [B, n_levels, T]
    const long long* s_idx;           // [B]
    const float* anchor_noise;        // [B, 2, Kmax, 2] or nullptr
    const float* path_noise;          // [B, 2, T, 2] or nullptr
    float* anchor_noise_out;          // export (Philox mode) or nullptr
    float* path_noise_out;
    unsigned long long seed, offset;
    float sigma[kMaxLevels + 1];
    float anchor_sigma[kMaxLevels + 1];
    long long B;
    int T, D, n_levels, Kmax, mode_dist, clamp_endpoints, vel, philox, adjacent;
    float dt;
    float* x_s;
    float* x_prev;
    unsigned char* mask_s;
    unsigned char* mask_prev;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// two independent N(0,1) draws for (row, slot, kind, pos): counter = (pos, kind | slot << 1 | offset_lo << 2, row_lo, row_hi ^ offset_hi)
__device__ __forceinline__ float2 philox_normal2(const AdjParams& p, long long row, int slot, int kind, int pos) {
    const uint4 ctr = make_uint4(static_cast<unsigned>(pos), static_cast<unsigned>(kind | (slot << 1)) | (static_cast<unsigned>(p.offset) << 2),
                                 static_cast<unsigned>(row), static_cast<unsigned>(static_cast<unsigned long long>(row) >> 32) ^ static_cast<unsigned>(p.offset >> 30));
    const uint4 r = philox4x32_10(ctr, make_uint2(static_cast<unsigned>(p.seed), static_cast<unsigned>(p.seed >> 32)));
    const float u1 = __fmul_rn(__fadd_rn(static_cast<float>(r.x >> 8), 0.5f), 5.9604644775390625e-8f);     // (0, 1): 24 bits + 1/2 ulp
    const float u2 = __fmul_rn(__fadd_rn(static_cast<float>(r.y >> 8), 0.5f), 5.9604644775390625e-8f);
    const float rad = sqrtf(__fmul_rn(-2.0f, logf(u1)));
    float sn, cs;
    sincospif(__fmul_rn(2.0f, u2), &sn, &cs);
    return make_float2(__fmul_rn(rad, cs), __fmul_rn(rad, sn));
}

__global__ void __launch_bounds__(128) corrupt_adjacent_kernel(const AdjParams p) {
    extern __shared__ int smem_i[];
    int* sidx = smem_i;                                          // [T]
    float* svals = reinterpret_cast<float*>(smem_i + p.T);       // [T * D]
    float* spos = svals + p.T * p.D;                             // [T * 2]
    __shared__ int sK;
    const int T = p.T, D = p.D;
    for (long long row = blockIdx.x; row < p.B; row += gridDim.x) {
        const long long s = p.s_idx[row];
        const float* src = p.source + row * T * D;
        for (int slot = 0; slot < (p.adjacent ? 2 : 1); ++slot) {
            float* orow = (slot ? p.x_prev : p.x_s) + row * T * D;
            unsigned char* mrow_out = slot ? p.mask_prev : p.mask_s;
            const long long lvl = s - slot;
            if (s < 1 || s >= p.n_levels) {                      // rows outside 1..levels stay zero (:328 loops s = 1..levels)
                for (int i = threadIdx.x; i < T * D; i += blockDim.x) orow[i] = 0.0f;
                if (mrow_out) for (int t = threadIdx.x; t < T; t += blockDim.x) mrow_out[row * T + t] = 0;
                continue;
            }
            const unsigned char* mrow = p.masks + (row * p.n_levels + lvl) * T;
            __syncthreads();                                     // previous slot / row is done with the shared lists
            if (threadIdx.x < 32) {                              // ordered compaction of the set bits (ascending t)
                int base = 0;
                for (int c = 0; c < T; c += 32) {
                    const int t = c + threadIdx.x;
                    const bool bit = t < T && mrow[t] != 0;
                    const unsigned b = __ballot_sync(0xffffffffu, bit);
                    if (bit) sidx[base + __popc(b & ((1u << threadIdx.x) - 1u))] = t;
                    base += __popc(b);
                }
                if (threadIdx.x == 0) sK = base;
            }
            __syncthreads();
            const int K = sK;
            if (mrow_out) for (int t = threadIdx.x; t < T; t += blockDim.x) mrow_out[row * T + t] = mrow[t];
            if (K < 2) {                                         // cannot happen for masks with endpoints; keep the row defined
                for (int i = threadIdx.x; i < T * D; i += blockDim.x) orow[i] = 0.0f;
                continue;
            }
            const float sigma = p.sigma[lvl], asig = p.anchor_sigma[lvl];
            const long long nslot = row * 2 + slot;
            for (int k = threadIdx.x; k < K; k += blockDim.x) {  // :484-494 (index jitter is not supported here: host falls back)
                const int g = sidx[k];
                float nz[2] = {0.0f, 0.0f};
                if (asig > 0.0f) {
                    if (p.philox) {
                        const float2 n2 = philox_normal2(p, row, slot, 0, k);
                        nz[0] = n2.x; nz[1] = n2.y;
                        if (p.anchor_noise_out) { p.anchor_noise_out[(nslot * p.Kmax + k) * 2] = n2.x; p.anchor_noise_out[(nslot * p.Kmax + k) * 2 + 1] = n2.y; }
                    } else {
                        nz[0] = p.anchor_noise[(nslot * p.Kmax + k) * 2];
                        nz[1] = p.anchor_noise[(nslot * p.Kmax + k) * 2 + 1];
                    }
                }
                const bool endp = p.clamp_endpoints && (g == 0 || g == T - 1);
                for (int d = 0; d < D; ++d) {
                    float v = src[static_cast<long long>(g) * D + d];
                    if (asig > 0.0f && d < 2) v = __fadd_rn(v, endp ? 0.0f : __fmul_rn(nz[d], asig));
                    svals[k * D + d] = v;
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < T; t += blockDim.x) {
                float nz[2] = {0.0f, 0.0f};
                float alpha = 1.0f;
                if (sigma > 0.0f) {                              // :496-504
                    if (p.philox) {
                        const float2 n2 = philox_normal2(p, row, slot, 1, t);
                        nz[0] = n2.x; nz[1] = n2.y;
                        if (p.path_noise_out) { p.path_noise_out[(nslot * T + t) * 2] = n2.x; p.path_noise_out[(nslot * T + t) * 2 + 1] = n2.y; }
                    } else {
                        nz[0] = p.path_noise[(nslot * T + t) * 2];
                        nz[1] = p.path_noise[(nslot * T + t) * 2 + 1];
                    }
                    if (p.mode_dist) {                           // _distance_alpha :444-455
                        const int cnt = upper_bound_smem(sidx, K, t);
                        const int seg = min(max(cnt - 1, 0), K - 2);
                        const int l = sidx[seg], rr = sidx[seg + 1];
                        const int gap = max(rr - l, 1);
                        const int dist = min(t - l, rr - t);
                        alpha = __fdiv_rn(__fmul_rn(2.0f, static_cast<float>(dist)), static_cast<float>(gap));
                        alpha = fminf(fmaxf(alpha, 0.0f), 1.0f);
                    }
                }
                for (int d = 0; d < D; ++d) {
                    if (p.vel && d >= 2) continue;
                    float x = interp_eval(sidx, svals, K, D, t, d);
                    if (sigma > 0.0f && d < 2) x = __fadd_rn(x, __fmul_rn(__fmul_rn(nz[d], sigma), alpha));
                    if (p.vel && d < 2) spos[t * 2 + d] = x;
                    orow[t * D + d] = x;
                }
            }
            if (p.vel) {                                         // :505-509
                __syncthreads();
                for (int i = threadIdx.x; i < T * 2; i += blockDim.x) {
                    const int t = i >> 1, d = i & 1;
                    const float v = (t == T - 1) ? 0.0f : __fdiv_rn(__fsub_rn(spos[(t + 1) * 2 + d], spos[t * 2 + d]), p.dt);
                    orow[t * D + 2 + d] = v;
                }
            }
        }
    }
}

}  // namespace idb200

using namespace idb200;

extern "C" int idb200_nested_masks_interp(const float* x0, const float* scores, int64_t score_stride, int64_t B, int T,
                                          int D, int n_levels, const int* K_list, uint8_t* masks, int64_t* idx_out,
                                          float* x_levels, int64_t level_stride, int s_lo, int s_hi, int flags,
                                          idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0, IDB200_EINVAL, "B must be >= 0");
    IDB_REQUIRE(T >= 1, IDB200_EINVAL, "T must be positive");
    IDB_REQUIRE(T <= 256, IDB200_EUNSUPPORTED, "nested_masks_interp supports T <= 256 (got %d)", T);
    IDB_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, IDB200_EUNSUPPORTED, "n_levels must be in [1,%d]", kMaxLevels);
    IDB_REQUIRE(K_list != nullptr, IDB200_EINVAL, "K_list is NULL");
    const bool noend = (flags & IDB200_F_NO_ENDPOINTS) != 0;
    IDB_REQUIRE(noend || T >= 2, IDB200_EINVAL, "T must be >= 2 when using endpoints");
    NestedParams p{};
    p.n = noend ? T : T - 2;
    IDB_REQUIRE(B == 0 || p.n == 0 || scores != nullptr, IDB200_EINVAL, "scores is NULL");
    IDB_REQUIRE(score_stride >= p.n, IDB200_EINVAL, "score_stride smaller than the row length");
    if (x0 != nullptr) {
        IDB_REQUIRE(!noend, IDB200_EINVAL, "interpolation needs endpoint anchors");
        IDB_REQUIRE(D == 2 || D == 4, IDB200_EUNSUPPORTED, "fused interpolation supports D in {2,4} (got %d)", D);
        IDB_REQUIRE(x_levels != nullptr, IDB200_EINVAL, "x_levels is NULL");
        IDB_REQUIRE(s_lo >= 0 && s_hi < n_levels && s_lo <= s_hi, IDB200_EINVAL, "bad level range [%d,%d]", s_lo, s_hi);
        const size_t va = static_cast<size_t>(D) * 4;      // one float2 / float4 per timestep
        IDB_REQUIRE(aligned(x0, va) && aligned(x_levels, va) && (level_stride * 4) % va == 0, IDB200_EALIGN,
                    "x0 / x_levels must be %zu-byte aligned", va);
    }
    IDB_REQUIRE(masks == nullptr || aligned(masks, 8), IDB200_EALIGN, "masks must be 8-byte aligned");
    long long off = 0;
    for (int s = 0; s < n_levels; ++s) {
        const int Ks = K_list[s];
        IDB_REQUIRE(Ks >= 1, IDB200_EINVAL, "K_list[%d] must be positive", s);
        if (noend) {
            p.thr[s] = Ks < T ? Ks : T;
            p.width[s] = p.thr[s];
        } else {
            int thr = Ks - 2;
            if (thr < 0) thr = 0;
            if (thr > p.n) thr = p.n;
            p.thr[s] = thr;
            p.width[s] = (Ks <= 2 || T <= 2) ? 2 : (Ks < T ? Ks : T);
        }
        p.idx_off[s] = off;
        off += p.width[s];
    }
    if (B == 0) return IDB200_OK;
    p.x0 = x0; p.scores = scores; p.score_stride = score_stride; p.masks = masks;
    p.idx_out = reinterpret_cast<long long*>(idx_out);
    p.x_levels = x_levels; p.level_stride = level_stride; p.B = B; p.T = T; p.n_levels = n_levels;
    p.s_lo = s_lo; p.s_hi = s_hi; p.flags = flags;
    p.dt = static_cast<float>(1.0 / static_cast<double>(T));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (T == 64 && n_levels <= 4 && !noend && (x0 == nullptr || aligned(x0, 16))) {   // compile-time specialised headline shape
        if (x0 == nullptr) return dispatch_nested_t64<0>(p, st);
        if (D == 2) return dispatch_nested_t64<2>(p, st);
        return dispatch_nested_t64<4>(p, st);
    }
    if (x0 == nullptr) return dispatch_nested_E<0>(p, st);
    if (D == 2) return dispatch_nested_E<2>(p, st);
    return dispatch_nested_E<4>(p, st);
}

extern "C" int idb200_interpolate_from_indices(const int64_t* idx, const float* vals, int64_t B, int K, int T, int D,
                                               int recompute_velocity, float* y, idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && D >= 1 && T >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(B == 0 || (idx && vals && y), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(K >= 2, IDB200_EINVAL, "K must be >= 2");
    IDB_REQUIRE(K <= 8192 && T <= 65536, IDB200_EUNSUPPORTED, "K <= 8192 and T <= 65536 supported");
    if (B == 0) return IDB200_OK;
    const int vel = (recompute_velocity && D == 4) ? 1 : 0;
    const int grid = grid_for(B, 1, 16);
    interp_indices_kernel<<<grid, 128, K * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(idx), vals, B, K, T, D, vel, static_cast<float>(1.0 / static_cast<double>(T)), y);
    return check_launch("interp_indices_kernel");
}

extern "C" int idb200_corrupt_from_anchors(const float* source, const int64_t* idx, const int64_t* idx_gather,
                                           const float* anchor_noise, const float* path_noise, const int64_t* row_index,
                                           int64_t B, int K, int T, int D, float sigma, float anchor_sigma, int mode_dist,
                                           int clamp_endpoints, int recompute_velocity, float* out,
                                           idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && D >= 1 && T >= 1 && K >= 2, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(B == 0 || (source && idx && out), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((!anchor_noise && !path_noise) || D >= 2, IDB200_EINVAL, "noise needs D >= 2");
    const size_t smem = sizeof(int) * K + sizeof(float) * (static_cast<size_t>(K) * D + 2 * static_cast<size_t>(T));
    IDB_REQUIRE(smem <= 48 * 1024, IDB200_EUNSUPPORTED, "K*D + 2T too large for one CTA (%zu bytes)", smem);
    if (B == 0) return IDB200_OK;
    const int vel = (recompute_velocity && D == 4) ? 1 : 0;
    const float* an = (anchor_sigma > 0.0f) ? anchor_noise : nullptr;   // :486 `if anchor_sigma > 0.0`
    const float* pn = (sigma > 0.0f) ? path_noise : nullptr;            // :496 `if sigma > 0.0`
    const int grid = grid_for(B, 1, 16);
    corrupt_from_anchors_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        source, reinterpret_cast<const long long*>(idx), reinterpret_cast<const long long*>(idx_gather), an, pn,
        reinterpret_cast<const long long*>(row_index), B, K, T, D, sigma, anchor_sigma, mode_dist, clamp_endpoints, vel,
        static_cast<float>(1.0 / static_cast<double>(T)), out);
    return check_launch("corrupt_from_anchors_kernel");
}

extern "C" int idb200_corrupt_adjacent(const float* source, const uint8_t* masks_levels, const int64_t* s_idx, int64_t B, int T, int D,
                                       int n_levels, const float* sigma_levels, const float* anchor_sigma_levels,
                                       const float* anchor_noise, const float* path_noise, int Kmax, uint64_t seed, uint64_t offset,
                                       float* anchor_noise_out, float* path_noise_out, int mode_dist, int clamp_endpoints,
                                       int recompute_velocity, float* x_s, float* x_prev, uint8_t* mask_s, uint8_t* mask_prev,
                                       idb200_stream_t stream) {
    IDB_REQUIRE(B >= 0 && D >= 1 && T >= 2, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(n_levels >= 2 && n_levels <= kMaxLevels + 1, IDB200_EUNSUPPORTED, "n_levels must be in [2,%d]", kMaxLevels + 1);
    IDB_REQUIRE(B == 0 || (source && masks_levels && s_idx && x_s && sigma_levels && anchor_sigma_levels), IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE((anchor_noise == nullptr) == (path_noise == nullptr), IDB200_EINVAL, "anchor_noise and path_noise must both be given (parity mode) or both NULL (Philox mode)");
    IDB_REQUIRE(Kmax >= 2 && Kmax <= T, IDB200_EINVAL, "Kmax must be in [2, T]");
    bool any_noise = false;
    for (int l = 0; l < n_levels; ++l) any_noise = any_noise || sigma_levels[l] > 0.0f || anchor_sigma_levels[l] > 0.0f;
    IDB_REQUIRE(!any_noise || D >= 2, IDB200_EINVAL, "noise needs D >= 2");
    const size_t smem = sizeof(int) * T + sizeof(float) * (static_cast<size_t>(T) * D + 2 * static_cast<size_t>(T));
    IDB_REQUIRE(smem <= 48 * 1024, IDB200_EUNSUPPORTED, "T*D too large for one CTA (%zu bytes)", smem);
    if (B == 0) return IDB200_OK;
    AdjParams p{};
    p.source = source; p.masks = masks_levels; p.s_idx = reinterpret_cast<const long long*>(s_idx);
    p.anchor_noise = anchor_noise; p.path_noise = path_noise; p.anchor_noise_out = anchor_noise_out; p.path_noise_out = path_noise_out;
    p.seed = seed; p.offset = offset;
    for (int l = 0; l < n_levels; ++l) { p.sigma[l] = sigma_levels[l]; p.anchor_sigma[l] = anchor_sigma_levels[l]; }
    p.B = B; p.T = T; p.D = D; p.n_levels = n_levels; p.Kmax = Kmax; p.mode_dist = mode_dist; p.clamp_endpoints = clamp_endpoints;
    p.vel = (recompute_velocity && D == 4) ? 1 : 0;
    p.philox = anchor_noise == nullptr ? 1 : 0;
    p.adjacent = x_prev != nullptr ? 1 : 0;
    p.dt = static_cast<float>(1.0 / static_cast<double>(T));
    p.x_s = x_s; p.x_prev = x_prev; p.mask_s = mask_s; p.mask_prev = mask_prev;
    corrupt_adjacent_kernel<<<grid_for(B, 1, 16), 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("corrupt_adjacent_kernel");
}
