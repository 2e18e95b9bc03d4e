// Dynamic-programming anchor placement, src/selection/epiplexity_dp.py:171-228 (dp_select_indices[_batch]): choose K strictly
// increasing indices 0 = i_0 < ... < i_{K-1} = T-1 minimising sum_k C[i_{k-1}, i_k].  The reference runs K*T tiny launches with a
// python double loop; here one block per sample keeps the DP rows and the parent table in shared memory.
//   dp[k][j] = min_{i<j} dp[k-1][i] + C[i][j],  parent[k][j] = first minimising i (torch.argmin's tie rule), dp[0][0] = 0.
// fp32 additions in the reference's order, so the selected indices are bit-identical.
#include "common.cuh"

namespace idb200 {
namespace sel {

__global__ void __launch_bounds__(256) dp_select_kernel(const float* __restrict__ C, int T, int K, long long* __restrict__ idx,
                                                        int* __restrict__ status) {
    extern __shared__ unsigned char smem_raw[];
    float* dp_prev = reinterpret_cast<float*>(smem_raw);          // [T]
    float* dp_cur = dp_prev + T;                                   // [T]
    short* parent = reinterpret_cast<short*>(dp_cur + T);         // [K][T]
    const long long b = blockIdx.x;
    const float* Cb = C + b * T * T;
    const float inf = __int_as_float(0x7f800000);
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        dp_prev[j] = j == 0 ? 0.0f : inf;
        parent[j] = -1;
    }
    __syncthreads();
    for (int k = 1; k < K; ++k) {
        for (int j = threadIdx.x; j < T; j += blockDim.x) {
            float best = inf;
            int arg = -1;
            if (j >= 1) {
                best = __fadd_rn(dp_prev[0], Cb[j]);
                arg = 0;
                for (int i = 1; i < j; ++i) {
                    const float v = __fadd_rn(dp_prev[i], Cb[static_cast<long long>(i) * T + j]);
                    if (v < best) { best = v; arg = i; }
                }
            }
            dp_cur[j] = best;
            parent[k * T + j] = static_cast<short>(arg);
        }
        __syncthreads();
        float* t = dp_prev; dp_prev = dp_cur; dp_cur = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int st = 0;
        const float last = dp_prev[T - 1];
        if (!(last < inf) || last != last) st = 1;                 // "DP failed to find a valid path to T-1"
        long long* out = idx + b * K;
        out[K - 1] = T - 1;
        int cur = T - 1;
        for (int k = K - 1; k >= 1; --k) {
            cur = (cur >= 0) ? parent[k * T + cur] : -1;
            if (cur < 0 && st == 0) st = 2;                        // "DP backtrack failed"
            out[k - 1] = cur < 0 ? 0 : cur;
        }
        status[b] = st;
    }
}

}  // namespace sel
}  // namespace idb200

using namespace idb200;

extern "C" int idb200_dp_select(const float* C, int64_t B, int T, int K, int64_t* idx, int* status, idb200_stream_t stream) {
    IDB_REQUIRE(C && idx && status, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && T >= 2, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(K >= 2, IDB200_EINVAL, "K must be >= 2");
    IDB_REQUIRE(K <= T, IDB200_EINVAL, "K must be <= T (the host clamps K to T like the reference)");
    const size_t smem = static_cast<size_t>(2 * T) * sizeof(float) + static_cast<size_t>(K) * T * sizeof(short);
    IDB_REQUIRE(T <= 32767 && smem <= 200 * 1024, IDB200_EUNSUPPORTED, "K * T too large for the shared-memory parent table");
    if (B == 0) return IDB200_OK;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sel::dp_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(IDB200_ECUDA, "cudaFuncSetAttribute(dp_select): %s", cudaGetErrorString(e));
    }
    const int threads = T >= 256 ? 256 : ((T + 31) / 32) * 32;
    sel::dp_select_kernel<<<static_cast<unsigned>(B), threads, smem, static_cast<cudaStream_t>(stream)>>>(C, T, K, reinterpret_cast<long long*>(idx), status);
    return check_launch("dp_select_kernel");
}

// Segment costs, src/selection/epiplexity_dp.py:120-147 (compute_segment_costs_batch): for every sample b and segment s = (i, j),
//   cost = weight[s] * weight_scale * sum_n || x[t_idx[s,n]] - (x_i + alpha[s,n] (x_j - x_i)) ||^2   over the 2 position dims,
// the squared deviation of the trajectory from the chord at the segment's sample points.  One thread per (b, s); the reference
// materialises [B, S, n, 2] gathers.  fp32 in the reference's operation order.
namespace idb200 {
namespace sel {
__global__ void __launch_bounds__(256) segment_costs_kernel(const float* __restrict__ x, long long B, int T, int D, const long long* __restrict__ seg_i,
                                                            const long long* __restrict__ seg_j, const long long* __restrict__ t_idx,
                                                            const float* __restrict__ alpha, const float* __restrict__ weight, int S, int n,
                                                            float weight_scale, int apply_scale, float* __restrict__ out) {
    const long long total = B * S;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int s = static_cast<int>(e % S);
        const long long b = e / S;
        const float* xb = x + b * T * D;
        const float xi0 = xb[seg_i[s] * D], xi1 = xb[seg_i[s] * D + 1];
        const float d0 = __fsub_rn(xb[seg_j[s] * D], xi0), d1 = __fsub_rn(xb[seg_j[s] * D + 1], xi1);
        float acc = 0.0f;
        for (int k = 0; k < n; ++k) {
            const float a = alpha[s * n + k];
            const long long t = t_idx[s * n + k];
            const float e0 = __fsub_rn(xb[t * D], __fadd_rn(xi0, __fmul_rn(a, d0)));
            const float e1 = __fsub_rn(xb[t * D + 1], __fadd_rn(xi1, __fmul_rn(a, d1)));
            acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)));
        }
        float c = __fmul_rn(acc, weight[s]);
        if (apply_scale) c = __fmul_rn(c, weight_scale);
        out[e] = c;
    }
}
}  // namespace sel
}  // namespace idb200

extern "C" int idb200_segment_costs(const float* x_pos, int64_t B, int T, int D, const int64_t* seg_i, const int64_t* seg_j, const int64_t* t_idx,
                                    const float* alpha, const float* weight, int S, int n, float weight_scale, float* out,
                                    idb200_stream_t stream) {
    IDB_REQUIRE(x_pos && seg_i && seg_j && t_idx && alpha && weight && out, IDB200_EINVAL, "NULL pointer");
    IDB_REQUIRE(B >= 0 && T >= 2 && S >= 1 && n >= 1, IDB200_EINVAL, "bad shape");
    IDB_REQUIRE(D >= 2, IDB200_EINVAL, "x_pos must have at least 2 dims");
    if (B == 0) return IDB200_OK;
    sel::segment_costs_kernel<<<grid_for(B * S, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_pos, B, T, D, reinterpret_cast<const long long*>(seg_i), reinterpret_cast<const long long*>(seg_j),
        reinterpret_cast<const long long*>(t_idx), alpha, weight, S, n, weight_scale, weight_scale != 1.0f ? 1 : 0, out);
    return check_launch("segment_costs_kernel");
}
