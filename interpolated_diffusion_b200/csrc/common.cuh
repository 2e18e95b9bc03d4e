// Shared helpers for libidb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/idb200.h"

namespace idb200 {

// thread-local message returned by idb200_last_error()
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IDB200_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return IDB200_OK;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// grid for a grid-stride kernel: enough CTAs for `work` items of `per_cta`, capped at
// ctas_per_sm full waves of the device (148 SMs on B200).
inline int grid_for(int64_t work, int per_cta, int ctas_per_sm) {
    int64_t need = (work + per_cta - 1) / per_cta;
    int64_t cap = static_cast<int64_t>(num_sms()) * ctas_per_sm;
    if (need < 1) need = 1;
    return static_cast<int>(need < cap ? need : cap);
}

#ifdef __CUDACC__
// SiLU / SiLU' on the bf16 token tensors of the training step (ff.0 pre-activation u; transformer.py:23-26) with ONE MUFU op:
// sigmoid(x) = 1/2 + 1/2 tanh(x/2) (tanh.approx: relative error ~2^-11, below the bf16 rounding of what is stored).  Shared by the
// stand-alone passes (train_bwd.cu) and the GEMM epilogues that fold them in (gemm.cu) so both forms round identically; explicit
// _rn intrinsics keep ptxas from contracting differently at the two sites.
__device__ __forceinline__ float tanh_approx(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}
__device__ __forceinline__ float silu16_fwd(float x) {
    const float h = __fmul_rn(0.5f, x);
    return __fmaf_rn(h, tanh_approx(h), h);
}
__device__ __forceinline__ float silu16_grad(float x) {      // s + x s (1 - s), s = sigmoid(x); 4 s (1 - s) = 1 - tanh^2(x/2)
    const float h = __fmul_rn(0.5f, x);
    const float t = tanh_approx(h);
    return __fmaf_rn(__fmul_rn(0.5f, h), __fmaf_rn(-t, t, 1.0f), __fmaf_rn(0.5f, t, 0.5f));
}
#endif

}  // namespace idb200

#define IDB_REQUIRE(cond, code, ...)                          \
    do {                                                      \
        if (!(cond)) return ::idb200::fail((code), __VA_ARGS__); \
    } while (0)
