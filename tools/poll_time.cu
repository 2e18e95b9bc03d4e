// Dev: how long does a FAILED mbarrier.try_wait with a suspend-time hint take?  (sizes the poll budget of tc::mbar_wait)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t hint, int n, long long* out) {
    __shared__ uint64_t bar;
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    __syncthreads();
    long long t0 = clock64();
    uint32_t ok = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t o;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(o) : "r"(a), "r"(0u), "r"(hint) : "memory");
        ok += o;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = ok; }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    for (uint32_t hint : {0u, 1000u, 100000u, 1000000u, 10000000u}) {
        for (int threads : {32, 640}) {
            k<<<1, threads>>>(hint, 200, d);
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("hint %u ns, %d threads: %.0f cycles per failed poll (ok=%lld) %s\n", hint, threads, h[0] / 200.0, h[1], cudaGetErrorString(cudaGetLastError()));
        }
    }
}
