// Dev microbenchmark: issue rate of the legacy tensor path (mma.sync m16n8k16) on sm_100a for bf16 / f16 inputs with f32 / f16 accumulators.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hmma_rate.bin tools/hmma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int kMode>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned long long* out, float* sink) {
    unsigned a[4] = {threadIdx.x + 1u, threadIdx.x * 3u + 7u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u};
    float c[4][4] = {};
    unsigned h[4][2] = {};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {           // four independent accumulators per warp
            if (kMode == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (kMode == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(h[j][0]), "+r"(h[j][1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0;
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][3] + __uint_as_float(h[j][0]);
    if (s == 12345.f) sink[threadIdx.x] = s;
}
int main() {
    unsigned long long* d; float* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
    const int iters = 4000;
    const char* names[3] = {"bf16 -> f32", "f16 -> f32", "f16 -> f16"};
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(iters, d, sink);
                if (mode == 1) k<1><<<148, warps * 32>>>(iters, d, sink);
                if (mode == 2) k<2><<<148, warps * 32>>>(iters, d, sink);
                cudaDeviceSynchronize();
            }
            unsigned long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double per_smsp = double(warps) / 4.0 * iters * 4;      // mma instructions per sub-partition
            printf("%s, %2d warps: %.1f cycles per mma.sync per sub-partition (%s)\n", names[mode], warps, h / per_smsp, cudaGetErrorString(cudaGetLastError()));
        }
}
