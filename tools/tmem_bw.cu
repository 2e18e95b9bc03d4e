// Dev microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM for different warp counts / loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw tools/tmem_bw.cu && /tmp/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD32(taddr, r)                                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),       \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),       \
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                        \
                 : "r"(taddr) : "memory")

template <int kInFlight>
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
            uint32_t r[kInFlight][32];
#pragma unroll
            for (int f = 0; f < kInFlight; ++f) LD32(base + ((it * kInFlight + f) * 32) % 512, r[f]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int f = 0; f < kInFlight; ++f)
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= r[f][j];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
    unsigned long long* d; uint32_t* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
    const int iters = 2000;
    for (int nw : {1, 4, 8, 16}) {
        for (int fl : {1, 2, 4}) {
            unsigned long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (fl == 1) k<1><<<148, 512>>>(nw, iters, d, sink);
                if (fl == 2) k<2><<<148, 512>>>(nw, iters, d, sink);
                if (fl == 4) k<4><<<148, 512>>>(nw, iters, d, sink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = double(nw) * iters * fl * 4096.0;
            printf("warps %2d in-flight %d: %8llu cycles  %.1f B/clk/SM  (%.1f cycles per x32 load per warp)  err=%s\n", nw, fl, h, bytes / h,
                   double(h) / (iters * fl), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
