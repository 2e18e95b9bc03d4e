// Dev microbenchmark: L2 -> shared-memory bandwidth of TMA tile loads of an L2-resident weight set by all SMs,
// unicast vs cluster multicast (decides how the fused transformer kernels share weight tiles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2bw tools/l2bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try(b, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("timeout block %d\n", blockIdx.x); __trap(); }
    }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

constexpr int kSlots = 8;
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * 64 * 2;

template <int C>
__global__ void __launch_bounds__(128, 1) l2bw_kernel(const __grid_constant__ CUtensorMap tmap, int rows_total, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSlots * kTileBytes);
    uint64_t* empty = full + kSlots;
    const uint32_t rank = (C > 1) ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], C); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (C > 1) cluster_sync();
    if (threadIdx.x == 0) {
        const int ntiles = rows_total / kTileRows;
        constexpr int kLag = kSlots / 2;                 // consume tile i - kLag, refill the slot of tile i - kSlots
        for (int i = 0; i < iters + kLag; ++i) {
            if (i >= kLag) {
                const int j = i - kLag;                  // consume tile j
                const int cs = j % kSlots;
                mbar_wait(&full[cs], (j / kSlots) & 1);
                if (C > 1) for (uint32_t c = 0; c < C; ++c) mbar_arrive_remote(&empty[cs], c);
            }
            if (i < iters) {
                const int slot = i % kSlots;
                const uint32_t use = i / kSlots;
                if (C > 1 && use > 0) mbar_wait(&empty[slot], (use - 1) & 1);   // every CTA of the cluster consumed the previous use
                const int tile = (i + (C > 1 ? 0 : blockIdx.x)) % ntiles;      // unicast: CTAs stagger through the weight set
                mbar_expect(&full[slot], kTileBytes);
                const int row0 = tile * kTileRows + rank * (kTileRows / C);
                uint8_t* dst = smem + slot * kTileBytes + rank * (kTileBytes / C);
                if (C == 1) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&full[slot])), "r"(0), "r"(row0) : "memory");
                } else {
                    const uint16_t mask = static_cast<uint16_t>((1u << C) - 1);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&full[slot])), "r"(0), "r"(row0), "h"(mask) : "memory");
                }
            }
        }
    }
    __syncthreads();
    if (C > 1) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C>
float run(const CUtensorMap& tm, int rows, int iters, int grid) {
    const int smem = kSlots * kTileBytes + 1024 + 256;
    CK(cudaFuncSetAttribute(l2bw_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, l2bw_kernel<C>, tm, rows, iters / 4));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, l2bw_kernel<C>, tm, rows, iters));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int mb : {1, 2, 12}) {
        const int rows = mb * 8192;      // [rows, 64] bf16 = mb MiB
        void* w;
        CK(cudaMalloc(&w, static_cast<size_t>(rows) * 128));
        CK(cudaMemset(w, 1, static_cast<size_t>(rows) * 128));
        for (int C : {1, 2, 4}) {
            CUtensorMap tm;
            cuuint64_t gdim[2] = {64, static_cast<cuuint64_t>(rows)};
            cuuint64_t gstr[1] = {128};
            cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kTileRows / C)};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int div : {1, 2, 4}) {
            const int grid = (sms / div / C) * C;
            float ms = (C == 1) ? run<1>(tm, rows, iters, grid) : (C == 2) ? run<2>(tm, rows, iters, grid) : run<4>(tm, rows, iters, grid);
            const double smem_bytes = static_cast<double>(grid) * iters * kTileBytes;
            printf("weights %2d MiB  cluster %d  grid %3d: %.3f ms  smem fill %.2f TB/s  (L2 reads %.2f TB/s)  per-SM %.1f B/ns\n", mb, C, grid, ms,
                   smem_bytes / ms / 1e9, smem_bytes / C / ms / 1e9, smem_bytes / grid / ms / 1e6);
            }
        }
        CK(cudaFree(w));
    }
    return 0;
}
