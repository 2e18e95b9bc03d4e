"""Dev micro-benchmark of the tcgen05 GEMM at the denoiser's shapes."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200 import _lib as L

res = {}
SHAPES = [(524288, 768, 256, 0), (524288, 256, 256, 2), (524288, 1024, 256, 1), (524288, 256, 1024, 2),
          (4194304, 768, 256, 0), (4194304, 1024, 256, 1), (4194304, 256, 1024, 2)]
if len(sys.argv) > 1:                                   # "M,N,K,epi;M,N,K,epi;..."
    SHAPES = [tuple(int(v) for v in t.split(",")) for t in sys.argv[1].split(";")]
for (M, N, K, epi) in SHAPES:
    A = torch.randn((M, K), device="cuda").bfloat16()
    W = torch.randn((N, K), device="cuda").bfloat16()
    bias = torch.randn((N,), device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.float32 if epi in (2, 3) else torch.bfloat16)
    aux = torch.randn((M, N), device="cuda").bfloat16() if epi >= 4 else None      # epilogues 4 / 5: the training forms with SiLU folded in
    def run():
        if aux is not None:
            L.call("idb200_gemm_bf16_aux", A.data_ptr(), W.data_ptr(), bias.data_ptr() if epi == 4 else None, out.data_ptr(), aux.data_ptr(), M, N, K, epi,
                   L.stream(A.device))
        else:
            L.call("idb200_gemm_bf16", A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epi, L.stream(A.device))
    for _ in range(3): run()
    torch.cuda.synchronize()
    it = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    flops = 2.0 * M * N * K
    byts = M * K * 2 + N * K * 2 + M * N * {0: 2, 1: 2, 2: 8, 3: 4, 4: 4, 5: 4}[epi]
    # torch reference (cuBLAS) for context
    for _ in range(2): torch.matmul(A, W.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): torch.matmul(A, W.t())
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / it
    res[f"{M}x{N}x{K}/epi{epi}"] = {"ms": ms, "TFLOPs": flops / ms / 1e9, "GBps": byts / ms / 1e6, "cublas_ms": ms_t, "cublas_TFLOPs": flops / ms_t / 1e9}
    del A, W, out, aux
print(json.dumps(res, indent=1))
