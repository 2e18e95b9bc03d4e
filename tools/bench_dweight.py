"""Dev: the split-K weight-gradient GEMM (idb200_gemm_bf16_nn_splitk + idb200_reduce_rows) over split counts, at the shapes of
the cfg-4 training step.  python tools/bench_dweight.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200 import _lib as L

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda:0")
for (n_out, k_in) in [(384, 1536), (1536, 384), (1152, 384), (384, 1152), (384, 384)]:
    dy = torch.randn((M, n_out), device=dev).bfloat16()
    x = torch.randn((M, k_in), device=dev).bfloat16()
    out = torch.empty((n_out, k_in), device=dev)
    line = []
    for splits in (2, 4, 8, 16, 32, 64):
        if (M // 64) % splits:
            continue
        part = torch.empty((splits, n_out, k_in), device=dev)
        def run():
            L.call("idb200_gemm_bf16_nn_splitk", dy.data_ptr(), x.data_ptr(), part.data_ptr(), n_out, k_in, M, splits, L.stream(dev))
            L.call("idb200_reduce_rows", part.data_ptr(), splits, n_out * k_in, 1.0, 0, out.data_ptr(), L.stream(dev))
        for _ in range(2): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line.append(f"s={splits}: {ms:.3f} ms ({2.0 * M * n_out * k_in / ms / 1e9:.0f} TF/s)")
    print(f"dW [{n_out} x {k_in}], {M} tokens: " + "  ".join(line))
