"""Dev: weight-gradient GEMMs of all 12 layers -- per layer (split-K + reduce_rows, the split count of train/backward.py) against
ONE launch over the stacked token axis with splits = n_layers.  python tools/bench_dweight_grouped.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200 import _lib as L
from interpolated_diffusion_b200.train import backward as bw

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nl = 12
dev = torch.device("cuda:0")
sc = bw._Scratch()
for (n_out, k_in) in [(384, 1536), (1536, 384), (1152, 384), (384, 384)]:
    dy = torch.randn((nl, M, n_out), device=dev).bfloat16()
    x = torch.randn((nl, M, k_in), device=dev).bfloat16()
    out = torch.empty((nl, n_out, k_in), device=dev)
    part = torch.empty((nl, n_out, k_in), device=dev)
    def per_layer():
        for i in range(nl):
            sc.dweight(dy[i], x[i], out[i])
    def grouped():
        L.call("idb200_gemm_bf16_nn_splitk", dy.data_ptr(), x.data_ptr(), part.data_ptr(), n_out, k_in, nl * M, nl, L.stream(dev))
    res = []
    for fn in (per_layer, grouped):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    fl = 2.0 * nl * M * n_out * k_in
    err = float((out - part).abs().max() / out.abs().max())
    print(f"dW [{n_out} x {k_in}] x {nl} layers, {M} tokens: per layer {res[0]:.3f} ms ({fl / res[0] / 1e9:.0f} TF/s), grouped {res[1]:.3f} ms ({fl / res[1] / 1e9:.0f} TF/s), rel diff {err:.1e}")
