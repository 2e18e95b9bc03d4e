"""Dev / profiles: HBM throughput of the K2 elementwise kernels (DDIM step + known clamp, Stage-2 epilogue) at a size
where they are bandwidth-bound (the generation path runs them on 1-17 MB, where launch latency dominates)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.diffusion.ddpm import ddim_step_scalar
from interpolated_diffusion_b200.utils.clamp import stage2_epilogue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
K, T, D = 8, 64, 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
res = {}
def timed(fn, nbytes, name, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "bytes": nbytes}
    print(f"{name}: {ms:.4f} ms, {nbytes/ms/1e6:.0f} GB/s", flush=True)
# DDIM step + known-value clamp on [B, K, D] (z, eps read; known_mask 1 B/elem, known_values read; z' written)
z = torch.randn((B, K, D), generator=g, device=dev); eps = torch.randn((B, K, D), generator=g, device=dev)
km = torch.zeros((B, K, D), dtype=torch.bool, device=dev); km[:, 0] = km[:, -1] = True
kv = torch.randn((B, K, D), generator=g, device=dev); out = torch.empty_like(z)
n = B * K * D
timed(lambda: ddim_step_scalar(z, eps, 0.02555, 0.0512, known_mask=km, known_values=kv, out=out), n * (4 * 4 + 1), "ddim_step+known_clamp [B,8,2]")
# Stage-2 epilogue on [B/8, T, D]: x_in, delta, x_ref read, conf [B,T] read, out written
B2 = B // 8
x = torch.rand((B2, T, D), generator=g, device=dev); dl = torch.randn((B2, T, D), generator=g, device=dev) * 0.01
xr = torch.rand((B2, T, D), generator=g, device=dev); conf = torch.rand((B2, T), generator=g, device=dev); o2 = torch.empty_like(x)
n2 = B2 * T * D
timed(lambda: stage2_epilogue(x, dl, xr, conf, 0.5, "endpoints", None, "pos", out=o2), n2 * 16 + B2 * T * 4, "stage2_epilogue [B,64,2] soft+endpoints")
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bench_elementwise.json"), "w"), indent=1)
