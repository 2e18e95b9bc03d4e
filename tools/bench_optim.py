"""Dev / profiles: HBM throughput of the fused clip + AdamW + EMA step on the large model's parameter count (24.3 M)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.train.optim import FlatAdamW, stage2_loss
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24295042
p = [torch.nn.Parameter(torch.randn(n, device="cuda"))]
opt = FlatAdamW(p)
g = torch.randn(opt.n, device="cuda") * 0.01
def timed(fn, name, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.4f} ms, {nbytes/ms/1e6:.0f} GB/s", flush=True)
    return {"ms": ms, "GBps": nbytes / ms / 1e6}
res = {"clip+adamw+ema (24.3M params, 40 B/param)": timed(lambda: opt.step(g), "clip + adamw + ema", opt.n * 40)}
B, T, D = 4096, 64, 2
dh, tg, cf = torch.randn((B, T, D), device="cuda"), torch.randn((B, T, D), device="cuda"), torch.rand((B, T), device="cuda")
res["stage2_loss+grad (B=4096)"] = timed(lambda: stage2_loss(dh, tg, cf), "stage2 loss + grad", B * T * (D * 4 * 3 + 4) * 1.5)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bench_optim.json"), "w"), indent=1)
