"""ln_film timing (dev): python tools/bench_ln.py ; IDB200_LN_BULK=0 selects the register-resident one-row-per-warp kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpolated_diffusion_b200.models import _engine as E
for (B, L, d) in [(16384, 64, 384), (16384, 8, 384), (65536, 64, 256), (8192, 256, 256)]:
    M = B * L
    h = torch.randn((M, d), device="cuda")
    w, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    gb = torch.randn((B, 2 * d), device="cuda")
    out = torch.empty((M, d), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        E.ln_film(h, w, b, gb, out, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        E.ln_film(h, w, b, gb, out, L)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"M={M} d={d} L={L}: {ms:.3f} ms  {M * d * 6 / ms / 1e6:.0f} GB/s", flush=True)
