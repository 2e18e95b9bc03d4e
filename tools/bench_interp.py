"""Dev micro-benchmark of K1 (BASELINE.json config 2): B = 2^20, T = 64, D = 4, S = 3."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.corruptions import keyframes as kf

B, T, D, K, S = 1 << 20, 64, 4, 8, 3
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
gen = torch.Generator(device="cuda").manual_seed(1234)
scores = torch.rand((B, T - 2), generator=gen, device="cuda")
x0 = torch.rand((B, T, D), generator=gen, device="cuda")
K_list = kf._compute_k_schedule(T, K, S)
res = {}
for name, kw in (("masks+interp", dict(x0=x0, levels_out=(1, S), want_idx=False)),
                 ("masks+interp+idx", dict(x0=x0, levels_out=(1, S), want_idx=True)),
                 ("masks only", dict(want_idx=False))):
    for _ in range(3):
        kf.nested_masks_interp(scores, T, K_list, **kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        kf.nested_masks_interp(scores, T, K_list, **kw)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    med = ms[len(ms) // 2]
    bytes_alg = B * (T * D * 4 * (1 + S) + (T - 2) * 4 + (S + 1) * T) if "interp" in name else B * ((T - 2) * 4 + (S + 1) * T)
    if "idx" in name:
        bytes_alg += B * 8 * sum(K_list)
    res[name] = {"ms_median": med, "ms_min": ms[0], "GBps": bytes_alg / med / 1e6, "traj_per_s": B / med * 1e3}
print(json.dumps(res, indent=1))
