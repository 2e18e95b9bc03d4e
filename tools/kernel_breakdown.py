"""Dev: per-entry-point time of one eager pass (every library call bracketed by CUDA events on the launching stream; each call
enqueues one kernel), keyed by entry point + its small integer arguments (shapes / flags).

    python tools/kernel_breakdown.py gen-large|gen-small|train|train-small|long [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from interpolated_diffusion_b200 import _lib as L


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "gen-large"
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    if mode.startswith("gen") or mode == "long":
        from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
        large = mode == "gen-large"
        B = int(sys.argv[2]) if len(sys.argv) > 2 else (16384 if large else 65536)
        if mode == "long":
            B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
            from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
            from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
            kp = KeypointDenoiser(data_dim=bench.D).to(dev)
            il = InterpLevelCausalDenoiser(data_dim=bench.D, max_levels=4, mask_channels=2).to(dev)
            cfg = GenerationConfig(T=256, K_min=32, levels=4, data_dim=bench.D)
        else:
            kp, il = bench.build_models(dev, large=large)
            cfg = GenerationConfig(T=bench.T, K_min=bench.K_MIN, levels=bench.LEVELS, data_dim=bench.D)
        graph = GenerationGraph(kp, il, B, cfg, device=dev)
        cond = bench.synthetic_cond(B, 1, device=dev)
        graph._body()                                        # warm-up (packs weights)
        body = graph._body

    else:
        from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
        from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        kw = bench.LARGE if mode == "train" else {}
        model = InterpLevelDenoiser(data_dim=bench.D, max_levels=bench.LEVELS, mask_channels=3, **kw).to(dev)
        tr = Stage2Trainer(model, cuda_graph=False, batch_mode="fused")
        g = torch.Generator(device=dev).manual_seed(1)
        x0 = torch.rand((B, bench.T, bench.D), device=dev, generator=g)
        cond = {"occ": (torch.rand((B, 1, 21, 21), device=dev, generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), device=dev, generator=g)}
        gen = torch.Generator(device=dev).manual_seed(2)
        for _ in range(2):
            tr.step(x0, cond, gen)
        body = lambda: tr.step(x0, cond, gen)
    torch.cuda.synchronize()
    evs = []
    orig = L.call

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *a)
        e1.record()
        key = name + " " + ",".join(str(x) for x in a if isinstance(x, int) and not isinstance(x, bool) and 0 <= x < (1 << 25))
        evs.append((key, e0, e1))
        return r

    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.call = timed
    w0.record()
    body()
    w1.record()
    torch.cuda.synchronize()
    L.call = orig
    agg = {}
    for key, e0, e1 in evs:
        t = agg.setdefault(key, [0, 0.0])
        t[0] += 1
        t[1] += e0.elapsed_time(e1)
    tot = sum(t[1] for t in agg.values())
    print(f"{mode} B={B}: {len(evs)} library calls, {tot:.2f} ms in kernels, {w0.elapsed_time(w1):.2f} ms wall (eager)")
    byname = {}
    for key, (n, ms) in agg.items():
        b = byname.setdefault(key.split(" ")[0], [0, 0.0])
        b[0] += n
        b[1] += ms
    print("-- by entry point")
    for name, (n, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}%  x{n:<5d} {name}")
    print("-- by entry point + integer arguments (top 40)")
    for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}%  x{n:<5d} {key}")


if __name__ == "__main__":
    main()
