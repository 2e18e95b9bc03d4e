"""Secondary BASELINE configurations through the same generation path (not bench.py lines): cfg 3 with D = 4 (velocity channels),
cfg 5 (T = 256, K = 32, levels = 4, causal Stage-2 denoiser), and the batched causal chunk loop.  Device-resident inputs, CUDA
events, 2 warm-ups + 3 timed runs each.  python tools/bench_configs.py [--out profiles/x.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
    from interpolated_diffusion_b200.sample.sample_generate_causal import generate_causal_chunked
    res = {}
    gen = torch.Generator().manual_seed(1)

    def cond_of(B):
        return {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}

    # cfg 3, D = 4
    B = 32768
    torch.manual_seed(0)
    kp, il = KeypointDenoiser(data_dim=4).cuda(), InterpLevelDenoiser(data_dim=4, max_levels=3, mask_channels=2).cuda()
    cfg = GenerationConfig(data_dim=4)
    cond, z = cond_of(B), torch.randn((B, 8, 4), generator=gen).cuda()
    g = GenerationGraph(kp, il, B, cfg)
    ms = timed(lambda: g.run(cond, z))
    res["cfg3_D4_small"] = {"B": B, "ms": ms, "traj_per_s": B / ms * 1e3}
    del g, kp, il
    # cfg 5: T = 256, K = 32, levels = 4, causal Stage-2
    B = 8192
    torch.manual_seed(0)
    kp, il = KeypointDenoiser(data_dim=2).cuda(), InterpLevelCausalDenoiser(data_dim=2, max_levels=4, mask_channels=2).cuda()
    cfg = GenerationConfig(T=256, K_min=32, levels=4)
    cond, z = cond_of(B), torch.randn((B, 32, 2), generator=gen).cuda()
    g = GenerationGraph(kp, il, B, cfg)
    ms = timed(lambda: g.run(cond, z))
    gf = (19 * 413.1 + 3760.0 / 2 + 16.5) / 1000.0          # Stage-1 L=32 evals + causal Stage-2 (half of dense) + conv
    res["cfg5_T256_K32_causal_small"] = {"B": B, "ms": ms, "traj_per_s": B / ms * 1e3, "tflops": B * gf / ms}
    del g
    # batched causal chunk loop (sample_generate_causal), T = 256, chunk 16, mask_channels = 1
    il1 = InterpLevelCausalDenoiser(data_dim=2, max_levels=3, mask_channels=1).cuda()
    B = 2048
    cond = cond_of(B)
    gg = torch.Generator(device="cuda").manual_seed(3)
    ms = timed(lambda: generate_causal_chunked(kp, il1, cond, T=256, chunk=16, K_min=8, levels=3, logit_space=True, generator=gg), n=2, warm=1)
    res["causal_chunked_T256_chunk16_small"] = {"B": B, "ms": ms, "traj_per_s": B / ms * 1e3}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
