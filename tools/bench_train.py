"""Stage-2 training step (BASELINE cfg 4: T=64, K_min=8, levels=3, adj mode, large model, bf16 GEMMs, AdamW + EMA) on one GPU
or under torchrun (data parallel).  Prints per-phase device times (CUDA events) and trajectories/s.
    python tools/bench_train.py [--batch 512] [--model large|small] [--steps 5]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py --batch 512"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def bench_stage1(a, dev, rank, world):
    """Stage-1 step (train_keypoints.py:505-556): keypoint gather + q_sample + KeypointDenoiser forward / backward + AdamW + EMA."""
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.train.train_keypoints import Stage1Trainer
    kw = dict(d_model=384, n_layers=12, n_heads=12, d_ff=1536, maze_channels=(32, 64, 128, 128)) if a.model == "large" else {}
    model = KeypointDenoiser(data_dim=2, kp_feat_dim=0, **kw).to(dev)
    gf = 3 * ((345.6 + 211.6) if a.model == "large" else (103.3 + 16.5)) / 1000.0
    tr = Stage1Trainer(model, T=64, K=8, cuda_graph=bool(a.graph))
    B = a.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x0 = 0.1 + 0.8 * torch.rand((B, 64, 2), device=dev, generator=g)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device=dev, generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), device=dev, generator=g)}
    gen = torch.Generator(device=dev).manual_seed(23 + rank)
    for _ in range(a.warmup):
        loss = tr.step(x0, cond, gen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = tr.step(x0, cond, gen)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    res = {"config": f"stage1 train step, {a.model} model, B={B}/GPU x {world} GPU, K=8 of T=64", "ms_per_step": ms,
           "traj_per_s": B * world / ms * 1e3, "tflops_per_gpu": B * gf / ms, "loss": float(loss), "cuda_graph": bool(a.graph)}
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512, help="trajectories per GPU")
    ap.add_argument("--model", default="large")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--stage", type=int, default=2, help="2: Stage-2 interp-levels step (cfg 4); 1: Stage-1 keypoint step (K = 8 tokens)")
    ap.add_argument("--graph", type=int, default=0, help="1: forward + loss + backward replayed as one CUDA graph")
    ap.add_argument("--overlap", type=int, default=0, help="1: whole steps through trainer.step() with the next batch prefetched on a side stream")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    torch.manual_seed(0)
    if a.stage == 1:
        return bench_stage1(a, dev, rank, world)
    if a.model == "large":
        model = InterpLevelDenoiser(d_model=384, n_layers=12, n_heads=12, d_ff=1536, data_dim=2, max_levels=3, mask_channels=3,
                                    maze_channels=(32, 64, 128, 128))
        gf_per_traj = 3 * (2798.1 + 211.6) / 1000.0
    else:
        model = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=3)
        gf_per_traj = 3 * (841.0 + 16.5) / 1000.0
    model = model.to(dev)
    tr = Stage2Trainer(model)
    B, T = a.batch, 64
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x0 = torch.rand((B, T, 2), device=dev, generator=g)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device=dev, generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), device=dev, generator=g)}
    gen = torch.Generator(device=dev).manual_seed(23 + rank)
    if a.overlap:
        tr.cuda_graph = bool(a.graph)
        for _ in range(a.warmup):
            loss = tr.step(x0, cond, gen)
            tr.prefetch(x0, cond, gen)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = tr.step(x0, cond, gen)
            tr.prefetch(x0, cond, gen)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        res = {"config": f"stage2 train step, {a.model} model, B={B}/GPU x {world} GPU, T=64, adj, anchor_conf, next batch prefetched", "ms_per_step": ms,
               "traj_per_s": B * world / ms * 1e3, "tflops_per_gpu": B * gf_per_traj / ms, "loss": float(loss), "cuda_graph": bool(a.graph)}
        if rank == 0:
            print(json.dumps(res))
            if a.out:
                with open(a.out, "w") as f:
                    json.dump(res, f, indent=1)
        if world > 1:
            dist.destroy_process_group()
        return
    names = ["batch", "forward", "loss", "backward", "allreduce", "optimizer"]
    acc = {n: 0.0 for n in names}
    total = 0.0
    for it in range(a.warmup + a.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        x_s, s_idx, mask_in, target, wm = tr.build_batch(x0, gen)
        ev[1].record()
        if a.graph:
            ev[2].record()
            ev[3].record()
            loss = tr._graphed_loss_and_grads(x_s, s_idx, mask_in, cond, target, wm)      # phases: all under "backward"
        else:
            delta = tr.bp.forward(x_s, s_idx, mask_in, cond)
            ev[2].record()
            from interpolated_diffusion_b200.train.optim import stage2_loss
            loss, dgrad = stage2_loss(delta, target, wm, anchor_conf=True, w_anchor=0.1, w_missing=1.0, grad_accum=world)
            ev[3].record()
            tr.bp.backward(dgrad, tr.grads)
        ev[4].record()
        tr.reduce_gradients()
        ev[5].record()
        tr.opt.step(tr.flat_grad)
        ev[6].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1])
            total += ev[0].elapsed_time(ev[6])
    ms = total / a.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    res = {"config": f"stage2 train step, {a.model} model, B={B}/GPU x {world} GPU, T=64, adj, anchor_conf", "ms_per_step": ms,
           "traj_per_s": B * world / ms * 1e3, "tflops_per_gpu": B * gf_per_traj / ms, "loss": float(loss) * (1 if a.graph else world), "cuda_graph": bool(a.graph),
           "phases_ms": {n: acc[n] / a.steps for n in names}, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
