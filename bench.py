#!/usr/bin/env python
"""Benchmark of the generation hot path (BASELINE.json: end-to-end trajectories/sec, DDIM-20 + interp + Stage-2; interp GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference itself on the host cores (baseline/_ref; oracle port if absent)

One "step" = one batch of synthetic particle-maze conditioning through the whole path: Stage-1 K=8 keypoint DDIM
(20 timesteps = 19 denoiser evaluations), sigmoid, piecewise-linear interpolation to T=64, Stage-2 one-step jump,
soft + hard clamp (clamp_policy=endpoints) -- BASELINE.json configs[2], B = 65536 trajectories per GPU (weak scaling), small
random-init models (d=256, 8 layers, 8 heads, ff 1024, maze 32-64).  Prints ONE JSON line on rank 0.  The same line carries
the other BASELINE configurations as extra keys, each measured in this run at this N:
    interp        configs[1]: the Interp(x0|M_s) kernel alone, 2^20 trajectories (rank 0)
    strong        configs[2] with the GLOBAL batch fixed at 65536 (B/N per GPU)
    train         configs[3]: Stage-2 training step, large model, global batch 4096 over N ranks (DP all-reduce)
    long_horizon  configs[4]: T=256 K=32 levels=4 causal, global batch 8192 (+ interp GB/s at T=256)
    large_model   configs[2] with the trainer-default 384x12 models (9.58 GFLOP per trajectory), 16384 trajectories per GPU
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "end-to-end trajectories/sec (DDIM20+interp+stage2)"
UNIT = "trajectories/s"
T, K_MIN, LEVELS, D = 64, 8, 3, 2
# algorithmic FLOPs per trajectory (small model): 19 Stage-1 evals + 1 Stage-2 eval + conv encoders
D_MODEL, N_LAYERS, D_FF, D_COND = 256, 8, 1024, 128


def gemm_flops_per_traj() -> float:
    per_tok_layer = 2 * D_MODEL * 3 * D_MODEL + 2 * D_MODEL * D_MODEL + 4 * D_MODEL * D_FF      # QKV + out + MLP
    return float(N_LAYERS * per_tok_layer * (19 * K_MIN + T))


def _attn_block_flops(a):
    M, Lq = a[9], a[10]
    return float(M) * (2 * D_MODEL * 3 * D_MODEL + 2 * D_MODEL * D_MODEL + 4 * Lq * D_MODEL)


# algorithmic FLOPs of one call of each dense entry point, from its C-ABI arguments (include/idb200.h)
DENSE_FLOPS = {
    "idb200_gemm_bf16": lambda a: 2.0 * a[4] * a[5] * a[6],
    "idb200_mlp_fused": lambda a: 4.0 * a[6] * a[7] * a[8],
    "idb200_mlp_block": lambda a: 4.0 * a[9] * a[11] * a[12],
    "idb200_attn_block": _attn_block_flops,
    # every layer of the encoder in one launch: n_layers * M * (QKV + out-proj + attention + MLP)
    # the same kernel with the token assembly / out head fused in (the head adds 2 * M * 256 * D, negligible, not counted)
    "idb200_denoiser_fused": lambda a: float(a[18]) * a[13] * (2.0 * a[15] * 3 * a[15] + 2.0 * a[15] * a[15] + 4.0 * a[14] * a[15] + 4.0 * a[15] * a[17]),
    "idb200_encoder_fused": lambda a: float(a[16]) * a[11] * (2.0 * a[13] * 3 * a[13] + 2.0 * a[13] * a[13] + 4.0 * a[12] * a[13] + 4.0 * a[13] * a[15]),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "bf16": p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0)), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def synthetic_cond(B: int, seed: int, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    occ = (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float()
    sg = torch.rand((B, 4), generator=g)
    if pin:
        occ, sg = occ.pin_memory(), sg.pin_memory()
    return {"occ": occ.to(device), "start_goal": sg.to(device)}


# trainer defaults of the reference (train_interp_levels.py:57-62): the "large" models of SURVEY 8d (second line of cfg 3)
LARGE = dict(d_model=384, n_layers=12, n_heads=12, d_ff=1536, maze_channels=(32, 64, 128, 128))


def build_models(device, large: bool = False):
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    torch.manual_seed(0)
    kw = LARGE if large else {}
    kp = KeypointDenoiser(data_dim=D, **kw)
    il = InterpLevelDenoiser(data_dim=D, max_levels=LEVELS, mask_channels=2, **kw)
    return kp.to(device), il.to(device)


def cpu_reference_rate(sample_B: int, steps: int = 1, warmup: int = 0):
    """The reference's generation path on the host cores -> (trajectories/s, s/step, threads, kind).  kind = "reference": the
    UNMODIFIED reference modules vendored under baseline/_ref (oracle/vendor_reference.py; git-ignored, travels with the gpurun
    snapshot) driven through their own component functions (oracle/ref_driver.py); kind = "port": the oracle restatement
    (oracle/generate.py of sample_generate.py:974-1285) when the vendored copy is absent."""
    from oracle import ref_driver
    if ref_driver.reference_root() is not None and not os.environ.get("IDB200_BENCH_PORT"):
        rate, dt, cores = ref_driver.timed_rate(lambda B: synthetic_cond(B, 0), sample_B, steps=steps, warmup=warmup)
        return rate, dt, cores, "reference"
    import numpy as np
    from oracle import generate as og
    torch.manual_seed(0)
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    kp = KeypointDenoiser(data_dim=D)
    il = InterpLevelDenoiser(data_dim=D, max_levels=LEVELS, mask_channels=2)
    sd_kp = {k: v.detach() for k, v in kp.state_dict().items()}
    sd_il = {k: v.detach() for k, v in il.state_dict().items()}
    cond = synthetic_cond(sample_B, 0)
    rng = np.random.default_rng(0)
    z_T = rng.standard_normal((sample_B, K_MIN, D)).astype(np.float32)
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            og.generate(sd_kp, sd_il, 8, cond, z_T, T=T, K_min=K_MIN, levels=LEVELS, D=D)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return sample_B / dt, dt, torch.get_num_threads(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = args.ref_batch
    rate, dt, cores, kind = cpu_reference_rate(sample_B, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "generation: Stage-1 DDIM-20 K=8 + interp T=64 + Stage-2 one-step (x0), small model 256x8 ff1024, "
                               "clamp_policy=endpoints, D=2 (BASELINE.json configs[2])",
                   "sample": f"{sample_B} trajectories per step (bounded CPU sample of the same workload)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{sample_B} trajectories/step x {args.steps} steps"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class _Ctx:
    """Per-process state shared by the bench sections."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the ONE JSON line (NCCL prints its version banner there)
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup, sample_clocks=False, finish=None):
        """W warm-ups, then K steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks.
        ``finish`` (optional) is enqueued after the last step inside the timed region (joins side streams)."""
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        torch.cuda.synchronize(self.dev)
        self.barrier()
        sampler = ClockSampler(self.local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(self.dev)
        e0.record()
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record()
        torch.cuda.synchronize(self.dev)
        self.barrier()
        clocks = sampler.stop() if sampler else None
        return self.max_over_ranks(e0.elapsed_time(e1)), clocks


class GatherOverlap:
    """The final all-gather of the samples taken off the compute stream: the step's samples are copied into one of
    two staging buffers (33 MB device copy) and all-gathered by NCCL on a side stream while the next step's graph runs."""

    def __init__(self, ctx: _Ctx, B: int, Tn: int, Dn: int):
        self.ctx = ctx
        self.side = torch.cuda.Stream(device=ctx.dev)
        self.stage = [torch.empty((B, Tn, Dn), device=ctx.dev) for _ in range(2)]
        self.out = [torch.empty((ctx.world * B, Tn, Dn), device=ctx.dev) for _ in range(2)]
        self.done = [None, None]
        self.i = 0

    def submit(self, x: torch.Tensor) -> torch.Tensor:
        k = self.i & 1
        self.i += 1
        main = torch.cuda.current_stream(self.ctx.dev)
        if self.done[k] is not None:
            main.wait_event(self.done[k])               # the all-gather that last read this staging buffer has finished
        self.stage[k].copy_(x, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self.ctx.dist.all_gather_into_tensor(self.out[k], self.stage[k])
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        return self.out[k]

    def join(self):
        torch.cuda.current_stream(self.ctx.dev).wait_stream(self.side)


def section_interp(ctx, Ti: int, Kmin: int, Si: int, Bi: int, label: str):
    """"interp GB/s": the Interp(x0|M_s) corruption kernel alone, D = 4; algorithmic bytes per trajectory = x0 read + S level writes
    + scores + (S+1) byte masks (4600 B at T=64/S=3, 22 776 B at T=256/S=4); median of 20 launches."""
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    dev = ctx.dev
    Di = 4
    gi = torch.Generator(device=dev).manual_seed(1234)
    scores = torch.rand((Bi, Ti - 2), generator=gi, device=dev)
    x0 = torch.rand((Bi, Ti, Di), generator=gi, device=dev)
    K_list = kf._compute_k_schedule(Ti, Kmin, Si)
    for _ in range(3):
        kf.nested_masks_interp(scores, Ti, K_list, x0=x0, levels_out=(1, Si), want_idx=False)
    torch.cuda.synchronize(dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    evs[0].record()
    for i in range(20):
        kf.nested_masks_interp(scores, Ti, K_list, x0=x0, levels_out=(1, Si), want_idx=False)
        evs[i + 1].record()
    torch.cuda.synchronize(dev)
    ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(20))[10]
    nbytes = Bi * (Ti * Di * 4 * (1 + Si) + (Ti - 2) * 4 + (Si + 1) * Ti)
    pk = peaks()
    return {"workload": label, "ms": ms, "GBps": nbytes / ms / 1e6, "peak_GBps": pk["hbm_gbs"], "frac": nbytes / ms / 1e6 / pk["hbm_gbs"],
            "trajectories_per_s": Bi / ms * 1e3, "bytes_per_trajectory": nbytes // Bi, "K_list": list(K_list)}


def section_train(ctx, args):
    """BASELINE.json configs[3]: Stage-2 interp-levels training step, T=64, K_min=8, levels=3, adj mode, anchor_conf (C=3), dist
    corruption on the device, large model (trainer defaults, train_interp_levels.py:57-62), bf16 GEMMs, AdamW + EMA, GLOBAL batch
    4096 sharded over the ranks (data parallel, gradient all-reduce over NCCL).  Forward + loss + backward replay as one CUDA graph;
    the next batch's corruption is built on a side stream."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    GB = args.train_batch
    B = GB // world
    torch.manual_seed(0)
    model = InterpLevelDenoiser(data_dim=D, max_levels=LEVELS, mask_channels=3, **LARGE).to(dev)
    tr = Stage2Trainer(model, cuda_graph=True, batch_mode="fused")
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x0 = torch.rand((B, T, D), device=dev, generator=g)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device=dev, generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), device=dev, generator=g)}
    gen = torch.Generator(device=dev).manual_seed(23 + rank)
    loss = [None]

    def step():
        loss[0] = tr.step(x0, cond, gen)
        tr.prefetch(x0, cond, gen)

    ms, _ = ctx.timed(step, args.train_steps, 3)
    ms /= args.train_steps
    # exposed all-reduce: time of step() with and without the reduction is not separable inside a graph; measure the collective
    # itself on the (idle) compute stream, after the timed region
    ar_ms = 0.0
    if world > 1:
        torch.cuda.synchronize(dev)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tr.reduce_gradients()
        e1.record()
        torch.cuda.synchronize(dev)
        ar_ms = ctx.max_over_ranks(e0.elapsed_time(e1) / 5)
    gf = 3 * (2798.1 + 211.6) / 1000.0            # 3 x (Stage-2 forward + conv encoder) GFLOP per trajectory
    pk = peaks()
    res = {"workload": "Stage-2 interp-levels training step, large model 384x12 ff1536 maze 32-64-128-128, T=64 K_min=8 levels=3, adj, "
                       "anchor_conf, dist corruption on device, bf16 GEMMs, AdamW+EMA (BASELINE.json configs[3])",
           "global_batch": B * world, "batch_per_gpu": B, "n_gpus": world, "ms_per_step": ms, "trajectories_per_s": B * world / ms * 1e3,
           "tflops_per_gpu": B * gf / ms, "frac_of_sustained_bf16": B * gf / ms / pk["bf16"], "allreduce_ms_standalone": ar_ms,
           "allreduce_overlapped": bool(getattr(tr, "overlap_allreduce", False)), "grad_bytes": int(tr.flat_grad.numel()) * 4,
           "steps": args.train_steps, "loss": float(loss[0]), "cuda_graph": True, "prefetch": True}
    del tr, model
    return res


def section_large(ctx, args):
    """SURVEY 8d, second line of cfg 3: the same generation workload with the trainer-default ("large") models -- d_model 384, 12
    layers, 12 heads, ff 1536, maze 32-64-128-128 -- 9.58 GFLOP per trajectory; batch per GPU fixed (weak)."""
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    B = args.large_batch
    kp, il = build_models(dev, large=True)
    cfg = GenerationConfig(T=T, K_min=K_MIN, levels=LEVELS, data_dim=D)
    graph = GenerationGraph(kp, il, B, cfg, device=dev).capture()
    cond = synthetic_cond(B, 4000 + rank, device=dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def step():
        graph.run(cond, torch.randn((B, K_MIN, D), generator=gen, device=dev))

    ms, _ = ctx.timed(step, 3, 2)
    ms /= 3
    gf = (19 * 345.6 + 2798.1 + 211.6) / 1000.0
    pk = peaks()
    res = {"workload": "generation with the trainer-default models (384x12 ff1536, maze 32-64-128-128): Stage-1 DDIM-20 K=8 + interp T=64 + "
                       "Stage-2 one-step, 9.58 GFLOP per trajectory (SURVEY 8d, second line of cfg 3)",
           "batch_per_gpu": B, "global_batch": B * world, "n_gpus": world, "ms_per_step": ms, "trajectories_per_s": B * world / ms * 1e3,
           "tflops_per_gpu": B * gf / ms, "frac_of_sustained_bf16": B * gf / ms / pk["bf16"]}
    del graph, kp, il
    torch.cuda.empty_cache()
    return res


def section_long_horizon(ctx, args):
    """BASELINE.json configs[4]: T=256, K=32, levels=4, causal Stage-2 denoiser (small models), B = 8192 sharded over the ranks."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    B = args.lh_batch // world
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D).to(dev)
    il = InterpLevelCausalDenoiser(data_dim=D, max_levels=4, mask_channels=2).to(dev)
    cfg = GenerationConfig(T=256, K_min=32, levels=4, data_dim=D)
    graph = GenerationGraph(kp, il, B, cfg, device=dev)
    graph.capture()
    cond = synthetic_cond(B, 2000 + rank, device=dev)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)

    def step():
        z_T = torch.randn((B, 32, D), generator=gen, device=dev)
        graph.run(cond, z_T)

    ms, _ = ctx.timed(step, 3, 1)
    ms /= 3
    gf = (19 * 413.1 + 3760.0 / 2 + 16.5) / 1000.0    # Stage-1 L=32 evals + causal Stage-2 (half of dense) + conv
    pk = peaks()
    res = {"workload": "long horizon: T=256 K=32 levels=4, Stage-1 DDIM-20 + interp + causal Stage-2 one-step, small models (BASELINE.json configs[4])",
           "global_batch": B * world, "batch_per_gpu": B, "n_gpus": world, "ms_per_step": ms, "trajectories_per_s": B * world / ms * 1e3,
           "tflops_per_gpu": B * gf / ms, "frac_of_sustained_bf16": B * gf / ms / pk["bf16"]}
    del graph, kp, il
    torch.cuda.empty_cache()
    if rank == 0:
        res["interp"] = section_interp(ctx, 256, 32, 4, 1 << 18, "Interp(x0|M_s): 2^18 trajectories, T=256, D=4, 4 nested levels")
    return res


def run_ours(args):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    ctx = _Ctx()
    dist, world, rank, dev = ctx.dist, ctx.world, ctx.rank, ctx.dev
    from interpolated_diffusion_b200 import _lib as L
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
    sections = set(args.sections.split(",")) if args.sections != "all" else {"interp", "strong", "train", "long_horizon", "large"}
    extras = {}

    interp = None
    if rank == 0 and "interp" in sections:
        # second half of BASELINE's metric ("interp GB/s", configs[1])
        interp = section_interp(ctx, 64, K_MIN, 3, 1 << 20, "Interp(x0|M_s): 2^20 trajectories, T=64, D=4, 3 nested levels (BASELINE.json configs[1])")
        torch.cuda.empty_cache()

    B = args.batch
    cfg = GenerationConfig(T=T, K_min=K_MIN, levels=LEVELS, data_dim=D)
    kp, il = build_models(dev, large=(args.model == "large"))
    graph = GenerationGraph(kp, il, B, cfg, device=dev)
    # count this library's kernel launches in one step (every L.call enqueues exactly one kernel here)
    calls = [0]
    orig_call = L.call

    def counting_call(name, *a):
        calls[0] += 1
        return orig_call(name, *a)

    L.call = counting_call
    graph._body()
    L.call = orig_call
    launches_per_step = calls[0]
    graph.capture()

    host = synthetic_cond(B, 1000 + rank, pin=True)
    host_out = torch.empty((B, T, D), dtype=torch.float32).pin_memory()
    dcond = {k: v.to(dev) for k, v in host.items()}
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def make_steps(graph, B, dcond, host, host_out):
        """(device-resident step, end-to-end step, finish).  At N > 1 both include the all-gather of the samples (every rank ends
        up with all N*B trajectories); it runs on a side stream against the next step's replay and is joined inside the timed
        region.  The end-to-end step copies the conditioning H2D from pinned memory and this rank's samples D2H."""
        go = GatherOverlap(ctx, B, T, D) if world > 1 else None

        def step_device():
            z_T = torch.randn((B, K_MIN, D), generator=gen, device=dev)
            x = graph.run(dcond, z_T)
            if go:
                go.submit(x)
            return x

        def step_e2e():
            z_T = torch.randn((B, K_MIN, D), generator=gen, device=dev)
            x = graph.run(host, z_T)                      # H2D of occ + start_goal from pinned memory inside the step
            if go:
                go.submit(x)
            host_out.copy_(x, non_blocking=True)          # D2H of the samples
            return x

        return step_device, step_e2e, (go.join if go else None)

    step_device, step_e2e, finish = make_steps(graph, B, dcond, host, host_out)
    ms_dev, clocks = ctx.timed(step_device, args.steps, args.warmup, sample_clocks=(rank == 0), finish=finish)
    ms_e2e, _ = ctx.timed(step_e2e, args.steps, max(args.warmup // 2, 1), finish=finish)

    # roofline of the dominant kernel: one instrumented eager step, every library call bracketed by CUDA events on the
    # launching stream (each L.call enqueues exactly one kernel), grouped by entry point
    roof, breakdown = None, None
    if rank == 0:
        evs = []

        def timed_call(name, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig_call(name, *a)
            e1.record()
            evs.append((name, e0, e1, DENSE_FLOPS.get(name, lambda a: 0.0)(a)))
            return r

        L.call = timed_call
        graph._body()
        torch.cuda.synchronize(dev)
        L.call = orig_call
        agg = {}
        for name, e0, e1, fl in evs:
            t = agg.setdefault(name, [0, 0.0, 0.0])
            t[0] += 1
            t[1] += e0.elapsed_time(e1)
            t[2] += fl
        total_ms = sum(t[1] for t in agg.values())
        breakdown = {n: {"launches": t[0], "ms": round(t[1], 3), "tflops": (round(t[2] / (t[1] * 1e-3) / 1e12, 1) if t[2] and t[1] > 0 else None)}
                     for n, t in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        dense = {n: t for n, t in agg.items() if t[2] > 0 and t[1] > 0}
        pk = peaks()
        if dense:
            top = max(dense, key=lambda n: dense[n][1])
            cnt, ms, fl = dense[top]
            achieved = fl / (ms * 1e-3) / 1e12
            all_ms = sum(t[1] for t in dense.values())
            all_fl = sum(t[2] for t in dense.values())
            traffic = None
            tpath = next((p for p in (os.path.join(ROOT, "profiles", f) for f in ("r02_encoder_fused_traffic.json", "r01_encoder_fused_traffic.json"))
                          if os.path.exists(p)), None)
            if top in ("idb200_encoder_fused", "idb200_denoiser_fused") and tpath and B == 65536 and args.model == "small":
                # DRAM bytes (read + write) per launch from the committed ncu --set full captures of the same shapes:
                # 19 Stage-1 launches (L = 8) + 1 Stage-2 launch (L = 64), averaged per launch
                with open(tpath) as fh:
                    tl = json.load(fh)["launches"]
                t8 = tl["L8_M524288"]["dram_read_bytes"] + tl["L8_M524288"]["dram_write_bytes"]
                t64 = tl["L64_M4194304"]["dram_read_bytes"] + tl["L64_M4194304"]["dram_write_bytes"]
                traffic = (19 * t8 + t64) / 20.0
            roof = {"bound": "tensor", "kernel": f"{top} (tcgen05; all {cnt} launches of one step)", "achieved": achieved,
                    "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"], "traffic": traffic, "peak_source": pk["src"] + " sustained",
                    "launches_per_step": cnt, "ms_per_step_in_kernel": ms, "share_of_step": ms / total_ms,
                    "all_dense_kernels": {"achieved": all_fl / (all_ms * 1e-3) / 1e12, "frac": all_fl / (all_ms * 1e-3) / 1e12 / pk["bf16"],
                                          "share_of_step": all_ms / total_ms}}

    per_step = ms_dev / args.steps
    value = world * B / (per_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    def guarded(name, fn):
        """A secondary section must never take the headline line down with it."""
        try:
            extras[name] = fn()
        except Exception as exc:                           # noqa: BLE001 -- reported in the line
            extras[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()

    # strong scaling of the same workload (SURVEY 8d cfg 3: "B = 65 536 per run, sharded B/N per GPU")
    if "strong" in sections:
        def strong():
            GB = args.strong_batch
            if world == 1 and GB == B:
                return {"global_batch": GB, "batch_per_gpu": GB, "ms_per_step": per_step, "trajectories_per_s": value,
                        "e2e_trajectories_per_s": e2e_value, "note": "N = 1: the weak and strong workloads coincide"}
            Bs = GB // world
            g2 = GenerationGraph(kp, il, Bs, cfg, device=dev).capture()
            h2 = synthetic_cond(Bs, 3000 + rank, pin=True)
            ho2 = torch.empty((Bs, T, D), dtype=torch.float32).pin_memory()
            d2 = {k: v.to(dev) for k, v in h2.items()}
            sd, se, fin = make_steps(g2, Bs, d2, h2, ho2)
            m1, _ = ctx.timed(sd, args.steps, args.warmup, finish=fin)
            m2, _ = ctx.timed(se, args.steps, max(args.warmup // 2, 1), finish=fin)
            return {"global_batch": Bs * world, "batch_per_gpu": Bs, "ms_per_step": m1 / args.steps,
                    "trajectories_per_s": Bs * world / (m1 / args.steps * 1e-3), "e2e_trajectories_per_s": Bs * world / (m2 / args.steps * 1e-3),
                    "tiles_per_gpu_stage1": Bs * K_MIN // 128, "tiles_per_gpu_stage2": Bs * T // 128}
        guarded("strong", strong)
    del graph
    torch.cuda.empty_cache()
    if "long_horizon" in sections:
        guarded("long_horizon", lambda: section_long_horizon(ctx, args))
    if "large" in sections:
        guarded("large_model", lambda: section_large(ctx, args))
    if "train" in sections:
        guarded("train", lambda: section_train(ctx, args))

    if rank == 0:
        cpu = cpu_reference_rate(args.cpu_sample, steps=1, warmup=0) if world == 1 and not args.skip_cpu else None
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "generation: Stage-1 DDIM-20 K=8 + interp T=64 + Stage-2 one-step (x0), "
                                   + ("small model 256x8 ff1024" if args.model == "small" else "large model 384x12 ff1536") +
                                   ", clamp_policy=endpoints, D=2 (BASELINE.json configs[2])",
                       "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"trajectory-sharded x{world}" + (" + NCCL all-gather of samples on a side stream (joined inside the timed region)" if world > 1 else ""),
                       "cuda_graph": True, "l2": "per-step activation working set (~20 GB) >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_out.numel() * 4,
                    "includes_all_gather": world > 1},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roof, "kernels": breakdown, "interp": interp,
            "flops_per_traj_gemm": gemm_flops_per_traj(), "tflops_e2e": value * gemm_flops_per_traj() / 1e12,
        }
        line.update(extras)
        if cpu is not None:
            cpu_rate, cpu_dt, cores, kind = cpu
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_sample} trajectories, one pass of the " +
                                              ("unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port") + f" ({cpu_dt:.1f} s)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="trajectories per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="trajectories of the CPU baseline sample")
    ap.add_argument("--ref-batch", type=int, default=512, help="trajectories per step of the reference arm")
    ap.add_argument("--model", default="small", choices=["small", "large"], help="small = BASELINE configs[2] (default); large = trainer-default models (dev)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--sections", default="all", help="comma list of the secondary measurements carried in the line: interp (configs[1]), "
                    "strong (configs[2], global batch sharded), long_horizon (configs[4]), train (configs[3]), large (configs[2] with the "
                    "trainer-default models); 'none' = headline only")
    ap.add_argument("--strong-batch", type=int, default=65536, help="global batch of the strong-scaling run")
    ap.add_argument("--train-batch", type=int, default=4096, help="global batch of the training step (configs[3])")
    ap.add_argument("--train-steps", type=int, default=8)
    ap.add_argument("--lh-batch", type=int, default=8192, help="global batch of the long-horizon run (configs[4])")
    ap.add_argument("--large-batch", type=int, default=16384, help="trajectories per GPU of the large-model generation run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
