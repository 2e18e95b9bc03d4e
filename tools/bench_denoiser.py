"""Dev: time one KeypointDenoiser / InterpLevelDenoiser forward (the launch a generation step repeats) with the loop-invariant pieces hoisted;
IDB200_PROF=1 prints the whole-encoder kernel's in-kernel phase cycles for the denoiser form (token assembly prologue, head epilogue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T, K, D = 64, 8, 2
torch.manual_seed(0)
g = torch.Generator(device="cuda").manual_seed(1)
cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g, device="cuda") < 0.2).float(), "start_goal": torch.rand((B, 4), generator=g, device="cuda")}
kp = KeypointDenoiser(data_dim=D).cuda()
il = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=2).cuda()
idx = torch.sort(torch.rand((B, T), generator=g, device="cuda").argsort(1)[:, :K], dim=1).values
z = torch.randn((B, K, D), generator=g, device="cuda")
km = torch.rand((B, K, D), generator=g, device="cuda") < 0.3
t = torch.full((B,), 500, dtype=torch.long, device="cuda")
xs = torch.rand((B, T, D), generator=g, device="cuda")
mk = torch.rand((B, T, 2), generator=g, device="cuda")
s = torch.full((B,), 2, dtype=torch.long, device="cuda")
def timed(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, m, call, L in (("keypoints", kp, lambda kw: kp(z, t, idx, km, cond, T, **kw), K), ("interp", il, lambda kw: il(xs, s, mk, cond, **kw), T)):
    cv = m.encode_cond(cond)
    pk = m.transformer.packed()
    film = pk.film_params(cv, L, m.precision)
    kw = {"cond_vec": cv, "film": film}
    for fio in (0, 1):
        m.fuse_io = bool(fio)
        ms = timed(lambda: call(kw))
        M = B * L
        fl = 8 * M * (2.0 * 256 * 768 + 2.0 * 256 * 256 + 4.0 * L * 256 + 4.0 * 256 * 1024)
        print(f"{name} fuse_io={fio} L={L} ms={ms:.3f} TF/s={fl/ms/1e9:.0f}", flush=True)
