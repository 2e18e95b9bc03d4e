"""Dev: time idb200_encoder_fused alone (8 layers, small model) and against the two-kernels-per-layer path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models.transformer import TransformerEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
Ls = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 64]
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["encoder", "blocks"]
torch.manual_seed(0)
enc = TransformerEncoder(d_model=256, n_layers=8, n_heads=8, d_ff=1024, cond_dim=128).cuda()
pk = enc.packed()
for L in Ls:
    h = torch.randn((B * L, 256), device="cuda")
    cv = torch.randn((B, 128), device="cuda")
    for mode in modes:
        pk.fuse_encoder = mode == "encoder"
        film = pk.film_params(cv, L) if not os.environ.get("NOFILM") else None
        for _ in range(2): pk.forward(h.clone(), B, L, film)
        hh = h.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n): pk.forward(hh, B, L, film)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        M = B * L
        fl = 8 * M * (2.0 * 256 * 768 + 2.0 * 256 * 256 + 4.0 * L * 256 + 4.0 * 256 * 1024)
        print(f"{mode} L={L} M={M} ms={ms:.3f} TF/s={fl/ms/1e9:.0f} us_per_tile_layer={ms*1e3/((M/128)/148)/8:.2f}", flush=True)
    del h, film
