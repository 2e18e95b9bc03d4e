"""Dev check of the tcgen05 attention path (idb200_attention, bf16, force_simt = 0) against torch fp32 on the same bf16-rounded
inputs, plus timing against the legacy mma.sync path (force_simt = 2).  python tools/check_attention_tc5.py [--time]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ref_attention(qkv, B, L, H, causal):
    d = H * 32
    x = qkv.float().view(B, L, 3, H, 32)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = q @ k.transpose(-1, -2) / 32 ** 0.5
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=s.device), 1), float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    from interpolated_diffusion_b200.models import _engine as E
    from interpolated_diffusion_b200 import _lib as L_
    torch.manual_seed(0)
    worst = 0.0
    cases = [(37, 8, 8), (64, 8, 2), (5, 16, 4), (7, 33, 2), (9, 48, 8), (16, 64, 8), (3, 64, 12), (5, 100, 2), (4, 128, 4),
             (3, 129, 2), (3, 200, 8), (6, 256, 8), (1, 1, 2), (130, 2, 2)]
    for B, L, H in cases:
        for causal in (False, True):
            d = H * 32
            qkv = (torch.randn((B * L, 3 * d), device="cuda") * 1.5).bfloat16()
            out = torch.full((B * L, d), float("nan"), device="cuda", dtype=torch.bfloat16)
            E.attention(qkv, out, B, L, H, causal)
            torch.cuda.synchronize()
            ref = ref_attention(qkv, B, L, H, causal)
            err = (out.float() - ref).abs().max().item()
            worst = max(worst, err)
            print(f"B={B:4d} L={L:3d} H={H:2d} causal={int(causal)} max|err|={err:.4f}" + ("   <-- FAIL" if not err < 2e-2 else ""), flush=True)
    print("worst", worst)
    if a.time:
        for B, L, H, causal in [(65536, 8, 12, False), (8192, 64, 12, False), (65536, 64, 8, False), (8192, 256, 8, True), (2048, 256, 8, False)]:
            d = H * 32
            qkv = torch.randn((B * L, 3 * d), device="cuda").bfloat16()
            out = torch.empty((B * L, d), device="cuda", dtype=torch.bfloat16)
            res = {}
            for name, force in (("tc5", 0), ("mma.sync", 2)):
                for _ in range(2):
                    L_.call("idb200_attention", qkv.data_ptr(), out.data_ptr(), 1, B, L, H, int(causal), force, L_.stream(out.device))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    L_.call("idb200_attention", qkv.data_ptr(), out.data_ptr(), 1, B, L, H, int(causal), force, L_.stream(out.device))
                e1.record()
                torch.cuda.synchronize()
                res[name] = e0.elapsed_time(e1) / 5
            fl = 4.0 * B * L * L * d * (0.5 if causal else 1.0)
            gb = B * L * 4 * d * 2 / 1e9
            print(f"B={B} L={L} H={H} causal={int(causal)}: tc5 {res['tc5']:.3f} ms ({fl / res['tc5'] / 1e9:.0f} TF/s, {gb / res['tc5'] * 1e3:.0f} GB/s)"
                  f"   mma.sync {res['mma.sync']:.3f} ms", flush=True)
    return 0 if worst < 2e-2 else 1


if __name__ == "__main__":
    sys.exit(main())
