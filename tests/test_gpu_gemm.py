"""GPU numerics of the tcgen05/TMA token GEMM (K3a) against a plain PyTorch fp32 reference of the same op
(bf16-rounded operands, fp32 accumulate): the only differences are fp32 summation order and the final
bf16 rounding, so tolerances are tight."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def gemm(A, W, bias, out, epi):
    from interpolated_diffusion_b200 import _lib as L
    M, K = A.shape
    N = W.shape[0]
    L.call("idb200_gemm_bf16", A.data_ptr(), W.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(),
           M, N, K, epi, L.stream(A.device))
    return out


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (1000, 768, 256), (4096, 1024, 256), (333, 256, 1024),
                                   (128, 64, 64), (777, 1152, 384), (2048, 384, 1536), (1, 32, 64), (20000, 96, 128)])
def test_gemm_store_f32_and_bf16(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).bfloat16()
    W = (torch.randn((N, K), generator=g, device="cuda") * (K ** -0.5)).bfloat16()
    bias = torch.randn((N,), generator=g, device="cuda")
    ref = A.float() @ W.float().t() + bias
    out = gemm(A, W, bias, torch.empty((M, N), device="cuda"), 3)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    out_nb = gemm(A, W, None, torch.empty((M, N), device="cuda"), 3)
    assert (out_nb - (ref - bias)).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    ob = gemm(A, W, bias, torch.empty((M, N), device="cuda", dtype=torch.bfloat16), 0)
    assert (ob.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    os_ = gemm(A, W, bias, torch.empty((M, N), device="cuda", dtype=torch.bfloat16), 1)
    assert (os_.float() - torch.nn.functional.silu(ref)).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    h = torch.randn((M, N), generator=g, device="cuda")
    h0 = h.clone()
    gemm(A, W, bias, h, 2)
    assert (h - (h0 + ref)).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_gemm_many_tiles_exact_integers():
    """Integer-valued operands: every partial sum is exact in fp32, so the result must be bit-exact --
    catches any tile / k-block / swizzle mix-up independent of rounding."""
    M, N, K = 148 * 128 * 2 + 77, 512, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randint(-4, 5, (M, K), generator=g, device="cuda").bfloat16()
    W = torch.randint(-4, 5, (N, K), generator=g, device="cuda").bfloat16()
    out = gemm(A, W, None, torch.empty((M, N), device="cuda"), 3)
    ref = A.float() @ W.float().t()
    assert torch.equal(out, ref)


def test_gemm_argument_errors():
    A = torch.zeros((8, 48), device="cuda", dtype=torch.bfloat16)
    W = torch.zeros((32, 48), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        gemm(A, W, None, torch.empty((8, 32), device="cuda"), 3)


@pytest.mark.parametrize("M,ff", [(128, 128), (1000, 1024), (148 * 128 * 2 + 5, 1024), (4096, 256), (333, 2048), (7, 384)])
def test_mlp_fused_vs_unfused_reference(M, ff):
    from interpolated_diffusion_b200.models import _engine as E
    d = 256
    g = torch.Generator(device="cuda").manual_seed(M + ff)
    a = (torch.randn((M, d), generator=g, device="cuda")).bfloat16()
    W1 = (torch.randn((ff, d), generator=g, device="cuda") * d ** -0.5).bfloat16()
    W2 = (torch.randn((d, ff), generator=g, device="cuda") * ff ** -0.5).bfloat16()
    b1 = torch.randn((ff,), generator=g, device="cuda") * 0.1
    b2 = torch.randn((d,), generator=g, device="cuda") * 0.1
    h0 = torch.randn((M, d), generator=g, device="cuda")
    hid = torch.nn.functional.silu(a.float() @ W1.float().t() + b1).bfloat16().float()     # hidden is rounded to bf16
    ref = h0 + hid @ W2.float().t() + b2
    h = h0.clone()
    E.mlp_fused(a, W1, b1, W2, b2, h)
    torch.cuda.synchronize()
    assert torch.isfinite(h).all()
    # differences: tanh.approx SiLU (~5e-4 rel) before the bf16 rounding of the hidden activation
    assert (h - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item() / 4)
    # exact-integer variant: every product and partial sum is exact -> bit-exact (catches swizzle / slot-order bugs)
    ai = torch.randint(-2, 3, (M, d), generator=g, device="cuda").bfloat16()
    W1i = torch.zeros((ff, d), device="cuda")
    W1i[torch.arange(ff), torch.randint(0, d, (ff,), generator=g, device="cuda")] = 64.0     # hidden = 64 * a[:, perm]: SiLU saturates to {0, +-128..}
    W2i = torch.randint(-2, 3, (d, ff), generator=g, device="cuda").bfloat16()
    hz = torch.zeros((M, d), device="cuda")
    E.mlp_fused(ai, W1i.bfloat16(), torch.zeros_like(b1), W2i, torch.zeros_like(b2), hz)
    hid_i = torch.nn.functional.silu(ai.float() @ W1i.t()).bfloat16().float()
    ref_i = hid_i @ W2i.float().t()
    assert (hz - ref_i).abs().max().item() < 1e-3 * max(1.0, ref_i.abs().max().item())
