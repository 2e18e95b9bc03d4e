"""Host-side orchestration of the denoiser kernels: weight packing (once per parameter version), workspaces
and the per-layer launch sequence.  All arithmetic happens in libidb200 (tcgen05 token GEMMs, mma.sync /
SIMT attention, fused LN+FiLM, fp32 SIMT linears for the per-trajectory terms); PyTorch only owns memory."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Tuple

import torch

from .. import _lib as L

EPI_BF16, EPI_SILU_BF16, EPI_RESID_F32, EPI_F32 = 0, 1, 2, 3
EPI_OUT_F16 = 0x100                                            # with EPI_BF16: the 16-bit output is IEEE half
FILM_F16 = os.environ.get("IDB200_FILM_F16", "1") != "0"        # dev: A/B of the half-precision [scale | shift] table


def _sig(params) -> Tuple:
    return (L.PARAM_EPOCH,) + tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params)


def sgemm(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None, *, act: int = 0,
          accumulate: bool = False) -> torch.Tensor:
    """out[M,N] = act(A[M,K] @ W[N,K]^T + bias) in fp32 (idb200_sgemm).  A may be fp32 or bf16, row-strided."""
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K and A.stride(1) == 1 and W.is_contiguous()
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    L.call("idb200_sgemm", A.data_ptr(), int(A.dtype == torch.bfloat16), A.stride(0), W.data_ptr(), L.ptr(bias), out.data_ptr(),
           out.stride(0), M, N, K, act, int(accumulate), L.stream(A.device))
    return out


def gemm_bf16(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, epilogue: int) -> torch.Tensor:
    M, K = A.shape
    L.call("idb200_gemm_bf16", A.data_ptr(), W.data_ptr(), L.ptr(bias), out.data_ptr(), M, W.shape[0], K, epilogue,
           L.stream(A.device))
    return out


def mlp_fused(a: torch.Tensor, W1: torch.Tensor, b1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """h += W2 . SiLU(W1 . a + b1) + b2 in one kernel (idb200_mlp_fused, d_model = 256)."""
    M, d = a.shape
    L.call("idb200_mlp_fused", a.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), h.data_ptr(), M, d,
           W1.shape[0], L.stream(a.device))
    return h


def attn_block(h: torch.Tensor, ln_w, ln_b, gb: Optional[torch.Tensor], wqkv_g, bqkv_g, wo, bo, Lseq: int, H: int, causal: bool):
    """h += out_proj(MHA(LN(h) * (1 + gamma) + beta)) in one kernel (idb200_attn_block; d_model = 256, L | 128)."""
    M, d = h.shape
    L.call("idb200_attn_block", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0),
           wqkv_g.data_ptr(), bqkv_g.data_ptr(), wo.data_ptr(), bo.data_ptr(), M, Lseq, d, H, int(causal), L.stream(h.device))
    return h


def mlp_block(h: torch.Tensor, ln_w, ln_b, gb: Optional[torch.Tensor], W1, b1, W2, b2, Lseq: int):
    """h += ff.2(SiLU(ff.0(LN(h) * (1 + gamma) + beta))) in one kernel (idb200_mlp_block; d_model = 256)."""
    M, d = h.shape
    L.call("idb200_mlp_block", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0),
           W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, Lseq, d, W1.shape[0], L.stream(h.device))
    return h


class Film:
    """FiLM parameters of every LayerNorm of an encoder for a batch of trajectories: ``t`` [B, 2 * n_layers, 2d] fp32.
    folded = False: rows are [gamma | beta] (transformer.py:37,43: a = LN(h) * (1 + gamma) + beta).
    folded = True : rows are [scale | shift] with the LayerNorm affine folded in (the form idb200_encoder_fused reads)."""

    def __init__(self, t: torch.Tensor, folded: bool, ln_major: bool = False):
        self.t, self.folded, self.ln_major = t, folded, ln_major        # ln_major: t is [2 * n_layers, B, 2d]

    def code(self) -> int:
        """film_folded argument of idb200_encoder_fused: 0 raw, 1 folded fp32, 2 folded IEEE half."""
        return 0 if not self.folded else (2 if self.t.dtype == torch.float16 else 1)

    def strides(self):
        """(table, floats between trajectories, floats between LayerNorm slots) as idb200_encoder_fused takes them."""
        t = self.t
        if t.stride(2) != 1:
            t = self.t = t.contiguous()
        return (t, t.stride(1), t.stride(0)) if self.ln_major else (t, t.stride(0), t.stride(1))


def encoder_fused(h: torch.Tensor, pk: "PackedEncoder", film: Optional["Film"], Lseq: int, causal: bool):
    """Every layer of the encoder in one persistent kernel (idb200_encoder_fused): h stays in tensor memory."""
    M, d = h.shape
    f = pk.fused
    ft, fts, ftl = (None, 0, 0) if film is None else film.strides()
    L.call("idb200_encoder_fused", h.data_ptr(), f["params"].data_ptr(), f["bias_last"].data_ptr(), L.ptr(ft),
           fts, ftl, 0 if film is None else film.code(), f["wqkv"].data_ptr(), f["wo"].data_ptr(),
           f["w1"].data_ptr(), f["w2"].data_ptr(), M, Lseq, d, pk.n_heads, pk.ff, len(pk.layers), int(causal), L.stream(h.device))
    return h


def encoder_fused_head(h: torch.Tensor, pk: "PackedEncoder", film: Optional["Film"], Lseq: int, causal: bool, W_out, b_out, y):
    """Every encoder layer + the out head in one launch (idb200_denoiser_fused with embed = NULL): h is read once and never
    written back; y [M, D] = out(h_final)."""
    M, d = h.shape
    f = pk.fused
    ft, fts, ftl = (None, 0, 0) if film is None else film.strides()
    head = L.HeadDesc(W_out.data_ptr(), b_out.data_ptr(), y.data_ptr(), W_out.shape[0])
    L.call("idb200_denoiser_fused", None, ctypes.byref(head), h.data_ptr(), f["params"].data_ptr(), f["bias_last"].data_ptr(),
           L.ptr(ft), fts, ftl, 0 if film is None else film.code(), f["wqkv"].data_ptr(), f["wo"].data_ptr(),
           f["w1"].data_ptr(), f["w2"].data_ptr(), M, Lseq, d, pk.n_heads, pk.ff, len(pk.layers), int(causal), L.stream(y.device))
    return y


def denoiser_fused(pk: "PackedEncoder", film: Optional["Film"], Lseq: int, causal: bool, M: int, src0, src1, src2, Wf, tab, tab_idx,
                   row_a, row_b, W_out, b_out, y):
    """embed_tokens + every encoder layer + out head in ONE launch (idb200_denoiser_fused): the residual stream exists
    only in tensor memory.  Arguments as embed_tokens / out_head; y [M, D] is written."""
    f = pk.fused
    d = pk.d
    ft, fts, ftl = (None, 0, 0) if film is None else film.strides()
    emb = L.EmbedDesc(src0.data_ptr(), src0.shape[-1], L.ptr(src1), 0 if src1 is None else src1.shape[-1], L.ptr(src2),
                      0 if src2 is None else src2.shape[-1], Wf.data_ptr(), tab.data_ptr(), L.ptr(tab_idx), row_a.data_ptr(),
                      0 if row_a.shape[0] == 1 else row_a.stride(0), row_b.data_ptr(), int(tab.shape[0]))
    head = L.HeadDesc(W_out.data_ptr(), b_out.data_ptr(), y.data_ptr(), W_out.shape[0])
    L.call("idb200_denoiser_fused", ctypes.byref(emb), ctypes.byref(head), None, f["params"].data_ptr(), f["bias_last"].data_ptr(),
           L.ptr(ft), fts, ftl, 0 if film is None else film.code(), f["wqkv"].data_ptr(), f["wo"].data_ptr(),
           f["w1"].data_ptr(), f["w2"].data_ptr(), M, Lseq, d, pk.n_heads, pk.ff, len(pk.layers), int(causal), L.stream(y.device))
    return y


def qkv_attention(a: torch.Tensor, wqkv_g: torch.Tensor, bqkv_g: torch.Tensor, out: torch.Tensor, Lseq: int, H: int, causal: bool) -> torch.Tensor:
    """out[M, d] = MHA(a) without the out_proj: in_proj + attention in one kernel (idb200_qkv_attention); out may alias a."""
    M, d = a.shape
    L.call("idb200_qkv_attention", a.data_ptr(), wqkv_g.data_ptr(), bqkv_g.data_ptr(), out.data_ptr(), M, Lseq, d, H, int(causal), L.stream(a.device))
    return out


def ln_qkv_attention(h: torch.Tensor, ln_w, ln_b, gb: Optional[torch.Tensor], wqkv_g: torch.Tensor, bqkv_g: torch.Tensor, out: torch.Tensor,
                     Lseq: int, H: int, causal: bool) -> torch.Tensor:
    """out[M, d] = MHA(LayerNorm(h) * (1 + gamma) + beta) without the out_proj, one kernel (idb200_ln_qkv_attention); gb: raw FiLM rows
    [B, 2d] (any row stride that is a multiple of 4 floats) or None."""
    M, d = h.shape
    L.call("idb200_ln_qkv_attention", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0),
           wqkv_g.data_ptr(), bqkv_g.data_ptr(), out.data_ptr(), M, Lseq, d, H, int(causal), L.stream(h.device))
    return out


def mlp_pair_w2_order(d: int, device) -> torch.Tensor:
    """Row order of ff.2.weight for idb200_mlp_pair (each CTA of a pair holds the rows its half of the pair MMAs produces)."""
    buf = (ctypes.c_int * d)()
    rc = L.lib().idb200_mlp_pair_w2_order(d, buf)
    if rc:
        raise RuntimeError(f"idb200_mlp_pair_w2_order failed ({rc})")
    return torch.tensor(list(buf), dtype=torch.long, device=device)


def mlp_pair(a: torch.Tensor, W1: torch.Tensor, b1: torch.Tensor, W2_packed: torch.Tensor, b2: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """h += ff.2(SiLU(ff.0(a))) in one pair-mode kernel (d_model 384 / 256); W2_packed = ff.2.weight[mlp_pair_w2_order]."""
    M, d = a.shape
    L.call("idb200_mlp_pair", a.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2_packed.data_ptr(), b2.data_ptr(), h.data_ptr(), M, d, W1.shape[0],
           L.stream(a.device))
    return h


def ln_mlp_pair(h: torch.Tensor, ln_w, ln_b, gb: Optional[torch.Tensor], W1, b1, W2_packed, b2, Lseq: int) -> torch.Tensor:
    """h += ff.2(SiLU(ff.0(LayerNorm(h) * (1 + gamma) + beta))) in one pair-mode kernel (idb200_ln_mlp_pair)."""
    M, d = h.shape
    L.call("idb200_ln_mlp_pair", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0), Lseq,
           W1.data_ptr(), b1.data_ptr(), W2_packed.data_ptr(), b2.data_ptr(), M, d, W1.shape[0], L.stream(h.device))
    return h


def sinusoid(rows: int, dim: int, device, args: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty((rows, dim), device=device, dtype=torch.float32)
    L.call("idb200_sinusoid", L.ptr(args), rows, dim, 0 if args is None else 1, out.data_ptr(), L.stream(out.device))
    return out


def ln_film(h: torch.Tensor, ln_w, ln_b, gb: Optional[torch.Tensor], out: torch.Tensor, Lseq: int, h_copy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = LN(h) * (1 + gamma) + beta; ``h_copy`` (fp32 [M, d]) also receives the rows of h that were read (training forward)."""
    M, d = h.shape
    if h_copy is not None:
        L.call("idb200_ln_film_save", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0),
               out.data_ptr(), int(out.dtype == torch.bfloat16), h_copy.data_ptr(), M, Lseq, d, L.stream(h.device))
        return out
    L.call("idb200_ln_film", h.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), L.ptr(gb), 0 if gb is None else gb.stride(0),
           out.data_ptr(), int(out.dtype == torch.bfloat16), M, Lseq, d, L.stream(h.device))
    return out


def attention(qkv: torch.Tensor, out: torch.Tensor, B: int, Lseq: int, H: int, causal: bool, force_simt=False):
    """softmax(Q K^T / sqrt(32) [+ causal mask]) V per (trajectory, head) on packed qkv [B*L, 3d].  bf16 inputs run on the
    tcgen05 path (any L <= 256); fp32 inputs / ``force_simt=True`` on the fp32-arithmetic SIMT kernel (check mode);
    ``force_simt=2`` selects the legacy mma.sync kernel (cross-check)."""
    L.call("idb200_attention", qkv.data_ptr(), out.data_ptr(), int(qkv.dtype == torch.bfloat16), B, Lseq, H, int(causal),
           int(force_simt), L.stream(qkv.device))
    return out


# number of CUDA graphs captured so far in this process (sample.GenerationGraph / train.GraphedStep bump it).  A captured
# graph bakes device addresses in: once one exists, a Workspace never frees a buffer it outgrows (see Workspace.get).
GRAPHS_CAPTURED = 0


def note_graph_captured() -> None:
    global GRAPHS_CAPTURED
    GRAPHS_CAPTURED += 1


class Workspace:
    """Grow-only device buffers keyed by name (stable addresses once sized: CUDA-graph friendly).  When a larger request
    replaces a buffer after a CUDA graph has been captured, the old buffer is retired, not freed: a graph captured at the
    smaller shape keeps replaying into memory that still belongs to this workspace (never into re-allocated memory)."""

    def __init__(self):
        self.bufs: Dict[str, torch.Tensor] = {}
        self.retired: List[torch.Tensor] = []

    def get(self, name: str, shape, dtype, device) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < n or buf.dtype != dtype or buf.device != device:
            if buf is not None and GRAPHS_CAPTURED > 0:
                self.retired.append(buf)
            buf = torch.empty((max(n, 1),), dtype=dtype, device=device)
            self.bufs[name] = buf
        return buf[:n].view(*shape)


class PackedEncoder:
    """bf16 (tensor-core) and fp32 (check-mode) views of a TransformerEncoder's weights + the launch sequence of
    transformer.py:35-46 / 73-82."""

    def __init__(self, encoder):
        self.enc = encoder
        self._key = None
        self._plist = None
        self.ws = Workspace()
        self.fuse_mlp = True            # d_model == 256: FF1 + SiLU + FF2 + residual in one kernel
        self.fuse_blocks = True         # d_model == 256, 8 heads, L | 128: two kernels per layer (attn_block, mlp_block)
        self.fuse_encoder = True        # ... and d_ff <= 1024: ONE kernel for all layers (encoder_fused)
        self.fuse_qkv_attn = True       # per-op path, d_model 256 / 384, L | 128: in_proj + attention in one kernel (qkv never in HBM)
        self.fuse_mlp_pair = True       # per-op path, d_model 384: FF1 + SiLU + FF2 + residual in one pair-mode kernel (idb200_mlp_pair)
        self.fuse_ln_mlp = True         # idb200_ln_mlp_pair: the MLP kernel with its LayerNorm + FiLM prologue computed in the kernel
        self.fuse_ln = False            # ... with the LayerNorm + FiLM prologue computed in the kernel (idb200_ln_qkv_attention).  Correct, but
        #                                 measured slower: the kernel is bound by its compute warps and the prologue lands on their critical
        #                                 path (large-model generation: 75.5 ms against 48.1 + 22.6 ms for the two launches)

    def _pack(self):
        layers = self.enc.layers
        if self._plist is None:
            # the Parameter objects of a module tree are stable (load_state_dict / .to() / the flat optimiser arena mutate them in
            # place; _sig sees that through data_ptr / _version): walk the tree once, not on every forward (the walk was 200 ms of a
            # 280 ms chunked generation)
            self._plist = [p for l in layers for p in l.parameters()]
        key = _sig(self._plist)
        if key == self._key:
            return
        self.layers = []
        film_w, film_b = [], []
        for l in layers:
            d = l.norm1.weight.shape[0]
            e = {
                "wqkv32": l.attn.in_proj_weight.detach().float().contiguous(), "bqkv": l.attn.in_proj_bias.detach().float().contiguous(),
                "wo32": l.attn.out_proj.weight.detach().float().contiguous(), "bo": l.attn.out_proj.bias.detach().float().contiguous(),
                "w132": l.ff[0].weight.detach().float().contiguous(), "b1": l.ff[0].bias.detach().float().contiguous(),
                "w232": l.ff[2].weight.detach().float().contiguous(), "b2": l.ff[2].bias.detach().float().contiguous(),
                "n1w": l.norm1.weight.detach().float().contiguous(), "n1b": l.norm1.bias.detach().float().contiguous(),
                "n2w": l.norm2.weight.detach().float().contiguous(), "n2b": l.norm2.bias.detach().float().contiguous(),
            }
            for k in ("wqkv", "wo", "w1", "w2"):
                e[k] = e[k + "32"].to(torch.bfloat16).contiguous()
            if d in (256, 384) and e["w1"].shape[0] % 128 == 0 and 128 <= e["w1"].shape[0] <= 2048:
                # ff.2 rows as idb200_mlp_pair stages them (the identity at d = 256)
                e["w2p"] = e["w2"] if d == 256 else e["w2"][mlp_pair_w2_order(d, e["w2"].device)].contiguous()
            if d % 64 == 0:
                # head-group-major in_proj for the fused attention block: group g = [Wq[64g:64g+64]; Wk[..]; Wv[..]]
                order = torch.cat([torch.arange(64) + part * d + g * 64 for g in range(d // 64) for part in range(3)]).to(e["wqkv"].device)
                e["wqkv_g"] = e["wqkv"][order].contiguous()
                e["bqkv_g"] = e["bqkv"][order].contiguous()
            self.layers.append(e)
            if l.film1 is not None:
                film_w += [l.film1.weight.detach().float(), l.film2.weight.detach().float()]
                film_b += [l.film1.bias.detach().float(), l.film2.bias.detach().float()]
        self.has_film = len(film_w) > 0
        if self.has_film:
            self.film_w = torch.cat(film_w, dim=0).contiguous()      # [n_layers * 2 * 2d, d_cond]
            self.film_b = torch.cat(film_b, dim=0).contiguous()
            # LayerNorm affine folded into the FiLM linear (both are linear in cond_vec):
            #   scale = ln_w * (1 + gamma) = (ln_w * W_g) c + ln_w * (1 + b_g)
            #   shift = ln_b * (1 + gamma) + beta = (ln_b * W_g + W_b) c + ln_b * (1 + b_g) + b_b
            d_ = layers[0].norm1.weight.shape[0]
            fw, fb = [], []
            for l, e in zip(layers, self.layers):
                for film, nw, nb in ((l.film1, e["n1w"], e["n1b"]), (l.film2, e["n2w"], e["n2b"])):
                    W, b = film.weight.detach().float(), film.bias.detach().float()
                    Wg, Wb, bg, bb = W[:d_], W[d_:], b[:d_], b[d_:]
                    fw += [nw[:, None] * Wg, nb[:, None] * Wg + Wb]
                    fb += [nw * (1.0 + bg), nb * (1.0 + bg) + bb]
            self.film_w_folded = torch.cat(fw, dim=0).contiguous()
            self.film_b_folded = torch.cat(fb, dim=0).contiguous()
            self.film_w16 = self.film_w.to(torch.bfloat16).contiguous()
            self.film_w_folded16 = self.film_w_folded.to(torch.bfloat16).contiguous()
        self.d = layers[0].norm1.weight.shape[0]
        self.ff = layers[0].ff[0].weight.shape[0]
        self.n_heads = layers[0].attn.num_heads
        self.fused = None
        if self.d == 256 and self.n_heads == 8 and self.ff % 128 == 0 and self.ff <= 1024:
            # stacked weights + per-layer parameter blobs of idb200_encoder_fused (layout: include/idb200.h)
            pend = torch.zeros(self.d, device=self.layers[0]["bo"].device, dtype=torch.float32)
            blobs = []
            d_ = self.d
            for e in self.layers:
                # the kernel adds the in_proj bias to q only (include/idb200.h): the k bias is softmax-invariant, the v bias passes
                # through the attention (rows of P sum to 1) and is folded, through out_proj, into the pending bias LayerNorm 2 adds
                bo_eff = e["bo"] + e["wo32"] @ e["bqkv"][2 * d_:]
                blobs += [e["n1w"], e["n1b"], pend, e["bqkv_g"], e["n2w"], e["n2b"], bo_eff, 0.5 * e["b1"]]
                pend = e["b2"]
            self.fused = {
                "params": torch.cat(blobs).contiguous(), "bias_last": pend.contiguous(),
                "wqkv": torch.cat([e["wqkv_g"] for e in self.layers], dim=0).contiguous(),
                "wo": torch.cat([e["wo"] for e in self.layers], dim=0).contiguous(),
                "w1": torch.cat([e["w1"] for e in self.layers], dim=0).contiguous(),
                "w2": torch.cat([e["w2"] for e in self.layers], dim=0).contiguous(),
            }
        self._key = key

    def fused_path(self, Lseq: int, precision: str = "bf16") -> bool:
        """True when forward() will run the whole encoder as one kernel (idb200_encoder_fused)."""
        self._pack()
        return (precision == "bf16" and self.fuse_encoder and self.fuse_blocks and self.fuse_mlp and self.fused is not None
                and 1 <= Lseq <= 128 and 128 % Lseq == 0)

    def film_params(self, cond_vec: torch.Tensor, Lseq: Optional[int] = None, precision: str = "bf16") -> Optional[Film]:
        """FiLM parameters of all layers in one GEMM: [B, n_layers*2, 2d] (loop-invariant across DDIM steps).  With Lseq
        given and the fused-encoder path selected, the folded [scale | shift] form is produced (same GEMM, folded weights).
        bf16 mode runs the GEMM on the tensor cores (bf16 operands, fp32 accumulate and output), fp32 mode on CUDA cores."""
        self._pack()
        if not self.has_film or cond_vec is None:
            return None
        folded = Lseq is not None and self.fused_path(Lseq, precision)
        B = cond_vec.shape[0]
        nln, d2 = len(self.layers) * 2, 2 * self.d
        tc = precision == "bf16" and cond_vec.shape[1] % 64 == 0
        if folded:
            # LayerNorm-major table [2 * n_layers, B, 2d]: the rows of a 128-token tile are contiguous for every LayerNorm
            # half-precision table (tensor-core mode, trajectories of >= 8 tokens: the kernel stages the rows in shared memory): the
            # LayerNorm of the whole-encoder kernel is bound by shared-memory return bandwidth, and half the bytes per column is ~1 k
            # cycles per LayerNorm.  IEEE half, not bf16: its 2^-12 rounding of a scale near 1 is invisible next to the bf16 rounding
            # of the LayerNorm output (a bf16 table moved the outputs by up to 1.2e-2)
            t16 = tc and Lseq >= 8 and FILM_F16
            out = torch.empty((nln, B, d2), device=cond_vec.device, dtype=torch.float16 if t16 else torch.float32)
            a16 = cond_vec.to(torch.bfloat16).contiguous() if tc else None
            for j in range(nln):
                if tc:
                    gemm_bf16(a16, self.film_w_folded16[j * d2:(j + 1) * d2], self.film_b_folded[j * d2:(j + 1) * d2], out[j],
                              (EPI_BF16 | EPI_OUT_F16) if t16 else EPI_F32)
                else:
                    sgemm(cond_vec, self.film_w_folded[j * d2:(j + 1) * d2], self.film_b_folded[j * d2:(j + 1) * d2], out[j])
            return Film(out, True, ln_major=True)
        if tc:
            out = torch.empty((B, self.film_w.shape[0]), device=cond_vec.device, dtype=torch.float32)
            gemm_bf16(cond_vec.to(torch.bfloat16).contiguous(), self.film_w16, self.film_b, out, EPI_F32)
        else:
            out = sgemm(cond_vec, self.film_w, self.film_b)
        return Film(out.view(B, nln, d2), False)

    def forward(self, h: torch.Tensor, B: int, Lseq: int, film, precision: str = "bf16") -> torch.Tensor:
        """In-place on the fp32 residual stream h [B*L, d].  film: Film, a raw [gamma | beta] tensor, or None."""
        self._pack()
        if isinstance(film, torch.Tensor):
            film = Film(film, False)
        if film is not None and film.folded and not self.fused_path(Lseq, precision):
            raise ValueError("folded FiLM parameters can only be consumed by the fused-encoder path")
        M, d = h.shape
        dev = h.device
        H, ff = self.n_heads, self.ff
        causal = bool(self.enc.causal)
        if precision == "bf16":
            fuse_mlp = self.fuse_mlp and d == 256 and ff % 128 == 0 and ff <= 2048
            fuse_attn = self.fuse_blocks and d == 256 and H == 8 and 128 % Lseq == 0 and M % Lseq == 0
            fuse_ln_mlp = self.fuse_blocks and fuse_mlp and (Lseq % 8 == 0 or 8 % Lseq == 0)
            fused = fuse_attn and fuse_ln_mlp
            if fused and self.fuse_encoder and self.fused is not None:
                return encoder_fused(h, self, film, Lseq, causal)
            film = None if film is None else film.t
            qkv_attn = (not fused and self.fuse_qkv_attn and d in (256, 384) and H * 32 == d and 128 % Lseq == 0 and M % Lseq == 0
                        and "wqkv_g" in self.layers[0])
            if not fused:
                a = self.ws.get("a", (M, d), torch.bfloat16, dev)
                qkv = None if qkv_attn else self.ws.get("qkv", (M, 3 * d), torch.bfloat16, dev)
            mlp_pair_ok = self.fuse_mlp_pair and "w2p" in self.layers[0]      # (preferred over the single-CTA mlp_fused at d = 256: 0.62 vs 0.83 ms at M = 512 k)
            f = None if (fuse_mlp or mlp_pair_ok) else self.ws.get("f", (M, ff), torch.bfloat16, dev)
            for i, e in enumerate(self.layers):
                g1 = film[:, 2 * i] if film is not None else None
                g2 = film[:, 2 * i + 1] if film is not None else None
                if fused:
                    attn_block(h, e["n1w"], e["n1b"], g1, e["wqkv_g"], e["bqkv_g"], e["wo"], e["bo"], Lseq, H, causal)
                    mlp_block(h, e["n2w"], e["n2b"], g2, e["w1"], e["b1"], e["w2"], e["b2"], Lseq)
                    continue
                if qkv_attn and self.fuse_ln:
                    ln_qkv_attention(h, e["n1w"], e["n1b"], g1, e["wqkv_g"], e["bqkv_g"], a, Lseq, H, causal)   # LN + FiLM + in_proj + attention
                elif qkv_attn:
                    ln_film(h, e["n1w"], e["n1b"], g1, a, Lseq)
                    qkv_attention(a, e["wqkv_g"], e["bqkv_g"], a, Lseq, H, causal)   # in place: a tile's rows are read before they are written
                else:
                    ln_film(h, e["n1w"], e["n1b"], g1, a, Lseq)
                    gemm_bf16(a, e["wqkv"], e["bqkv"], qkv, EPI_BF16)
                    attention(qkv, a, B, Lseq, H, causal)                # `a` is free again: reuse as attention output
                gemm_bf16(a, e["wo"], e["bo"], h, EPI_RESID_F32)
                if mlp_pair_ok and self.fuse_ln_mlp and M % Lseq == 0:
                    ln_mlp_pair(h, e["n2w"], e["n2b"], g2, e["w1"], e["b1"], e["w2p"], e["b2"], Lseq)      # LN + FiLM + FF1 + SiLU + FF2 + residual
                    continue
                ln_film(h, e["n2w"], e["n2b"], g2, a, Lseq)
                if mlp_pair_ok:
                    mlp_pair(a, e["w1"], e["b1"], e["w2p"], e["b2"], h)
                elif fuse_mlp:
                    mlp_fused(a, e["w1"], e["b1"], e["w2"], e["b2"], h)
                else:
                    gemm_bf16(a, e["w1"], e["b1"], f, EPI_SILU_BF16)
                    gemm_bf16(f, e["w2"], e["b2"], h, EPI_RESID_F32)
        elif precision == "fp32":
            film = None if film is None else film.t
            a = self.ws.get("a32", (M, d), torch.float32, dev)
            o = self.ws.get("o32", (M, d), torch.float32, dev)
            qkv = self.ws.get("qkv32", (M, 3 * d), torch.float32, dev)
            f = self.ws.get("f32", (M, ff), torch.float32, dev)
            for i, e in enumerate(self.layers):
                g1 = film[:, 2 * i] if film is not None else None
                g2 = film[:, 2 * i + 1] if film is not None else None
                ln_film(h, e["n1w"], e["n1b"], g1, a, Lseq)
                sgemm(a, e["wqkv32"], e["bqkv"], qkv)
                attention(qkv, o, B, Lseq, H, causal)
                sgemm(o, e["wo32"], e["bo"], h, accumulate=True)
                ln_film(h, e["n2w"], e["n2b"], g2, a, Lseq)
                sgemm(a, e["w132"], e["b1"], f, act=1)
                sgemm(f, e["w232"], e["b2"], h, accumulate=True)
        else:
            raise ValueError(f"unknown precision {precision!r} (use 'bf16' or 'fp32')")
        return h


_CONV_WS = Workspace()


def conv_encoder(occ: torch.Tensor, sdf: Optional[torch.Tensor], weights: List[torch.Tensor], biases: List[torch.Tensor]) -> torch.Tensor:
    B, _, Hh, Ww = occ.shape
    n = len(weights)
    chans = [weights[0].shape[1]] + [w.shape[0] for w in weights]
    ch = (ctypes.c_int * (n + 1))(*chans)
    wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in weights])
    bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in biases])
    pooled = torch.empty((B, chans[-1]), device=occ.device, dtype=torch.float32)
    cmid = max(chans[1:-1]) if n > 1 else 0
    scratch = _CONV_WS.get(f"conv{occ.device.index}", (2, B, max(cmid, 1), Hh, Ww), torch.float32, occ.device) if n > 1 else None
    L.call("idb200_conv_encoder", occ.data_ptr(), L.ptr(sdf), B, Hh, Ww, n, ch, wp, bp, L.ptr(scratch), pooled.data_ptr(),
           L.stream(occ.device))
    return pooled


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def conv_weight_matrix(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, 3, 3] -> bf16 [C_out, pad64(9 C_in)], k = (ky * 3 + kx) * C_in + c (zero K padding)."""
    co, ci = w.shape[0], w.shape[1]
    wm = torch.zeros((co, _pad64(9 * ci)), device=w.device, dtype=torch.bfloat16)
    wm[:, :9 * ci] = w.detach().float().permute(0, 2, 3, 1).reshape(co, 9 * ci).to(torch.bfloat16)
    return wm


def im2col3x3(src: torch.Tensor, B: int, Hh: int, Ww: int, C: int, act: bool, out: torch.Tensor) -> torch.Tensor:
    L.call("idb200_im2col3x3", src.data_ptr(), B, Hh, Ww, C, out.shape[1], int(act), out.data_ptr(), L.stream(src.device))
    return out


def conv_stack_gemm(x: torch.Tensor, weights: List[torch.Tensor], biases: List[torch.Tensor], ws: "Workspace", keep: bool = False,
                    chunk: int = 4096):
    """MazeEncoder conv stack (encoders.py:15-24) of any depth as im2col + tcgen05 GEMM on NHWC bf16 PRE-activations (SiLU is
    applied when the next layer gathers its patches / by the pooling kernel).  x fp32 [B, C0, H, W] -> pooled fp32 [B, C_last].
    keep=True also returns (x0 NHWC bf16, [u_l], weight matrices, [patch matrix of layer l]) for the backward (training batches fit one chunk); otherwise the batch is
    processed in chunks so the patch matrix stays bounded (4096 trajectories x 441 x 9 C x 2 B)."""
    B, C0, Hh, Ww = x.shape
    dev = x.device
    P = Hh * Ww
    wms = [conv_weight_matrix(w) for w in weights]
    x0 = x.permute(0, 2, 3, 1).reshape(B, P, C0).to(torch.bfloat16).contiguous()
    C_last = weights[-1].shape[0]
    pooled = torch.empty((B, C_last), device=dev, dtype=torch.float32)
    if keep:
        chunk = B
    us_all = []
    cols = []
    for lo in range(0, B, chunk):
        n = min(chunk, B - lo)
        src, C = x0[lo:lo + n], C0
        us = []
        for li, (wm, b) in enumerate(zip(wms, biases)):
            # keep: one patch matrix per layer, handed to the backward (its weight-gradient GEMM reads the same patches: 4 of the
            # 11 im2col launches of a training step were rebuilding them)
            col = ws.get(f"col_keep{li}" if keep else "col", (n * P, wm.shape[1]), torch.bfloat16, dev)
            im2col3x3(src, n, Hh, Ww, C, li > 0, col)
            if keep:
                cols.append(col)
            co = wm.shape[0]
            u = torch.empty((n * P, co), device=dev, dtype=torch.bfloat16) if keep else ws.get(f"u{li % 2}", (n * P, co), torch.bfloat16, dev)
            gemm_bf16(col, wm, b, u, EPI_BF16)
            us.append(u)
            src, C = u, co
        L.call("idb200_pool_silu", src.data_ptr(), n, P, C, pooled[lo:lo + n].data_ptr(), L.stream(dev))
        us_all = us
    if keep:
        return pooled, x0, us_all, wms, cols
    return pooled


def conv_tap_weight(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, 3, 3] -> bf16 [C_out, nkb * 64] in the k-block order of idb200_conv3x3_gemm (csrc/gemm.cu): C_in % 64 == 0:
    k = (ky * 3 + kx) * C_in + c; C_in == 32: k-block (ky, j) = [tap (ky, 2j) | tap (ky, 2j + 1)] with zeros for the missing tap."""
    co, ci = w.shape[0], w.shape[1]
    wt = w.detach().float().permute(0, 2, 3, 1)                          # [co, ky, kx, ci]
    if ci == 32:
        wm = torch.zeros((co, 3, 4, 32), device=w.device, dtype=torch.float32)
        wm[:, :, :3] = wt
        return wm.reshape(co, 384).to(torch.bfloat16).contiguous()
    return wt.reshape(co, 9 * ci).to(torch.bfloat16).contiguous()


def conv_implicit_supported(convs) -> bool:
    """Stacks the tap-shifted implicit GEMM takes: first layer C_in <= 2 -> 32 or 64 channels (CUDA cores), every further layer
    C_in in {32} or a multiple of 64 (<= 256) and C_out a multiple of 64."""
    if len(convs) < 2 or any(tuple(c.weight.shape[2:]) != (3, 3) for c in convs):
        return False
    if convs[0].weight.shape[1] > 2 or convs[0].weight.shape[0] not in (32, 64):
        return False
    for c in convs[1:]:
        ci, co = c.weight.shape[1], c.weight.shape[0]
        if not (ci == 32 or (ci % 64 == 0 and ci <= 256)) or co % 64 != 0:
            return False
    return True


def conv_stack_implicit(x: torch.Tensor, convs, ws: "Workspace", chunk: int = 8192) -> torch.Tensor:
    """MazeEncoder conv stack (encoders.py:15-24) + mean pool without an im2col matrix: zero-bordered NHWC bf16 activations
    [B, (H+2)(W+2), C], first layer on CUDA cores, every further layer as ONE tap-shifted tcgen05 GEMM (SiLU in the epilogue,
    border positions re-zeroed), pooling over the bordered layout.  x fp32 [B, C0, H, W] -> pooled fp32 [B, C_last]."""
    B, C0, Hh, Ww = x.shape
    dev = x.device
    P2 = (Hh + 2) * (Ww + 2)
    key = _sig([c.weight for c in convs[1:]])
    if ws.bufs.get("_tapw_key") != key:
        ws.bufs["_tapw"] = [conv_tap_weight(c.weight) for c in convs[1:]]
        ws.bufs["_tapw_key"] = key
    tapw = ws.bufs["_tapw"]
    f = lambda t: t.detach().float().contiguous()
    C_last = convs[-1].weight.shape[0]
    pooled = torch.empty((B, C_last), device=dev, dtype=torch.float32)
    st = L.stream(dev)
    x = x.contiguous()
    for lo in range(0, B, chunk):
        n = min(chunk, B - lo)
        c1 = convs[0].weight.shape[0]
        # one extra (zero) row: the overlapping-row view of a 32-channel activation reads pixel r + 1 with pixel r
        a = ws.get("act0", (n * P2 + 1, c1), torch.bfloat16, dev)
        a[n * P2:].zero_()
        L.call("idb200_conv_first_nhwc", x[lo:lo + n].data_ptr(), n, C0, Hh, Ww, f(convs[0].weight).data_ptr(), f(convs[0].bias).data_ptr(),
               c1, a.data_ptr(), st)
        C = c1
        for li, c in enumerate(convs[1:]):
            co = c.weight.shape[0]
            o = ws.get(f"act{1 + li % 2}", (n * P2 + 1, co), torch.bfloat16, dev)
            if co == 32:
                o[n * P2:].zero_()
            L.call("idb200_conv3x3_gemm", a.data_ptr(), C, tapw[li].data_ptr(), f(c.bias).data_ptr(), o.data_ptr(), co, n, Hh, Ww,
                   EPI_SILU_BF16, st)
            a, C = o, co
        L.call("idb200_pool_bordered", a.data_ptr(), n, Hh, Ww, C, pooled[lo:lo + n].data_ptr(), st)
    return pooled


def conv_gemm_supported(convs) -> bool:
    return all(c.weight.shape[0] % 32 == 0 and tuple(c.weight.shape[2:]) == (3, 3) for c in convs)


def conv_encoder_tc(occ: torch.Tensor, sdf: Optional[torch.Tensor], w0, b0, w1_packed, b1) -> torch.Tensor:
    """Two-layer conv stack on tensor cores (idb200_conv_encoder_tc); w1_packed = bf16 [c2, 9*c1], k = tap*c1 + c."""
    B, _, Hh, Ww = occ.shape
    c1, cin = w0.shape[0], w0.shape[1]
    c2 = w1_packed.shape[0]
    pooled = torch.empty((B, c2), device=occ.device, dtype=torch.float32)
    L.call("idb200_conv_encoder_tc", occ.data_ptr(), L.ptr(sdf), B, Hh, Ww, cin, c1, c2, w0.data_ptr(), b0.data_ptr(),
           w1_packed.data_ptr(), b1.data_ptr(), pooled.data_ptr(), L.stream(occ.device))
    return pooled


def conv_encoder_tc5(occ: torch.Tensor, sdf: Optional[torch.Tensor], w0, b0, w1_packed, b1) -> torch.Tensor:
    """Two-layer conv stack with the second conv on tcgen05 (idb200_conv_encoder_tc5; maze_channels = (32, 64))."""
    B, _, Hh, Ww = occ.shape
    c1, cin = w0.shape[0], w0.shape[1]
    c2 = w1_packed.shape[0]
    pooled = torch.empty((B, c2), device=occ.device, dtype=torch.float32)
    L.call("idb200_conv_encoder_tc5", occ.data_ptr(), L.ptr(sdf), B, Hh, Ww, cin, c1, c2, w0.data_ptr(), b0.data_ptr(),
           w1_packed.data_ptr(), b1.data_ptr(), pooled.data_ptr(), L.stream(occ.device))
    return pooled


def conv_tc5_supported(convs, Hh: int, Ww: int) -> bool:
    if len(convs) != 2 or convs[0].weight.shape[0] != 32 or convs[1].weight.shape[0] != 64:
        return False
    pw = (Ww + 2 + 7) & ~7
    return Hh * pw <= 512 and convs[0].weight.shape[1] * (Hh + 2) * (Ww + 2) <= 2 * 26 * 26


def conv_tc_supported(convs, Hh: int, Ww: int) -> bool:
    if len(convs) != 2:
        return False
    c1, c2 = convs[0].weight.shape[0], convs[1].weight.shape[0]
    if c1 % 16 or not (16 <= c1 <= 64) or c2 not in (32, 64):
        return False
    pp = (Hh + 2) * (Ww + 2)
    smem = pp * (c1 + 8) * 2 + c2 * (9 * c1 + 8) * 2 + (convs[0].weight.shape[1] * pp + c1 * 18 + c1 + 5 * c2 + 8) * 4
    return smem <= 220 * 1024


def embed_tokens(src0, src1, src2, Wf, tab, tab_idx, row_a, row_b, h, M, Lseq, d):
    n0 = src0.shape[-1]
    n1 = 0 if src1 is None else src1.shape[-1]
    n2 = 0 if src2 is None else src2.shape[-1]
    stride = 0 if row_a.shape[0] == 1 else row_a.stride(0)
    L.call("idb200_embed_tokens", src0.data_ptr(), n0, L.ptr(src1), n1, L.ptr(src2), n2, Wf.data_ptr(), tab.data_ptr(),
           L.ptr(tab_idx), row_a.data_ptr(), stride, row_b.data_ptr(), h.data_ptr(), M, Lseq, d, L.stream(h.device))
    return h


def out_head(h: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    M, d = h.shape
    L.call("idb200_out_head", h.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), M, d, W.shape[0], L.stream(h.device))
    return y
