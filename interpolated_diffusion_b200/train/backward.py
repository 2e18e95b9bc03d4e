"""Training-form forward and hand-written backward of the Stage-2 denoiser on libidb200 kernels: what
``loss.backward()`` computes in ``src/train/train_interp_levels.py:1142-1161`` for ``InterpLevelDenoiser``
(``src/models/denoiser_interp_levels.py:64-84``, ``transformer.py:28-46``, ``encoders.py:8-71``) under bf16 autocast.

Forward keeps what the backward needs (per layer: the residual stream before each LayerNorm in fp32; LN+FiLM outputs, packed
qkv, attention output, MLP pre-activation and activation in bf16).  Backward walks the layers in reverse:

* dense contractions on the tcgen05 GEMM: ``dX = dY W`` (transposed weight as the weight operand; the MLP's SiLU / SiLU' passes ride
  in the ff.0 / dU GEMM epilogues), ``dW = dY^T X`` read in place as MN-major operands -- for the encoder ONE split-K launch per
  weight kind over the token axis stacked across the layers (split s = layer s: no partial sums to reduce);
* everything else in ``csrc/train_bwd.cu``: LayerNorm+FiLM backward (two passes: row scalars, then a thread per column group),
  attention backward (one block per trajectory x head, also the in_proj bias gradient), bias / LayerNorm-affine column sums
  (per-trajectory partials of all layers reduced once per pass), out-head and in_proj gradients, the conv stack as im2col + GEMM
  (dgrad = the same conv with flipped weights).

PyTorch owns memory and reshapes (one-hot level rows, conv weight packing); no arithmetic of the step runs in torch.
Gradients are written into caller-provided fp32 tensors keyed by the reference's parameter names (``grads[name]``)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from .. import _lib as L
from ..models import _engine as E

BF16, F32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------------------------------------------------
# thin wrappers
def transpose_bf16(src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    M, N = src.shape
    L.call("idb200_transpose_bf16", src.data_ptr(), int(src.dtype == F32), M, N, out.data_ptr(), L.stream(src.device))
    return out


EPI_BF16_SILU_DUAL, EPI_BF16_DSILU = 4, 5


def gemm_bf16_aux(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, aux: torch.Tensor, epilogue: int,
                  fused: Optional[bool] = None) -> torch.Tensor:
    """The token GEMM + the MLP's SiLU pass (4: out = u, aux = SiLU(u); 5: out = (A W^T) * SiLU'(aux)) as ONE launch
    (idb200_gemm_bf16_aux), bit-identical to the two-launch form (``fused=False`` or IDB200_TRAIN_FUSED_SILU=0: token GEMM, then
    idb200_silu_bf16 -- kept as the cross-check).  History: the first fused epilogue (exp + IEEE division, two staging passes, the
    u slab fetched and awaited inside each slab) made the cfg-4 step 10 % SLOWER than the separate HBM-bound passes; with one-MUFU
    SiLU / SiLU', both outputs staged from one TMEM read and the u slab requested half a slab ahead, the fused GEMMs take 0.41 /
    0.44 ms at M = 262 144 (plain GEMM 0.29 + pass 0.34) and the B = 4096 step went 94.2 -> 90.4 ms."""
    M, K = A.shape
    if fused is None:
        fused = os.environ.get("IDB200_TRAIN_FUSED_SILU", "1") != "0" and W.shape[0] % 64 == 0      # the aux epilogues stage 64-column slabs
    if not fused:
        E.gemm_bf16(A, W, bias, out, E.EPI_BF16)
        return silu_bf16(out, aux) if epilogue == EPI_BF16_SILU_DUAL else silu_bf16(aux, out, g=out)
    L.call("idb200_gemm_bf16_aux", A.data_ptr(), W.data_ptr(), L.ptr(bias), out.data_ptr(), aux.data_ptr(), M, W.shape[0], K, epilogue,
           L.stream(A.device))
    return out


def silu_bf16(u: torch.Tensor, out: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
    L.call("idb200_silu_bf16", u.data_ptr(), L.ptr(g), u.numel(), 0 if g is None else 1, out.data_ptr(), L.stream(u.device))
    return out


def silu_f32(u: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty_like(u)
    L.call("idb200_silu_f32", u.data_ptr(), L.ptr(g), u.numel(), 0 if g is None else 1, out.data_ptr(), L.stream(u.device))
    return out


def sgemm_strided(A: torch.Tensor, a_t: bool, Bm: torch.Tensor, b_t: bool, out: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    """out[i, j] (+)= sum_k A'[i, k] * B'[j, k] with A' = A^T if a_t, B' = B^T if b_t (A, B contiguous 2-D fp32)."""
    sa = (1, A.shape[1]) if a_t else (A.shape[1], 1)
    sb = (1, Bm.shape[1]) if b_t else (Bm.shape[1], 1)
    M, K = (A.shape[1], A.shape[0]) if a_t else A.shape
    N = Bm.shape[1] if b_t else Bm.shape[0]
    assert (Bm.shape[0] if b_t else Bm.shape[1]) == K and out.shape == (M, N) and out.is_contiguous()
    L.call("idb200_sgemm_strided", A.data_ptr(), sa[0], sa[1], Bm.data_ptr(), sb[0], sb[1], out.data_ptr(), N, M, N, K,
           int(accumulate), L.stream(A.device))
    return out


def multi_copy(pairs) -> None:
    """dst.copy_(src) for a list of (dst, src) fp32 tensor pairs in ONE launch per 96 pairs (idb200_multi_copy_f32): the gradient
    slices a backward pass scatters into the flat arena were ~130 separate 2-3 us copy kernels per step.  Pairs that are not
    contiguous fp32 of equal size fall back to ``copy_``."""
    import ctypes
    fast = []
    for dst, src in pairs:
        if (dst.dtype == F32 and src.dtype == F32 and dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
                and dst.device == src.device and dst.is_cuda):
            fast.append((dst, src))
        else:
            dst.copy_(src)
    if not fast:
        return
    n = len(fast)
    srcs = (ctypes.c_void_p * n)(*[s.data_ptr() for _, s in fast])
    dsts = (ctypes.c_void_p * n)(*[d.data_ptr() for d, _ in fast])
    cnts = (ctypes.c_int64 * n)(*[d.numel() for d, _ in fast])
    L.call("idb200_multi_copy_f32", srcs, dsts, cnts, n, L.stream(fast[0][0].device))


class _Scratch:
    """Workspaces shared by the backward helpers."""

    def __init__(self):
        self.ws = E.Workspace()

    def colsum(self, src: torch.Tensor, out: torch.Tensor, scale: float = 1.0, accumulate: bool = False) -> torch.Tensor:
        M, N = src.shape
        n = L.lib().idb200_colsum_scratch_floats(M, N)
        sc = self.ws.get("colsum", (n,), F32, src.device)
        L.call("idb200_colsum", src.data_ptr(), int(src.dtype == BF16), M, N, sc.data_ptr(), float(scale), int(accumulate),
               out.data_ptr(), L.stream(src.device))
        return out

    def colsum_segments(self, src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[s, :] = column sums of src[s] for src [segs, M, N] (two launches for all segments)."""
        segs, M, N = src.shape
        n = L.lib().idb200_colsum_scratch_floats(M, N) * segs
        sc = self.ws.get("colsum_seg", (n,), F32, src.device)
        L.call("idb200_colsum_segments", src.data_ptr(), int(src.dtype == BF16), segs, M, N, sc.data_ptr(), 1.0, 0, out.data_ptr(),
               L.stream(src.device))
        return out

    def narrow_outer(self, A: torch.Tensor, X: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[n, K] = A[M, n]^T X[M, K] (n <= 8)."""
        M, n = A.shape
        K = X.shape[1]
        cnt = L.lib().idb200_narrow_outer_scratch_floats(M, n, K)
        sc = self.ws.get("narrow", (cnt,), F32, A.device)
        L.call("idb200_narrow_outer", A.data_ptr(), n, X.data_ptr(), M, K, sc.data_ptr(), 0, out.data_ptr(), L.stream(A.device))
        return out

    def _splits(self, n_out: int, k_in: int, m_tok: int, bns) -> int:
        if m_tok % 64 != 0:
            raise ValueError(f"the weight-gradient GEMM reduces over the tokens in blocks of 64 (got {m_tok} rows)")
        kb = m_tok // 64
        bn = next(c for c in bns if k_in % c == 0)
        tiles = ((n_out + 127) // 128) * (k_in // bn)
        want = max(1, min(32, 296 // tiles))
        return max(s for s in range(1, want + 1) if kb % s == 0)

    def dweight(self, dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[N_out, K_in] = dY^T X from dy [M, N_out], x [M, K_in] (bf16, row-major): split-K tcgen05 GEMM over the M tokens
        with both operands read as MN-major UMMA tiles straight from their row-major storage (no transposes); partial
        products reduced in a fixed order."""
        m_tok, n_out = dy.shape
        k_in = x.shape[1]
        assert dy.dtype == BF16 and x.dtype == BF16 and dy.is_contiguous() and x.is_contiguous() and x.shape[0] == m_tok
        splits = self._splits(n_out, k_in, m_tok, (256, 192, 128, 64))
        part = self.ws.get("splitk", (splits, n_out, k_in), F32, dy.device)
        L.call("idb200_gemm_bf16_nn_splitk", dy.data_ptr(), x.data_ptr(), part.data_ptr(), n_out, k_in, m_tok, splits,
               L.stream(dy.device))
        L.call("idb200_reduce_rows", part.data_ptr(), splits, n_out * k_in, 1.0, 0, out.data_ptr(), L.stream(dy.device))
        return out

    def dweight_t(self, dy_t: torch.Tensor, x_t: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Same product from bf16 transposes dy_t [N_out, M], x_t [K_in, M] (K-major operands; kept as a cross-check)."""
        n_out, m_tok = dy_t.shape
        k_in = x_t.shape[0]
        splits = self._splits(n_out, k_in, m_tok, (256, 192, 128, 96, 64, 32))
        part = self.ws.get("splitk", (splits, n_out, k_in), F32, dy_t.device)
        L.call("idb200_gemm_bf16_splitk", dy_t.data_ptr(), x_t.data_ptr(), part.data_ptr(), n_out, k_in, m_tok, splits,
               L.stream(dy_t.device))
        L.call("idb200_reduce_rows", part.data_ptr(), splits, n_out * k_in, 1.0, 0, out.data_ptr(), L.stream(dy_t.device))
        return out


# ----------------------------------------------------------------------------------------------------------------------
class EncoderBackprop:
    """``TransformerEncoder`` (transformer.py:49-82) in training form."""

    def __init__(self, encoder, scratch: Optional[_Scratch] = None):
        self.enc = encoder
        self.sc = scratch or _Scratch()
        self.saved = None

    def _weights(self):
        """Per layer: fp32 views of the parameters + the bf16 operand copies of the four token-GEMM weights (W for the forward, W^T as
        the weight operand of dX = dY W).  The 8 n_layers casts / transposes are ONE launch (idb200_cast_weights_bf16) into buffers
        that persist across steps (stable addresses: CUDA-graph friendly)."""
        import ctypes
        out, mats = [], []
        cache = getattr(self, "_w16", None)
        if cache is None:
            cache = self._w16 = {}
        for li, l in enumerate(self.enc.layers):
            f = lambda t: t.detach().float().contiguous()
            w = {"wqkv": f(l.attn.in_proj_weight), "bqkv": f(l.attn.in_proj_bias), "wo": f(l.attn.out_proj.weight), "bo": f(l.attn.out_proj.bias),
                 "w1": f(l.ff[0].weight), "b1": f(l.ff[0].bias), "w2": f(l.ff[2].weight), "b2": f(l.ff[2].bias),
                 "n1w": f(l.norm1.weight), "n1b": f(l.norm1.bias), "n2w": f(l.norm2.weight), "n2b": f(l.norm2.bias)}
            for k in ("wqkv", "wo", "w1", "w2"):
                src = w[k]
                key = (li, k, tuple(src.shape), src.device)
                bufs = cache.get(key)
                if bufs is None:
                    bufs = cache[key] = (torch.empty(src.shape, device=src.device, dtype=BF16),
                                         torch.empty((src.shape[1], src.shape[0]), device=src.device, dtype=BF16))
                w[k + "16"], w[k + "t16"] = bufs
                mats.append((src, bufs[0], bufs[1]))
            out.append(w)
        n = len(mats)
        if n:
            srcs = (ctypes.c_void_p * n)(*[m[0].data_ptr() for m in mats])
            dsts = (ctypes.c_void_p * n)(*[m[1].data_ptr() for m in mats])
            dsts_t = (ctypes.c_void_p * n)(*[m[2].data_ptr() for m in mats])
            rows = (ctypes.c_int * n)(*[m[0].shape[0] for m in mats])
            cols = (ctypes.c_int * n)(*[m[0].shape[1] for m in mats])
            L.call("idb200_cast_weights_bf16", srcs, dsts, dsts_t, rows, cols, n, L.stream(mats[0][0].device))
        return out

    def forward(self, h: torch.Tensor, B: int, Lseq: int, cond_vec: Optional[torch.Tensor]) -> torch.Tensor:
        """In place on the fp32 residual stream h [B*L, d]; keeps the activations the backward needs."""
        M, d = h.shape
        dev = h.device
        layers = self.enc.layers
        nl, H, ff = len(layers), layers[0].attn.num_heads, layers[0].ff[0].weight.shape[0]
        W = self._weights()
        film = film_w = None
        if layers[0].film1 is not None and cond_vec is not None:
            film_w = torch.cat([m.weight.detach().float() for l in layers for m in (l.film1, l.film2)], dim=0).contiguous()
            film_b = torch.cat([m.bias.detach().float() for l in layers for m in (l.film1, l.film2)], dim=0).contiguous()
            film = torch.empty((B, 2 * nl, 2 * d), device=dev, dtype=F32)
            if cond_vec.shape[1] % 64 == 0:
                E.gemm_bf16(cond_vec.to(BF16).contiguous(), film_w.to(BF16), film_b, film.view(B, -1), E.EPI_F32)
            else:
                E.sgemm(cond_vec, film_w, film_b, film.view(B, -1))
        sv = {"h_in": torch.empty((nl, M, d), device=dev, dtype=F32), "h_mid": torch.empty((nl, M, d), device=dev, dtype=F32),
              "a1": torch.empty((nl, M, d), device=dev, dtype=BF16), "qkv": torch.empty((nl, M, 3 * d), device=dev, dtype=BF16),
              "o": torch.empty((nl, M, d), device=dev, dtype=BF16), "a2": torch.empty((nl, M, d), device=dev, dtype=BF16),
              "u": torch.empty((nl, M, ff), device=dev, dtype=BF16), "f": torch.empty((nl, M, ff), device=dev, dtype=BF16)}
        causal = bool(self.enc.causal)
        for i, w in enumerate(W):
            g1 = film[:, 2 * i] if film is not None else None
            g2 = film[:, 2 * i + 1] if film is not None else None
            E.ln_film(h, w["n1w"], w["n1b"], g1, sv["a1"][i], Lseq, h_copy=sv["h_in"][i])      # also saves the rows it read
            E.gemm_bf16(sv["a1"][i], w["wqkv16"], w["bqkv"], sv["qkv"][i], E.EPI_BF16)
            E.attention(sv["qkv"][i], sv["o"][i], B, Lseq, H, causal)
            E.gemm_bf16(sv["o"][i], w["wo16"], w["bo"], h, E.EPI_RESID_F32)
            E.ln_film(h, w["n2w"], w["n2b"], g2, sv["a2"][i], Lseq, h_copy=sv["h_mid"][i])
            # ff.0 with both outputs from one epilogue: u (pre-activation, for the backward) and f = SiLU(u)
            gemm_bf16_aux(sv["a2"][i], w["w116"], w["b1"], sv["u"][i], sv["f"][i], EPI_BF16_SILU_DUAL)
            E.gemm_bf16(sv["f"][i], w["w216"], w["b2"], h, E.EPI_RESID_F32)
        self.saved = {"sv": sv, "W": W, "film": film, "film_w": film_w, "cond_vec": cond_vec, "B": B, "L": Lseq, "H": H, "ff": ff,
                      "causal": causal}
        return h

    def backward(self, dh: torch.Tensor, dh16: torch.Tensor, grads: Dict[str, torch.Tensor], prefix: str = "transformer.") -> Optional[torch.Tensor]:
        """dh [M, d] fp32 (in/out: gradient w.r.t. the encoder output -> w.r.t. its input), dh16 its bf16 copy (kept in sync).
        Writes the parameter gradients into ``grads`` and returns d loss / d cond_vec [B, d_cond] (None without FiLM).
        = ``backward_layers`` (every per-layer gradient except the FiLM linears: final when it returns) + ``backward_film``."""
        self.backward_layers(dh, dh16, grads, prefix)
        return self.backward_film(grads, prefix)

    def backward_layers(self, dh: torch.Tensor, dh16: torch.Tensor, grads: Dict[str, torch.Tensor], prefix: str = "transformer.") -> None:
        """Every per-layer gradient except the FiLM linears.

        The dX chain (dU GEMM -> da GEMM -> LayerNorm backward -> dO GEMM -> attention backward -> da GEMM -> LayerNorm backward) runs
        layer by layer.  The weight gradients dW = dY^T X are NOT computed per layer: every layer's dY (bf16 dh before the MLP and
        before the attention, dU, dqkv) is kept in a buffer stacked over the layers, like the saved activations X already are, and
        each of the four weight kinds is ONE launch of the split-K GEMM over the stacked token axis [n_layers * M] with splits =
        n_layers: "split" s then covers exactly layer s and its partial product IS that layer's gradient -- no partial sums, no
        reduce_rows, 4 GEMM launches instead of 48 + 48, reduction loops of M tokens instead of M / 16 (which is what made the
        per-layer form inefficient at the cfg-4 per-GPU shape: 32 k-blocks per work unit, 3.2 waves of quantisation per launch).
        ``IDB200_TRAIN_DW_GROUPED=0`` selects the per-layer form (on a side stream with double-buffered operands; kept as the
        cross-check)."""
        S = self.saved
        sv, W, film = S["sv"], S["W"], S["film"]
        B, Lseq, H, ff, causal = S["B"], S["L"], S["H"], S["ff"], S["causal"]
        M, d = dh.shape
        dev = dh.device
        nl = len(W)
        sc, ws = self.sc, self.sc.ws
        st = L.stream(dev)
        grouped = os.environ.get("IDB200_TRAIN_DW_GROUPED", "1") != "0" and M % 64 == 0 and d % 64 == 0 and ff % 64 == 0
        da = ws.get("da16", (M, d), BF16, dev)            # gradient w.r.t. the LayerNorm outputs, bf16 (half the bytes of 3 passes)
        do16 = ws.get("do16", (M, d), BF16, dev)
        # per-trajectory partials of every LayerNorm ([dw | db | sum_t dh]) and of every in_proj bias: reduced over the batch by
        # ONE segmented column sum each at the end (was two launches per LayerNorm / layer: 0.4 ms of a 14.6 ms step at B = 512)
        dwb_all = ws.get("dwb_all", (2 * nl, B, 3 * d), F32, dev)
        dqkv_sum_all = ws.get("dqkv_sum_all", (nl, B, 3 * d), F32, dev)
        fused_qkv_sums = 32 < Lseq <= 64            # the tensor-core attention backward emits the per-trajectory sums
        fused_du_sums = ff % 64 == 0 and d % 64 == 0 and os.environ.get("IDB200_TRAIN_FUSED_SILU", "1") != "0"
        du_part_all = ws.get("du_part", (nl, 8 * ((M + 255) // 256), ff), F32, dev) if fused_du_sums else None   # per-warp column sums of dU
        stats = ws.get("ln_stats", (M, 4), F32, dev)
        dgb = torch.zeros((B, 2 * nl, 2 * d), device=dev, dtype=F32) if film is not None else None

        # ---- where each dY lives
        main = torch.cuda.current_stream(dev)
        side = None
        readers: Dict[int, "torch.cuda.Event"] = {}
        if grouped:
            g_mlp = ws.get("g16_mlp", (nl, M, d), BF16, dev)        # bf16 dh entering layer i's MLP backward   (dY of ff.2)
            g_att = ws.get("g16_att", (nl, M, d), BF16, dev)        # bf16 dh entering layer i's attention backward (dY of out_proj)
            du_all = ws.get("du16_all", (nl, M, ff), BF16, dev)     # dY of ff.0
            dqkv_all = ws.get("dqkv16_all", (nl, M, 3 * d), BF16, dev)
            g_mlp[nl - 1].copy_(dh16)
        else:
            if os.environ.get("IDB200_TRAIN_DW_STREAM", "1") != "0":
                if getattr(self, "_dw_stream", None) is None:
                    self._dw_stream = torch.cuda.Stream(device=dev)
                side = self._dw_stream
            dh16_bufs = [dh16, ws.get("dh16_alt", (M, d), BF16, dev)]      # LayerNorm backward k writes buffer (k + 1) % 2: 2 nl calls end in dh16
        cur16 = [0]

        def dw_async(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor) -> None:
            """Per-layer form: the weight-gradient GEMM on the side stream, ordered after everything enqueued on the main stream."""
            if grouped:
                return
            if side is None:
                sc.dweight(dy, x, out)
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                sc.dweight(dy, x, out)
                ev = torch.cuda.Event()
                ev.record(side)
            readers[dy.data_ptr()] = ev

        def before_write(buf: torch.Tensor) -> None:
            ev = readers.pop(buf.data_ptr(), None)
            if ev is not None:
                main.wait_event(ev)

        def ln_bwd(h_saved, nw, nb, j, out16):
            """LayerNorm + FiLM backward of LayerNorm slot j; its partials (incl. the column sums of the UPDATED dh = the bias gradient
            of the GEMM that accumulated into the residual stream just below this LayerNorm) land in dwb_all[j]; the bf16 copy of the
            updated dh goes to ``out16``."""
            gb = film[:, j] if film is not None else None
            dg = dgb[:, j] if film is not None else None
            before_write(out16)
            L.call("idb200_ln_film_bwd2", da.data_ptr(), 1, h_saved.data_ptr(), nw.data_ptr(), nb.data_ptr(), L.ptr(gb),
                   0 if gb is None else gb.stride(0), B, Lseq, d, dh.data_ptr(), out16.data_ptr(), L.ptr(dg),
                   0 if dg is None else dg.stride(0), dwb_all[j].data_ptr(), 1, stats.data_ptr(), st)

        def next16():
            """Per-layer form: the buffer the next LayerNorm backward writes (ping-pong, or dh16 itself without the side stream)."""
            if side is None:
                return dh16
            cur16[0] = 1 - cur16[0]
            return dh16_bufs[cur16[0]]

        g16 = g_mlp[nl - 1] if grouped else dh16
        for i in range(nl - 1, -1, -1):
            w = W[i]
            p = f"{prefix}layers.{i}."
            # ---- MLP: h_out = h_mid + ff.2(silu(ff.0(a2)))
            dw_async(g16, sv["f"][i], grads[p + "ff.2.weight"])
            if i == nl - 1:
                sc.colsum(dh, grads[p + "ff.2.bias"])           # (the other layers' come out of the LayerNorm backward above them)
            du = du_all[i] if grouped else ws.get(f"du16_{i & 1}" if side is not None else "du16", (M, ff), BF16, dev)
            before_write(du)
            if fused_du_sums:          # du = (dh W2) * silu'(u) and its per-warp column sums (-> ff.0 bias gradient) in one launch
                L.call("idb200_gemm_bf16_dsilu_sums", g16.data_ptr(), w["w2t16"].data_ptr(), du.data_ptr(), sv["u"][i].data_ptr(),
                       du_part_all[i].data_ptr(), M, ff, d, st)                          # (reduced for all layers at the end)
            else:
                gemm_bf16_aux(g16, w["w2t16"], None, du, sv["u"][i], EPI_BF16_DSILU)   # du = (dh W2) * silu'(u)
                sc.colsum(du, grads[p + "ff.0.bias"])
            dw_async(du, sv["a2"][i], grads[p + "ff.0.weight"])
            E.gemm_bf16(du, w["w1t16"], None, da, E.EPI_BF16)                           # da2 = du W1
            g16 = g_att[i] if grouped else next16()
            ln_bwd(sv["h_mid"][i], w["n2w"], w["n2b"], 2 * i + 1, g16)
            # ---- attention: h_mid = h_in + out_proj(MHA(a1))
            dw_async(g16, sv["o"][i], grads[p + "attn.out_proj.weight"])
            E.gemm_bf16(g16, w["wot16"], None, do16, E.EPI_BF16)                        # dO = dh Wo
            dqkv = dqkv_all[i] if grouped else ws.get(f"dqkv16_{i & 1}" if side is not None else "dqkv16", (M, 3 * d), BF16, dev)
            before_write(dqkv)
            if fused_qkv_sums:         # tensor-core kernel: also emits the per-trajectory column sums (= the in_proj bias gradient)
                L.call("idb200_attention_bwd_sums", sv["qkv"][i].data_ptr(), do16.data_ptr(), dqkv.data_ptr(), dqkv_sum_all[i].data_ptr(), B, Lseq,
                       H, int(causal), st)
            else:
                L.call("idb200_attention_bwd", sv["qkv"][i].data_ptr(), do16.data_ptr(), dqkv.data_ptr(), B, Lseq, H, int(causal), 0, st)
                sc.colsum(dqkv, grads[p + "attn.in_proj_bias"])
            dw_async(dqkv, sv["a1"][i], grads[p + "attn.in_proj_weight"])
            E.gemm_bf16(dqkv, w["wqkvt16"], None, da, E.EPI_BF16)                       # da1 = dqkv Wqkv
            if grouped:
                g16 = g_mlp[i - 1] if i > 0 else dh16                                   # the last one lands in the caller's dh16
            else:
                g16 = next16()
            ln_bwd(sv["h_in"][i], w["n1w"], w["n1b"], 2 * i, g16)
        if side is not None:
            main.wait_stream(side)                           # every weight gradient is final (and the capture's fork is joined)
            assert cur16[0] == 0                             # 2 nl LayerNorm backward calls: the last one wrote the caller's dh16
        pairs = []
        if grouped:
            # the four weight kinds, all layers at once: split s of the stacked token axis = layer s
            for key, dy_all, x_all, n_out, k_in in (("ff.2.weight", g_mlp, sv["f"], d, ff), ("ff.0.weight", du_all, sv["a2"], ff, d),
                                                    ("attn.out_proj.weight", g_att, sv["o"], d, d),
                                                    ("attn.in_proj_weight", dqkv_all, sv["a1"], 3 * d, d)):
                part = ws.get("dw_" + key, (nl, n_out, k_in), F32, dev)
                L.call("idb200_gemm_bf16_nn_splitk", dy_all.data_ptr(), x_all.data_ptr(), part.data_ptr(), n_out, k_in, nl * M, nl, st)
                pairs += [(grads[f"{prefix}layers.{i}.{key}"], part[i]) for i in range(nl)]
        sums = self.sc.colsum_segments(dwb_all, ws.get("dwb_sums", (2 * nl, 3 * d), F32, dev))
        for i in range(nl):
            p = f"{prefix}layers.{i}."
            pairs += [(grads[p + "norm1.weight"], sums[2 * i, :d]), (grads[p + "norm1.bias"], sums[2 * i, d:2 * d]),
                      (grads[p + "norm2.weight"], sums[2 * i + 1, :d]), (grads[p + "norm2.bias"], sums[2 * i + 1, d:2 * d]),
                      (grads[p + "attn.out_proj.bias"], sums[2 * i + 1, 2 * d:])]
            if i > 0:
                pairs.append((grads[f"{prefix}layers.{i - 1}.ff.2.bias"], sums[2 * i, 2 * d:]))
        if fused_qkv_sums:
            qs = self.sc.colsum_segments(dqkv_sum_all, ws.get("dqkv_sums", (nl, 3 * d), F32, dev))
            pairs += [(grads[f"{prefix}layers.{i}.attn.in_proj_bias"], qs[i]) for i in range(nl)]
        if fused_du_sums:
            us = self.sc.colsum_segments(du_part_all, ws.get("du_sums", (nl, ff), F32, dev))
            pairs += [(grads[f"{prefix}layers.{i}.ff.0.bias"], us[i]) for i in range(nl)]
        multi_copy(pairs)                                    # one launch per 96 slices instead of a copy kernel each
        self._dgb = dgb

    def backward_film(self, grads: Dict[str, torch.Tensor], prefix: str = "transformer.") -> Optional[torch.Tensor]:
        """FiLM linears of all LayerNorms (gradients w.r.t. film1 / film2 of every layer, and d cond_vec)."""
        S = self.saved
        film, dgb, W = S["film"], self._dgb, S["W"]
        B = S["B"]
        nl = len(W)
        ws, sc = self.sc.ws, self.sc
        # the saved activations (3.6 GB per layer at B = 4096) are dead from here on: release them so that the next forward -- in
        # particular the CUDA-graph capture that follows the eager warm-up pass -- does not hold two sets (117 -> 75 GB peak at B = 4096)
        self.saved = None
        self._dgb = None
        if film is None:
            return None
        d = dgb.shape[2] // 2
        dev = dgb.device
        # FiLM linears of all LayerNorms at once: gb = cond_vec W_all^T + b_all  (W_all [2 nl * 2d, d_cond])
        cond_vec, film_w = S["cond_vec"], S["film_w"]
        dc = cond_vec.shape[1]
        dgb2 = dgb.view(B, -1)
        n_all = dgb2.shape[1]
        db_all = ws.get("film_db", (n_all,), F32, dev)
        sc.colsum(dgb2, db_all)
        dW_all = ws.get("film_dw", (n_all, dc), F32, dev)
        dcond = torch.empty((B, dc), device=dev, dtype=F32)
        if B % 64 == 0 and dc % 64 == 0:
            dgb16 = dgb2.to(BF16)
            sc.dweight(dgb16, cond_vec.to(BF16).contiguous(), dW_all)
            E.gemm_bf16(dgb16, film_w.t().contiguous().to(BF16), None, dcond, E.EPI_F32)
        else:
            sgemm_strided(dgb2, True, cond_vec, True, dW_all)
            sgemm_strided(dgb2, False, film_w, True, dcond)
        two_d = 2 * d
        pairs = []
        for i in range(nl):
            for k, nm in enumerate(("film1", "film2")):
                r0 = (2 * i + k) * two_d
                pairs.append((grads[f"{prefix}layers.{i}.{nm}.weight"], dW_all[r0:r0 + two_d]))
                pairs.append((grads[f"{prefix}layers.{i}.{nm}.bias"], db_all[r0:r0 + two_d]))
        multi_copy(pairs)
        return dcond


# ----------------------------------------------------------------------------------------------------------------------
class _MLP2:
    """Linear -> SiLU -> Linear on per-trajectory rows [B, *] in fp32 (level_proj, sg.mlp, t_embed)."""

    def __init__(self, seq, scratch: _Scratch):
        self.l0, self.l2, self.sc = seq[0], seq[2], scratch

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        f = lambda t: t.detach().float().contiguous()
        self.x = x
        self.pre = E.sgemm(x, f(self.l0.weight), f(self.l0.bias))
        self.hid = silu_f32(self.pre)
        return E.sgemm(self.hid, f(self.l2.weight), f(self.l2.bias), out, accumulate=out is not None)

    def backward(self, dy: torch.Tensor, grads: Dict[str, torch.Tensor], prefix: str, want_dx: bool = True) -> Optional[torch.Tensor]:
        f = lambda t: t.detach().float().contiguous()
        sgemm_strided(dy, True, self.hid, True, grads[prefix + "2.weight"])
        self.sc.colsum(dy, grads[prefix + "2.bias"])
        dhid = torch.empty_like(self.hid)
        sgemm_strided(dy, False, f(self.l2.weight), True, dhid)
        dpre = silu_f32(self.pre, g=dhid)
        sgemm_strided(dpre, True, self.x, True, grads[prefix + "0.weight"])
        self.sc.colsum(dpre, grads[prefix + "0.bias"])
        if not want_dx:
            return None
        dx = torch.empty_like(self.x)
        sgemm_strided(dpre, False, f(self.l0.weight), True, dx)
        return dx


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


class CondEncoderBackprop:
    """``MazeConditionEncoder`` (encoders.py:41-71) in training form: the conv stack as im2col + tcgen05 GEMM on NHWC bf16
    pre-activations, mean pool, fc, start/goal MLP."""

    def __init__(self, cond_enc, scratch: _Scratch):
        self.m, self.sc = cond_enc, scratch
        self.sg = _MLP2(cond_enc.sg.mlp, scratch) if cond_enc.sg is not None else None

    def _col(self, src: torch.Tensor, B: int, Hh: int, Ww: int, C: int, act: bool, name: str) -> torch.Tensor:
        col = self.sc.ws.get(name, (B * Hh * Ww, _pad64(9 * C)), BF16, src.device)
        return E.im2col3x3(src, B, Hh, Ww, C, act, col)

    def forward(self, cond: Dict[str, torch.Tensor]) -> torch.Tensor:
        m = self.m
        occ = L.f32c(cond["occ"])
        if m.use_sdf:
            if cond.get("sdf") is None:
                raise ValueError("use_sdf is True but sdf missing from cond")
            x = torch.cat([occ, L.f32c(cond["sdf"])], dim=1)
        else:
            x = occ
        B, _, Hh, Ww = x.shape
        convs = [c for c in m.maze.convs if isinstance(c, torch.nn.Conv2d)]
        self.dims = (B, Hh, Ww)
        self.pooled, self.x0, self.us, self.wmats, self.cols = E.conv_stack_gemm(
            x, [c.weight for c in convs], [c.bias.detach().float().contiguous() for c in convs], self.sc.ws, keep=True)
        emb = E.sgemm(self.pooled, m.maze.fc.weight.detach().float().contiguous(), m.maze.fc.bias.detach().float().contiguous())
        if m.use_start_goal:
            if "start_goal" not in cond:
                raise ValueError("use_start_goal is True but start_goal missing from cond")
            self.sg.forward(L.f32c(cond["start_goal"]), out=emb)
        return emb

    def backward(self, dcv: torch.Tensor, grads: Dict[str, torch.Tensor], prefix: str = "cond_enc.") -> None:
        m, sc, ws = self.m, self.sc, self.sc.ws
        B, Hh, Ww = self.dims
        P = Hh * Ww
        dev = dcv.device
        st = L.stream(dev)
        if self.sg is not None:
            self.sg.backward(dcv, grads, prefix + "sg.mlp.", want_dx=False)
        fcw = m.maze.fc.weight.detach().float().contiguous()
        sgemm_strided(dcv, True, self.pooled, True, grads[prefix + "maze.fc.weight"])
        sc.colsum(dcv, grads[prefix + "maze.fc.bias"])
        dpooled = torch.empty_like(self.pooled)
        sgemm_strided(dcv, False, fcw, True, dpooled)
        convs = [(k, c) for k, c in enumerate(m.maze.convs) if isinstance(c, torch.nn.Conv2d)]
        n = len(convs)
        C_last = self.us[-1].shape[1]
        du = torch.empty_like(self.us[-1])
        L.call("idb200_pool_silu_bwd", self.us[-1].data_ptr(), dpooled.data_ptr(), B, P, C_last, du.data_ptr(), st)
        for li in range(n - 1, -1, -1):
            seq_idx, c = convs[li]
            co, ci = c.weight.shape[0], c.weight.shape[1]
            col = self.cols[li]                                 # the forward's patch matrix of this layer
            kpad = col.shape[1]
            dwm = ws.get("dwm", (co, kpad), F32, dev)
            sc.dweight(du, col, dwm)
            grads[f"{prefix}maze.convs.{seq_idx}.weight"].copy_(dwm[:, :9 * ci].reshape(co, 3, 3, ci).permute(0, 3, 1, 2))
            sc.colsum(du, grads[f"{prefix}maze.convs.{seq_idx}.bias"])
            if li == 0:
                break
            # dgrad = conv3x3(du, W flipped and transposed):  Wflip[ci, (ky, kx, co)] = W[co, ci, 2 - ky, 2 - kx]
            w = c.weight.detach().float()
            kp2 = _pad64(9 * co)
            wf = torch.zeros((ci, kp2), device=dev, dtype=BF16)
            wf[:, :9 * co] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).to(BF16)
            colg = self._col(du, B, Hh, Ww, co, False, "colg")
            dact = torch.empty((B * P, ci), device=dev, dtype=BF16)
            E.gemm_bf16(colg, wf, None, dact, E.EPI_BF16)
            du = silu_bf16(self.us[li - 1], dact, g=dact)


# ----------------------------------------------------------------------------------------------------------------------
class _DenoiserBackprop:
    """Shared by both denoisers: cond encoder -> token assembly -> encoder -> out head, and the way back."""

    def __init__(self, model):
        self.model = model
        self.sc = _Scratch()
        self.enc = EncoderBackprop(model.transformer, self.sc)
        self.cond = CondEncoderBackprop(model.cond_enc, self.sc) if hasattr(model.cond_enc, "maze") else None
        self.split_hook = None                  # called once per backward at the split point (see _backward_to_tokens)

    def early_param_names(self) -> List[str]:
        """Parameters whose gradients are final at the split point of the backward."""
        return [n for n, _ in self.model.named_parameters()
                if n.startswith("out.") or (n.startswith("transformer.") and ".film" not in n)]

    def param_names(self) -> List[str]:
        return [n for n, _ in self.model.named_parameters()]

    def new_grads(self) -> Dict[str, torch.Tensor]:
        return {n: torch.zeros_like(p, dtype=F32) for n, p in self.model.named_parameters()}

    def _cond_rows(self, cond: Dict[str, torch.Tensor]):
        m = self.model
        if self.cond is None:
            raise ValueError("training needs the built-in MazeConditionEncoder")
        f = lambda t: t.detach().float().contiguous()
        self.cond_vec = self.cond.forward(cond)
        return E.sgemm(self.cond_vec, f(m.cond_proj.weight), (f(m.cond_proj.bias) + f(m.in_proj.bias)).contiguous())

    def _encode_and_head(self, h: torch.Tensor, B: int, Lseq: int, D: int) -> torch.Tensor:
        m = self.model
        f = lambda t: t.detach().float().contiguous()
        M, d = h.shape
        self.enc.forward(h, B, Lseq, self.cond_vec)
        self.h_final = h
        self.shape = (B, Lseq, D, d, M)
        out = torch.empty((B, Lseq, D), device=h.device, dtype=F32)
        E.out_head(h, f(m.out.weight), f(m.out.bias), out.view(M, D))
        return out

    def _backward_to_tokens(self, d_out: torch.Tensor, grads: Dict[str, torch.Tensor]):
        """out head + encoder backward; returns (dh0 fp32 [M,d], its bf16 copy, per-trajectory token sums [B,d], d cond_vec)."""
        m, sc = self.model, self.sc
        B, Lseq, D, d, M = self.shape
        dev = d_out.device
        st = L.stream(dev)
        f = lambda t: t.detach().float().contiguous()
        dy = L.f32c(d_out).view(M, D)
        sc.narrow_outer(dy, self.h_final, grads["out.weight"])
        sc.colsum(dy, grads["out.bias"])
        dh = torch.empty((M, d), device=dev, dtype=F32)
        dh16 = torch.empty((M, d), device=dev, dtype=BF16)
        L.call("idb200_head_bwd", dy.data_ptr(), f(m.out.weight).data_ptr(), M, d, D, dh.data_ptr(), dh16.data_ptr(), st)
        self.enc.backward_layers(dh, dh16, grads, "transformer.")
        # ---- split point of the data-parallel step: out.* and every transformer gradient except the FiLM linears are final here
        # (``early_param_names``); what follows (FiLM, token assembly, level / timestep MLP, conditioning encoder) takes ~2.5 ms
        # at the cfg-4 shapes, enough to hide the all-reduce of the first group (Stage2Trainer.overlap_allreduce)
        if self.split_hook is not None:
            self.split_hook()
        dcond = self.enc.backward_film(grads, "transformer.")
        tok = torch.empty((B, d), device=dev, dtype=F32)
        L.call("idb200_token_sum", dh.data_ptr(), B, Lseq, d, tok.data_ptr(), st)
        # every token carries in_proj.bias and cond_proj(cond_vec): their gradients are the per-trajectory token sums
        sc.colsum(tok, grads["in_proj.bias"])
        grads["cond_proj.bias"].copy_(grads["in_proj.bias"])
        sgemm_strided(tok, True, self.cond_vec, True, grads["cond_proj.weight"])
        if dcond is None:
            dcond = torch.zeros_like(self.cond_vec)
        sgemm_strided(tok, False, f(m.cond_proj.weight), True, dcond, accumulate=True)
        return dh, dh16, tok, dcond


class InterpLevelBackprop(_DenoiserBackprop):
    """``InterpLevelDenoiser.forward`` (denoiser_interp_levels.py:64-84) + its backward."""

    def __init__(self, model):
        super().__init__(model)
        self.level = _MLP2(model.level_proj, self.sc)

    def forward(self, x_s: torch.Tensor, s: torch.Tensor, mask: torch.Tensor, cond: Dict[str, torch.Tensor]) -> torch.Tensor:
        m = self.model
        dev = L.require_cuda(x_s, s, mask)
        B, T, D = x_s.shape
        d = m.in_proj.weight.shape[0]
        C = 1 if mask.dim() == 2 else mask.shape[-1]
        if C != m.mask_channels:
            raise ValueError(f"mask has {C} channels, expected {m.mask_channels}")
        M = B * T
        f = lambda t: t.detach().float().contiguous()
        # token features [x_s | mask] as fp32 (the in_proj weight gradient reads them back)
        self.feat = torch.cat([L.f32c(x_s).view(M, D), mask.reshape(M, C).to(F32)], dim=1).contiguous()
        row_b = self._cond_rows(cond)
        Wf = f(m.in_proj.weight).t().contiguous()
        if getattr(self, "_tab_key", None) != (T, d, dev):              # input-independent table: built once (host linspace)
            self._tab, self._tab_key = m._positional_embedding(T, dev, d), (T, d, dev)
        tab = self._tab
        s64 = L.i64c(s)
        self.onehot = torch.zeros((B, m.level_emb.weight.shape[0]), device=dev, dtype=F32)
        self.onehot.scatter_(1, s64.view(B, 1), 1.0)
        emb = f(m.level_emb.weight)[s64].contiguous()
        level_vec = self.level.forward(emb)
        h = torch.empty((M, d), device=dev, dtype=F32)
        E.embed_tokens(self.feat, None, None, Wf, tab, None, level_vec, row_b, h, M, T, d)
        return self._encode_and_head(h, B, T, D)

    def backward(self, d_out: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
        """d_out = d loss / d delta_hat [B, T, D]; fills ``grads`` (every parameter of the model)."""
        dh, _, tok, dcond = self._backward_to_tokens(d_out, grads)
        d = dh.shape[1]
        # token assembly: h0 = feat Wf + pos[t] + level_vec[b] + (cond_proj(cond_vec) + biases)[b]
        dWf = torch.empty((self.feat.shape[1], d), device=dh.device, dtype=F32)
        self.sc.narrow_outer(self.feat, dh, dWf)
        grads["in_proj.weight"].copy_(dWf.t())
        demb = self.level.backward(tok, grads, "level_proj.")
        sgemm_strided(self.onehot, True, demb, True, grads["level_emb.weight"])
        self.cond.backward(dcond, grads, "cond_enc.")


class KeypointBackprop(_DenoiserBackprop):
    """``KeypointDenoiser.forward`` (denoiser_keypoints.py:82-113) + its backward (Stage-1 training, train_keypoints.py:505-556)."""

    def __init__(self, model):
        super().__init__(model)
        self.t_embed = _MLP2(model.t_embed, self.sc)

    def forward(self, z_t: torch.Tensor, t: torch.Tensor, idx: torch.Tensor, known_mask: torch.Tensor, cond: Dict[str, torch.Tensor],
                T: int) -> torch.Tensor:
        from ..models.denoiser_keypoints import timestep_embedding
        m = self.model
        dev = L.require_cuda(z_t, t, idx, known_mask)
        B, K, D = z_t.shape
        d = m.in_proj.weight.shape[0]
        P, Fk = m.pos_dim, m.kp_feat_dim
        M = B * K
        f = lambda x: x.detach().float().contiguous()
        z = L.f32c(z_t).view(M, D)
        km = known_mask.reshape(M, D).to(F32)
        if Fk > 0:
            kp = L.f32c(cond["kp_feat"]).view(M, Fk) if (cond is not None and "kp_feat" in cond) else torch.zeros((M, Fk), device=dev)
        else:
            kp = None
        pos_tab = E.sinusoid(T, P - (P % 2), device=dev)                      # sinusoid(r / max(1, T - 1)), r = 0..T-1
        if P % 2 == 1:
            pos_tab = torch.nn.functional.pad(pos_tab, (0, 1))
        idx64 = L.i64c(idx).view(M)
        # reference feature order (denoiser_keypoints.py:99-102): [z_t | pos_emb | known_mask | kp_feat]
        parts = [z, pos_tab[idx64], km] + ([kp] if kp is not None else [])
        x_full = torch.cat(parts, dim=1)
        fan_in = x_full.shape[1]
        self.x16 = torch.zeros((M, _pad64(fan_in)), device=dev, dtype=BF16)
        self.x16[:, :fan_in] = x_full
        self.fan_in = fan_in
        row_b = self._cond_rows(cond)
        t_vec = self.t_embed.forward(timestep_embedding(L.i64c(t), d))
        # token assembly by the inference kernel: h0 = [z | kp | km] Wf + (pos_tab W_pos^T)[idx] + t_vec[b] + row_b[b]
        W = f(m.in_proj.weight)
        Wf = torch.cat([W[:, :D], W[:, 2 * D + P:], W[:, D + P: 2 * D + P]], dim=1).t().contiguous()
        tab = E.sgemm(pos_tab[:, :P - (P % 2)].contiguous(), W[:, D: D + P - (P % 2)].contiguous(), None)
        h = torch.empty((M, d), device=dev, dtype=F32)
        E.embed_tokens(z, kp, L.u8c(known_mask).view(M, D), Wf, tab, idx64, t_vec, row_b, h, M, K, d)
        return self._encode_and_head(h, B, K, D)

    def backward(self, d_out: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
        """d_out = d loss / d eps_hat [B, K, D]; fills ``grads``."""
        dh, dh16, tok, dcond = self._backward_to_tokens(d_out, grads)
        d = dh.shape[1]
        dW = self.sc.ws.get("dw_in", (d, self.x16.shape[1]), F32, dh.device)
        self.sc.dweight(dh16, self.x16, dW)                                   # in_proj.weight: dh0^T [z | pos | km | kp]
        grads["in_proj.weight"].copy_(dW[:, :self.fan_in])
        self.t_embed.backward(tok, grads, "t_embed.", want_dx=False)
        self.cond.backward(dcond, grads, "cond_enc.")
