"""Mirror of the reference's ``src/models/denoiser_interp_levels.py`` (and, through ``causal=True``, of
``denoiser_interp_levels_causal.py``): same constructor, parameter tree and ``forward(x_s, s, mask, cond)``."""
from typing import Dict, Optional

import os

import torch
import torch.nn as nn

from .. import _lib as L
from . import _engine as E
from .encoders import MazeConditionEncoder
from .transformer import TransformerEncoder


class InterpLevelDenoiser(nn.Module):
    _causal = False

    def __init__(self, d_model: int = 256, n_layers: int = 8, n_heads: int = 8, d_ff: int = 1024, dropout: float = 0.0,
                 d_cond: int = 128, use_sdf: bool = False, use_start_goal: bool = True, data_dim: int = 2, max_levels: int = 8,
                 use_checkpoint: bool = False, mask_channels: int = 1, cond_encoder: Optional[nn.Module] = None,
                 maze_channels: tuple = (32, 64)):
        super().__init__()
        self.data_dim = data_dim
        self.d_cond = d_cond
        self.mask_channels = mask_channels
        self.in_proj = nn.Linear(data_dim + mask_channels, d_model)
        self.level_emb = nn.Embedding(max_levels + 1, d_model)
        self.level_proj = nn.Sequential(nn.Linear(d_model, d_model), nn.SiLU(), nn.Linear(d_model, d_model))
        if cond_encoder is None:
            cond_encoder = MazeConditionEncoder(use_sdf=use_sdf, d_cond=d_cond, use_start_goal=use_start_goal,
                                                maze_channels=maze_channels)
        self.cond_enc = cond_encoder
        self.cond_proj = nn.Linear(d_cond, d_model)
        self.transformer = TransformerEncoder(d_model=d_model, n_layers=n_layers, n_heads=n_heads, d_ff=d_ff, dropout=dropout,
                                              cond_dim=d_cond, causal=self._causal, use_checkpoint=use_checkpoint)
        self.out = nn.Linear(d_model, data_dim)
        self.precision = "bf16"
        # token assembly + out head inside the fused-encoder launch (idb200_denoiser_fused): h never exists in HBM (saves the
        # [M, 256] fp32 buffer, 4.3 GB at B = 65536, T = 64).  On by default since round 2: the prologue's operands are staged in
        # shared memory by TMA (<= 64 table rows, <= 8 features, >= 8 tokens per trajectory; other shapes gather from global memory,
        # which round 1 measured 2 % slower than the dedicated embed kernel).  IDB200_FUSE_IO=0 selects the separate kernels.
        self.fuse_io = os.environ.get("IDB200_FUSE_IO", "1") != "0"
        self.fuse_head = True            # the out head as the tile epilogue of the fused-encoder launch (h is not written back)
        self.pad_causal = True           # causal models: right-pad T < 128 to a divisor of 128 so the whole-encoder kernel applies
        self._cache = {}
        self._ws = E.Workspace()

    def _positional_embedding(self, T: int, device: torch.device, dim: int) -> torch.Tensor:
        """denoiser_interp_levels.py:54-62: sinusoid(linspace(0, 1, T)) -> [T, dim] (input-independent table)."""
        pos = torch.linspace(0.0, 1.0, T).to(device)                 # same primitive as the reference, host side
        emb = E.sinusoid(T, dim - (dim % 2), device, args=pos.contiguous())
        if dim % 2 == 1:
            emb = torch.nn.functional.pad(emb, (0, 1))
        return emb

    def _derived(self, T: int, device):
        key = (T, E._sig([self.in_proj.weight, self.in_proj.bias, self.cond_proj.bias]))
        hit = self._cache.get("derived")
        if hit is not None and hit[0] == key:
            return hit[1]
        d = self.in_proj.weight.shape[0]
        val = {
            "Wf": self.in_proj.weight.detach().float().t().contiguous(),          # [D + C, d]
            "tab": self._positional_embedding(T, device, d),
            "bias_b": (self.cond_proj.bias.detach().float() + self.in_proj.bias.detach().float()).contiguous(),
        }
        self._cache["derived"] = (key, val)
        return val

    @torch.no_grad()
    def _forward_padded(self, xs_p, s, mk_p, cond, cond_vec, T_real: int) -> torch.Tensor:
        """forward() on a right-padded causal sequence with the positional table of the real length."""
        Tp = xs_p.shape[1]
        dev = xs_p.device
        d = self.in_proj.weight.shape[0]
        der = dict(self._derived(T_real, dev))
        der["tab"] = torch.nn.functional.pad(der["tab"], (0, 0, 0, Tp - T_real)).contiguous()
        key = self._cache.get("derived")
        self._cache["derived"] = ((Tp, E._sig([self.in_proj.weight, self.in_proj.bias, self.cond_proj.bias])), der)
        try:
            self.pad_causal = False
            return self.forward(xs_p, s, mk_p, cond, cond_vec=cond_vec)
        finally:
            self.pad_causal = True
            self._cache["derived"] = key

    @torch.no_grad()
    def _sync_precision(self):
        if hasattr(self.cond_enc, "maze"):
            self.cond_enc.precision = self.precision

    @torch.no_grad()
    def encode_cond(self, cond: Dict[str, torch.Tensor]) -> torch.Tensor:
        self._sync_precision()
        return self.cond_enc(cond)

    @torch.no_grad()
    def level_vector(self, s: torch.Tensor) -> torch.Tensor:
        """level_proj(level_emb(s)) -> [len(s), d] (denoiser_interp_levels.py:75)."""
        emb = self.level_emb.weight.detach().float()[s].contiguous()
        hdn = E.sgemm(emb, self.level_proj[0].weight.detach().float().contiguous(), self.level_proj[0].bias.detach().float().contiguous(), act=1)
        return E.sgemm(hdn, self.level_proj[2].weight.detach().float().contiguous(), self.level_proj[2].bias.detach().float().contiguous())

    @torch.no_grad()
    def forward(self, x_s: torch.Tensor, s: torch.Tensor, mask: torch.Tensor, cond: Dict[str, torch.Tensor], *,
                cond_vec: Optional[torch.Tensor] = None, film: Optional[torch.Tensor] = None,
                level_vec: Optional[torch.Tensor] = None, row_b: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """denoiser_interp_levels.py:64-84 -> delta [B, T, D]."""
        dev = L.require_cuda(x_s, s, mask)
        B, T, D = x_s.shape
        d = self.in_proj.weight.shape[0]
        C = 1 if mask.dim() == 2 else mask.shape[-1]
        if C != self.mask_channels:
            raise ValueError(f"mask has {C} channels, expected {self.mask_channels}")
        pad_fused = T < 128 and 128 % T != 0 and self.transformer.packed().fused_path(128, self.precision)
        pad_mma = T > 128 and T % 16 != 0 and self.precision == "bf16"
        if self._causal and self.pad_causal and (pad_fused or pad_mma) and out is None and level_vec is None and row_b is None \
                and film is None:
            # Causal attention never looks to the right, so right-padding the sequence leaves every real token's result unchanged
            # (the positional table is built for the real length).  Short sequences are padded to a divisor of 128, the lengths the
            # whole-encoder kernel takes; long ones to a multiple of 16, the tensor-core attention kernel's granularity -- at other
            # lengths the attention falls back to the fp32 SIMT kernel (measured: 1.2 s of a 1.5 s T = 256 chunked generation).
            Tp = (1 << (T - 1).bit_length()) if pad_fused else (T + 15) // 16 * 16
            pad = Tp - T
            xs_p = torch.nn.functional.pad(L.f32c(x_s), (0, 0, 0, pad))
            mk_p = torch.nn.functional.pad(mask, (0, pad) if mask.dim() == 2 else (0, 0, 0, pad))
            return self._forward_padded(xs_p, s, mk_p, cond, cond_vec, T)[:, :T].contiguous()
        der = self._derived(T, dev)
        M = B * T
        src1 = src2 = None
        if mask.dtype in (torch.bool, torch.uint8):
            src2 = L.u8c(mask).view(M, C)
        else:
            src1 = L.f32c(mask).view(M, C)
        if cond_vec is None:
            if cond and self.cond_enc is not None:
                cond_vec = self.encode_cond(cond)
            else:
                cond_vec = torch.zeros((B, self.d_cond), device=dev, dtype=torch.float32)
        if row_b is None:
            row_b = E.sgemm(cond_vec, self.cond_proj.weight.detach().float().contiguous(), der["bias_b"])
        if level_vec is None:
            level_vec = self.level_vector(L.i64c(s))
        pk = self.transformer.packed()
        if film is None:
            film = pk.film_params(cond_vec, T, self.precision)
        if out is None:
            out = torch.empty((B, T, D), device=dev, dtype=torch.float32)
        W_out, b_out = self.out.weight.detach().float().contiguous(), self.out.bias.detach().float().contiguous()
        # (the staged prologue needs <= 64 table rows, <= 8 features, T >= 8; other shapes are faster through the separate embed kernel
        # unless fuse_io == "always")
        staged = der["tab"].shape[0] <= 64 and der["Wf"].shape[0] <= 8 and T >= 8
        if self.fuse_io and (staged or self.fuse_io == "always") and pk.fused_path(T, self.precision) and W_out.shape[0] <= 4 and (film is None or isinstance(film, E.Film)):
            # token assembly, all encoder layers and the out head in one launch: h never exists in HBM
            E.denoiser_fused(pk, film, T, bool(self.transformer.causal), M, L.f32c(x_s).view(M, D), src1, src2, der["Wf"], der["tab"], None,
                             level_vec, row_b, W_out, b_out, out.view(M, D))
            return out
        h = self._ws.get("h", (M, d), torch.float32, dev)
        E.embed_tokens(L.f32c(x_s).view(M, D), src1, src2, der["Wf"], der["tab"], None, level_vec, row_b, h, M, T, d)
        if self.fuse_head and pk.fused_path(T, self.precision) and W_out.shape[0] <= 4 and (film is None or isinstance(film, E.Film)):
            # all encoder layers + the out head in one launch: the final residual stream is never written back
            E.encoder_fused_head(h, pk, film, T, bool(self.transformer.causal), W_out, b_out, out.view(M, D))
            return out
        pk.forward(h, B, T, film, self.precision)
        E.out_head(h, W_out, b_out, out.view(M, D))
        return out
