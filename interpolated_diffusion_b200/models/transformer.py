"""Mirror of the reference's ``src/models/transformer.py``: same module tree / parameter names (so reference
checkpoints and seeded random init load unchanged), forward executed by libidb200 kernels.  Dropout is 0 on
the hot path (reference default) and is not applied."""
from typing import Optional

import torch
import torch.nn as nn

from . import _engine as E


class TransformerBlock(nn.Module):
    def __init__(self, d_model: int, n_heads: int, d_ff: int, dropout: float = 0.0, cond_dim: Optional[int] = None):
        super().__init__()
        if d_model // n_heads != 32 or d_model % n_heads != 0:
            raise ValueError("the B200 attention kernels are specialised for head_dim == 32 (the reference's setting)")
        self.attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout, batch_first=True)
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.SiLU(), nn.Linear(d_ff, d_model))
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.cond_dim = cond_dim
        if cond_dim is not None:
            self.film1 = nn.Linear(cond_dim, d_model * 2)
            self.film2 = nn.Linear(cond_dim, d_model * 2)
        else:
            self.film1 = None
            self.film2 = None


class TransformerEncoder(nn.Module):
    def __init__(self, d_model: int = 256, n_layers: int = 8, n_heads: int = 8, d_ff: int = 1024, dropout: float = 0.0,
                 cond_dim: Optional[int] = None, causal: bool = False, use_checkpoint: bool = False):
        super().__init__()
        self.layers = nn.ModuleList(
            [TransformerBlock(d_model, n_heads, d_ff, dropout=dropout, cond_dim=cond_dim) for _ in range(n_layers)])
        self.causal = causal
        self.use_checkpoint = use_checkpoint
        self.precision = "bf16"
        self._packed = None

    def flops_per_token(self, Lseq: int) -> float:
        """Dense FLOPs (2 per multiply-add) one token costs in this encoder at sequence length ``Lseq``: per layer the packed QKV
        projection, the attention products, the out projection and the two MLP GEMMs (SURVEY 8d's per-layer formula, without
        the per-trajectory FiLM linears).  Used by the benches to turn a measured time into tensor-pipe utilisation."""
        blk = self.layers[0]
        d = blk.norm1.weight.shape[0]
        ff = blk.ff[0].weight.shape[0]
        per_layer = 2.0 * d * 3 * d + 2.0 * d * d + 4.0 * Lseq * d + 4.0 * d * ff
        return per_layer * len(self.layers)

    def packed(self) -> E.PackedEncoder:
        if self._packed is None:
            self._packed = E.PackedEncoder(self)
        return self._packed

    @torch.no_grad()
    def forward(self, x: torch.Tensor, cond: Optional[torch.Tensor] = None, *, film: Optional[torch.Tensor] = None):
        """transformer.py:73-82: x [B, L, d] fp32 -> [B, L, d].  ``film`` may carry precomputed FiLM parameters."""
        B, Lseq, d = x.shape
        h = x.detach().float().contiguous().clone().view(B * Lseq, d)
        pk = self.packed()
        if film is None and cond is not None:
            film = pk.film_params(cond.detach().float().contiguous(), Lseq, self.precision)
        pk.forward(h, B, Lseq, film, self.precision)
        return h.view(B, Lseq, d)
