"""Mirror of the reference's ``src/models/segment_cost.py`` (``SegmentCostPredictor`` :11-57, the d_phi model that scores every
(i, j) segment for the DP anchor placement): same parameters, forward on libidb200.

The first Linear acts on ``[cond_vec | seg_feat]``; it is split into its conditioning half (one row per sample) and its
segment-feature half (one row per segment, or per (sample, segment)) and assembled by the token-embedding kernel, so the
[B, S, d_cond + F] concatenation of the reference is never built.  Hidden layers run on the tcgen05 GEMM (bf16, fp32 accumulate)."""
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib as L
from . import _engine as E
from .encoders import MazeConditionEncoder

BF16, F32 = torch.bfloat16, torch.float32


class SegmentCostPredictor(nn.Module):
    def __init__(self, d_cond: int = 128, seg_feat_dim: int = 3, hidden_dim: int = 256, n_layers: int = 3, dropout: float = 0.0,
                 cond_encoder: Optional[nn.Module] = None, use_sdf: bool = False, use_start_goal: bool = True,
                 maze_channels: tuple = (32, 64)) -> None:
        super().__init__()
        if hidden_dim % 64 != 0:
            raise ValueError("hidden_dim must be a multiple of 64 on the B200 path")
        if cond_encoder is None:
            cond_encoder = MazeConditionEncoder(use_sdf=use_sdf, d_cond=d_cond, use_start_goal=use_start_goal, maze_channels=maze_channels)
        self.cond_enc = cond_encoder
        self.d_cond = d_cond
        self.seg_feat_dim = seg_feat_dim
        layers, in_dim = [], d_cond + seg_feat_dim
        for _ in range(max(1, n_layers - 1)):
            layers += [nn.Linear(in_dim, hidden_dim), nn.SiLU()]
            if dropout > 0:
                layers.append(nn.Dropout(dropout))
            in_dim = hidden_dim
        layers.append(nn.Linear(in_dim, 1))
        self.mlp = nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, cond: Dict[str, torch.Tensor], seg_feat: torch.Tensor) -> torch.Tensor:
        """segment_cost.py:43-57: seg_feat [S, F] (shared) or [B, S, F] -> predicted cost [B, S]."""
        if cond is None:
            raise ValueError("cond is required for SegmentCostPredictor")
        cond_vec = self.cond_enc(cond)
        B = cond_vec.shape[0]
        dev = cond_vec.device
        if seg_feat.dim() not in (2, 3):
            raise ValueError("seg_feat must be [S,F] or [B,S,F]")
        if seg_feat.shape[-1] != self.seg_feat_dim:
            raise ValueError("seg_feat_dim mismatch")
        f = lambda t: t.detach().float().contiguous()
        lin = [m for m in self.mlp if isinstance(m, nn.Linear)]
        W0, b0 = f(lin[0].weight), f(lin[0].bias)
        hid = W0.shape[0]
        S = seg_feat.shape[-2]
        M = B * S
        row_cond = E.sgemm(cond_vec, W0[:, :self.d_cond].contiguous(), b0)                  # [B, hid]
        Wf = W0[:, self.d_cond:].t().contiguous()                                           # [F, hid]
        zero_row = torch.zeros((B, hid), device=dev, dtype=F32)
        h = torch.empty((M, hid), device=dev, dtype=F32)
        sf = L.f32c(seg_feat).to(dev)
        if sf.dim() == 2:                  # shared features: a [S, hid] table indexed by the segment, plus the sample's row
            tab = E.sgemm(sf, W0[:, self.d_cond:].contiguous(), None)
            dummy = torch.zeros((M, 1), device=dev, dtype=F32)
            E.embed_tokens(dummy, None, None, torch.zeros((1, hid), device=dev, dtype=F32), tab, None, row_cond, zero_row, h, M, S, hid)
        else:
            tab = torch.zeros((S, hid), device=dev, dtype=F32)
            E.embed_tokens(sf.reshape(M, self.seg_feat_dim).contiguous(), None, None, Wf, tab, None, row_cond, zero_row, h, M, S, hid)
        a = torch.empty((M, hid), device=dev, dtype=F32)
        L.call("idb200_silu_f32", h.data_ptr(), None, h.numel(), 0, a.data_ptr(), L.stream(dev))
        a16 = a.to(BF16)
        for m in lin[1:-1]:
            nxt = torch.empty((M, m.weight.shape[0]), device=dev, dtype=BF16)
            E.gemm_bf16(a16, m.weight.detach().to(BF16).contiguous(), f(m.bias), nxt, E.EPI_SILU_BF16)
            a16 = nxt
        # last Linear(hidden -> 1): its single row padded to a 32-row tile of the tensor-core GEMM
        w_last = torch.zeros((32, a16.shape[1]), device=dev, dtype=BF16)
        w_last[0] = lin[-1].weight.detach().to(BF16)[0]
        b_last = torch.zeros((32,), device=dev, dtype=F32)
        b_last[0] = lin[-1].bias.detach().float()[0]
        out = torch.empty((M, 32), device=dev, dtype=F32)
        E.gemm_bf16(a16, w_last, b_last, out, E.EPI_F32)
        return out[:, 0].reshape(B, S).contiguous()
