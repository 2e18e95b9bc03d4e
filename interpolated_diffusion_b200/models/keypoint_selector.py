"""Mirror of the reference's ``src/models/keypoint_selector.py`` (``KeypointSelector`` :41-188, ``select_topk_indices`` :191-):
same constructor, parameter tree (``state_dict`` keys / shapes) and ``forward(cond) -> logits [B, T]``; the forward runs on
libidb200: Gaussian start/goal maps, the spatial conv stack as im2col + tcgen05 GEMM, the 1x1 projection / K|V / Q / out / MLP
token GEMMs on tcgen05, LayerNorm, and cross attention over the H*W (+ extra) memory tokens by ``idb200_cross_attention``.
bf16 tensor-core path only (fp32 accumulate).  ``use_cond_bias`` (both ``cond_bias_mode``s, :101-111, :170-175) and one-hot
start / goal maps (``sg_map_sigma <= 0``, :129-139) are covered."""
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib as L
from . import _engine as E
from .encoders import MazeConditionEncoder  # noqa: F401  (parity of the reference module's import surface)

BF16, F32 = torch.bfloat16, torch.float32


class CrossAttnBlock(nn.Module):
    def __init__(self, d_model: int, n_heads: int, d_ff: int, dropout: float = 0.0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout, batch_first=True)
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.SiLU(), nn.Linear(d_ff, d_model))
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)


class KeypointSelector(nn.Module):
    def __init__(self, T: int, d_model: int = 256, n_heads: int = 8, d_ff: int = 512, n_layers: int = 2, pos_dim: int = 64,
                 dropout: float = 0.0, use_sdf: bool = False, use_start_goal: bool = True, use_sg_map: bool = True,
                 use_sg_token: bool = True, use_goal_dist_token: bool = False, use_cond_bias: bool = False,
                 cond_bias_mode: str = "memory", use_level: bool = False, level_mode: str = "k_norm", sg_map_sigma: float = 1.5,
                 maze_channels: tuple = (32, 64)) -> None:
        super().__init__()
        if d_model // n_heads != 32 or d_model % n_heads != 0:
            raise ValueError("the B200 attention kernels are specialised for head_dim == 32")
        if d_model % 64 != 0 or d_ff % 64 != 0 or any(c % 32 != 0 for c in maze_channels) or maze_channels[-1] % 64 != 0:
            raise ValueError("d_model, d_ff and the last conv width must be multiples of 64, conv widths multiples of 32")
        self.T, self.d_model, self.n_heads, self.pos_dim = int(T), int(d_model), int(n_heads), int(pos_dim)
        self.use_sdf, self.use_start_goal, self.use_sg_map = bool(use_sdf), bool(use_start_goal), bool(use_sg_map)
        self.use_sg_token, self.use_goal_dist_token, self.use_cond_bias = bool(use_sg_token), bool(use_goal_dist_token), bool(use_cond_bias)
        self.cond_bias_mode, self.use_level, self.level_mode = str(cond_bias_mode), bool(use_level), str(level_mode)
        self.sg_map_sigma = float(sg_map_sigma)
        in_channels = 1 + (1 if self.use_sdf else 0) + (2 if self.use_sg_map else 0)
        convs, c_in = [], in_channels
        for c_out in maze_channels:
            convs += [nn.Conv2d(c_in, c_out, kernel_size=3, padding=1), nn.SiLU()]
            c_in = c_out
        self.spatial_conv = nn.Sequential(*convs)
        self.spatial_proj = nn.Conv2d(c_in, d_model, kernel_size=1) if c_in != d_model else nn.Identity()
        self.sg_token = None
        if self.use_start_goal and self.use_sg_token:
            self.sg_token = nn.Sequential(nn.Linear(4, d_model), nn.SiLU(), nn.Linear(d_model, d_model))
        self.goal_dist_token = None
        if self.use_goal_dist_token:
            self.goal_dist_token = nn.Sequential(nn.Linear(1, d_model), nn.SiLU(), nn.Linear(d_model, d_model))
        self.time_proj = nn.Linear(pos_dim, d_model)
        self.level_mlp = None
        if self.use_level:
            self.level_mlp = nn.Sequential(nn.Linear(1, d_model), nn.SiLU(), nn.Linear(d_model, d_model))
        self.cond_bias = None
        self.cond_enc = None
        if self.use_cond_bias:                                        # keypoint_selector.py:101-111
            self.cond_bias = nn.Sequential(nn.Linear(d_model, d_model), nn.SiLU(), nn.Linear(d_model, d_model))
            if self.cond_bias_mode not in {"memory", "encoder"}:
                raise ValueError(f"cond_bias_mode must be 'memory' or 'encoder', got {self.cond_bias_mode}")
            if self.cond_bias_mode == "encoder":
                self.cond_enc = MazeConditionEncoder(use_sdf=self.use_sdf, d_cond=d_model, use_start_goal=self.use_start_goal,
                                                     maze_channels=maze_channels)
        self.blocks = nn.ModuleList([CrossAttnBlock(d_model, n_heads, d_ff, dropout=dropout) for _ in range(max(1, n_layers))])
        self.out = nn.Linear(d_model, 1)
        self._ws = E.Workspace()

    @staticmethod
    def _mlp2(seq, x: torch.Tensor) -> torch.Tensor:
        f = lambda t: t.detach().float().contiguous()
        hdn = E.sgemm(x, f(seq[0].weight), f(seq[0].bias), act=1)
        return E.sgemm(hdn, f(seq[2].weight), f(seq[2].bias))

    @torch.no_grad()
    def forward(self, cond: Dict[str, torch.Tensor]) -> torch.Tensor:
        """keypoint_selector.py:148-188 -> logits fp32 [B, T]."""
        occ = L.f32c(cond["occ"])
        dev = L.require_cuda(occ)
        B, _, Hh, Ww = occ.shape
        d, T, H = self.d_model, self.T, self.n_heads
        f = lambda t: t.detach().float().contiguous()
        feats = [occ]
        if self.use_sdf:
            if cond.get("sdf") is None:
                raise ValueError("use_sdf is True but sdf missing from cond")
            feats.append(L.f32c(cond["sdf"]))
        need_sg = self.use_start_goal and (self.use_sg_map or self.use_sg_token)
        if (need_sg or self.use_goal_dist_token) and "start_goal" not in cond:
            raise ValueError("use_start_goal is True but start_goal missing from cond" if need_sg else "use_goal_dist_token requires start_goal")
        sg = L.f32c(cond["start_goal"]) if "start_goal" in cond else None
        if self.use_start_goal and self.use_sg_map:
            m = torch.empty((B, 2, Hh, Ww), device=dev, dtype=F32)
            L.call("idb200_sg_map", sg.data_ptr(), B, Hh, Ww, self.sg_map_sigma, m.data_ptr(), L.stream(dev))
            feats.append(m)
        x = torch.cat(feats, dim=1)
        convs = [c for c in self.spatial_conv if isinstance(c, nn.Conv2d)]
        _, _, us, _, _ = E.conv_stack_gemm(x, [c.weight for c in convs], [f(c.bias) for c in convs], self._ws, keep=True)
        P = Hh * Ww
        act = torch.empty_like(us[-1])                                   # SiLU of the last conv layer: the spatial features [B*P, C]
        L.call("idb200_silu_bf16", us[-1].data_ptr(), None, act.numel(), 0, act.data_ptr(), L.stream(dev))
        if isinstance(self.spatial_proj, nn.Conv2d):
            wp = self.spatial_proj.weight.detach().float().reshape(d, -1).to(BF16).contiguous()
            mem = torch.empty((B * P, d), device=dev, dtype=BF16)
            E.gemm_bf16(act, wp, f(self.spatial_proj.bias), mem, E.EPI_BF16)
        else:
            mem = act
        extras = []                                                       # the reference prepends them; key order is irrelevant
        if self.sg_token is not None:
            extras.append(self._mlp2(self.sg_token, sg))
        if self.goal_dist_token is not None:
            gd = torch.norm(sg[:, :2] - sg[:, 2:], dim=-1, keepdim=True).contiguous()
            extras.append(self._mlp2(self.goal_dist_token, gd))
        n_ex = len(extras)
        mem_ex = torch.stack(extras, dim=1).reshape(B * n_ex, d).to(BF16).contiguous() if n_ex else None
        # queries: time_proj(sinusoid(linspace(0, 1, T))) for every sample (+ the level embedding)
        pos = torch.linspace(0.0, 1.0, T).to(dev)
        emb = E.sinusoid(T, self.pos_dim - (self.pos_dim % 2), dev, args=pos.contiguous())
        if self.pos_dim % 2 == 1:
            emb = torch.nn.functional.pad(emb, (0, 1))
        q0 = E.sgemm(emb, f(self.time_proj.weight), f(self.time_proj.bias))                      # [T, d]
        q = q0.unsqueeze(0).expand(B, T, d).contiguous()
        if self.use_cond_bias:                                            # :170-175
            if self.cond_bias_mode == "encoder":
                cond_vec = self.cond_enc(cond)
            else:
                # mean over ALL memory tokens (extras + the H*W spatial tokens): fp32 token sums of the bf16 spatial memory
                tot = torch.empty((B, d), device=dev, dtype=F32)
                mem32 = mem.float()
                L.call("idb200_token_sum", mem32.data_ptr(), B, P, d, tot.data_ptr(), L.stream(dev))
                for e in extras:
                    tot = tot + e
                cond_vec = (tot / float(P + n_ex)).contiguous()
            q = q + self._mlp2(self.cond_bias, cond_vec).unsqueeze(1)
        if self.use_level:
            if "level" not in cond:
                raise ValueError("use_level is True but level missing from cond")
            level = L.f32c(cond["level"])
            if level.dim() == 1:
                level = level.unsqueeze(1)
            q = q + self._mlp2(self.level_mlp, level.contiguous()).unsqueeze(1)
        q = q.view(B * T, d).contiguous()
        h16 = torch.empty((B * T, d), device=dev, dtype=BF16)
        qp = torch.empty((B * T, d), device=dev, dtype=BF16)
        ao = torch.empty((B * T, d), device=dev, dtype=BF16)
        kv = torch.empty((B * P, 2 * d), device=dev, dtype=BF16)
        kv_ex = torch.empty((B * n_ex, 2 * d), device=dev, dtype=BF16) if n_ex else None
        for blk in self.blocks:
            W, bias = blk.attn.in_proj_weight.detach().float(), blk.attn.in_proj_bias.detach().float()
            wq, wkv = W[:d].to(BF16).contiguous(), W[d:].to(BF16).contiguous()
            E.ln_film(q, f(blk.norm1.weight), f(blk.norm1.bias), None, h16, T)
            E.gemm_bf16(h16, wq, bias[:d].contiguous(), qp, E.EPI_BF16)
            E.gemm_bf16(mem, wkv, bias[d:].contiguous(), kv, E.EPI_BF16)
            if n_ex:
                E.gemm_bf16(mem_ex, wkv, bias[d:].contiguous(), kv_ex, E.EPI_BF16)
            L.call("idb200_cross_attention", qp.data_ptr(), kv.data_ptr(), L.ptr(kv_ex), ao.data_ptr(), B, T, P, n_ex, H, L.stream(dev))
            E.gemm_bf16(ao, blk.attn.out_proj.weight.detach().to(BF16).contiguous(), f(blk.attn.out_proj.bias), q, E.EPI_RESID_F32)
            E.ln_film(q, f(blk.norm2.weight), f(blk.norm2.bias), None, h16, T)
            ff = torch.empty((B * T, blk.ff[0].weight.shape[0]), device=dev, dtype=BF16)
            E.gemm_bf16(h16, blk.ff[0].weight.detach().to(BF16).contiguous(), f(blk.ff[0].bias), ff, E.EPI_SILU_BF16)
            E.gemm_bf16(ff, blk.ff[2].weight.detach().to(BF16).contiguous(), f(blk.ff[2].bias), q, E.EPI_RESID_F32)
        logits = torch.empty((B * T, 1), device=dev, dtype=F32)
        E.out_head(q, f(self.out.weight), f(self.out.bias), logits)
        return logits.view(B, T)


def select_topk_indices(logits: torch.Tensor, K: int, stochastic: bool = False, tau: float = 1.0) -> torch.Tensor:
    """keypoint_selector.py:191-230: endpoints + the K - 2 best interior positions (optionally Gumbel-perturbed), sorted.
    Index selection only (torch.topk / sort on the logits' device), identical to the reference by construction."""
    if logits.dim() != 2:
        raise ValueError("logits must be [B,T]")
    B, T = logits.shape
    if K < 2:
        raise ValueError("K must be >= 2")
    if K > T:
        K = T
    if K == 2:
        idx = torch.zeros((B, 2), device=logits.device, dtype=torch.long)
        idx[:, 1] = T - 1
        return idx
    interior = logits[:, 1:-1]
    if stochastic:
        eps = 1e-6
        gumbel = -torch.log(-torch.log(torch.rand_like(interior).clamp_min(eps)).clamp_min(eps))
        tau = float(tau)
        if tau <= 0.0:
            tau = 1.0
        scores = (interior + gumbel) / tau
    else:
        scores = interior
    topk = torch.topk(scores, K - 2, dim=1).indices + 1
    idx = torch.cat([torch.zeros((B, 1), device=logits.device, dtype=torch.long), topk,
                     torch.full((B, 1), T - 1, device=logits.device, dtype=torch.long)], dim=1)
    return torch.sort(idx, dim=1).values
