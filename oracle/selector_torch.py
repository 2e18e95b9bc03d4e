"""torch-CPU fp32 functional restatement of ``src/models/keypoint_selector.py`` (``KeypointSelector.forward`` :148-188,
``CrossAttnBlock`` :22-38, ``_sg_map`` :113-146, ``select_topk_indices`` :191-): the producer of anchor logits.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Driven by a reference-layout ``state_dict``; pinned against the live
reference by ``tests/golden/make_golden_selector.py`` -> ``tests/golden/selector.npz``."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .denoiser_torch import _lin, sinusoid

SD = Dict[str, torch.Tensor]


def sg_map(start_goal: torch.Tensor, H: int, W: int, sigma: float) -> torch.Tensor:
    y = torch.arange(H, dtype=torch.float32)
    x = torch.arange(W, dtype=torch.float32)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    xx, yy = xx.unsqueeze(0), yy.unsqueeze(0)
    c = [start_goal[:, i].clamp(0.0, 1.0) * float((W if i % 2 == 0 else H) - 1) for i in range(4)]
    sx, sy, gx, gy = [v.view(-1, 1, 1) for v in c]
    s2 = float(sigma) ** 2
    s_map = torch.exp(-((xx - sx) ** 2 + (yy - sy) ** 2) / (2.0 * s2))
    g_map = torch.exp(-((xx - gx) ** 2 + (yy - gy) ** 2) / (2.0 * s2))
    return torch.stack([s_map, g_map], dim=1)


def _mlp2(x, sd, name):
    return _lin(F.silu(_lin(x, sd, name + ".0")), sd, name + ".2")


def cross_attn_block(q: torch.Tensor, kv: torch.Tensor, sd: SD, pre: str, n_heads: int) -> torch.Tensor:
    B, Lq, d = q.shape
    Lk = kv.shape[1]
    hd = d // n_heads
    W, bias = sd[pre + "attn.in_proj_weight"], sd[pre + "attn.in_proj_bias"]
    h = F.layer_norm(q, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    qq = h @ W[:d].t() + bias[:d]
    kk = kv @ W[d:2 * d].t() + bias[d:2 * d]
    vv = kv @ W[2 * d:].t() + bias[2 * d:]
    sh = lambda t, L: t.view(B, L, n_heads, hd).transpose(1, 2)
    p = torch.softmax((sh(qq, Lq) / math.sqrt(hd)) @ sh(kk, Lk).transpose(-1, -2), dim=-1)
    o = (p @ sh(vv, Lk)).transpose(1, 2).reshape(B, Lq, d)
    x = q + _lin(o, sd, pre + "attn.out_proj")
    h = F.layer_norm(x, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    return x + _lin(F.silu(_lin(h, sd, pre + "ff.0")), sd, pre + "ff.2")


def keypoint_selector(sd: SD, cond: Dict[str, torch.Tensor], *, T: int, n_heads: int, pos_dim: int, use_sdf: bool = False,
                      use_sg_map: bool = True, sg_map_sigma: float = 1.5) -> torch.Tensor:
    """-> logits [B, T].  Optional parts are detected from the state dict (sg_token, goal_dist_token, level_mlp); cond_bias is not
    restated (unsupported by the CUDA mirror)."""
    occ = cond["occ"].float()
    feats = [occ]
    if use_sdf:
        feats.append(cond["sdf"].float())
    if use_sg_map:
        feats.append(sg_map(cond["start_goal"].float(), occ.shape[-2], occ.shape[-1], sg_map_sigma))
    x = torch.cat(feats, dim=1)
    ids = sorted({int(k.split(".")[1]) for k in sd if k.startswith("spatial_conv.")})
    for i in ids:
        x = F.silu(F.conv2d(x, sd[f"spatial_conv.{i}.weight"], sd[f"spatial_conv.{i}.bias"], padding=1))
    if "spatial_proj.weight" in sd:
        x = F.conv2d(x, sd["spatial_proj.weight"], sd["spatial_proj.bias"])
    B = x.shape[0]
    tokens = [x.flatten(2).transpose(1, 2)]
    if "sg_token.0.weight" in sd:
        tokens.insert(0, _mlp2(cond["start_goal"].float(), sd, "sg_token").unsqueeze(1))
    if "goal_dist_token.0.weight" in sd:
        sg = cond["start_goal"].float()
        tokens.insert(0, _mlp2(torch.norm(sg[:, :2] - sg[:, 2:], dim=-1, keepdim=True), sd, "goal_dist_token").unsqueeze(1))
    memory = torch.cat(tokens, dim=1)
    q = _lin(sinusoid(torch.linspace(0.0, 1.0, T), pos_dim), sd, "time_proj").unsqueeze(0).expand(B, -1, -1)
    if "level_mlp.0.weight" in sd:
        level = cond["level"].float()
        if level.dim() == 1:
            level = level.unsqueeze(1)
        q = q + _mlp2(level, sd, "level_mlp").unsqueeze(1)
    n_blocks = max(int(k.split(".")[1]) for k in sd if k.startswith("blocks.")) + 1
    for i in range(n_blocks):
        q = cross_attn_block(q, memory, sd, f"blocks.{i}.", n_heads)
    return _lin(q, sd, "out").squeeze(-1)


def segment_cost_predictor(sd: SD, cond: Dict[str, torch.Tensor], seg_feat: torch.Tensor) -> torch.Tensor:
    """``src/models/segment_cost.py:43-57``: MLP over [cond_vec | seg_feat] for every segment -> [B, S]."""
    from .denoiser_torch import cond_encoder
    cond_vec = cond_encoder(sd, cond)
    if seg_feat.dim() == 2:
        seg_feat = seg_feat.unsqueeze(0).expand(cond_vec.shape[0], -1, -1)
    x = torch.cat([cond_vec.unsqueeze(1).expand(-1, seg_feat.shape[1], -1), seg_feat.float()], dim=-1)
    ids = sorted({int(k.split(".")[1]) for k in sd if k.startswith("mlp.") and k.endswith(".weight")})
    for n, i in enumerate(ids):
        x = _lin(x, sd, f"mlp.{i}")
        if n < len(ids) - 1:
            x = F.silu(x)
    return x.squeeze(-1)
