"""torch-CPU fp32 functional restatement of the two denoisers and their sub-modules:
``src/models/{transformer,encoders,denoiser_keypoints,denoiser_interp_levels,
denoiser_interp_levels_causal}.py``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Everything is driven by a reference-layout ``state_dict`` so the same
weights feed the oracle and the CUDA path.  ``nn.MultiheadAttention`` is spelled out
(packed QKV projection, per-head scaled softmax, output projection) instead of called.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _lin(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def sinusoid(args_scalar: torch.Tensor, dim: int) -> torch.Tensor:
    """Shared body of ``timestep_embedding`` / ``continuous_time_embedding`` /
    ``_positional_embedding`` (denoiser_keypoints.py:11-34, denoiser_interp_levels.py:54-62):
    freqs = exp(-ln(1e4) * i / half), emb = [sin(a*f), cos(a*f)] (+ zero pad if dim is odd)."""
    half = dim // 2
    freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(0, half) / half)
    args = args_scalar.float().unsqueeze(-1) * freqs
    emb = torch.cat([torch.sin(args), torch.cos(args)], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def n_layers_of(sd: SD, prefix: str = "transformer.layers.") -> int:
    ids = {int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)}
    return max(ids) + 1 if ids else 0


def cond_encoder(sd: SD, cond: Dict[str, torch.Tensor], prefix: str = "cond_enc.") -> torch.Tensor:
    """encoders.py:8-71: [conv3x3 pad1 + SiLU]*n -> mean over HxW -> fc, + start/goal MLP."""
    conv_ids = sorted({int(k[len(prefix + "maze.convs."):].split(".")[0])
                       for k in sd if k.startswith(prefix + "maze.convs.")})
    in_ch = sd[f"{prefix}maze.convs.{conv_ids[0]}.weight"].shape[1]
    x = cond["occ"].float()
    if in_ch == 2:                                            # use_sdf (encoders.py:58-63)
        if cond.get("sdf") is None:
            raise ValueError("use_sdf is True but sdf missing from cond")
        x = torch.cat([x, cond["sdf"].float()], dim=1)
    for i in conv_ids:
        x = F.silu(F.conv2d(x, sd[f"{prefix}maze.convs.{i}.weight"], sd[f"{prefix}maze.convs.{i}.bias"], padding=1))
    x = x.mean(dim=[2, 3])
    emb = _lin(x, sd, prefix + "maze.fc")
    if (prefix + "sg.mlp.0.weight") in sd:                    # use_start_goal (encoders.py:66-70)
        if "start_goal" not in cond:
            raise ValueError("use_start_goal is True but start_goal missing from cond")
        sg = _lin(F.silu(_lin(cond["start_goal"].float(), sd, prefix + "sg.mlp.0")), sd, prefix + "sg.mlp.2")
        emb = emb + sg
    return emb


def mha(h: torch.Tensor, sd: SD, pre: str, n_heads: int, attn_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.MultiheadAttention(batch_first=True) self-attention (transformer.py:11,39)."""
    B, L, d = h.shape
    hd = d // n_heads
    qkv = h @ sd[pre + "attn.in_proj_weight"].t() + sd[pre + "attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(B, L, n_heads, hd).transpose(1, 2)
    k = k.view(B, L, n_heads, hd).transpose(1, 2)
    v = v.view(B, L, n_heads, hd).transpose(1, 2)
    scores = (q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2)
    if attn_mask is not None:
        scores = scores + attn_mask
    p = torch.softmax(scores, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, d)
    return o @ sd[pre + "attn.out_proj.weight"].t() + sd[pre + "attn.out_proj.bias"]


def film(x: torch.Tensor, cond_vec: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    """transformer.py:28-33: x*(1+gamma)+beta, [gamma|beta] = Linear(cond)."""
    gb = _lin(cond_vec, sd, name)
    gamma, beta = gb.chunk(2, dim=-1)
    return x * (1.0 + gamma.unsqueeze(1)) + beta.unsqueeze(1)


def transformer_encoder(h: torch.Tensor, cond_vec: torch.Tensor, sd: SD, n_heads: int, causal: bool,
                        prefix: str = "transformer.layers.") -> torch.Tensor:
    """transformer.py:35-46, 73-82 (pre-LN, FiLM, MHA, SiLU-MLP; eps=1e-5; no final norm)."""
    L, d = h.shape[1], h.shape[2]
    attn_mask = None
    if causal:                                               # transformer.py:68-71
        attn_mask = torch.triu(torch.ones(L, L), diagonal=1)
        attn_mask = attn_mask.masked_fill(attn_mask == 1, float("-inf"))
    for i in range(n_layers_of(sd, prefix)):
        pre = f"{prefix}{i}."
        a = F.layer_norm(h, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
        if (pre + "film1.weight") in sd:
            a = film(a, cond_vec, sd, pre + "film1")
        h = h + mha(a, sd, pre, n_heads, attn_mask)
        a = F.layer_norm(h, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
        if (pre + "film2.weight") in sd:
            a = film(a, cond_vec, sd, pre + "film2")
        h = h + _lin(F.silu(_lin(a, sd, pre + "ff.0")), sd, pre + "ff.2")
    return h


def keypoint_denoiser(sd: SD, n_heads: int, z_t: torch.Tensor, t: torch.Tensor, idx: torch.Tensor,
                      known_mask: torch.Tensor, cond: Dict[str, torch.Tensor], T: int,
                      cond_vec: Optional[torch.Tensor] = None) -> torch.Tensor:
    """denoiser_keypoints.py:82-113.  ``cond_vec`` may be passed to reuse a hoisted encoder output."""
    B, K, D = z_t.shape
    d = sd["in_proj.weight"].shape[0]
    fan_in = sd["in_proj.weight"].shape[1]
    kp_feat_dim = 0
    pos_dim = d // 2
    if fan_in != 2 * D + pos_dim:
        kp_feat_dim = fan_in - 2 * D - pos_dim
    pos = idx.float() / max(1.0, float(T - 1))
    pos_emb = sinusoid(pos, pos_dim)
    if kp_feat_dim > 0 and cond is not None and "kp_feat" in cond:
        kp_feat = cond["kp_feat"].float()
    else:
        kp_feat = torch.zeros((B, K, kp_feat_dim))
    x = torch.cat([z_t.float(), pos_emb, known_mask.float(), kp_feat], dim=-1)
    h = _lin(x, sd, "in_proj")
    t_emb = sinusoid(t, d)
    t_emb = _lin(F.silu(_lin(t_emb, sd, "t_embed.0")), sd, "t_embed.2")
    h = h + t_emb.unsqueeze(1)
    if cond_vec is None:
        cond_vec = cond_encoder(sd, cond)
    h = h + _lin(cond_vec, sd, "cond_proj").unsqueeze(1)
    h = transformer_encoder(h, cond_vec, sd, n_heads, causal=False)
    return _lin(h, sd, "out")


def interp_level_denoiser(sd: SD, n_heads: int, x_s: torch.Tensor, s: torch.Tensor, mask: torch.Tensor,
                          cond: Dict[str, torch.Tensor], causal: bool = False,
                          cond_vec: Optional[torch.Tensor] = None) -> torch.Tensor:
    """denoiser_interp_levels.py:64-84 (causal twin: denoiser_interp_levels_causal.py:49)."""
    B, T, D = x_s.shape
    d = sd["in_proj.weight"].shape[0]
    mask_in = mask.unsqueeze(-1).float() if mask.dim() == 2 else mask.float()
    C = sd["in_proj.weight"].shape[1] - D
    if mask_in.shape[-1] != C:
        raise ValueError(f"mask has {mask_in.shape[-1]} channels, expected {C}")
    h = _lin(torch.cat([x_s.float(), mask_in], dim=-1), sd, "in_proj")
    h = h + sinusoid(torch.linspace(0.0, 1.0, T), d).unsqueeze(0)
    level = _lin(F.silu(_lin(sd["level_emb.weight"][s], sd, "level_proj.0")), sd, "level_proj.2")
    h = h + level.unsqueeze(1)
    if cond_vec is None:
        cond_vec = cond_encoder(sd, cond)
    h = h + _lin(cond_vec, sd, "cond_proj").unsqueeze(1)
    h = transformer_encoder(h, cond_vec, sd, n_heads, causal=causal)
    return _lin(h, sd, "out")
