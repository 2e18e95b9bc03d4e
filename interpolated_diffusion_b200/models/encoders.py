"""Mirror of the reference's ``src/models/encoders.py`` (MazeEncoder / StartGoalEncoder / MazeConditionEncoder):
same parameters, forward on libidb200 (fp32 conv stack with planes resident in shared memory, fp32 SIMT linears)."""
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib as L
from . import _engine as E


class MazeEncoder(nn.Module):
    def __init__(self, in_channels: int, d_cond: int = 128, channels: tuple = (32, 64)):
        super().__init__()
        if len(channels) == 0:
            raise ValueError("channels must be non-empty")
        layers = []
        c_in = in_channels
        for c_out in channels:
            layers.append(nn.Conv2d(c_in, c_out, kernel_size=3, padding=1))
            layers.append(nn.SiLU())
            c_in = c_out
        self.convs = nn.Sequential(*layers)
        self.fc = nn.Linear(channels[-1], d_cond)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """encoders.py:22-25; x [B, C_in, H, W].  ``self.precision == "bf16"`` (default) runs the two-layer stack as a
        tensor-core implicit GEMM; "fp32" (check mode) and other depths use the fp32 direct-conv kernels."""
        L.require_cuda(x)
        x = L.f32c(x)
        occ = x[:, 0:1].contiguous()
        sdf = x[:, 1:2].contiguous() if x.shape[1] > 1 else None
        convs = [m for m in self.convs if isinstance(m, nn.Conv2d)]
        fcw, fcb = self.fc.weight.detach().float().contiguous(), self.fc.bias.detach().float().contiguous()
        if getattr(self, "precision", "bf16") == "bf16" and E.conv_tc_supported(convs, x.shape[2], x.shape[3]):
            key = E._sig([convs[1].weight])
            if getattr(self, "_w1_key", None) != key:
                w = convs[1].weight.detach().float()
                self._w1_packed = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
                self._w1_key = key
            conv = E.conv_encoder_tc5 if (getattr(self, "use_tc5", True) and E.conv_tc5_supported(convs, x.shape[2], x.shape[3])) else E.conv_encoder_tc
            pooled = conv(occ, sdf, convs[0].weight.detach().float().contiguous(), convs[0].bias.detach().float().contiguous(),
                          self._w1_packed, convs[1].bias.detach().float().contiguous())
        elif getattr(self, "precision", "bf16") == "bf16" and getattr(self, "use_implicit", True) and E.conv_implicit_supported(convs):
            # deeper / wider stacks (the trainer default 32,64,128,128): tap-shifted implicit GEMM per layer, no im2col matrix
            if not hasattr(self, "_gemm_ws"):
                self._gemm_ws = E.Workspace()
            pooled = E.conv_stack_implicit(x, convs, self._gemm_ws)
        elif getattr(self, "precision", "bf16") == "bf16" and E.conv_gemm_supported(convs):
            # deeper / wider stacks (the trainer default 32,64,128,128): im2col + tcgen05 GEMM per layer
            if not hasattr(self, "_gemm_ws"):
                self._gemm_ws = E.Workspace()
            pooled = E.conv_stack_gemm(x, [c.weight for c in convs], [c.bias.detach().float().contiguous() for c in convs], self._gemm_ws)
        else:
            pooled = E.conv_encoder(occ, sdf, [c.weight.detach().float().contiguous() for c in convs],
                                    [c.bias.detach().float().contiguous() for c in convs])
        return E.sgemm(pooled, fcw, fcb)


class StartGoalEncoder(nn.Module):
    def __init__(self, d_cond: int = 128):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(4, d_cond), nn.SiLU(), nn.Linear(d_cond, d_cond))

    @torch.no_grad()
    def forward(self, start_goal: torch.Tensor, *, out: torch.Tensor = None) -> torch.Tensor:
        """encoders.py:37-38 (``out=``: accumulate into an existing embedding)."""
        L.require_cuda(start_goal)
        hdn = E.sgemm(L.f32c(start_goal), self.mlp[0].weight.detach().float().contiguous(),
                      self.mlp[0].bias.detach().float().contiguous(), act=1)
        return E.sgemm(hdn, self.mlp[2].weight.detach().float().contiguous(), self.mlp[2].bias.detach().float().contiguous(),
                       out, accumulate=out is not None)


class MazeConditionEncoder(nn.Module):
    def __init__(self, use_sdf: bool = False, d_cond: int = 128, use_start_goal: bool = True, maze_channels: tuple = (32, 64)):
        super().__init__()
        in_channels = 2 if use_sdf else 1
        self.use_sdf = use_sdf
        self.use_start_goal = use_start_goal
        self.maze = MazeEncoder(in_channels, d_cond, channels=maze_channels)
        self.sg = StartGoalEncoder(d_cond) if use_start_goal else None

    @torch.no_grad()
    def forward(self, cond: Dict[str, torch.Tensor]) -> torch.Tensor:
        """encoders.py:56-71 -> cond_vec fp32 [B, d_cond]."""
        occ = cond["occ"]
        if self.use_sdf:
            sdf = cond.get("sdf")
            if sdf is None:
                raise ValueError("use_sdf is True but sdf missing from cond")
            x = torch.cat([occ, sdf], dim=1)
        else:
            x = occ
        self.maze.precision = getattr(self, "precision", "bf16")
        emb = self.maze(x)
        if self.use_start_goal:
            if "start_goal" not in cond:
                raise ValueError("use_start_goal is True but start_goal missing from cond")
            self.sg(cond["start_goal"], out=emb)
        return emb
