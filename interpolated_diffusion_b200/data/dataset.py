"""Reader of the reference's prepared-dataset format (``src/data/dataset.py:682-747``, ``PreparedTrajectoryDataset``): an ``.npz`` with
``x`` [N,T,D], ``start_goal`` [N,4], ``occ`` ([H,W] shared or [N,H,W]) and optionally ``sdf``, ``kp_idx`` [N,K], ``kp_feat``
[N,K,F], ``kp_mask_levels`` ([S+1,T] shared or [N,S+1,T]), ``difficulty`` [N].  ``__getitem__`` returns the reference's sample dict
(so a torch ``DataLoader`` works unchanged); ``batch()`` is the GPU-side path: one gather per array from pinned host memory and
asynchronous copies to the device, no per-sample python objects."""
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import Dataset


class PreparedTrajectoryDataset(Dataset):
    def __init__(self, path: str, use_sdf: bool = False, pin_memory: bool = False):
        data = np.load(path)
        has = lambda k: k in data.files
        self.x = torch.from_numpy(np.ascontiguousarray(data["x"].astype(np.float32)))
        self.start_goal = torch.from_numpy(np.ascontiguousarray(data["start_goal"].astype(np.float32)))
        n = self.x.shape[0]

        def plane(a: np.ndarray, dtype) -> torch.Tensor:            # [H,W] shared map -> [1,H,W]
            a = a.astype(dtype)
            return torch.from_numpy(np.ascontiguousarray(a[None, ...] if a.ndim == 2 else a))

        self.occ = plane(data["occ"], np.float32)
        self.occ_per_sample = self.occ.shape[0] == n and self.occ.dim() == 3
        self.sdf = plane(data["sdf"], np.float32) if (use_sdf and has("sdf")) else None
        self.sdf_per_sample = self.sdf is not None and self.sdf.dim() == 3 and self.sdf.shape[0] == n
        self.difficulty = torch.from_numpy(data["difficulty"].astype(np.int64)) if has("difficulty") else None
        self.kp_idx = torch.from_numpy(data["kp_idx"].astype(np.int64)) if has("kp_idx") else None
        self.kp_feat = torch.from_numpy(data["kp_feat"].astype(np.float32)) if has("kp_feat") else None
        self.kp_mask_levels = torch.from_numpy(data["kp_mask_levels"].astype(np.bool_)) if has("kp_mask_levels") else None
        self.kp_mask_levels_per_sample = (self.kp_mask_levels is not None and self.kp_mask_levels.dim() == 3
                                          and self.kp_mask_levels.shape[0] == n)
        if pin_memory and torch.cuda.is_available():
            for k in ("x", "start_goal", "occ", "sdf", "kp_idx", "kp_feat", "kp_mask_levels"):
                t = getattr(self, k)
                if t is not None:
                    setattr(self, k, t.pin_memory())

    def __len__(self) -> int:
        return self.x.shape[0]

    def __getitem__(self, idx: int) -> Dict:
        occ = self.occ[idx] if self.occ_per_sample else self.occ[0]
        sample = {"x": self.x[idx], "cond": {"occ": occ[None, ...], "start_goal": self.start_goal[idx]}}
        cond = sample["cond"]
        if self.kp_idx is not None:
            cond["kp_idx"] = self.kp_idx[idx]
        if self.kp_feat is not None:
            cond["kp_feat"] = self.kp_feat[idx]
        if self.kp_mask_levels is not None:
            cond["kp_mask_levels"] = self.kp_mask_levels[idx] if self.kp_mask_levels_per_sample else self.kp_mask_levels
        if self.difficulty is not None:
            sample["difficulty"] = torch.tensor(int(self.difficulty[idx]), dtype=torch.int64)
        if self.sdf is not None:
            sdf = self.sdf[idx] if self.sdf_per_sample else self.sdf[0]
            cond["sdf"] = sdf[None, ...]
        return sample

    def batch(self, indices: Union[Sequence[int], torch.Tensor], device: Optional[torch.device] = None) -> Dict:
        """{"x": [B,T,D], "cond": {"occ": [B,1,H,W], "start_goal": [B,4], ...}} for the given rows, on ``device`` if given (what the
        default collate of a DataLoader would produce from ``__getitem__``)."""
        ii = torch.as_tensor(indices, dtype=torch.long)
        B = ii.numel()
        mv = (lambda t: t.to(device, non_blocking=True)) if device is not None else (lambda t: t)

        def maps(t: torch.Tensor, per_sample: bool) -> torch.Tensor:
            m = t[ii] if per_sample else t[:1].expand(B, -1, -1)
            return mv(m.unsqueeze(1).contiguous())

        out = {"x": mv(self.x[ii]), "cond": {"occ": maps(self.occ, self.occ_per_sample), "start_goal": mv(self.start_goal[ii])}}
        cond = out["cond"]
        if self.kp_idx is not None:
            cond["kp_idx"] = mv(self.kp_idx[ii])
        if self.kp_feat is not None:
            cond["kp_feat"] = mv(self.kp_feat[ii])
        if self.kp_mask_levels is not None:
            m = self.kp_mask_levels[ii] if self.kp_mask_levels_per_sample else self.kp_mask_levels[None].expand(B, -1, -1)
            cond["kp_mask_levels"] = mv(m.contiguous())
        if self.difficulty is not None:
            out["difficulty"] = mv(self.difficulty[ii])
        if self.sdf is not None:
            cond["sdf"] = maps(self.sdf, self.sdf_per_sample)
        return out
