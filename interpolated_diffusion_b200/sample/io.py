"""The sampler's on-disk output (``samples.npz``, sample_generate.py:1665-1689): same keys, shapes and dtypes, written from the
batched tensors of ``generate(..., return_all=True)`` instead of per-sample python lists."""
import os
from typing import Optional

import numpy as np
import torch


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save_samples_npz(out_dir: str, *, interp, refined, keypoints, idx, mask, start_goal, occ=None, sdf=None, gt=None,
                     difficulty=None, interp_steps=None, filename: str = "samples.npz") -> str:
    """interp [n,T,D] (Stage-1 keypoints interpolated), refined [n,T,D] (after Stage 2 + clamp), keypoints [n,K,D], idx i64 [n,K],
    mask bool [n,T], start_goal [n,4]; optional occ [n,H,W] (or one [H,W] map), sdf, gt [n,T,D], difficulty i64 [n]."""
    os.makedirs(out_dir, exist_ok=True)
    kw = {"interp": _np(interp), "refined": _np(refined), "keypoints": _np(keypoints), "idx": _np(idx).astype(np.int64),
          "mask": _np(mask), "start_goal": _np(start_goal)}
    if gt is not None:
        kw["gt"] = _np(gt)
    if difficulty is not None:
        kw["difficulty"] = _np(difficulty).astype(np.int64)
    if occ is not None:
        kw["occ"] = _np(occ)
    if sdf is not None:
        kw["sdf"] = _np(sdf)
    if interp_steps is not None:
        kw["interp_steps"] = _np(interp_steps)
    path = os.path.join(out_dir, filename)
    np.savez_compressed(path, **kw)
    return path
