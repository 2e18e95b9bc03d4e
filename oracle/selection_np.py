"""numpy restatement of ``src/selection/epiplexity_dp.py:200-228`` (``dp_select_indices_batch``).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the live reference by
``tests/golden/make_golden_dp.py`` -> ``tests/golden/dp_select.npz``."""
import numpy as np


def dp_select_indices_batch(C: np.ndarray, K: int) -> np.ndarray:
    B, T, _ = C.shape
    if K < 2:
        raise ValueError("K must be >= 2")
    K = min(K, T)
    C = C.astype(np.float32)
    dp = np.full((B, K, T), np.inf, np.float32)
    parent = np.full((B, K, T), -1, np.int64)
    dp[:, 0, 0] = 0.0
    rows = np.arange(B)
    for k in range(1, K):                                           # :210-215
        for j in range(1, T):
            prev = dp[:, k - 1, :j] + C[:, :j, j]
            best = np.argmin(prev, axis=1)                          # first minimum, like torch.argmin
            dp[:, k, j] = prev[rows, best]
            parent[:, k, j] = best
    if not np.isfinite(dp[:, K - 1, T - 1]).all():
        raise RuntimeError("DP failed to find a valid path to T-1 for some samples.")
    idx = np.zeros((B, K), np.int64)
    idx[:, -1] = T - 1
    cur = np.full((B,), T - 1, np.int64)
    for k in range(K - 1, 0, -1):                                   # :221-226
        cur = parent[rows, k, cur]
        if (cur < 0).any():
            raise RuntimeError("DP backtrack failed.")
        idx[:, k - 1] = cur
    return idx


def compute_segment_costs_batch(x_pos: np.ndarray, seg_i, seg_j, t_idx, alpha, weight, weight_scale: float = 1.0) -> np.ndarray:
    """``epiplexity_dp.py:120-147``: squared deviation of the trajectory from the chord (i, j) at the sample points."""
    x = x_pos[..., :2].astype(np.float32)
    x_i, x_j = x[:, seg_i], x[:, seg_j]                                     # [B, S, 2]
    mu = x_i[:, :, None, :] + alpha[None, :, :, None].astype(np.float32) * (x_j - x_i)[:, :, None, :]
    x_t = x[:, t_idx.reshape(-1)].reshape(x.shape[0], t_idx.shape[0], t_idx.shape[1], 2)
    diff = x_t - mu
    cost = (diff * diff).sum(-1).sum(-1) * weight[None, :].astype(np.float32)
    if weight_scale != 1.0:
        cost = cost * np.float32(weight_scale)
    return cost.astype(np.float32)
