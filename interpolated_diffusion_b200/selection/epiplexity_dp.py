"""Mirror of the DP anchor placement of the reference's ``src/selection/epiplexity_dp.py`` (``dp_select_indices`` :171-197,
``dp_select_indices_batch`` :200-228): the producer of ``idx`` when ``kp_index_mode = dp``.  One kernel, one block per sample."""
from __future__ import annotations

import torch

from .. import _lib as L


def dp_select_indices_batch(C: torch.Tensor, K: int) -> torch.Tensor:
    """C fp32 [B, T, T] segment costs (inf where no segment) -> idx i64 [B, min(K, T)], idx[:, 0] = 0, idx[:, -1] = T - 1."""
    if C.dim() != 3:
        raise ValueError("C must be [B,T,T]")
    dev = L.require_cuda(C)
    B, T, _ = C.shape
    if K < 2:
        raise ValueError("K must be >= 2")
    if K > T:
        K = T
    Cc = L.f32c(C)
    idx = torch.empty((B, K), device=dev, dtype=torch.long)
    status = torch.empty((B,), device=dev, dtype=torch.int32)
    L.call("idb200_dp_select", Cc.data_ptr(), B, T, K, idx.data_ptr(), status.data_ptr(), L.stream(dev))
    st = int(status.max().item()) if B > 0 else 0                  # the reference raises here too (:216, :223)
    if st == 1:
        raise RuntimeError("DP failed to find a valid path to T-1 for some samples.")
    if st == 2:
        raise RuntimeError("DP backtrack failed.")
    return idx


def dp_select_indices(C: torch.Tensor, K: int) -> torch.Tensor:
    """Single cost matrix [T, T] -> idx [K] (:171-197)."""
    if C.dim() != 2:
        raise ValueError("C must be [T,T]")
    try:
        return dp_select_indices_batch(C.unsqueeze(0), K)[0]
    except RuntimeError as e:
        raise RuntimeError(str(e).replace(" for some samples", "")) from None
