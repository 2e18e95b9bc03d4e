"""ctypes binding of libidb200.so (C ABI in include/idb200.h).

There is NO fallback: if the shared library is missing or a tensor is not on a CUDA device the
call raises.  PyTorch is used for device memory, streams and RNG only.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDB200_LIB") or os.path.join(_HERE, "libidb200.so")      # IDB200_LIB: dev A/B of two builds in one run

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_f = ctypes.c_float

# name -> argtypes (restype is always int unless listed in _SPECIAL)
_SIGS = {
    "idb200_nested_masks_interp": [c_p, c_p, c_l, c_l, c_i, c_i, c_i, ctypes.POINTER(c_i), c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_interpolate_from_indices": [c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p, c_p],
    "idb200_corrupt_from_anchors": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_f, c_f, c_i, c_i, c_i, c_p, c_p],
    "idb200_corrupt_adjacent": [c_p, c_p, c_p, c_l, c_i, c_i, c_i, ctypes.POINTER(c_f), ctypes.POINTER(c_f), c_p, c_p, c_i, ctypes.c_uint64,
                                ctypes.c_uint64, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "idb200_ddim_step": [c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_f, c_l, c_l, c_i, c_p, c_p, c_i, c_f, c_f, c_p, c_p],
    "idb200_q_sample": [c_p, c_p, c_p, c_p, c_p, c_i, c_l, c_l, c_p, c_p],
    "idb200_known_mask_values": [c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p],
    "idb200_logit_pos": [c_p, c_l, c_i, c_f, c_p, c_p],
    "idb200_sigmoid_pos": [c_p, c_l, c_i, c_p, c_p],
    "idb200_stage2_epilogue": [c_p, c_p, c_p, c_p, c_f, c_i, c_p, c_i, c_i, c_f, c_f, c_l, c_i, c_i, c_p, c_p],
    "idb200_gemm_bf16": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_conv_first_nhwc": [c_p, c_l, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_p],
    "idb200_conv3x3_gemm": [c_p, c_i, c_p, c_p, c_p, c_i, c_l, c_i, c_i, c_i, c_p],
    "idb200_pool_bordered": [c_p, c_l, c_i, c_i, c_i, c_p, c_p],
    "idb200_gemm_bf16_aux": [c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_gemm_bf16_dsilu_sums": [c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_sgemm": [c_p, c_i, c_l, c_p, c_p, c_p, c_l, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_conv_encoder": [c_p, c_p, c_l, c_i, c_i, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_p), ctypes.POINTER(c_p), c_p, c_p, c_p],
    "idb200_sinusoid": [c_p, c_i, c_i, c_i, c_p, c_p],
    "idb200_embed_tokens": [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_ln_film": [c_p, c_p, c_p, c_p, c_l, c_p, c_i, c_l, c_i, c_i, c_p],
    "idb200_ln_film_save": [c_p, c_p, c_p, c_p, c_l, c_p, c_i, c_p, c_l, c_i, c_i, c_p],
    "idb200_out_head": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_attention": [c_p, c_p, c_i, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_mlp_fused": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_attn_block": [c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_mlp_block": [c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_mlp_pair_w2_order": [c_i, ctypes.POINTER(c_i)],
    "idb200_ln_mlp_pair": [c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_mlp_pair": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_p],
    "idb200_ln_qkv_attention": [c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_qkv_attention": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_encoder_fused": [c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "idb200_denoiser_fused": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "idb200_conv_encoder_tc": [c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "idb200_traj_metrics": [c_p, c_l, c_p, c_p, c_l, c_p, c_l, c_l, c_i, c_i, c_i, c_i, c_f, c_p, c_i, c_p],
    "idb200_conv_encoder_tc5": [c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "idb200_stage2_loss": [c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_l, c_i, c_i, c_p, c_p, c_p, c_p],
    "idb200_grad_clip_coef": [c_p, c_l, c_f, c_p, c_p, c_p],
    "idb200_adamw_ema_step": [c_p, c_p, c_p, c_p, c_p, c_l, c_f, c_f, c_f, c_f, c_f, c_l, c_f, c_p, c_p],
    "idb200_gemm_bf16_splitk": [c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_gemm_bf16_nn_splitk": [c_p, c_p, c_p, c_l, c_i, c_l, c_i, c_p],
    "idb200_transpose_bf16": [c_p, c_i, c_l, c_i, c_p, c_p],
    "idb200_colsum_scratch_floats": [c_l, c_i],
    "idb200_tail_scratch_doubles": [],
    "idb200_colsum": [c_p, c_i, c_l, c_i, c_p, c_f, c_i, c_p, c_p],
    "idb200_multi_copy_f32": [ctypes.POINTER(c_p), ctypes.POINTER(c_p), ctypes.POINTER(c_l), c_i, c_p],
    "idb200_cast_weights_bf16": [ctypes.POINTER(c_p), ctypes.POINTER(c_p), ctypes.POINTER(c_p), ctypes.POINTER(c_i), ctypes.POINTER(c_i), c_i, c_p],
    "idb200_colsum_segments": [c_p, c_i, c_i, c_l, c_i, c_p, c_f, c_i, c_p, c_p],
    "idb200_reduce_rows": [c_p, c_i, c_l, c_f, c_i, c_p, c_p],
    "idb200_silu_bf16": [c_p, c_p, c_l, c_i, c_p, c_p],
    "idb200_silu_f32": [c_p, c_p, c_l, c_i, c_p, c_p],
    "idb200_ln_film_bwd": [c_p, c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_i, c_p, c_p, c_p, c_l, c_p, c_p, c_p],
    "idb200_ln_film_bwd2": [c_p, c_i, c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_i, c_p, c_p, c_p, c_l, c_p, c_i, c_p, c_p],
    "idb200_attention_bwd": [c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_attention_bwd_sums": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p],
    "idb200_head_bwd": [c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p],
    "idb200_narrow_outer_scratch_floats": [c_l, c_i, c_i],
    "idb200_narrow_outer": [c_p, c_i, c_p, c_l, c_i, c_p, c_i, c_p, c_p],
    "idb200_token_sum": [c_p, c_l, c_i, c_i, c_p, c_p],
    "idb200_sgemm_strided": [c_p, c_l, c_l, c_p, c_l, c_l, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_im2col3x3": [c_p, c_l, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "idb200_pool_silu": [c_p, c_l, c_i, c_i, c_p, c_p],
    "idb200_pool_silu_bwd": [c_p, c_p, c_l, c_i, c_i, c_p, c_p],
    "idb200_cross_attention": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_i, c_p],
    "idb200_sg_map": [c_p, c_l, c_i, c_i, c_f, c_p, c_p],
    "idb200_segment_costs": [c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p],
    "idb200_dp_select": [c_p, c_l, c_i, c_i, c_p, c_p, c_p],
    "idb200_anchor_conf": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_i, c_l, c_i, c_i, c_p, c_p, c_p],
}

EINVAL, EALIGN, EUNSUPPORTED, ECUDA = -1, -2, -3, -4

# bumped by every in-place parameter update made outside torch's version counters (train/optim.py: the fused AdamW kernel);
# part of the key of every packed-weight cache (models/_engine.py::_sig)
PARAM_EPOCH = 0


class EmbedDesc(ctypes.Structure):
    """idb200_embed_t (include/idb200.h)."""
    _fields_ = [("src0", c_p), ("n0", c_i), ("src1", c_p), ("n1", c_i), ("src2", c_p), ("n2", c_i), ("Wf", c_p), ("tab", c_p),
                ("tab_idx", c_p), ("row_a", c_p), ("row_a_stride", c_l), ("row_b", c_p), ("tab_rows", c_i)]


class HeadDesc(ctypes.Structure):
    """idb200_head_t (include/idb200.h)."""
    _fields_ = [("W", c_p), ("bias", c_p), ("y", c_p), ("D", c_i)]

_lib: Optional[ctypes.CDLL] = None


def declared_symbols():
    """Every symbol include/idb200.h declares (used by the CPU-side export test)."""
    return ["idb200_version", "idb200_last_error"] + list(_SIGS)


def register(name: str, argtypes) -> None:
    """Other modules of the package (denoiser kernels) add their entry points here."""
    _SIGS[name] = argtypes
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = c_i


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m interpolated_diffusion_b200.csrc.build` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.idb200_version.restype = c_i
        L.idb200_last_error.restype = ctypes.c_char_p
        for name, argtypes in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = c_i
        _lib = L
    return _lib


def last_error() -> str:
    return lib().idb200_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        msg = last_error()
        if rc == EINVAL:
            raise ValueError(msg)
        raise RuntimeError(f"{name} failed ({rc}): {msg}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("interpolated_diffusion_b200 runs on CUDA (sm_100a) only; got a tensor on "
                               f"{t.device}.  There is no CPU fallback.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


def resolve_device(device) -> torch.device:
    """Reference default is CPU; this package is CUDA-only, so None means the current CUDA device."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if device is None:
        raise RuntimeError("interpolated_diffusion_b200 needs a CUDA device (no CPU fallback)")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"interpolated_diffusion_b200 runs on CUDA only (got device={device}); no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


def u8c(t: torch.Tensor) -> torch.Tensor:
    """bool / uint8 tensor as contiguous 1-byte storage (torch.bool is 0/1 bytes)."""
    if t.dtype not in (torch.bool, torch.uint8):
        t = t != 0
    return t.contiguous()
