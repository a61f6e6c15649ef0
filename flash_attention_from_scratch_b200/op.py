"""The operator: `forward(kernel_cfg, q, k, v, o=None)` / `forward_timed(...)`.

Same signature and error behaviour as the reference's Python operator
(/root/reference/flash_attention/__init__.py:7-17) and the checks of its C++ launcher
(/root/reference/src/flash_attention.cu:34-98), but the work is done by libfa_sm100.so through
its C ABI (include/fa_sm100.h).  PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_DTYPE_CODE = {torch.float16: _lib.FA_DTYPE_FP16, torch.bfloat16: _lib.FA_DTYPE_BF16}


def _check_input(t: torch.Tensor, name: str) -> None:
    # CHECK_INPUT, /root/reference/src/include/cuda_utils.cuh:5-11
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _cfg_dtype(kernel_cfg):
    """The reference reads `cfg.dtype.to_torch_dtype()` (flash_attention.cu:16-19).  Every other
    field of the config describes mma.sync tiling and is ignored here; `None` is accepted."""
    if kernel_cfg is None:
        return None
    dt = getattr(kernel_cfg, "dtype", None)
    if dt is None:
        return None
    if isinstance(dt, torch.dtype):
        return dt
    return dt.to_torch_dtype()


def _prepare(kernel_cfg, q, k, v, o):
    _check_input(q, "q")
    _check_input(k, "k")
    _check_input(v, "v")
    if q.dtype not in _DTYPE_CODE:
        raise RuntimeError("Only fp16 and bf16 are supported")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise RuntimeError("Input tensors must have the same data type")
    if q.dim() != 4:
        raise RuntimeError("Expected (batch, seq_len, n_heads, d_head) tensors")
    cfg_dtype = _cfg_dtype(kernel_cfg)
    if cfg_dtype is not None and cfg_dtype != q.dtype:
        raise RuntimeError("Kernel configuration dtype does not match input dtype")
    cfg_d = getattr(kernel_cfg, "d_head", None) if kernel_cfg is not None else None
    if q.size(3) != 128 or (cfg_d is not None and cfg_d != q.size(3)):
        raise RuntimeError("Kernel configuration was not found in flash_kernels.cuh")
    if q.shape != k.shape:
        raise RuntimeError("Query and key tensors have same shape")
    if q.shape != v.shape:
        raise RuntimeError("Query and value tensors have same shape")
    # The reference requires seq_len % B_r == 0 and % B_c == 0 (flash_attention.cu:79-82).  Kept
    # verbatim when a reference-style config is passed; with kernel_cfg=None any seq_len >= 1 runs
    # (the kernel masks the ragged tail), an extension beyond the reference.
    b_r = getattr(kernel_cfg, "B_r", None) if kernel_cfg is not None else None
    b_c = getattr(kernel_cfg, "B_c", None) if kernel_cfg is not None else None
    if b_r and q.size(1) % b_r != 0:
        raise RuntimeError("Only multiples of B_r are supported for seq_len Q currently")
    if b_c and q.size(1) % b_c != 0:
        raise RuntimeError("Only multiples of B_c are supported for seq_len K currently")
    if q.size(1) < 1:
        raise RuntimeError("seq_len must be positive")
    if o is not None:
        if o.dtype != q.dtype:
            raise RuntimeError("Output tensor must have the same dtype as inputs")
        # the reference re-checks q vs v here by mistake (flash_attention.cu:94); check o properly
        if o.shape != q.shape or o.stride() != q.stride() or not o.is_cuda:
            raise RuntimeError("Query and output tensors have same shape")
    else:
        o = torch.empty_like(q)
    return o


def _args(q, k, v, o):
    B, N, H, D = q.shape
    sb, sn, sh, _ = q.stride()
    return (q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, D, sb, sn, sh,
            _DTYPE_CODE[q.dtype])


def forward(kernel_cfg, q, k, v, o=None):
    """O = softmax(Q K^T / sqrt(d)) V; asynchronous on the current stream; returns O."""
    lib = _lib.load()
    o = _prepare(kernel_cfg, q, k, v, o)
    with torch.cuda.device(q.device), _lib.thread_kernel_mode(getattr(kernel_cfg, "cta_group", 0)):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.fa_fwd(*_args(q, k, v, o), stream)
    if rc != 0:
        raise RuntimeError(_lib.last_error())
    return o


def forward_timed(kernel_cfg, q, k, v, o=None):
    """Like `forward` but synchronises and also returns the kernel time in ms (cudaEvent pair
    around the launch inside the library, as flash_attention.cu:119-132)."""
    lib = _lib.load()
    o = _prepare(kernel_cfg, q, k, v, o)
    ms = C.c_float(0.0)
    with torch.cuda.device(q.device), _lib.thread_kernel_mode(getattr(kernel_cfg, "cta_group", 0)):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.fa_fwd_timed(*_args(q, k, v, o), stream, C.byref(ms))
    if rc != 0:
        raise RuntimeError(_lib.last_error())
    return o, ms.value


def forward_host(q, k, v, o=None, device: int = 0):
    """End-to-end entry for HOST (ideally pinned) tensors: H2D, kernel, D2H inside the library
    (fa_fwd_host), pipelined over the batch dimension.  Returns the host output tensor."""
    lib = _lib.load()
    for name, t in (("q", q), ("k", k), ("v", v)):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous host tensor")
    if q.dtype not in _DTYPE_CODE or k.dtype != q.dtype or v.dtype != q.dtype:
        raise RuntimeError("Only fp16 and bf16 are supported")
    if q.shape != k.shape or q.shape != v.shape or q.dim() != 4:
        raise RuntimeError("Query, key and value tensors have same shape")
    if o is None:
        o = torch.empty_like(q, pin_memory=q.is_pinned())
    elif o.is_cuda or not o.is_contiguous() or o.shape != q.shape or o.dtype != q.dtype:
        # the library copies batch*seq*heads*128 elements into o: anything else would be written out of bounds
        raise RuntimeError("o must be a contiguous host tensor with the shape and dtype of q")
    if q.size(3) != 128:
        raise RuntimeError("Kernel configuration was not found in flash_kernels.cuh")
    B, N, H, D = q.shape
    rc = lib.fa_fwd_host(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, D,
                         _DTYPE_CODE[q.dtype], device)
    if rc != 0:
        raise RuntimeError(_lib.last_error())
    return o
