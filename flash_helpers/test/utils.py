"""`flash_helpers.test.utils` as the reference's test/bench scripts import it
(/root/reference/py/flash_helpers/test/utils.py:9-17,102-187) -- TEST INFRASTRUCTURE.

Differences: the CUDA-only comparator imports (`flash_attn_2_cuda`, `flash_attn_3_cuda`,
utils.py:6-7) are optional instead of mandatory, and `generate_qkv` takes an optional seed
(the reference is unseeded).  `py_flash_attention` is the oracle restated in `oracle/`.
"""
from dataclasses import dataclass

import torch

from oracle.attention_ref import py_flash_attention  # noqa: F401  (test oracle)

BATCH_SIZE_FOR_SEQ_LEN = {512: 16, 1024: 16, 2048: 16, 4096: 16, 8192: 8, 16384: 4}
BENCHMARK_N_HEADS = 16
BENCHMARK_BATCH_SIZE = 16  # imported by the reference's ncu_bench.py:18 (missing from its own utils.py)


@dataclass(frozen=True)
class QKVConfig:
    n_heads: int
    d_head: int
    batch_size: int
    seq_len: int
    dtype: torch.dtype
    device: torch.device


def generate_qkv(cfg: QKVConfig, seed=None):
    gen = None
    if seed is not None:
        gen = torch.Generator(device=cfg.device).manual_seed(seed)
    shape = (cfg.batch_size, cfg.seq_len, cfg.n_heads, cfg.d_head)
    q, k, v = (torch.randn(shape, dtype=cfg.dtype, device=cfg.device, generator=gen)
               for _ in range(3))
    return q, k, v


def generate_qkvo(cfg: QKVConfig, seed=None):
    """One [4, B, N, H, d] allocation sliced into q, o, k, v (utils.py:124-134)."""
    gen = None
    if seed is not None:
        gen = torch.Generator(device=cfg.device).manual_seed(seed)
    buf = torch.empty((4, cfg.batch_size, cfg.seq_len, cfg.n_heads, cfg.d_head),
                      dtype=cfg.dtype, device=cfg.device)
    q, o, k, v = (buf[i] for i in range(4))
    for t in (q, k, v):
        t.normal_(generator=gen)
    return q, k, v, o


def reference_forward_kernel_v2(q, k, v, o=None):
    """Same-box comparator: stock flash-attn 2 (the reference calls a patched 14-argument
    `flash_attn_2_cuda.fwd`, utils.py:58-77; the public API is used here)."""
    from flash_attn import flash_attn_func

    return flash_attn_func(q, k, v, causal=False)


def reference_forward_kernel_v2_timed(q, k, v, o=None):
    """(out, ms) like the reference's patched flash-attn build (utils.py:80-99); the stock wheel has no timed
    entry, so the call is bracketed with CUDA events here."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = reference_forward_kernel_v2(q, k, v, o)
    e1.record()
    e1.synchronize()
    return out, e0.elapsed_time(e1)


def reference_forward_kernel_v3(q, k, v, o=None):
    """The reference's second comparator is flash-attn 3 (`flash_attn_3_cuda.fwd`, utils.py:20-55), a Hopper
    (wgmma) build that does not exist for sm_100.  Used when it imports; otherwise the strongest comparator this
    box has stands in: cuDNN fused attention through torch SDPA."""
    try:
        import flash_attn_3_cuda  # noqa: F401

        from flash_attn_interface import flash_attn_func as fa3_func

        out = fa3_func(q, k, v, causal=False)
        return out[0] if isinstance(out, tuple) else out
    except Exception:  # noqa: BLE001
        from torch.nn.attention import SDPBackend, sdpa_kernel

        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            return torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)


def is_a100(device_idx=0):
    return torch.cuda.is_available() and "A100" in torch.cuda.get_device_name(device_idx)


def error_stats(expected, actual, atol=1e-5, rtol=1e-3):
    close = torch.isclose(expected, actual, atol=atol, rtol=rtol)
    mismatched = close.numel() - close.sum()
    return mismatched, mismatched / expected.numel() * 100, (expected - actual).abs().max()


def evaluate_kernel(cfg, out_ref, out):
    mismatched, pct, max_diff = error_stats(out_ref, out)
    print(f"{cfg}")
    print(f"  Mismatched elements: {mismatched} / {out.numel()} ({pct:.1f}%)")
    print(f"  Greatest absolute difference: {max_diff}")


def get_cuda_device_info(device_idx=0):
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA not available")
    p = torch.cuda.get_device_properties(device_idx)
    return {"name": p.name, "compute_capability": f"{p.major}.{p.minor}",
            "total_memory": f"{p.total_memory / 2**30:.2f} GB",
            "multi_processor_count": p.multi_processor_count}
