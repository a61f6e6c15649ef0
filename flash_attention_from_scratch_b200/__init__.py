"""B200-native fused attention forward behind the operator API of
sonnyli/flash_attention_from_scratch (`flash_attention.forward(kernel_cfg, q, k, v, o=None)`).

Only what the hot path needs lives here: `csrc/` (the sm_100a kernel + C-ABI host library),
`_lib` (ctypes binding), `op` (the operator), `kernel_configs` (config dataclass + FLOP model).
"""
from .op import forward, forward_host, forward_timed  # noqa: F401

__all__ = ["forward", "forward_timed", "forward_host"]
