"""ctypes binding of libfa_sm100.so (C ABI: include/fa_sm100.h).

There is deliberately NO fallback: if the shared library is missing or fails to load the
import of the operator raises, so a GPU box can never silently run something else.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from .build import LIB_PATH as _DEFAULT_LIB_PATH

# FA_SM100_LIB selects another build of the SAME library (e.g. the hang-guard bring-up build).
LIB_PATH = Path(os.environ.get("FA_SM100_LIB", str(_DEFAULT_LIB_PATH)))

FA_DTYPE_FP16 = 5
FA_DTYPE_BF16 = 15

# name -> (restype, argtypes); must list every symbol include/fa_sm100.h declares
_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
_FWD_ARGS = [_P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _I]
SYMBOLS = {
    "fa_fwd": (_I, _FWD_ARGS + [_P]),
    "fa_fwd_timed": (_I, _FWD_ARGS + [_P, C.POINTER(C.c_float)]),
    "fa_fwd_host": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I]),
    "fa_host_workspace_free": (_I, [_I]),
    "fa_last_error_string": (C.c_char_p, []),
    "fa_device_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "fa_kernel_info": (_I, [C.POINTER(_I)] * 4),
    "fa_launch_count": (_L, []),
    "fa_last_kernel": (_I, []),
    "fa_pick_kernel": (_I, [_I, _I, _I, _I]),
    "fa_tensor_map_cache_stats": (_I, [C.POINTER(_L), C.POINTER(_L)]),
    "fa_set_kernel_mode": (_I, [_I]),
    "fa_set_thread_kernel_mode": (_I, [_I]),
    "fa_fwd_debug": (_I, _FWD_ARGS + [_P, C.POINTER(C.c_uint32), _P]),
}

_lib = None


def lib_path() -> Path:
    return LIB_PATH


def load() -> C.CDLL:
    """Load (once) and type the shared library. Raises RuntimeError if it is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is not built; run `python -m flash_attention_from_scratch_b200.build` "
                "(there is no fallback implementation)"
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().fa_last_error_string().decode("utf-8", "replace")


def launch_count() -> int:
    return int(load().fa_launch_count())


MODE_AUTO, MODE_SINGLE, MODE_PAIR, MODE_PINGPONG = 0, 1, 2, 3
KERNEL_NAMES = {MODE_SINGLE: "fa_fwd_kernel", MODE_PAIR: "fa_fwd_kernel_pair", MODE_PINGPONG: "fa_fwd_kernel_pp"}


def last_kernel() -> str:
    """Name of the kernel the calling thread's last launch used (what AUTO picked)."""
    return KERNEL_NAMES.get(int(load().fa_last_kernel()), "none")


def pick_kernel(seq_len: int, batch: int, n_heads: int, n_sms: int = 0) -> str:
    """The kernel the library would launch for this problem (no launch): the C side's AUTO rule."""
    return {1: "fa_fwd_kernel", 2: "fa_fwd_kernel_pair", 3: "fa_fwd_kernel_pp"}.get(
        int(load().fa_pick_kernel(seq_len, batch, n_heads, n_sms)), "?")


def tensor_map_cache_stats() -> dict:
    h, m = C.c_int64(), C.c_int64()
    load().fa_tensor_map_cache_stats(C.byref(h), C.byref(m))
    return {"hits": h.value, "misses": m.value}


def set_kernel_mode(mode: int) -> int:
    """Select the machine mapping (include/fa_sm100.h: fa_set_kernel_mode); returns the previous one."""
    prev = int(load().fa_set_kernel_mode(int(mode)))
    if prev < 0:
        raise ValueError(f"invalid kernel mode {mode}")
    return prev


class thread_kernel_mode:
    """Context manager: machine mapping for launches of the calling thread only
    (include/fa_sm100.h: fa_set_thread_kernel_mode).  `mode` 0 / None leaves the choice alone."""

    def __init__(self, mode):
        self.mode = int(mode) if mode else 0
        self.prev = -1

    def __enter__(self):
        if self.mode in (MODE_SINGLE, MODE_PAIR, MODE_PINGPONG):
            self.prev = int(load().fa_set_thread_kernel_mode(self.mode))
        elif self.mode != 0:
            raise ValueError(f"invalid cta_group {self.mode} (0 = auto, 1 = single CTAs, 2 = CTA pairs, "
                             "3 = CTA pairs, ping-pong kernel)")
        return self

    def __exit__(self, *exc):
        if self.mode in (MODE_SINGLE, MODE_PAIR, MODE_PINGPONG):
            load().fa_set_thread_kernel_mode(self.prev)
        return False


def kernel_info() -> dict:
    v = [C.c_int() for _ in range(4)]
    load().fa_kernel_info(*[C.byref(x) for x in v])
    return dict(zip(("smem_bytes", "threads", "rows_per_cta", "tmem_cols"), (x.value for x in v)))
