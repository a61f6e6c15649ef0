"""Kernel-config surface compatible with the reference's `flash_helpers.kernel_configs`
(/root/reference/py/flash_helpers/kernel_configs.py:9-175, 389-485).

The reference uses a 13-field frozen dataclass as the key into a registry of 85 Ampere template
instantiations.  The B200 library has ONE kernel per dtype, so here the dataclass is a
description/compatibility object: the operator only reads `dtype` and `d_head`
(see op.py); the mma.sync tiling knobs are carried so that reference scripts that print, sort or
parse configs keep working, but they select nothing.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from enum import IntEnum

ELEM_SIZE = 2  # bytes per element (fp16 / bf16)


class DType(IntEnum):
    """Values are torch ScalarType ints, as in the reference (kernel_configs.py:9-14); they are
    also the dtype codes of the C ABI (include/fa_sm100.h)."""

    FP16 = 5
    BF16 = 15

    def to_torch_dtype(self):
        import torch

        return {DType.FP16: torch.float16, DType.BF16: torch.bfloat16}[self]

    def to_cpp_str(self) -> str:
        return {DType.FP16: "FA_DTYPE_FP16", DType.BF16: "FA_DTYPE_BF16"}[self]

    @classmethod
    def from_string(cls, text: str) -> "DType":
        text = text.strip()
        if text.lstrip("-").isdigit():
            return cls(int(text))
        try:
            return cls[text.upper()]
        except KeyError:
            raise ValueError(
                f"Invalid dtype string '{text}'. Valid options: "
                + ", ".join(f"{m.name} ({m.value})" for m in cls)
            ) from None

    @classmethod
    def from_torch_dtype(cls, dt) -> "DType":
        import torch

        return {torch.float16: cls.FP16, torch.bfloat16: cls.BF16}[dt]


def calc_self_attn_flop(n_samples: int, n_heads: int, seq_len: int, d_head: int) -> int:
    """The reference's README/benchmark FLOP model, B*H*(4 N^2 d + 6 N^2)
    (kernel_configs.py:102-103).  Used only for README-comparable columns."""
    return n_samples * n_heads * (4 * seq_len**2 * d_head + 6 * seq_len**2)


def calc_matmul_flop(n_samples: int, n_heads: int, seq_len: int, d_head: int) -> int:
    """Algorithmic work used for the roofline: 4*B*H*N^2*d (QK^T + PV, 2 flop per MAC)."""
    return 4 * n_samples * n_heads * seq_len**2 * d_head


def algorithmic_bytes(n_samples: int, n_heads: int, seq_len: int, d_head: int) -> int:
    """Q, K, V read once and O written once."""
    return 4 * n_samples * seq_len * n_heads * d_head * ELEM_SIZE


@dataclass(frozen=True, order=True)
class FlashForwardKernelConfig:
    """Field-compatible with kernel_configs.py:106-120 of the reference."""

    dtype: DType
    d_head: int = 128
    B_r: int = 128
    B_c: int = 128
    n_warps: int = 12
    async_copy: bool = True
    eager_load_blocks: bool = True
    swizzled: bool = True
    Q_mma_load_K_tiles: int = 0
    K_mma_load_K_tiles: int = 0
    V_mma_load_K_tiles: int = 0
    mma_double_buffer_loads: bool = False
    optimized_softmax: bool = True
    # Blackwell knob (not in the reference): machine mapping of the sm_100a kernel.  0 = the library's
    # choice by sequence length, 1 = single CTAs (tcgen05 cta_group::1), 2 = 2-CTA clusters (cta_group::2),
    # 3 = 2-CTA clusters running the KV ping-pong kernel (one Q tile per CTA, csrc/fa_fwd_pp_sm100.cuh).
    cta_group: int = 0

    def __str__(self) -> str:
        return self.short_form()

    def short_form(self, include_d_head: bool = True, include_tup: bool = True) -> str:
        feats = [name for flag, name in ((self.async_copy, "async"), (self.eager_load_blocks, "eager"),
                                         (self.swizzled, "swizzled")) if flag]
        feats.append(f"load_{self.Q_mma_load_K_tiles}_{self.K_mma_load_K_tiles}_"
                     f"{self.V_mma_load_K_tiles}_tiles")
        if self.mma_double_buffer_loads:
            feats.append("buffer")
        if self.optimized_softmax:
            feats.append("opt_softmax")
        if self.cta_group:
            feats.append(f"cta{self.cta_group}")
        head = ""
        if include_tup:
            d = f"{self.d_head}, " if include_d_head else ""
            head = f"({self.dtype.name}, {d}{self.B_r}, {self.B_c}, {self.n_warps}): "
        return head + "+".join(feats)

    def kernel_name(self) -> str:
        return {2: "fa_fwd_kernel_pair", 3: "fa_fwd_kernel_pp"}.get(self.cta_group, "fa_fwd_kernel")

    def attn_flop(self, n_samples: int, n_heads: int, seq_len: int) -> int:
        return calc_self_attn_flop(n_samples, n_heads, seq_len, self.d_head)

    def total_flop(self, n_samples: int, n_heads: int, seq_len: int) -> int:
        return calc_matmul_flop(n_samples, n_heads, seq_len, self.d_head)


_SHORT_RE = re.compile(r"\(\s*(\w+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)\s*:\s*(\S*)")


def parse_kernel_name_into_config(text: str) -> FlashForwardKernelConfig:
    """Parses the short form `(BF16, 128, 128, 128, 12): async+...+load_a_b_c_tiles+...`
    (the format of the reference's tables, kernel_configs.py:253-331)."""
    m = _SHORT_RE.search(text)
    if not m:
        raise ValueError(f"Invalid kernel name: {text}")
    dtype, d_head, b_r, b_c, n_warps, feats = m.groups()
    feats = feats.split("+")
    load = next((f for f in feats if f.startswith("load_")), "load_0_0_0_tiles")
    qt, kt, vt = (int(x) for x in load[len("load_"):-len("_tiles")].split("_"))
    return FlashForwardKernelConfig(
        dtype=DType.from_string(dtype), d_head=int(d_head), B_r=int(b_r), B_c=int(b_c),
        n_warps=int(n_warps), async_copy="async" in feats, eager_load_blocks="eager" in feats,
        swizzled="swizzled" in feats, Q_mma_load_K_tiles=qt, K_mma_load_K_tiles=kt,
        V_mma_load_K_tiles=vt, mma_double_buffer_loads="buffer" in feats,
        optimized_softmax="opt_softmax" in feats,
        cta_group=next((int(f[3:]) for f in feats if f in ("cta1", "cta2", "cta3")), 0))


_DEMANGLED_RE = re.compile(r"fa::(?:pp::)?(fa_fwd_kernel(?:_pair|_pp)?)<\(bool\)([01]), \(bool\)([01]), \(bool\)([01])>")


def parse_flash_forward_kernel_config(kernel_name: str) -> FlashForwardKernelConfig:
    """Kernel name as ncu prints it -> config.  The reference's ncu_bench.py imports this name
    (/root/reference/tools/benchmark/ncu_bench.py:15,104; its own library only has
    `parse_kernel_name_into_config`, kernel_configs.py:323-335, which is why that script is stale).  Accepts
    the demangled sm_100a kernels (`void fa::fa_fwd_kernel_pair<(bool)1, (bool)0, (bool)0>(...)`: bf16?, debug?,
    ragged?) and the short form."""
    m = _DEMANGLED_RE.search(kernel_name)
    if m:
        kern, bf16, _debug, _ragged = m.groups()
        return FlashForwardKernelConfig(dtype=DType.BF16 if bf16 == "1" else DType.FP16,
                                        cta_group={"fa_fwd_kernel": 1, "fa_fwd_kernel_pair": 2, "fa_fwd_kernel_pp": 3}[kern])
    return parse_kernel_name_into_config(kernel_name)


def tile_softmax_flop(B_r: int, B_c: int, d_head: int) -> int:
    return B_r * (4 * B_c + d_head + 4)  # kernel_configs.py:61-63 (kernels 6-16)


def kv_tile_flop(B_r: int, B_c: int, d_head: int) -> int:
    return 2 * B_r * d_head * B_c + 2 * B_r * B_c * d_head + tile_softmax_flop(B_r, B_c, d_head)


def calc_total_flop(n_samples: int, n_heads: int, seq_len: int, B_r: int, B_c: int, d_head: int) -> int:
    """The reference's tile-level FLOP model including the softmax arithmetic (kernel_configs.py:87-99),
    used by its ncu_bench.py for the "total" column."""
    assert seq_len % B_r == 0 and seq_len % B_c == 0
    t_r, t_c = seq_len // B_r, seq_len // B_c
    return t_r * (t_c * kv_tile_flop(B_r, B_c, d_head) + B_r * d_head) * n_samples * n_heads


def get_kernels_to_build():
    """One kernel per dtype (the reference returns its 80-config autotune grid here,
    kernel_configs.py:457-462; Ampere tile knobs have no meaning for tcgen05)."""
    return sorted(FlashForwardKernelConfig(dtype=dt) for dt in DType)


def get_autotuning_kernel_configs(dtypes=(DType.BF16, DType.FP16)):
    """The tuning grid of the sm_100a kernel: both machine mappings per dtype (the reference's grid of
    Ampere tile knobs, kernel_configs.py:364-455, has no meaning here).  Build-time knobs (exp2 emulation
    fraction, K/V ring depth, register split) are swept with tools/build_variants.py instead."""
    return [FlashForwardKernelConfig(dtype=dt, cta_group=cg) for dt in dtypes for cg in (1, 2, 3)]


def get_kernel_configs(kernels_key: str = ""):
    """Same `KERNELS` env UX as the reference (kernel_configs.py:465-485); every key maps onto the
    per-dtype kernels except "tune" (both machine mappings per dtype); "B_r,B_c" keeps only matching
    tile sizes (i.e. "128,128")."""
    if kernels_key == "":
        kernels_key = os.environ.get("KERNELS", "all")
    if kernels_key == "tune":
        return get_autotuning_kernel_configs()
    if kernels_key.startswith("prog") or kernels_key == "all":
        return get_kernels_to_build()
    if "," in kernels_key:
        b_r, b_c = map(int, kernels_key.split(","))
        return [c for c in get_kernels_to_build() if c.B_r == b_r and c.B_c == b_c]
    raise ValueError(f"Invalid kernels env key: {kernels_key}")
