#!/usr/bin/env python
"""Launches the operator `--n_runs` times on one problem -- the target ncu wraps
(/root/reference/tools/benchmark/run_kernels.py:39-158 and ncu_bench.py:319-330):

  ncu --set full --clock-control none -k regex:fa_fwd -s 3 -c 1 -o prof \\
      python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 5
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import flash_attention  # noqa: E402
from flash_helpers.kernel_configs import DType, FlashForwardKernelConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq_len", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--n_heads", type=int, default=32)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--n_runs", type=int, required=True)
    a = ap.parse_args()
    cfg = FlashForwardKernelConfig(dtype=DType.from_string(a.dtype))
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(a.batch, a.seq_len, a.n_heads, 128, device="cuda", generator=g).to(
        cfg.dtype.to_torch_dtype()) for _ in range(3))
    o = torch.empty_like(q)
    for _ in range(a.n_runs):
        flash_attention.forward(cfg, q, k, v, o)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
