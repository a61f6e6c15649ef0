#!/usr/bin/env python
"""Quick diff of the operator against a same-box comparator -- the role of the reference's
/root/reference/tools/debug/sanity_check.py:15-73 (same flags `--small --diff --kernel`): run every
selected kernel config on one problem and print the mismatch statistics of `error_stats`
(atol 1e-5, rtol 1e-3: a diagnostic printout, not a pass criterion -- the pass criteria live in tests/).

Comparator: flash-attn 2 when it imports (as the reference does), else torch SDPA on the GPU.
`KERNELS=tune` diffs both machine mappings.  Also the target of tools/debug/check_race.sh.

    python tools/debug/sanity_check.py --small
    KERNELS=tune python tools/debug/sanity_check.py --kernel 1 --diff
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import flash_attention  # noqa: E402
from flash_helpers.kernel_configs import get_kernel_configs  # noqa: E402
from flash_helpers.test.utils import QKVConfig, evaluate_kernel, generate_qkv  # noqa: E402


def comparator(q, k, v):
    try:
        from flash_helpers.test.utils import reference_forward_kernel_v2

        return "flash-attn 2", reference_forward_kernel_v2(q, k, v).reshape(q.shape)
    except Exception:  # noqa: BLE001  (not installed / no sm_100 build)
        out = torch.nn.functional.scaled_dot_product_attention(
            q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).contiguous()
        return "torch SDPA", out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--small", action="store_true", help="(1, 512, 1, 128) instead of (16, 2048, 16, 128)")
    ap.add_argument("--diff", action="store_true", dest="print_diffs", help="print per-row counts of |diff| > 1e-3")
    ap.add_argument("--kernel", type=int, default=-1, help="index into get_kernel_configs() (default: all)")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    batch, seq_len, n_heads = (1, 512, 1) if a.small else (16, 2048, 16)
    cfgs = get_kernel_configs()
    if a.kernel >= 0:
        cfgs = [cfgs[a.kernel]]
    dev = torch.device("cuda:0")
    for d_head in (128,):
        print("d_head:", d_head)
        for kcfg in (c for c in cfgs if c.d_head == d_head):
            qkv = QKVConfig(n_heads=n_heads, d_head=d_head, batch_size=batch, seq_len=seq_len,
                            dtype=kcfg.dtype.to_torch_dtype(), device=dev)
            q, k, v = generate_qkv(qkv, seed=a.seed)
            name, out_ref = comparator(q, k, v)
            out = flash_attention.forward(kcfg, q, k, v)
            torch.cuda.synchronize()
            print(f"vs {name}:")
            evaluate_kernel(kcfg, out_ref, out)
            if a.print_diffs:
                diff = (out - out_ref).abs() > 1e-3
                print(diff.reshape((-1, diff.shape[-1])).sum(dim=-1, keepdim=True))


if __name__ == "__main__":
    main()
