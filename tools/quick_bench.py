#!/usr/bin/env python
"""Quick single-GPU timing of the operator on a few shapes (development aid, not bench.py).

Timing follows the reference harness (/root/reference/tools/benchmark/pt_bench.py:98-174):
warm-ups, then per repeat an L2 flush (zero-fill > L2) and the library's own cudaEvent timing.
Optionally times same-box comparators (torch SDPA backends, flash_attn 2) for context.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import flash_attention_from_scratch_b200 as fa  # noqa: E402


def flops(B, N, H, D):
    return 4.0 * B * H * N * N * D


def time_fn(fn, flush, reps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4,4096,32;16,512,16;16,1024,16;16,2048,16;16,4096,16;8,8192,16;4,16384,16")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--comparators", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for spec in args.shapes.split(";"):
        B, N, H = map(int, spec.split(","))
        torch.manual_seed(0)
        q = torch.randn(B, N, H, 128, device="cuda", dtype=dt)
        k = torch.randn_like(q)
        v = torch.randn_like(q)
        o = torch.empty_like(q)
        fl = flops(B, N, H, 128)
        ts = []
        for _ in range(args.warmup):
            fa.forward_timed(None, q, k, v, o)
        for _ in range(args.reps):
            flush.zero_()
            torch.cuda.synchronize()
            _, ms = fa.forward_timed(None, q, k, v, o)
            ts.append(ms)
        mean = sum(ts) / len(ts)
        row = {"shape": [B, N, H, 128], "dtype": args.dtype, "ms_mean": mean, "ms_min": min(ts),
               "tflops_mean": fl / mean / 1e9, "tflops_best": fl / min(ts) / 1e9}
        if args.check:
            ref = torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
            row["maxdiff_vs_sdpa16"] = (o.float() - ref.float()).abs().max().item()
            hs = slice(0, min(H, 2))
            ref32 = torch.nn.functional.scaled_dot_product_attention(
                q[:1, :, hs].float().transpose(1, 2), k[:1, :, hs].float().transpose(1, 2),
                v[:1, :, hs].float().transpose(1, 2)).transpose(1, 2)
            row["maxdiff_vs_sdpa32_subset"] = (o[:1, :, hs].float() - ref32).abs().max().item()
        if args.comparators:
            qt, kt, vt = (x.transpose(1, 2) for x in (q, k, v))
            from torch.nn.attention import SDPBackend, sdpa_kernel
            for name, be in (("sdpa_flash", SDPBackend.FLASH_ATTENTION),
                             ("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION)):
                try:
                    with sdpa_kernel(be):
                        t = time_fn(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt),
                                    flush, max(3, args.reps // 2), 3)
                    row[name + "_tflops"] = fl / (sum(t) / len(t)) / 1e9
                except Exception as e:  # noqa: BLE001
                    row[name + "_err"] = str(e)[:100]
            try:
                from flash_attn import flash_attn_func
                t = time_fn(lambda: flash_attn_func(q, k, v), flush, max(3, args.reps // 2), 3)
                row["flash_attn2_tflops"] = fl / (sum(t) / len(t)) / 1e9
            except Exception as e:  # noqa: BLE001
                row["flash_attn2_err"] = str(e)[:100]
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
