#!/usr/bin/env python
"""Kernel benchmark with the reference's CLI and methodology
(/root/reference/tools/benchmark/pt_bench.py:98-174,180-222): for every seq_len the batch size of
`BATCH_SIZE_FOR_SEQ_LEN`, 16 heads, one [4,B,N,H,d] allocation sliced into q,o,k,v; `num_warmups`
warm-ups, then per repeat an L2 flush (zero-fill of a buffer larger than L2), a short device sleep
and the operator's own cudaEvent timing (`forward_timed`).  TFLOP/s uses the reference FLOP model
B*H*(4 N^2 d + 6 N^2) for README comparability and, in a second column, the matmul-only model
4*B*H*N^2*d used for the roofline.  The last row is the harmonic mean over the seq_lens
(BASELINE.json configs[2]).  Differences: no prettytable, no `sudo nvidia-smi -lgc` clock locking
(never change clocks on the shared B200 boxes); comparators are optional and use public APIs.
"""
import argparse
import csv
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import flash_attention  # noqa: E402
from flash_helpers.kernel_configs import calc_self_attn_flop, get_kernel_configs  # noqa: E402
from flash_helpers.test.utils import (BATCH_SIZE_FOR_SEQ_LEN, BENCHMARK_N_HEADS, QKVConfig,  # noqa: E402
                                      generate_qkvo)

L2_FLUSH_BYTES = 512 << 20  # > 126 MB L2 of B200 (the reference used 100 MB against A100's 40 MB)


def benchmark_kernel(fn, flush, n_repeats, n_warmups, stabilize=True):
    for _ in range(n_warmups):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(n_repeats):
        if stabilize:
            flush.zero_()
            torch.cuda._sleep(1_000_000)
        torch.cuda.synchronize()
        _, ms = fn()
        times.append(ms)
    return times


def timed_sdpa_cudnn(q, k, v, o):
    from torch.nn.attention import SDPBackend, sdpa_kernel
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        e0.record()
        out = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2),
                                                               v.transpose(1, 2))
        e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def timed_fa2(q, k, v, o):
    from flash_attn import flash_attn_func
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = flash_attn_func(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--d_heads", type=str, default="128")
    ap.add_argument("--seq_lens", type=str, default="512,1024,2048,4096,8192,16384")
    ap.add_argument("--num_warmups", type=int, default=10)
    ap.add_argument("--num_repeats", type=int, default=64)
    ap.add_argument("--noncu", action="store_true", help="do not flush L2 / sleep between repeats")
    ap.add_argument("--comparators", action="store_true", help="also time cuDNN SDPA and flash-attn 2")
    ap.add_argument("--csv", type=str, default="")
    args = ap.parse_args()
    d_heads = [int(x) for x in args.d_heads.split(",")]
    seq_lens = [int(x) for x in args.seq_lens.split(",")]
    dev = torch.device("cuda:0")
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    rows = []
    for d_head in d_heads:
        for cfg in get_kernel_configs():  # KERNELS env: all (default) | tune (both machine mappings) | "128,128"
            dt = cfg.dtype.to_torch_dtype()
            per_seq = []
            for n in seq_lens:
                b = BATCH_SIZE_FOR_SEQ_LEN[n]
                q, k, v, o = generate_qkvo(QKVConfig(BENCHMARK_N_HEADS, d_head, b, n, dt, dev), seed=0)
                ts = benchmark_kernel(lambda: flash_attention.forward_timed(cfg, q, k, v, o), flush,
                                      args.num_repeats, args.num_warmups, not args.noncu)
                mean = sum(ts) / len(ts)
                ref_flop = calc_self_attn_flop(b, BENCHMARK_N_HEADS, n, d_head)
                mm_flop = 4.0 * b * BENCHMARK_N_HEADS * n * n * d_head
                row = {"kernel": str(cfg), "seq_len": n, "batch": b, "ms_mean": mean, "ms_min": min(ts),
                       "tflops_ref_model": ref_flop / mean / 1e9, "tflops_matmul": mm_flop / mean / 1e9}
                if args.comparators:
                    for name, fn in (("cudnn_sdpa", timed_sdpa_cudnn), ("flash_attn2", timed_fa2)):
                        try:
                            t = benchmark_kernel(lambda: fn(q, k, v, o), flush, max(4, args.num_repeats // 4),
                                                 3, not args.noncu)
                            row[name + "_tflops_matmul"] = mm_flop / (sum(t) / len(t)) / 1e9
                        except Exception as e:  # noqa: BLE001
                            row[name + "_err"] = str(e)[:80]
                rows.append(row)
                per_seq.append(row)
                print(f"{row['kernel']:70s} N={n:6d} B={b:3d}  {mean:8.4f} ms  "
                      f"{row['tflops_ref_model']:8.1f} TFLOP/s (ref model)  {row['tflops_matmul']:8.1f} (matmul)"
                      + "".join(f"  {k_}={v_:.1f}" for k_, v_ in row.items() if k_.endswith("_tflops_matmul")),
                      flush=True)
            hm = len(per_seq) / sum(1.0 / r["tflops_matmul"] for r in per_seq)
            print(f"{str(cfg):70s} harmonic mean over {seq_lens}: {hm:.1f} TFLOP/s (matmul model)")
            rows.append({"kernel": str(cfg), "seq_len": "harmonic_mean", "tflops_matmul": hm})
    if args.csv:
        keys = sorted({k_ for r in rows for k_ in r})
        with open(args.csv, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            w.writerows(rows)


if __name__ == "__main__":
    main()
