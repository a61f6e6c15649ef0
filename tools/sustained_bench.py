#!/usr/bin/env python
"""Sustained (power-capped) throughput: back-to-back launches for several seconds with NVML clocks and
power sampled every 10 ms, for this kernel and for same-box comparators (cuDNN fused attention, cuBLAS
bf16 GEMM).  Answers "who burns the watts": at the power cap the figure of merit is FLOP per joule.

    python tools/sustained_bench.py --seconds 3 --shape 4,4096,32 --what fa,cudnn,gemm
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Sampler(threading.Thread):
    def __init__(self, period=0.01):
        super().__init__(daemon=True)
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.period = period
        self.rows = []
        self.stop_flag = False

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((time.time(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)


def run_region(fn, seconds, chunk):
    """Back-to-back launches for `seconds`; returns (launches, device ms, per-chunk ms list)."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True)]
    evs[0].record()
    t0 = time.time()
    n = 0
    while time.time() - t0 < seconds:
        for _ in range(chunk):
            fn()
        n += chunk
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        evs.append(e)
        e.synchronize() if len(evs) % 4 == 0 else None   # keep the launch queue short but never empty
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    return n, sum(per), per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4,4096,32")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--chunk", type=int, default=25)
    ap.add_argument("--what", default="fa,cudnn,gemm")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    B, N, H = map(int, args.shape.split(","))
    torch.manual_seed(0)
    q = torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn_like(q)
    v = torch.randn_like(q)
    o = torch.empty_like(q)
    fl = 4.0 * B * H * N * N * 128
    rows = []
    for what in args.what.split(","):
        if what == "fa":
            import flash_attention_from_scratch_b200 as fa
            fn = lambda: fa.forward(None, q, k, v, o)  # noqa: E731
            work = fl
        elif what == "cudnn":
            from torch.nn.attention import SDPBackend, sdpa_kernel
            qt, kt, vt = (x.transpose(1, 2) for x in (q, k, v))
            ctx = sdpa_kernel(SDPBackend.CUDNN_ATTENTION)
            ctx.__enter__()
            fn = lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt)  # noqa: E731
            work = fl
        elif what == "gemm":
            a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
            b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
            c = torch.empty_like(a)
            fn = lambda: torch.matmul(a, b, out=c)  # noqa: E731
            work = 2.0 * 8192 ** 3
        else:
            continue
        time.sleep(1.0)  # let the board cool between arms
        smp = Sampler()
        smp.start()
        n, ms, per = run_region(fn, args.seconds, args.chunk)
        smp.stop_flag = True
        smp.join()
        if what == "cudnn":
            ctx.__exit__(None, None, None)
        clk = sorted(r[1] for r in smp.rows[len(smp.rows) // 4:])
        pw = [r[2] for r in smp.rows[len(smp.rows) // 4:]]
        reasons = 0
        for r in smp.rows:
            reasons |= r[3]
        first = per[0] / args.chunk
        last = sum(per[-4:]) / (4 * args.chunk) if len(per) >= 4 else per[-1] / args.chunk
        row = {"what": what, "shape": [B, N, H, 128], "launches": n, "region_ms": ms,
               "tflops_region": work * n / ms / 1e9, "tflops_first_chunk": work / first / 1e9,
               "tflops_last_chunks": work / last / 1e9, "sm_mhz_median": clk[len(clk) // 2] if clk else None,
               "power_w_mean": sum(pw) / len(pw) if pw else None, "power_w_max": max(pw) if pw else None,
               "nvml_samples": len(smp.rows), "reasons_or": hex(reasons),
               "gflop_per_joule": (work * n / (ms / 1e3)) / (sum(pw) / len(pw)) / 1e9 if pw else None}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
