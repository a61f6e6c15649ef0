#!/usr/bin/env python
"""How much of a short launch is fixed cost?  Event-pair time (as quick_bench / bench.py measure it) of
  * an empty-ish torch kernel (one-element fill): the floor of the measurement itself,
  * this library on problems of ONE KV block per work tile (all three kernels), with one tile per CTA and with
    two, cold (L2 flushed before every launch, as the benchmarks do) and warm (no flush),
  * cuDNN fused attention on the same problems.
Development aid; results under gpurun_out/.

    python tools/launch_overhead.py --out gpurun_out/launch_overhead.json
"""
import argparse
import json
import os
import statistics as st
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flash_attention_from_scratch_b200 as fa  # noqa: E402
from flash_attention_from_scratch_b200 import _lib  # noqa: E402


def timed(fn, flush, reps=40, warmup=10):
    ts = []
    for i in range(warmup + reps):
        if flush is not None:
            flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(a.elapsed_time(b) * 1e3)
    return {"us_median": st.median(ts), "us_min": min(ts)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "launch_overhead.json"))
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    one = torch.zeros(1, device="cuda")
    rows = []

    def rec(name, fn):
        r = {"what": name, "cold": timed(fn, flush), "warm": timed(fn, None)}
        rows.append(r)
        print(f"{name:44s} cold {r['cold']['us_median']:7.1f} us (min {r['cold']['us_min']:6.1f})   "
              f"warm {r['warm']['us_median']:7.1f} us (min {r['warm']['us_min']:6.1f})", flush=True)

    rec("torch fill of one element", lambda: one.fill_(1.0))
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for (B, N, H) in ((1, 128, 1), (37, 128, 4), (37, 512, 4), (16, 512, 16), (16, 1024, 16)):
        q, k, v = (torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
        o = torch.empty_like(q)
        for mode in ("pp", "pair", "single"):
            with _lib.thread_kernel_mode({"single": 1, "pair": 2, "pp": 3}[mode]):
                rec(f"fa {mode:6s} (B,N,H)=({B},{N},{H})", lambda: fa.forward(None, q, k, v, o))
        qt, kt, vt = (x.transpose(1, 2) for x in (q, k, v))
        try:
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                rec(f"cudnn     (B,N,H)=({B},{N},{H})",
                    lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt))
        except Exception as e:  # noqa: BLE001
            print("cudnn failed", str(e)[:100])
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
