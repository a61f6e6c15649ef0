#!/usr/bin/env python
"""Optional same-box comparator: the FlashAttention-4 CuTe-DSL forward shipped inside vllm
(library code, JIT-compiled on the box).  Prints TFLOP/s at the headline shape or the reason it
could not run.  Context only -- never part of the product path."""
import sys
import time

import torch


def main():
    B, N, H, D = 4, 4096, 32, 128
    q, k, v = (torch.randn(B, N, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    try:
        from vllm.vllm_flash_attn.cute.interface import _flash_attn_fwd as fwd  # noqa: PLC0415
    except Exception as e:  # noqa: BLE001
        print("FA4 import failed:", repr(e)[:300])
        return
    try:
        t0 = time.time()
        out = fwd(q, k, v, causal=False)
        torch.cuda.synchronize()
        print(f"FA4 first call (JIT) {time.time() - t0:.1f} s")
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fwd(q, k, v, causal=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts[3:]) / len(ts[3:])
        print(f"FA4 (vllm cute) headline: {ms:.4f} ms  {4.0 * B * H * N * N * D / ms / 1e9:.1f} TFLOP/s")
        o = out[0] if isinstance(out, (tuple, list)) else out
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2),
                                                               v.transpose(1, 2)).transpose(1, 2)
        print("maxdiff vs sdpa16", (o.float() - ref.float()).abs().max().item())
    except Exception as e:  # noqa: BLE001
        print("FA4 run failed:", repr(e)[:600])


if __name__ == "__main__":
    sys.exit(main())
