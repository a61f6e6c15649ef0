#!/usr/bin/env python
"""Cycle trace of the ping-pong kernel (debug level 5 of fa_fwd_kernel_pp, csrc/fa_fwd_pp_sm100.cuh): one CTA's
second work tile at the headline shape, all events on one time axis (clk relative to the first stamp).

    FA_SM100_MODE=pp python tools/gpu_pp_trace.py --out gpurun_out/pp_trace.json
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flash_attention_from_scratch_b200 import _lib  # noqa: E402

TRACE_BASE = 2 * 128 * 128 + 512


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4,4096,32")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    B, N, H = map(int, args.shape.split(","))
    lib = _lib.load()
    torch.manual_seed(0)
    q = torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn_like(q)
    v = torch.randn_like(q)
    o = torch.zeros_like(q)
    dump = torch.zeros(TRACE_BASE + 1024, dtype=torch.float32).pin_memory()
    diag = torch.zeros(256, dtype=torch.int32).pin_memory()
    knobs = (C.c_uint32 * 8)(0, 0, 0, 0, 0, 0, 0, 5)
    sb, sn, sh, _ = q.stride()
    for _ in range(2):  # second run is warm
        dump.zero_()
        rc = lib.fa_fwd_debug(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, 128, sb, sn, sh, 15,
                              dump.data_ptr(), knobs, diag.data_ptr())
        if rc != 0:
            print("rc", rc, _lib.last_error())
            return 1
    t = dump.view(torch.int32)[TRACE_BASE:].tolist()
    t = [x & 0xFFFFFFFF for x in t]
    nb = min((N + 127) // 128, 32)
    sm = [[t[(w * 16 + kk) * 8:(w * 16 + kk) * 8 + 8] for kk in range(16)] for w in range(2)]
    pv = [t[256 + j * 4:256 + j * 4 + 4] for j in range(32)]
    ss = [t[384 + j * 4:384 + j * 4 + 4] for j in range(32)]
    ob = [t[512 + 2 * j:512 + 2 * j + 2] for j in range(32)]
    base = min(x for x in t[:576] if x)
    rel = lambda x: (x - base) & 0xFFFFFFFF if x else -1  # noqa: E731
    print("block | softmax: wait_S  S_ready  S_in_regs  m_pub  token  P_free  p_full  p_last | S issue: start K_ok acc_free done"
          " | PV issue: start P96 Pall done | pipe: S_ret PV_ret")
    rows = []
    for j in range(nb):
        w, kk = j & 1, j >> 1
        r = {"j": j, "wg": w, "softmax": [rel(x) for x in sm[w][kk]], "s_issue": [rel(x) for x in ss[j]],
             "pv_issue": [rel(x) for x in pv[j]], "pipe": [rel(x) for x in ob[j]]}
        rows.append(r)
        print(f"{j:3d} w{w} | " + " ".join(f"{x:6d}" for x in r["softmax"]) + " | " +
              " ".join(f"{x:6d}" for x in r["s_issue"]) + " | " + " ".join(f"{x:6d}" for x in r["pv_issue"]) +
              " | " + " ".join(f"{x:6d}" for x in r["pipe"]))
    # summary: per-block period and where each warpgroup's time goes (blocks 8.. of the tile)
    lo = 8
    if nb > lo + 4:
        per = (rows[nb - 1]["softmax"][7] - rows[lo]["softmax"][7]) / (nb - 1 - lo)
        def avg(f):
            xs = [f(r) for r in rows[lo:nb]]
            return sum(xs) / len(xs)
        print(f"period per block {per:.0f} clk;  softmax: wait S {avg(lambda r: r['softmax'][1]-r['softmax'][0]):.0f}"
              f"  ld S {avg(lambda r: r['softmax'][2]-r['softmax'][1]):.0f}  max+m {avg(lambda r: r['softmax'][3]-r['softmax'][2]):.0f}"
              f"  token {avg(lambda r: r['softmax'][4]-r['softmax'][3]):.0f}  frag0+P_free {avg(lambda r: r['softmax'][5]-r['softmax'][4]):.0f}"
              f"  ->p_full {avg(lambda r: r['softmax'][6]-r['softmax'][5]):.0f}  ->p_last {avg(lambda r: r['softmax'][7]-r['softmax'][6]):.0f}")
        print(f"MMA warp: PV wait P96 {avg(lambda r: r['pv_issue'][1]-r['pv_issue'][0]):.0f}  wait Pall {avg(lambda r: r['pv_issue'][2]-r['pv_issue'][1]):.0f}"
              f"  issue {avg(lambda r: r['pv_issue'][3]-r['pv_issue'][2]):.0f};  S wait K + acc {avg(lambda r: r['s_issue'][2]-r['s_issue'][0]):.0f}"
              f"  issue {avg(lambda r: r['s_issue'][3]-r['s_issue'][2]):.0f}")
        print(f"p_full signalled -> PV warp saw it {avg(lambda r: r['pv_issue'][1]-r['softmax'][6]):.0f};  "
              f"S committed -> softmax saw it {avg(lambda r: r['softmax'][1]-r['s_issue'][3]):.0f} (negative = S waited for the softmax)")
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
