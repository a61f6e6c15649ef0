#!/usr/bin/env python
"""Cycle-level timeline of one CTA of the attention kernel (debug instantiation, level 5).

Records %clock at the hand-off points of the softmax warpgroups and the MMA-issuing thread for
the first 32 KV blocks of work tile 0 on SM of CTA 0, while the whole grid runs the headline
problem, and prints per-block phase durations (cycles).  Development aid for tuning the ping-pong."""
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
    B, N, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4, 4096, 32)))
    level = int(os.environ.get("FA_TRACE_LEVEL", "5"))  # 6..9: ablations, see FwdDebug::level
    lib = _lib.load()
    torch.manual_seed(0)
    q = torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn_like(q)
    v = torch.randn_like(q)
    o = torch.empty_like(q)
    dump = torch.zeros(TRACE_BASE + 1152, device="cuda", dtype=torch.float32)
    diag = torch.zeros(256, device="cuda", dtype=torch.int32)
    knobs = (C.c_uint32 * 8)(0, 0, 0, 0, 0, 0, 0, level)
    sb, sn, sh, _ = q.stride()
    for _ in range(2):
        rc = lib.fa_fwd_debug(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, 128, sb, sn,
                              sh, 15, dump.data_ptr(), knobs, diag.data_ptr())
        assert rc == 0, _lib.last_error()
    tr = dump.view(torch.int32)[TRACE_BASE:].cpu().numpy().astype("int64") & 0xFFFFFFFF
    nb = min(32, N // 128)
    sm = tr[:512].reshape(2, 32, 8)[:, :nb]
    mma = tr[512:768].reshape(32, 2, 4)[:nb]
    sis = tr[768:1024].reshape(32, 2, 4)[:nb]  # generation 6: S issue (enter, deps met, issued)
    t0 = int(sm[0, 0, 0])
    d = lambda a, b: int((a - b) & 0xFFFFFFFF)  # noqa: E731
    rows = []
    print("blk st |  S seen   ld     max    exp96  exp32 | next S seen - P_last (wait) | MMA: pfull_seen-parrive  pv6 issue  plast_seen-parrive  tail issue")
    for j in range(1, nb - 1):
        for s in range(2):
            e = sm[s, j]
            nxt = sm[s, j + 1]
            m = mma[j, s]
            r = {
                "j": j, "s": s, "t_S": d(e[0], t0), "ld": d(e[1], e[0]), "max": d(e[2], e[1]),
                "exp96": d(e[3], e[2]), "exp32": d(e[4], e[3]), "wait_next_S": d(nxt[0], e[4]),
                "mma_pfull_lat": d(m[0], e[3]), "mma_pv6_issue": d(m[1], m[0]),
                "mma_plast_lat": d(m[2], e[4]), "mma_tail_issue": d(m[3], m[2]),
                "period": d(nxt[0], e[0]),
                # generation 6 only (zero otherwise): wait for PV_s(j-1) before the first P store,
                # and the issue of S_s(j): time blocked on K / s_free, S issued -> softmax sees it
                "pv_wait": d(e[6], e[5]), "s_issue_blocked": d(sis[j, s][1], sis[j, s][0]),
                "s_issue_to_seen": d(e[0], sis[j, s][2]), "s_issued_at": d(sis[j, s][2], t0),
            }
            rows.append(r)
            print(f"{j:3d} {s}  | {r['t_S']:7d} {r['ld']:5d} {r['max']:6d} {r['exp96']:6d} {r['exp32']:6d} | "
                  f"{r['wait_next_S']:6d} | {r['mma_pfull_lat']:6d} {r['mma_pv6_issue']:6d} "
                  f"{r['mma_plast_lat']:6d} {r['mma_tail_issue']:6d} | period {r['period']} | pvwait {r['pv_wait']} "
                  f"s_blocked {r['s_issue_blocked']} s_issue->seen {r['s_issue_to_seen']}")
    import statistics as st
    # tensor-pipe observer: retirement stamps of S_0(j), PV_0(j-1), S_1(j), PV_1(j-1); the delta to the
    # previous stamp is the pipe time of that 8-MMA group while the pipe is backlogged
    ob = tr[1024:1152].reshape(32, 4)[:nb]
    pipe_rows = []
    print("blk | retire-to-retire: S_0(j)  PV_0(j-1)  S_1(j)  PV_1(j-1) | S_0(j) seen by softmax after retire")
    for j in range(2, nb - 1):
        if 0 in ob[j] or 0 in ob[j - 1]:
            continue
        pr = {"j": j, "dS0": d(ob[j, 0], ob[j - 1, 3]), "dPV0": d(ob[j, 1], ob[j, 0]), "dS1": d(ob[j, 2], ob[j, 1]),
              "dPV1": d(ob[j, 3], ob[j, 2]), "s0_seen_lag": d(sm[0, j, 0], ob[j, 0])}
        pipe_rows.append(pr)
        print(f"{j:3d} | {pr['dS0']:6d} {pr['dPV0']:6d} {pr['dS1']:6d} {pr['dPV1']:6d} | {pr['s0_seen_lag']:6d}")
    if pipe_rows:
        pk = ["dS0", "dPV0", "dS1", "dPV1", "s0_seen_lag"]
        print("PIPE_MEDIANS", json.dumps({k_: st.median(r[k_] for r in pipe_rows if r["j"] >= 4) for k_ in pk}))
    keys = ["ld", "max", "exp96", "exp32", "wait_next_S", "mma_pfull_lat", "mma_pv6_issue", "mma_plast_lat",
            "mma_tail_issue", "period", "pv_wait", "s_issue_blocked", "s_issue_to_seen"]
    summ = {k_: st.median(r[k_] for r in rows if r["j"] >= 4) for k_ in keys}
    print("MEDIANS", json.dumps(summ))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", os.environ.get("FA_TRACE_OUT", "trace.json")), "w") as f:
        json.dump({"shape": [B, N, H, 128], "level": level, "rows": rows, "medians": summ, "pipe_rows": pipe_rows}, f)


if __name__ == "__main__":
    main()
