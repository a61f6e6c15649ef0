#!/usr/bin/env python
"""Cycle timeline of the shared-S kernels' inner loop (debug instantiation, level 5): %clock stamps of CTA 0 for
its SECOND work tile, from the two softmax warpgroups (warp 0 of each), the S-issuing warp and the PV-issuing warp,
while the whole grid runs the problem.  Prints, per KV block and as medians, every leg of the serial chain through
the shared S accumulator
    tcgen05.ld S_s -> s_free arrive -> issuing warp sees it -> S of the other tile issued -> ... -> its warpgroup sees S
and where each PV group went.  Development aid (successor of tools/gpu_trace.py for generation 14+).

    FA_SM100_MODE=pair python tools/gpu_trace2.py [B N H]        (default 4 4096 32)
"""
import ctypes as C
import json
import os
import statistics as st
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flash_attention_from_scratch_b200 import _lib  # noqa: E402

TRACE_BASE = 2 * 128 * 128 + 512


def main():
    B, N, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4, 4096, 32)))
    lib = _lib.load()
    torch.manual_seed(0)
    q = torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    o = torch.empty_like(q)
    dump = torch.zeros(TRACE_BASE + 2048, device="cuda", dtype=torch.float32)
    diag = torch.zeros(256, device="cuda", dtype=torch.int32)
    knobs = (C.c_uint32 * 8)(0, 0, 0, 0, 0, 0, 0, int(os.environ.get('FA_TRACE_LEVEL', '5')))  # 40: FA_TRACE builds
    sb, sn, sh, _ = q.stride()
    for _ in range(2):
        rc = lib.fa_fwd_debug(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, 128, sb, sn, sh, 15,
                              dump.data_ptr(), knobs, diag.data_ptr())
        assert rc == 0, _lib.last_error()
    tr = dump.view(torch.int32)[TRACE_BASE:].cpu().numpy().astype("int64") & 0xFFFFFFFF
    nb = min(32, N // 128)
    sm = tr[:512].reshape(2, 32, 8)[:, :nb]
    pv = tr[512:768].reshape(32, 2, 4)[:nb]
    qk = tr[768:1024].reshape(32, 2, 4)[:nb]
    t0 = int(sm[0, 0, 0])
    rel = lambda x: int((int(x) - t0) & 0xFFFFFFFF)  # noqa: E731
    d = lambda a, b: int((int(a) - int(b)) & 0xFFFFFFFF)  # noqa: E731
    rows = []
    print("blk s | S_seen  ld  max  exp96 exp32 | waitS period | sfree->met  met->issued  issued->Sseen(next in chain) "
          "| pfull->pvmet  pv1->plast_seen  plast->seen  tail | qk_met-pv_met")
    for j in range(1, nb - 1):
        for s in range(2):
            e, nx = sm[s, j], sm[s, j + 1]
            # next S in the shared accumulator's order after S_s(j): S_1(j) if s == 0 else S_0(j+1)
            ns, nj = (1, j) if s == 0 else (0, j + 1)
            cq = qk[nj, ns]
            r = {"j": j, "s": s, "S_seen": rel(e[0]), "ld": d(e[1], e[0]), "max": d(e[2], e[1]), "exp96": d(e[3], e[2]),
                 "exp32": d(e[4], e[3]), "waitS": d(nx[0], e[4]), "period": d(nx[0], e[0]),
                 "sfree_to_met": d(cq[1], e[1]), "met_to_issued": d(cq[2], cq[1]),
                 "issued_to_seen": d(sm[ns, nj, 0], cq[2]),
                 "pfull_to_pvmet": d(pv[j, s][1], e[3]), "pv1_to_plastseen": d(pv[j, s][2], pv[j, s][1]),
                 "plast_to_seen": d(pv[j, s][2], e[4]), "tail": d(pv[j, s][3], pv[j, s][2]),
                 # did S_s(j+1) get its dependencies before PV_s(j) did?  (> 0: PV first)
                 "qkmet_minus_pvmet": (rel(qk[j + 1, s][1]) - rel(pv[j, s][1]))}
            for key in ("sfree_to_met", "issued_to_seen", "pfull_to_pvmet", "plast_to_seen"):
                if r[key] > (1 << 31):
                    r[key] -= (1 << 32)
            rows.append(r)
            print(f"{j:3d} {s} | {r['S_seen']:7d} {r['ld']:4d} {r['max']:4d} {r['exp96']:5d} {r['exp32']:5d} | "
                  f"{r['waitS']:5d} {r['period']:6d} | {r['sfree_to_met']:6d} {r['met_to_issued']:6d} {r['issued_to_seen']:6d} | "
                  f"{r['pfull_to_pvmet']:6d} {r['pv1_to_plastseen']:6d} {r['plast_to_seen']:6d} {r['tail']:5d} | "
                  f"{r['qkmet_minus_pvmet']:6d}")
    keys = [k_ for k_ in rows[0] if k_ not in ("j", "s", "S_seen")]
    for s in range(2):
        med = {k_: st.median(r[k_] for r in rows if r["j"] >= 6 and r["s"] == s) for k_ in keys}
        print(f"MEDIANS s={s}", json.dumps(med))
    out = os.path.join(ROOT, "gpurun_out", os.environ.get("FA_TRACE_OUT", "trace2.json"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"shape": [B, N, H, 128], "rows": rows}, f)


if __name__ == "__main__":
    main()
