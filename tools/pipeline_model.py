#!/usr/bin/env python
"""Dataflow model of the attention kernel's inner loop (planning aid, CPU only).

Three kinds of agents advance their own clocks and meet at mbarrier-like events, exactly as in
csrc/fa_fwd_sm100.cuh: two softmax warpgroups (one per Q tile), the single MMA-issuing warp (an in-order
program of waits and issues) and the in-order tensor pipe.  All latencies are measured numbers of this
round (profiles/r01_g9_notes.md, r01_mma_probe_latency.jsonl, ncu warp-state samples):

    tcgen05.ld of S + hand-off arrive   370 clk     softmax of a block after the ld   ~1630 clk (W ~ 2000)
    8 MMAs of a group                   512 clk     group issued on an idle pipe      +140 clk before it starts
    commit -> waiter sees the barrier   100 clk     barrier arrive -> MMA warp issues  150 clk

Every schedule variant is a different MMA-warp program / dependency set:

    gen9     one S accumulator shared by both Q tiles (production)
    psmem    one S accumulator per Q tile, P through shared memory (-DFA_P_SMEM=1, not yet run)
    psmem2   same, both S of a block issued before the two PV (-DFA_P_SMEM=2)
    g4b      per-tile S with P aliased on it: S_s(j+1) behind PV_s(j) (generation 4b)

The model reproduces the measured loop periods of the variants that were run (gen9 ~2690 clk, g4b slower)
and is used to rank the ones that were not.  It knows nothing about power, clocks or smem bandwidth.

    python tools/pipeline_model.py            # table of all variants
"""
import argparse
from collections import defaultdict

P = dict(ld=370, softmax_to_pfull=1230, softmax_to_plast=400, mma_group=512, pv_first=384, pv_tail=128,
         cold_fill=140, commit_seen=100, arrive_to_issue=150, issue_per_group=100)


class Sim:
    def __init__(self, n_blocks, prm):
        self.n = n_blocks
        self.p = prm
        self.ev = {}                      # event key -> time it becomes visible to waiters
        self.pipe_free = 0.0              # tensor pipe: time the last queued MMA retires
        self.busy = 0.0

    # ---- tensor pipe: in-order; a group issued while the pipe is idle pays the operand-fetch fill
    def pipe(self, t_issue, clk):
        start = t_issue + self.p["cold_fill"] if self.pipe_free <= t_issue else self.pipe_free
        self.pipe_free = start + clk
        self.busy += clk
        return self.pipe_free + self.p["commit_seen"]

    def seen(self, key):
        return self.ev.get(key)


def run(variant, n_blocks=32, prm=None):
    prm = dict(P, **(prm or {}))
    sim = Sim(n_blocks, prm)
    ev = sim.ev
    n = n_blocks
    # MMA-warp program: list of (kind, tile, block)
    prog = [("S", 0, 0), ("S", 1, 0)]
    if variant == "gen9":
        prog.append(("S", 0, 1))
        for j in range(n):
            prog += [("PV", 0, j), ("S", 1, j + 1), ("PV", 1, j), ("S", 0, j + 2)]
    elif variant == "psmem":
        for j in range(n):
            prog += [("S", 0, j + 1), ("PV", 0, j), ("S", 1, j + 1), ("PV", 1, j)]
    elif variant == "psmem2":
        for j in range(n):
            prog += [("S", 0, j + 1), ("S", 1, j + 1), ("PV", 0, j), ("PV", 1, j)]
    elif variant == "g4b":
        for j in range(n):
            prog += [("PVS", 0, j), ("PVS", 1, j)]
    else:
        raise ValueError(variant)
    prog = [x for x in prog if x[2] < n]

    # softmax warpgroups as lazily evaluated event chains: block j of tile s
    def softmax_block(s, j):
        """Fills ev for block j of tile s once S_s(j) is visible; returns False if it is not yet."""
        if ("sm_done", s, j) in ev:
            return True
        t_s = ev.get(("s_full", s, j))
        prev = ev.get(("sm_done", s, j - 1), 0.0) if j > 0 else 0.0
        if t_s is None or (j > 0 and ("sm_done", s, j - 1) not in ev):
            return False
        t = max(t_s, prev)
        ev[("s_read", s, j)] = t + prm["ld"]                       # S is in registers -> s_free arrive
        t_exp = t + prm["ld"]
        if j > 0:                                                    # P_s buffer reuse: PV_s(j-1) retired
            pv = ev.get(("pv_done", s, j - 1))
            if pv is None:
                return False
            # the wait sits a quarter into the exp phase; it only bites if PV is very late
            t_exp = max(t_exp, pv - 0.25 * prm["softmax_to_pfull"])
        ev[("p_full", s, j)] = t_exp + prm["softmax_to_pfull"]
        ev[("p_last", s, j)] = ev[("p_full", s, j)] + prm["softmax_to_plast"]
        ev[("sm_done", s, j)] = ev[("p_last", s, j)]
        return True

    t_mma = 0.0
    pc = 0
    s_count = 0                        # gen9: S accumulators issued so far (order of the shared buffer)
    s_order = []
    guard = 0
    while pc < len(prog):
        guard += 1
        if guard > 100000:
            raise RuntimeError("model deadlock")
        for s in (0, 1):               # let the softmax agents catch up
            for j in range(n):
                if not softmax_block(s, j):
                    break
        kind, s, j = prog[pc]
        deps = []
        if kind == "S":
            if variant == "gen9":
                if s_count > 0:
                    ps, pj = s_order[s_count - 1]
                    deps.append(("s_read", ps, pj))
            elif j > 0:
                deps.append(("s_read", s, j - 1))
            if variant == "g4b" and j > 0:
                deps.append(("p_last", s, j - 1))
        elif kind in ("PV", "PVS"):
            deps.append(("p_full", s, j))
        if any(d not in ev for d in deps):
            continue                   # softmax has to advance first (loop again)
        t_ready = max([t_mma] + [ev[d] + prm["arrive_to_issue"] for d in deps])
        if kind == "S":
            t_issue = t_ready
            ev[("s_full", s, j)] = sim.pipe(t_issue, prm["mma_group"])
            t_mma = t_issue + prm["issue_per_group"]
            s_order.append((s, j))
            s_count += 1
        else:
            sim.pipe(t_ready, prm["pv_first"])
            t_mid = t_ready + prm["issue_per_group"]
            if ("p_last", s, j) not in ev:
                continue
            t_tail = max(t_mid, ev[("p_last", s, j)] + prm["arrive_to_issue"])
            ev[("pv_done", s, j)] = sim.pipe(t_tail, prm["pv_tail"])
            t_mma = t_tail + 40
            if kind == "PVS" and j + 1 < n:      # generation 4b: S_s(j+1) right behind PV_s(j)
                ev[("s_full", s, j + 1)] = sim.pipe(t_mma, prm["mma_group"])
                t_mma += prm["issue_per_group"]
        pc += 1
    for s in (0, 1):
        for j in range(n):
            softmax_block(s, j)
    # steady-state period: spacing of block completions of tile 0 in the second half of the run
    a, b = n // 2, n - 2
    period = (ev[("sm_done", 0, b)] - ev[("sm_done", 0, a)]) / (b - a)
    wait = period - (prm["ld"] + prm["softmax_to_pfull"] + prm["softmax_to_plast"])
    return {"variant": variant, "period": period, "mma_bound": 4 * prm["mma_group"],
            "pipe_util": 4 * prm["mma_group"] / period, "softmax_wait": wait}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--blocks", type=int, default=32)
    a = ap.parse_args()
    print(f"{'variant':8s} {'period clk/block':>17s} {'tensor pipe util':>17s} {'softmax idle clk':>17s}")
    for v in ("gen9", "g4b", "psmem", "psmem2"):
        r = run(v, a.blocks)
        print(f"{v:8s} {r['period']:17.0f} {r['pipe_util']:17.2f} {r['softmax_wait']:17.0f}")


if __name__ == "__main__":
    main()
