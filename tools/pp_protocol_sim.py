#!/usr/bin/env python
"""Deadlock check of the ping-pong kernel's barrier protocol (csrc/fa_fwd_pp_sm100.cuh) on the CPU.

Every agent of one CTA pair (the two TMA producers, the two MMA-issuing warps, the two softmax warpgroups, the
epilogue warpgroup) is a generator that yields
the conditions it waits for, in the order the kernel's code does; barriers are completion counters.  The
scheduler runs agents round-robin until all finish or nobody can move (deadlock -> the stuck waits are
printed).  It does not model time, only ordering: what it proves is that the in-order streams cannot block
each other for a given (n_blocks, n_tiles, ring depths).  Written after an earlier version of the kernel (one
issuing warp, S issued four blocks ahead) hung a B200: the in-order warp had parked S(g+4) in front of the PV an
epilogue was waiting for.

    python tools/pp_protocol_sim.py            # sweep; exit code 1 on any deadlock
"""
import itertools
import sys


def simulate(n, tiles, k_stages=4, v_stages=4, verbose=False):
    total = n * tiles
    done = {}  # barrier name -> completions

    def cnt(name):
        return done.get(name, 0)

    def sig(name):
        done[name] = cnt(name) + 1

    def producer_k():
        for x in range(total):
            it, j = divmod(x, n)
            if j == 0:
                if it >= 2:
                    yield ("q_empty%d" % (it & 1), (it >> 1))      # (it>>1)-th release of that Q buffer
                sig("q_full%d" % (it & 1))
            if x >= k_stages:
                yield ("k_empty%d" % (x % k_stages), x // k_stages)
            sig("k_full%d" % (x % k_stages))

    def producer_v():
        for g in range(total):
            if g >= v_stages:
                yield ("v_empty%d" % (g % v_stages), g // v_stages)
            sig("v_full%d" % (g % v_stages))

    def mma_s():
        for x in range(total):
            it, j = divmod(x, n)
            w = x & 1
            if j == 0:
                yield ("q_full%d" % (it & 1), (it >> 1) + 1)
            yield ("k_full%d" % (x % k_stages), x // k_stages + 1)
            yield ("s_free%d" % w, x >> 1)                         # S(x-2) read out
            sig("s_full%d" % w)
            sig("k_empty%d" % (x % k_stages))
            if j == n - 1:
                sig("q_empty%d" % (it & 1))

    def mma_pv():
        for g in range(total):
            it, j = divmod(g, n)
            w = g & 1
            yield ("v_full%d" % (g % v_stages), g // v_stages + 1)
            yield ("p_full%d" % w, (g >> 1) + 1)
            if j == 0:
                yield ("o_free", it)                               # the previous tile's epilogue has read O
            yield ("p_last%d" % w, (g >> 1) + 1)
            sig("pv_done%d" % w)
            sig("v_empty%d" % (g % v_stages))

    def softmax(w):
        for it in range(tiles):
            g0 = it * n
            for j in range((g0 + w) & 1, n, 2):
                g = g0 + j
                yield ("s_full%d" % w, (g >> 1) + 1)
                sig("s_free%d" % w)
                if g > 0:  # the row-state slot has one writer at a time, in block order
                    yield ("m_ready%d" % (w ^ 1), ((g - 1) >> 1) + 1)
                sig("m_ready%d" % w)
                if g > 0:
                    yield ("exp_done%d" % (w ^ 1), ((g - 1) >> 1) + 1)  # exp2 phases alternate in block order
                yield ("pv_done%d" % w, g >> 1)                    # P(g-2) consumed
                sig("exp_done%d" % w)
                sig("p_full%d" % w)
                sig("p_last%d" % w)
            if it > 0:
                yield ("lm_free", it)
            sig("lm_ready%d" % w)

    def epilogue():
        for it in range(tiles):
            gl = (it + 1) * n - 1
            yield ("lm_ready0", it + 1)
            yield ("lm_ready1", it + 1)
            sig("lm_free")
            yield ("pv_done%d" % (gl & 1), (gl >> 1) + 1)
            sig("o_free")

    agents = {"producer_k": producer_k(), "producer_v": producer_v(), "mma_s": mma_s(), "mma_pv": mma_pv(),
              "wg0": softmax(0), "wg1": softmax(1), "epilogue": epilogue()}
    waiting = {}
    for name, gen in list(agents.items()):
        try:
            waiting[name] = next(gen)
        except StopIteration:
            del agents[name]
    progress = True
    while agents and progress:
        progress = False
        for name in list(agents):
            while name in agents and cnt(waiting[name][0]) >= waiting[name][1]:
                progress = True
                try:
                    waiting[name] = next(agents[name])
                except StopIteration:
                    del agents[name]
    if agents:
        if verbose:
            print("DEADLOCK n=%d tiles=%d:" % (n, tiles),
                  {a: (waiting[a], cnt(waiting[a][0])) for a in agents})
        return False
    return True


def main():
    bad = 0
    for n, tiles, ks, vs in itertools.product(range(1, 9), range(1, 6), (2, 4), (2, 4)):
        if not simulate(n, tiles, ks, vs, verbose=True):
            bad += 1
    print("deadlocks:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
