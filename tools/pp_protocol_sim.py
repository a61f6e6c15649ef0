#!/usr/bin/env python
"""Deadlock check of the kernels' barrier protocols on the CPU: `simulate` = the ping-pong kernel
(csrc/fa_fwd_pp_sm100.cuh), `simulate_shared_s` = the CTA-pair / single-CTA kernels (csrc/fa_fwd_sm100.cuh).

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


def simulate_shared_s(n, tiles, ring=4, rescale=False, epi_wg=True, verbose=False, lazy=None):
    """The shared-S kernels of csrc/fa_fwd_sm100.cuh (generation 15: CTA pairs and single CTAs): one TMA producer,
    the S-issuing warp, the PV-issuing warp, two softmax warpgroups (one per Q tile) and the epilogue warpgroup.
    `ring` = K (and V) ring slots; `rescale` = every block takes the lazy-rescale path (it waits for the previous PV
    of its tile: data dependent in the kernel, so both extremes are checked); `epi_wg=False` = generation 14 (the
    softmax warpgroups run their own epilogue).  Besides deadlocks it checks the PARITY discipline: a wait that
    finds its barrier more than one completion ahead would, with the kernel's parity waits, block for ever -- the
    scheduler lets every agent run as far ahead as the protocol allows, so such a barrier shows up here.
    `lazy` names an agent that only moves when nobody else can (the adversarial schedule for that agent's waits).
    Returns True when every agent finishes and no barrier ever ran ahead of a waiter."""
    done = {}
    ahead = []

    def cnt(name):
        return done.get(name, 0)

    def sig(name):
        done[name] = cnt(name) + 1

    def producer():
        x_k = x_v = 0  # K / V blocks loaded so far (all tiles): ring slot and use count

        def load(kind, x):
            slot, use = x % ring, x // ring
            if use > 0:
                yield ("%s_empty%d" % (kind, slot), use)
            sig("%s_full%d" % (kind, slot))

        for it in range(tiles):
            if it > 0:
                yield ("q_empty0", it)
            sig("q_full0")
            yield from load("k", x_k); x_k += 1
            if it > 0:
                yield ("q_empty1", it)
            sig("q_full1")
            yield from load("v", x_v); x_v += 1
            for _ in range(1, n):
                yield from load("k", x_k); x_k += 1
                yield from load("v", x_v); x_v += 1

    def mma_s():
        kb = u = 0
        for it in range(tiles):
            for jj in range(n):
                yield ("k_full%d" % (kb % ring), kb // ring + 1)
                for s in (0, 1):
                    if jj == 0:
                        yield ("q_full%d" % s, it + 1)
                    if u > 0:
                        yield ("s_free", u)          # the previous S (either tile) was read out
                    sig("s_full%d" % s)
                    if jj == n - 1:
                        sig("q_empty%d" % s)
                    if s == 1:
                        sig("k_empty%d" % (kb % ring))
                    u += 1
                kb += 1

    def mma_pv():
        vb = 0
        for it in range(tiles):
            for j in range(n):
                g = it * n + j
                yield ("v_full%d" % (vb % ring), vb // ring + 1)
                for s in (0, 1):
                    yield ("p_full%d" % s, g + 1)
                    if j == 0 and it > 0:
                        yield ("o_free%d" % s, it)   # the previous tile's O_s was read out
                    yield ("p_last%d" % s, g + 1)
                    sig("pv_done%d" % s)
                    if s == 1:
                        sig("v_empty%d" % (vb % ring))
                vb += 1

    def softmax(s):
        for it in range(tiles):
            for j in range(n):
                g = it * n + j
                yield ("s_full%d" % s, g + 1)
                sig("s_free")
                if rescale and j > 0:
                    yield ("pv_done%d" % s, g)       # O_s quiescent before it is rescaled
                if (g > 0) if epi_wg else (j > 0):
                    yield ("pv_done%d" % s, g)       # P_s(g-1) consumed before P_s(g) overwrites it
                sig("p_full%d" % s)
                sig("p_last%d" % s)
            if epi_wg:
                if it > 0:
                    yield ("l_free%d" % s, it)       # named barrier 12 + s
                sig("l_ready%d" % s)                 # named barrier 10 + s
            else:
                yield ("pv_done%d" % s, (it + 1) * n)
                sig("o_free%d" % s)

    def epilogue():
        for it in range(tiles):
            for s in (0, 1):
                yield ("l_ready%d" % s, it + 1)
                sig("l_free%d" % s)
                yield ("pv_done%d" % s, (it + 1) * n)
                sig("o_free%d" % s)

    agents = {"producer": producer(), "mma_s": mma_s(), "mma_pv": mma_pv(), "wg0": softmax(0), "wg1": softmax(1)}
    if epi_wg:
        agents["epilogue"] = epilogue()
    waiting = {}
    for name, gen in list(agents.items()):
        try:
            waiting[name] = next(gen)
        except StopIteration:
            del agents[name]
    def step(name, once=False):
        moved = False
        while name in agents and cnt(waiting[name][0]) >= waiting[name][1]:
            bar, need = waiting[name]
            if cnt(bar) > need:  # the barrier ran ahead of this waiter: a parity wait could miss the phase
                ahead.append((name, bar, need, cnt(bar)))
            moved = True
            try:
                waiting[name] = next(agents[name])
            except StopIteration:
                del agents[name]
            if once:
                break
        return moved

    progress = True
    while agents and progress:
        progress = False
        for name in list(agents):
            if name != lazy and step(name):
                progress = True
        if not progress and lazy in agents:
            progress = step(lazy, once=True)  # one wait at a time, then everybody else runs ahead again
    if agents or ahead:
        if verbose:
            print("FAIL n=%d tiles=%d ring=%d:" % (n, tiles, ring),
                  {a: (waiting[a], cnt(waiting[a][0])) for a in agents}, ahead[:4])
        return False
    return True


def main():
    bad = 0
    for n, tiles, ks, vs in itertools.product(range(1, 9), range(1, 6), (2, 4), (2, 4)):
        if not simulate(n, tiles, ks, vs, verbose=True):
            bad += 1
    for n, tiles, ring, rescale, epi in itertools.product(range(1, 9), range(1, 6), (2, 4), (False, True), (True, False)):
        for lazy in (None, "producer", "mma_s", "mma_pv", "wg0", "wg1", "epilogue"):
            if not simulate_shared_s(n, tiles, ring, rescale, epi, verbose=True, lazy=lazy):
                bad += 1
    print("deadlocks:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
