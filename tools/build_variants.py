#!/usr/bin/env python
"""Builds tuning variants of libfa_sm100.so (different -D knobs) into csrc/variants/ so one GPU
round trip can time all of them (tools/sweep_variants.py).  Development aid.

    python tools/build_variants.py base emu6            # csrc/variants/libfa_{base,emu6}.so
    python tools/build_variants.py psmem --guard        # + libfa_guard_psmem.so for tools/gpu_variant_check.sh
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from flash_attention_from_scratch_b200 import build as fa_build  # noqa: E402

VARIANTS = {
    # name: defines   (every library holds the single-CTA kernel and the CTA-pair kernel; FA_SM100_MODE picks)
    "base": {},
    "nouwarp": {"FA_UNIFORM_WARP": 0},    # warp index straight from threadIdx (generation 7 code shape)
    "emu6": {"FA_EMU_PAIRS": 6},
    "emu8": {"FA_EMU_PAIRS": 8},
    "emu6_4": {"FA_EMU_PAIRS": 6, "FA_EMU_PAIRS_LAST": 4},
    "noldsplit": {"FA_LD_SPLIT": 0},      # S fetched with one wait, row max afterwards
    "g4b": {"FA_SHARED_S": 0},            # single-CTA kernel = generation 4b instead of 6
    "emu4_4": {"FA_EMU_PAIRS_LAST": 4},
    "emu2": {"FA_EMU_PAIRS": 2},
    "emu0": {"FA_EMU_PAIRS": 0},
    "emu6_6": {"FA_EMU_PAIRS": 6, "FA_EMU_PAIRS_LAST": 6},
    "nosplit": {"FA_SPLIT_P": 0},
    "pptoken": {"FA_PP_TOKEN": 1},         # ping-pong kernel with the exp2-phase token
    "ppnoskew": {"FA_PP_SKEW_NS": 0},
    "noasm": {"FA_QK_ONE_ASM": 0},         # CTA-pair kernel: one asm statement per tcgen05.mma instead of per group
    "ppnoprobe": {"FA_PP_PROBE": 0},
    "ppexpv1": {"FA_EXP_VARIANT": 1},
    "hint": {"FA_WAIT_HINT": 10000000},   # CUTLASS-style 10 ms suspend-time hint on every try_wait
    "sleep32": {"FA_WAIT_SLEEP": 32},     # nanosleep back-off in the wait loops
    "kvpf": {"FA_Q_PREFETCH": 2},         # ... plus the next tile's first two K/V blocks when (batch, head) changes
    "r200": {"FA_REGS_SOFTMAX": 200, "FA_REGS_CTRL": 56, "FA_REGS_EPI": 56},
    "trace": {"FA_TRACE": 1},             # production instantiations record the cycle trace (tools/gpu_trace2.py)
    "noqpf": {"FA_Q_PREFETCH": 0},        # generation 15 without the L2 prefetch of the next tile's Q
}


def main():
    args = [a for a in sys.argv[1:] if a != "--guard"]
    guard = "--guard" in sys.argv[1:]  # also build libfa_guard_<name>.so (bounded mbarrier spins that trap)
    names = args or list(VARIANTS)
    out_dir = fa_build.CSRC / "variants"
    out_dir.mkdir(exist_ok=True)

    def one(name):
        flags = tuple(f"-D{k}={v}" for k, v in VARIANTS[name].items())
        path = fa_build.build(force=True, extra_flags=flags, out=out_dir / f"libfa_{name}.so")
        if guard:
            fa_build.build(force=True, extra_flags=flags + ("-DFA_HANG_GUARD=1",),
                           out=out_dir / f"libfa_guard_{name}.so")
        return name, path

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for name, path in ex.map(one, names):
            print(name, path)


if __name__ == "__main__":
    main()
