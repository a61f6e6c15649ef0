#!/usr/bin/env python
"""Before / after comparison of two SASS listings written by tools/dump_sass.sh (the role of the reference's
tools/analysis/* diff scripts, for sm_100a mnemonics): instruction count, histogram delta of the mnemonics that
matter here (UTCHMMA, R2UR, SYNCS, LDTM/STTM, MUFU, FFMA2/FADD2, F2FP, FMNMX, UTMALDG/UTMASTG, spills), and the
R2UR count between every barrier wait and the first UTCHMMA of an issue group.

    python tools/sass_diff.py profiles/r01_g9_fa_fwd_kernel_pair_bf16.sass profiles/r02_g14_fa_fwd_kernel_pair_bf16.sass
"""
import collections
import re
import sys

WATCH = ["UTCHMMA", "UTCBAR", "R2UR", "SYNCS", "LDTM", "STTM", "MUFU", "FFMA2", "FADD2", "F2FP", "FMNMX", "UTMALDG",
         "UTMASTG", "STL", "LDL", "BAR", "LDS", "STS", "ELECT"]


def load(path):
    ops = []
    for ln in open(path):
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            o = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().split(";")[0].split()
            if o:
                ops.append(o[1] if o[0].startswith("@") and len(o) > 1 else o[0])
    return ops


def issue_gaps(ops):
    gaps, since, in_group = [], None, False
    for o in ops:
        if o.startswith("SYNCS.PHASECHK"):
            since, in_group = 0, False
        elif o.startswith("UTCHMMA"):
            if not in_group and since is not None:
                gaps.append(since)
            in_group = True
        elif since is not None and not in_group and o.startswith("R2UR"):
            since += 1
    return gaps


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    print(f"{'':10s} {'before':>8s} {'after':>8s} {'delta':>7s}")
    print(f"{'instr':10s} {len(a):8d} {len(b):8d} {len(b) - len(a):+7d}")
    ha = collections.Counter(next((w for w in WATCH if o.startswith(w)), None) for o in a)
    hb = collections.Counter(next((w for w in WATCH if o.startswith(w)), None) for o in b)
    for w in WATCH:
        if ha[w] or hb[w]:
            print(f"{w:10s} {ha[w]:8d} {hb[w]:8d} {hb[w] - ha[w]:+7d}")
    print("R2UR between a barrier wait and the first UTCHMMA of a group:")
    print("  before", issue_gaps(a))
    print("  after ", issue_gaps(b))


if __name__ == "__main__":
    main()
