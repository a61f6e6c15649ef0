#!/usr/bin/env python
"""Times every library variant under csrc/variants/ on the given shapes (one subprocess each)."""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4,4096,32")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=120, help="per variant and mode; a timeout stops the sweep")
    ap.add_argument("--modes", default="auto", help="comma list of FA_SM100_MODE values (auto, single, pair)")
    args = ap.parse_args()
    vdir = ROOT / "flash_attention_from_scratch_b200" / "csrc" / "variants"
    rows = []
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    for lib in sorted(vdir.glob("libfa_*.so")):
        name = lib.stem[len("libfa_"):]
        if name.startswith("guard_"):  # bring-up builds (tools/build_variants.py --guard), not for timing
            continue
        if args.only and name not in args.only.split(","):
            continue
        for mode in args.modes.split(","):
            env = dict(os.environ, FA_SM100_LIB=str(lib), FA_SM100_MODE=mode)
            try:
                p = subprocess.run([sys.executable, str(ROOT / "tools" / "quick_bench.py"), "--shapes", args.shapes,
                                    "--reps", str(args.reps), "--check"], capture_output=True, text=True, env=env,
                                   timeout=args.timeout)
            except subprocess.TimeoutExpired:
                print(name, mode, "TIMEOUT (hang?) -- stopping the sweep", flush=True)
                rows.append({"variant": name, "mode": mode, "timeout": True})
                with open(args.out, "w") as f:
                    json.dump(rows, f, indent=1)
                return 3
            for line in p.stdout.splitlines():
                try:
                    r = json.loads(line)
                except Exception:  # noqa: BLE001
                    continue
                r["variant"] = name
                r["mode"] = mode
                rows.append(r)
                print(f"{name:14s} {mode:6s} {r['shape']} mean {r['tflops_mean']:.1f} best {r['tflops_best']:.1f} "
                      f"TF/s  maxdiff16 {r.get('maxdiff_vs_sdpa16')}", flush=True)
            if p.returncode != 0:
                print(name, mode, "FAILED", p.stderr[-500:], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
