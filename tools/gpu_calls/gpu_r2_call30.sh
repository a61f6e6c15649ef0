#!/bin/bash
# grid-aware AUTO: small / awkward grids in the three modes (does the cost model pick the faster kernel?)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for M in pair pp auto; do
  FA_SM100_MODE=$M timeout 200 python tools/quick_bench.py --reps 30 --warmup 5 --check --shapes "1,2048,4;2,4096,25;1,8192,12;1,4096,8;3,2048,20;1,16384,9;5,1536,16" --out gpurun_out/r02_g15_auto_grid_$M.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('$M', r['shape'], 'ms', round(r['ms_mean'],4), 'TF', round(r['tflops_mean'],1), 'maxdiff', r.get('maxdiff_vs_sdpa16'))
"
done
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -2
