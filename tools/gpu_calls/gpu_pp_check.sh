#!/bin/bash
# Validates the ping-pong kernel (FA_SM100_MODE=pp): guarded bring-up, A/B against the pair kernel, GPU suite.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-pp}
FA_SM100_MODE=pp timeout 400 python tools/gpu_bringup.py --levels 1,4 --out gpurun_out/bringup_$TAG.json > gpurun_out/bringup_$TAG.log 2>&1
echo "bringup rc=$?"; cut -c1-420 gpurun_out/bringup_$TAG.log | tail -14
if ! grep -q '"passed_level": 4' gpurun_out/bringup_$TAG.json; then echo "GATE FAILED"; exit 1; fi
for M in pair pp; do
  FA_SM100_MODE=$M timeout 300 python tools/quick_bench.py --reps 15 --check --shapes "4,4096,32;16,512,16;16,1024,16;16,2048,16;8,8192,16;4,16384,16" --out gpurun_out/qb_${TAG}_$M.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$M', r['shape'], 'mean', round(r['tflops_mean'],1), 'best', round(r['tflops_best'],1), 'maxdiff16', r.get('maxdiff_vs_sdpa16'), 'maxdiff32', r.get('maxdiff_vs_sdpa32_subset'))
"
done
FA_SM100_MODE=pp timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.txt
FA_SM100_MODE=pp timeout 200 python tools/sustained_bench.py --seconds 3 --what fa --out gpurun_out/r02_sustained_$TAG.json 2>&1 | tail -1
