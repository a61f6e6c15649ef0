#!/bin/bash
# generation 15 with AUTO = ping-pong up to 1024: launch-overhead probe, GPU suite, sanitizers on the two kernels that changed
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g15
timeout 300 python tools/launch_overhead.py --out gpurun_out/${T}_launch_overhead.json 2>&1 | tail -24
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
for M in pair single; do
  for TOOL in racecheck synccheck memcheck; do
    # 640 = 5 KV blocks, ragged (masked tail); 3 heads
    FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_${M}.txt 2>&1
    echo "$TOOL $M 640 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_${M}.txt | tail -1)"
  done
  # several work tiles per CTA (pair: 160 tiles on 74 pairs): the epilogue warpgroup's hand-overs across tiles
  FA_SM100_MODE=$M timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 256 --batch 10 --n_heads 16 --n_runs 1 > gpurun_out/${T}_racecheck_${M}_multitile.txt 2>&1
  echo "racecheck $M multitile rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${T}_racecheck_${M}_multitile.txt | tail -1)"
  FA_SM100_MODE=$M timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 256 --batch 10 --n_heads 16 --n_runs 1 > gpurun_out/${T}_synccheck_${M}_multitile.txt 2>&1
  echo "synccheck $M multitile rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/${T}_synccheck_${M}_multitile.txt | tail -1)"
done
