#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g14b
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
FA_SM100_MODE=pp timeout 200 python tools/gpu_bringup.py --levels 1,4 --quick --out gpurun_out/bringup_$T.json > gpurun_out/bringup_$T.log 2>&1
echo "bringup rc=$? $(grep passed_level gpurun_out/bringup_$T.log)"
if ! grep -q '"passed_level": 4' gpurun_out/bringup_$T.json; then echo "GATE FAILED"; cut -c1-500 gpurun_out/bringup_$T.log | tail -4; exit 1; fi
FA_SM100_MODE=pp timeout 100 python tools/quick_bench.py --reps 3 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;1,384,2" 2>&1 | cut -c1-120 | tail -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "PP RUN FAILED"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
for M in pair pingpong; do
  for TOOL in racecheck synccheck memcheck; do
    FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_${M}.txt 2>&1
    echo "$TOOL $M rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_${M}.txt | tail -1)"
  done
  FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 2304 --batch 1 --n_heads 2 --n_runs 1 > gpurun_out/${T}_racecheck_${M}_2304.txt 2>&1
  echo "racecheck $M 2304 rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${T}_racecheck_${M}_2304.txt | tail -1)"
done
for M in single; do for TOOL in racecheck synccheck memcheck; do
    FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_${M}.txt 2>&1
    echo "$TOOL $M rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_${M}.txt | tail -1)"
done; done
timeout 300 python tools/sweep_variants.py --timeout 100 --only base --shapes "4,4096,32;16,512,16;16,1024,16;16,2048,16" --modes pp,auto --reps 12 --out gpurun_out/${T}_sweep.json 2>&1 | tail -9
