#!/bin/bash
# validation of the cleaned-up tree + A/B of the one-asm MMA groups + the reference-style bench CLIs on the GPU
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g14c
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
for M in pair single; do
  FA_SM100_MODE=$M timeout 300 python tools/gpu_bringup.py --quick --out gpurun_out/bringup_${T}_$M.json > gpurun_out/bringup_${T}_$M.log 2>&1
  echo "bringup $M rc=$? $(grep passed_level gpurun_out/bringup_${T}_$M.log)"
  if ! grep -q '"passed_level": 4' gpurun_out/bringup_${T}_$M.json; then echo "GATE FAILED"; cut -c1-500 gpurun_out/bringup_${T}_$M.log | tail -5; exit 1; fi
done
FA_SM100_LIB=$G timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;2,2304,3" 2>&1 | cut -c1-120 | tail -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,noasm --shapes "4,4096,32;16,4096,16;8,8192,16;4,16384,16" --modes pair --reps 15 --out gpurun_out/${T}_sweep.json 2>&1 | tail -9
for TOOL in racecheck synccheck memcheck; do
  FA_SM100_MODE=pair timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_pair.txt 2>&1
  echo "$TOOL pair rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_pair.txt | tail -1)"
done
KERNELS=tune timeout 600 python tools/benchmark/pt_bench.py --seq_lens 512,1024,2048,4096,8192,16384 --num_repeats 20 --comparators --csv gpurun_out/${T}_pt_bench_tune.csv > gpurun_out/${T}_pt_bench_tune.txt 2>&1; tail -30 gpurun_out/${T}_pt_bench_tune.txt | cut -c1-200
timeout 600 python tools/benchmark/ncu_bench.py --seq_lens 1024,4096 --runs 2 > gpurun_out/${T}_ncu_bench.txt 2>&1; tail -12 gpurun_out/${T}_ncu_bench.txt | cut -c1-220
timeout 300 python tools/debug/sanity_check.py --small > gpurun_out/${T}_sanity_small.txt 2>&1; tail -6 gpurun_out/${T}_sanity_small.txt
timeout 300 bash tools/debug/check_race.sh > gpurun_out/${T}_check_race.txt 2>&1; grep -E "==|SUMMARY" gpurun_out/${T}_check_race.txt | tail -8
