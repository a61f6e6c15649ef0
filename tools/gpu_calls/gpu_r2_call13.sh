#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
for M in pair single; do
  FA_SM100_MODE=$M timeout 300 python tools/gpu_bringup.py --quick --out gpurun_out/bringup_g14_$M.json > gpurun_out/bringup_g14_$M.log 2>&1
  echo "bringup $M rc=$? $(grep passed_level gpurun_out/bringup_g14_$M.log)"
  if ! grep -q '"passed_level": 4' gpurun_out/bringup_g14_$M.json; then echo "GATE FAILED"; cut -c1-500 gpurun_out/bringup_g14_$M.log | tail -5; exit 1; fi
  FA_SM100_LIB=$G FA_SM100_MODE=$M timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3" 2>&1 | cut -c1-120 | tail -4
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
done
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,nodual --shapes "4,4096,32;16,512,16;16,1024,16;16,2048,16;8,8192,16;4,16384,16" --modes pair,single --reps 12 --out gpurun_out/r02_sweep_g14.json 2>&1 | tail -26
timeout 200 python tools/sweep_variants.py --timeout 100 --only base --shapes "4,4096,32;16,512,16;16,1024,16;16,2048,16;8,8192,16;4,16384,16" --modes pp --reps 12 --out gpurun_out/r02_sweep_g14_pp.json 2>&1 | tail -7
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
