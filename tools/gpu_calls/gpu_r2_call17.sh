#!/bin/bash
# generation 15 (epilogue warpgroup in the shared-S kernels): guarded bring-up, A/B against generation 14, GPU suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g15
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
for M in pair single; do
  FA_SM100_MODE=$M timeout 300 python tools/gpu_bringup.py --quick --out gpurun_out/bringup_${T}_$M.json > gpurun_out/bringup_${T}_$M.log 2>&1
  echo "bringup $M rc=$? $(grep passed_level gpurun_out/bringup_${T}_$M.log)"
  if ! grep -q '"passed_level": 4' gpurun_out/bringup_${T}_$M.json; then echo "GATE FAILED"; cut -c1-500 gpurun_out/bringup_${T}_$M.log | tail -5; exit 1; fi
done
for M in pair single; do
  FA_SM100_MODE=$M FA_SM100_LIB=$G timeout 150 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;2,2304,3;40,100,7" 2>&1 | cut -c1-140 | tail -6
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED ($M)"; exit 1; fi
done
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,noepi,noqpf --shapes "4,4096,32;16,4096,16;16,2048,16;8,8192,16" --modes pair --reps 15 --out gpurun_out/${T}_sweep.json 2>&1 | tail -13
timeout 300 python tools/sweep_variants.py --timeout 100 --only base,noepi --shapes "4,4096,32;16,1024,16" --modes single --reps 15 --out gpurun_out/${T}_sweep_single.json 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
