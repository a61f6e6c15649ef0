#!/bin/bash
# generation 16 candidate (PV_s(j) issued behind S_s(j+1)): guarded runs, then A/B against generation 15
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g16
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
for M in pair single; do
  FA_SM100_MODE=$M FA_SM100_LIB=$G timeout 200 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;2,2304,3;40,100,7;9,128,41;1,128,1" 2>&1 | cut -c1-100,230-330 | tail -8
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED ($M)"; exit 1; fi
done
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,noqkf --shapes "4,4096,32;16,4096,16;16,2048,16;8,8192,16;16,1024,16" --modes pair --reps 15 --out gpurun_out/${T}_sweep.json 2>&1 | tail -11
timeout 300 python tools/sweep_variants.py --timeout 100 --only base,noqkf --shapes "4,4096,32;16,1024,16" --modes single --reps 15 --out gpurun_out/${T}_sweep_single.json 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
