#!/bin/bash
# ping-pong epilogue with the early o_free: guarded runs, A/B against the generation-13 epilogue, GPU suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g15
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
FA_SM100_MODE=pp FA_SM100_LIB=$G timeout 200 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "16,512,16;16,1024,16;3,640,5;2,128,3;9,128,41;2,2304,3;40,100,7" 2>&1 | cut -c1-60,230-330 | tail -7
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,ppepi0 --shapes "16,512,16;16,1024,16;1,2048,4;3,2048,20;16,256,16;4,4096,32" --modes pp --reps 30 --out gpurun_out/${T}_sweep_ppepi.json 2>&1 | tail -13
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_final6_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_final6_pytest_gpu.txt
