#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
FA_SM100_LIB=$G FA_SM100_MODE=pp timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;1,384,2" 2>&1 | cut -c1-120 | tail -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
FA_SM100_MODE=pp timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -40
timeout 500 python tools/sweep_variants.py --timeout 100 --only base,ppnoprobe,ppmbar,ppmbarnoprobe --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pp --reps 12 --out gpurun_out/r02_sweep_pp8.json 2>&1 | tail -12
timeout 500 python tools/sweep_variants.py --timeout 100 --only base --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pair --reps 12 --out gpurun_out/r02_sweep_pp8_pair.json 2>&1 | tail -3
