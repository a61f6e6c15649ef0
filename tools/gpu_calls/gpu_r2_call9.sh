#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
FA_SM100_MODE=pp timeout 200 python tools/gpu_bringup.py --levels 1,4 --quick --out gpurun_out/bringup_pp9.json > gpurun_out/bringup_pp9.log 2>&1
echo "bringup rc=$?"; grep passed_level gpurun_out/bringup_pp9.log
if ! grep -q '"passed_level": 4' gpurun_out/bringup_pp9.json; then echo "GATE FAILED"; cut -c1-600 gpurun_out/bringup_pp9.log | tail; exit 1; fi
FA_SM100_LIB=$G FA_SM100_MODE=pp timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;1,384,2" 2>&1 | cut -c1-230 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
timeout 500 python tools/sweep_variants.py --timeout 100 --only base,ppnotoken,ppnoskew,emu2,emu6 --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pp --reps 10 --out gpurun_out/r02_sweep_pp5.json 2>&1 | tail -13
for X in ppnotoken; do echo "== $X"; FA_SM100_LIB=$V/libfa_$X.so FA_SM100_MODE=pp timeout 120 python tools/gpu_pp_trace.py --out gpurun_out/pp_trace3_$X.json > gpurun_out/pp_trace3_$X.txt 2>&1; tail -4 gpurun_out/pp_trace3_$X.txt; done
echo "== base"; FA_SM100_MODE=pp timeout 120 python tools/gpu_pp_trace.py --out gpurun_out/pp_trace3_base.json > gpurun_out/pp_trace3_base.txt 2>&1; tail -4 gpurun_out/pp_trace3_base.txt
FA_SM100_MODE=pp timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
