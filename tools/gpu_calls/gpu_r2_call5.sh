#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
FA_SM100_MODE=pp timeout 200 python tools/gpu_bringup.py --levels 1,4 --quick --out gpurun_out/bringup_pp5.json > gpurun_out/bringup_pp5.log 2>&1
echo "bringup rc=$?"; grep passed_level gpurun_out/bringup_pp5.log
if ! grep -q '"passed_level": 4' gpurun_out/bringup_pp5.json; then echo "GATE FAILED"; cut -c1-400 gpurun_out/bringup_pp5.log | tail; exit 1; fi
# full-size shapes first on the hang-guard build (bounded spins trap instead of hanging)
FA_SM100_LIB=$G FA_SM100_MODE=pp timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;1,384,2" 2>&1 | cut -c1-260 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
timeout 500 python tools/sweep_variants.py --timeout 100 --only base,ppahead3,emu2,emu6 --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pp --reps 10 --out gpurun_out/r02_sweep_pp2.json 2>&1 | tail -13
FA_SM100_MODE=pp timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/r02_pp_v2 python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/r02_pp_v2_ncu.log 2>&1; tail -1 gpurun_out/r02_pp_v2_ncu.log
FA_SM100_MODE=pp timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
FA_SM100_MODE=pp timeout 100 python tools/sustained_bench.py --seconds 3 --what fa 2>&1 | tail -1 | cut -c1-400
