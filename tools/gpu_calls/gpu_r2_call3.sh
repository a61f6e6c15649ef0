#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/sweep_variants.py --only base,ppahead2,emu0,emu8,emu2 --shapes "4,4096,32" --modes pp --reps 10 --out gpurun_out/r02_sweep_pp1.json 2>&1 | tail -8
FA_SM100_MODE=pp timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/r02_pp_v1 python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/r02_pp_v1_ncu.log 2>&1; tail -2 gpurun_out/r02_pp_v1_ncu.log
ls -la gpurun_out/*.ncu-rep
