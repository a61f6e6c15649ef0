#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
FA_SM100_MODE=pp timeout 300 python tools/gpu_bringup.py --levels 1,4 --quick --out gpurun_out/bringup_pp4.json > gpurun_out/bringup_pp4.log 2>&1
echo "bringup rc=$?"; grep -c '"full_maxerr"' gpurun_out/bringup_pp4.log; grep passed_level gpurun_out/bringup_pp4.log
if ! grep -q '"passed_level": 4' gpurun_out/bringup_pp4.json; then echo "GATE FAILED"; cut -c1-400 gpurun_out/bringup_pp4.log | tail; exit 1; fi
timeout 600 python tools/sweep_variants.py --only base,ppahead3,emu0,emu2,emu6 --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pp --reps 10 --out gpurun_out/r02_sweep_pp2.json 2>&1 | tail -16
timeout 200 python tools/sweep_variants.py --only base --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pair --reps 10 --out gpurun_out/r02_sweep_pp2_pair.json 2>&1 | tail -3
FA_SM100_MODE=pp timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/r02_pp_v2 python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/r02_pp_v2_ncu.log 2>&1; tail -1 gpurun_out/r02_pp_v2_ncu.log
FA_SM100_MODE=pp timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -2
