#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
FA_SM100_MODE=pp timeout 200 python tools/gpu_bringup.py --levels 1,4 --quick --out gpurun_out/bringup_pp11.json > gpurun_out/bringup_pp11.log 2>&1
echo "bringup rc=$?"; grep passed_level gpurun_out/bringup_pp11.log
if ! grep -q '"passed_level": 4' gpurun_out/bringup_pp11.json; then echo "GATE FAILED"; cut -c1-600 gpurun_out/bringup_pp11.log | tail; exit 1; fi
FA_SM100_LIB=$G FA_SM100_MODE=pp timeout 120 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,128,3;1,384,2" 2>&1 | cut -c1-200 | tail -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED"; exit 1; fi
timeout 500 python tools/sweep_variants.py --timeout 100 --only base,pptoken,ppnoprobe,emu6 --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pp --reps 10 --out gpurun_out/r02_sweep_pp7.json 2>&1 | tail -12
echo "== trace base"; FA_SM100_MODE=pp timeout 120 python tools/gpu_pp_trace.py --out gpurun_out/pp_trace5_base.json > gpurun_out/pp_trace5_base.txt 2>&1; tail -4 gpurun_out/pp_trace5_base.txt
for M in pp; do
  FA_SM100_MODE=$M timeout 300 python tools/quick_bench.py --reps 15 --check --out gpurun_out/r02_qb8_$M.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$M', r['shape'], 'mean', round(r['tflops_mean'],1), 'best', round(r['tflops_best'],1), 'maxdiff16', r.get('maxdiff_vs_sdpa16'))
"
done
FA_SM100_MODE=pp timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
FA_SM100_MODE=pp timeout 100 python tools/sustained_bench.py --seconds 3 --what fa --out gpurun_out/r02_sustained_pp11.json 2>&1 | tail -1 | cut -c1-420
FA_SM100_MODE=pp timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/r02_pp_v4 python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/r02_pp_v4_ncu.log 2>&1; tail -1 gpurun_out/r02_pp_v4_ncu.log
