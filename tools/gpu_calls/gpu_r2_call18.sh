#!/bin/bash
# generation 15: which kernel where (AUTO threshold), tuning variants, cuDNN on the same box, ncu capture, bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g15
timeout 500 python tools/sweep_variants.py --timeout 150 --only base --shapes "16,512,16;16,1024,16;16,2048,16;16,4096,16;8,8192,16;4,16384,16;4,4096,32" --modes pair,pp,single --reps 12 --out gpurun_out/${T}_sweep_modes.json 2>&1 | tail -22
timeout 500 python tools/sweep_variants.py --timeout 100 --only base,kvpf,r200,emu2,emu6 --shapes "4,4096,32;16,2048,16;8,8192,16" --modes pair --reps 15 --out gpurun_out/${T}_sweep_tune.json 2>&1 | tail -16
timeout 300 python tools/quick_bench.py --reps 15 --comparators --shapes "4,4096,32;16,4096,16;8,8192,16;4,16384,16" --out gpurun_out/${T}_qb_comparators.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['shape'], 'ours', round(r['tflops_mean'],1), 'cudnn', round(r.get('sdpa_cudnn_tflops',0),1), 'flash2', round(r.get('flash_attn2_tflops',0),1))
"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/${T}_headline python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/${T}_headline_ncu.log 2>&1; tail -1 gpurun_out/${T}_headline_ncu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-2500 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
