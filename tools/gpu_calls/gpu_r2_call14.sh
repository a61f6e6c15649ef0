#!/bin/bash
# Round 2 validation of the tree as it ships: GPU suite, smoke, both bench arms, sweep workload, sanitizers on the
# three kernels, launch list + full ncu capture of the headline kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_g14
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-1500 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload sweep --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_sweep.json 2> gpurun_out/${T}_bench_sweep.err; python - <<PY
import json
try:
    r = json.load(open('gpurun_out/${T}_bench_sweep.json'))
    print('sweep value (harmonic mean)', round(r['value'],1), [(x['seq_len'], round(x['tflops'],1), x['kernel'][:22]) for x in r['sweep']], 'sustained', r.get('sustained',{}).get('value'))
except Exception as e: print('sweep failed', e)
PY
for M in single pair pingpong; do
  for TOOL in memcheck racecheck synccheck; do
    FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_${M}.txt 2>&1
    echo "$TOOL $M rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_${M}.txt | tail -1)"
  done
done
for M in pair pingpong; do
  FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 2304 --batch 1 --n_heads 2 --n_runs 1 > gpurun_out/${T}_racecheck_${M}_2304.txt 2>&1
  echo "racecheck $M 2304 rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${T}_racecheck_${M}_2304.txt | tail -1)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/${T}_launches_bench.log 2>&1; grep -c fa_fwd gpurun_out/${T}_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/${T}_headline python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 4 > gpurun_out/${T}_headline_ncu.log 2>&1; tail -1 gpurun_out/${T}_headline_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 2 -c 1 -f -o gpurun_out/${T}_pp1024 python tools/benchmark/run_kernels.py --seq_len 1024 --batch 16 --n_heads 16 --n_runs 4 > gpurun_out/${T}_pp1024_ncu.log 2>&1; tail -1 gpurun_out/${T}_pp1024_ncu.log
timeout 300 python tools/quick_bench.py --reps 15 --comparators --out gpurun_out/${T}_qb_comparators.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['shape'], 'ours', round(r['tflops_mean'],1), 'cudnn', round(r.get('sdpa_cudnn_tflops',0),1), 'flash2', round(r.get('flash_attn2_tflops',0),1))
"
