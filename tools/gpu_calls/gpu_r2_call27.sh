#!/bin/bash
# the s_free arrive without the NaN-payload skew: guarded run, A/B against the generation-15 library, GPU suite, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_final2
G=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
for M in pair single; do
  FA_SM100_MODE=$M FA_SM100_LIB=$G timeout 200 python tools/quick_bench.py --reps 2 --warmup 1 --check --shapes "4,4096,32;16,512,16;3,640,5;2,2304,3;9,128,41" 2>&1 | cut -c1-60,230-330 | tail -5
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARD RUN FAILED ($M)"; exit 1; fi
done
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,g15 --shapes "4,4096,32;16,2048,16;8,8192,16" --modes pair --reps 20 --out gpurun_out/${T}_sweep.json 2>&1 | tail -7
timeout 300 python tools/sweep_variants.py --timeout 100 --only base,g15 --shapes "4,4096,32;16,1024,16" --modes single --reps 15 --out gpurun_out/${T}_sweep_single.json 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-330 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
r = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(r['value'],1), 'frac', round(r['roofline']['frac'],3), 'b2b', round(r['back_to_back']['value'],1), 'sust', round(r['sustained']['value'],1), 'host_us', round(r['host_us_per_call'],1), r['tensor_map_cache'], 'e2e', round(r['e2e']['value'],1))
PY
for TOOL in racecheck synccheck memcheck; do
  FA_SM100_MODE=pair timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 3 --n_runs 1 > gpurun_out/${T}_${TOOL}_pair.txt 2>&1
  echo "$TOOL pair 640 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${TOOL}_pair.txt | tail -1)"
done
