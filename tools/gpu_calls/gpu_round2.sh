#!/bin/bash
# GPU round trip used late in round 1: guarded bring-up, A/B sweep, cycle trace (with the tensor-pipe
# observer) for both machine mappings, GPU tests, bench lines, ncu launch list + full capture.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-g9}
GUARD=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python tools/gpu_bringup.py --quick > gpurun_out/bringup.log 2>&1
echo "bringup rc=$?"; tail -3 gpurun_out/bringup.log | cut -c1-300
python - <<'PY'
import json, sys
log = json.load(open('gpurun_out/bringup.json'))
res = [r for r in log if r['name'] == 'RESULT'][0]
bad = [r['name'] for r in log if r['name'].startswith('shape') and not (r.get('rc') == 0 and r.get('full_maxerr', 1) < 2e-2)]
print('passed_level', res['passed_level'], 'bad', bad)
sys.exit(0 if res['passed_level'] == 4 and not bad else 1)
PY
if [ $? -ne 0 ]; then echo "GATE FAILED"; exit 1; fi
FA_SM100_LIB=$GUARD timeout 300 python tools/quick_bench.py --shapes "2,512,4;1,128,2;4,4096,32" --reps 3 --warmup 1 --check 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARDED PRODUCTION RUN FAILED"; exit 1; fi
if [ -n "$SWEEP" ]; then
  timeout 900 python tools/sweep_variants.py --shapes "${SWEEP_SHAPES:-4,4096,32;16,1024,16}" --modes "${SWEEP_MODES:-pair}" --reps 20 --out gpurun_out/sweep_${TAG}.json 2>&1 | tail -40
fi
for V in ${TRACE_VARIANTS:-prod}; do
  if [ "$V" = none ]; then continue; fi
  if [ "$V" = prod ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/flash_attention_from_scratch_b200/csrc/variants/libfa_$V.so; fi
  FA_TRACE_OUT=trace_${TAG}_${V}_pair.json timeout 200 python tools/gpu_trace.py > gpurun_out/trace_${TAG}_${V}_pair.txt 2>&1; echo "$V pair: $(grep MEDIANS gpurun_out/trace_${TAG}_${V}_pair.txt | cut -c1-450 | tr '\n' ' ')"
  FA_SM100_MODE=single FA_TRACE_OUT=trace_${TAG}_${V}_single.json timeout 200 python tools/gpu_trace.py > gpurun_out/trace_${TAG}_${V}_single.txt 2>&1; echo "$V single: $(grep MEDIANS gpurun_out/trace_${TAG}_${V}_single.txt | cut -c1-450 | tr '\n' ' ')"
done
unset FA_SM100_LIB
if [ -n "$ABLATE" ]; then
  # timing ablations of the debug instantiation (FwdDebug::level 6..9), guarded library
  for L in 5 6 7 8 9; do
    for M in pair single; do
      FA_TRACE_LEVEL=$L FA_SM100_MODE=$M FA_SM100_LIB=$GUARD FA_TRACE_OUT=trace_${TAG}_L${L}_${M}.json timeout 120 python tools/gpu_trace.py > gpurun_out/trace_${TAG}_L${L}_${M}.txt 2>&1
      echo "level $L $M rc=$? $(grep PIPE_MEDIANS gpurun_out/trace_${TAG}_L${L}_${M}.txt | cut -c1-200) $(grep '^MEDIANS' gpurun_out/trace_${TAG}_L${L}_${M}.txt | cut -c1-400)"
    done
  done
fi
if [ -n "$FULL" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_${TAG}.txt 2>&1; tail -2 gpurun_out/pytest_gpu_${TAG}.txt
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; cut -c1-300 gpurun_out/bench_ref_${TAG}.json
  timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cut -c1-1500 gpurun_out/bench_${TAG}.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python tools/benchmark/run_kernels.py --seq_len 4096 --batch 4 --n_heads 32 --n_runs 5 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
  timeout 600 python tools/quick_bench.py --comparators --reps 10 --out gpurun_out/qb_sweep_${TAG}.json 2>&1 | tail -9
  timeout 300 python tools/quick_bench.py --shapes "8,8192,16" --dtype fp16 --comparators --reps 10 --check --out gpurun_out/qb_fp16_${TAG}.json 2>&1 | tail -2
fi
