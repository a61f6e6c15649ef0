#!/bin/bash
# Final-tree verification on one B200: GPU tests, smoke, both bench arms, sanitizer, sweep with AUTO mapping.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-final}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_${TAG}.txt 2>&1; tail -2 gpurun_out/pytest_gpu_${TAG}.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; cut -c1-200 gpurun_out/bench_ref_${TAG}.json
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cut -c1-260 gpurun_out/bench_${TAG}.json
for M in single pair; do
  FA_SM100_MODE=$M timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 2 --n_heads 3 --n_runs 1 > gpurun_out/memcheck_${M}_${TAG}.txt 2>&1; echo "memcheck $M rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/memcheck_${M}_${TAG}.txt)"
done
timeout 600 python tools/quick_bench.py --reps 10 --out gpurun_out/qb_sweep_${TAG}.json > gpurun_out/qb_sweep_${TAG}.txt 2>&1
python - <<PY
import json
rows = json.load(open('gpurun_out/qb_sweep_${TAG}.json'))
tf = [r['tflops_mean'] for r in rows[1:]]
print([ (r['shape'][1], round(r['tflops_mean'],1)) for r in rows ], 'harmonic mean of the 6 sweep shapes', round(len(tf)/sum(1/x for x in tf),1))
PY
