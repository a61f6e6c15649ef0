#!/bin/bash
# 8-GPU box: weak scaling of the headline, BASELINE config 5 (shard16k, strong scaling) at 1/2/4/8 GPUs, e2e per N,
# the 2-device GPU test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02_final
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "second_device or concurrent" 2>&1 | tail -2
for N in 8 4 2 1; do
  timeout 400 python bench.py --gpus $N --workload shard16k --steps 10 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/${T}_bench_n${N}_shard16k.json 2> gpurun_out/${T}_bench_n${N}_shard16k.err
  python - <<PY
import json
try:
    r = json.load(open('gpurun_out/${T}_bench_n${N}_shard16k.json'))
    print('shard16k N=$N', round(r['value'],1), 'TFLOP/s', round(r['ms_per_step'],3), 'ms/step', 'e2e', round(r['e2e']['value'],1), r['clocks'].get('sm_mhz'), r['clocks'].get('reasons'), r['e2e'].get('host_numa_binding'))
except Exception as e: print('shard16k N=$N failed', e)
PY
done
for N in 8 4 2 1; do
  timeout 400 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > gpurun_out/${T}_bench_n${N}_weak.json 2> gpurun_out/${T}_bench_n${N}_weak.err
  python - <<PY
import json
try:
    r = json.load(open('gpurun_out/${T}_bench_n${N}_weak.json'))
    print('weak N=$N', round(r['value'],1), 'TFLOP/s', round(r['ms_per_step'],3), 'ms/step', 'e2e', round(r['e2e']['value'],1), r['clocks'].get('sm_mhz'), r['clocks'].get('reasons'))
except Exception as e: print('weak N=$N failed', e)
PY
done
