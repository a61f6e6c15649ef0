#!/bin/bash
# last check of the tree as committed: GPU suite (with the seq_len-1280 golden fixture), smoke, bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=r02_final3
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
r = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(r['value'],1), 'frac', round(r['roofline']['frac'],3), 'b2b', round(r['back_to_back']['value'],1), 'sust', round(r['sustained']['value'],1), 'host_us', round(r['host_us_per_call'],1), r['tensor_map_cache'], 'e2e', round(r['e2e']['value'],1), r['clocks'])
PY
