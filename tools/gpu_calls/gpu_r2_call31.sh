#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_final5_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_final5_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > gpurun_out/r02_final5_bench.json 2> gpurun_out/r02_final5_bench.err; python -c "
import json
r = json.load(open('gpurun_out/r02_final5_bench.json'))
print(round(r['value'],1), r['roofline']['kernel'], r['clocks'], round(r['e2e']['value'],1))"
