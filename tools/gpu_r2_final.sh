#!/bin/bash
# Round 2 validation of the tree as it ships (generation 15; production SASS identical to the r02_g15 listings, ncu
# capture and sanitizer logs): GPU suite, smoke, both bench arms, every BASELINE config, launch list, reference CLI.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02_final
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
timeout 600 python bench.py --workload sweep --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_sweep.json 2> gpurun_out/${T}_bench_sweep.err
timeout 600 python bench.py --workload dtype16k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_dtype16k_bf16.json 2> gpurun_out/${T}_bench_dtype16k_bf16.err
timeout 600 python bench.py --workload dtype16k --dtype fp16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_dtype16k_fp16.json 2> gpurun_out/${T}_bench_dtype16k_fp16.err
timeout 600 python bench.py --workload shard16k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_n1_shard16k.json 2> gpurun_out/${T}_bench_n1_shard16k.err
python - <<PY
import json
def show(name):
    try:
        r = json.load(open(f'gpurun_out/${T}_{name}.json'))
        b = r.get('back_to_back') or {}
        s = r.get('sustained') or {}
        print(name, 'value', round(r['value'],1), 'ms', round(r['ms_per_step'],4), 'frac', round(r['roofline']['frac'],3), 'b2b', round(b.get('value',0),1), 'sust', round(s.get('value',0),1), s.get('sm_mhz_median'), 'e2e', round((r.get('e2e') or {}).get('value',0),1), r['clocks'].get('reasons'), r['roofline']['kernel'][:24])
        if 'sweep' in r: print('   ', [(x['seq_len'], round(x['tflops'],1), x['kernel'][:14]) for x in r['sweep']])
    except Exception as e: print(name, 'failed', e)
for n in ('bench','bench_default','bench_sweep','bench_dtype16k_bf16','bench_dtype16k_fp16','bench_n1_shard16k'): show(n)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/${T}_launches_bench.log 2>&1; grep -c fa_fwd gpurun_out/${T}_launches.csv
KERNELS=tune timeout 600 python tools/benchmark/pt_bench.py --seq_lens 512,1024,2048,4096,8192,16384 --num_repeats 20 --comparators --csv gpurun_out/${T}_pt_bench_tune.csv > gpurun_out/${T}_pt_bench_tune.txt 2>&1; tail -32 gpurun_out/${T}_pt_bench_tune.txt | cut -c1-200
