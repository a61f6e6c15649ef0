timeout 300 python -m pytest tests -m gpu -x -q -k "host or launch" 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_e2e.json 2>gpurun_out/bench_e2e.err
python -c "
import json; b=json.load(open('gpurun_out/bench_e2e.json')); print(b['value'], b['e2e'])"
for M in single pair; do
FA_SM100_MODE=$M timeout 300 python tools/quick_bench.py --shapes "16,512,16;16,1024,16;16,2048,16;8,256,16;2,128,8" --reps 20 > gpurun_out/qb_mode_$M.txt 2>&1
python - <<PY
import json
for l in open('gpurun_out/qb_mode_$M.txt'):
    try: r=json.loads(l)
    except Exception: continue
    print('$M', r['shape'], round(r['tflops_mean'],1), round(r['tflops_best'],1))
PY
done
