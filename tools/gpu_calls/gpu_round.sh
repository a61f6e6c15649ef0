#!/bin/bash
# One GPU round trip for a new kernel generation: guarded bring-up ladder (bounded mbarrier spins, so a
# protocol bug traps instead of hanging the box), then A/B sweep of the built variants, cycle trace, tests.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
GUARD=$PWD/flash_attention_from_scratch_b200/csrc/libfa_sm100_guard.so
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
timeout 600 python tools/gpu_bringup.py --quick > gpurun_out/bringup.log 2>&1
echo "bringup rc=$?"; tail -8 gpurun_out/bringup.log | cut -c1-600
python - <<'PY'
import json, sys
log = json.load(open('gpurun_out/bringup.json'))
res = [r for r in log if r['name'] == 'RESULT'][0]
bad = [r['name'] for r in log if r['name'].startswith('shape') and not (r.get('rc') == 0 and r.get('full_maxerr', 1) < 2e-2)]
print('passed_level', res['passed_level'], 'bad', bad)
sys.exit(0 if res['passed_level'] == 4 and not bad else 1)
PY
if [ $? -ne 0 ]; then echo "GATE FAILED"; exit 1; fi
FA_SM100_LIB=$GUARD timeout 300 python tools/quick_bench.py --shapes "2,512,4;1,128,2;3,1280,5;4,4096,32" --reps 3 --warmup 1 --check 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARDED PRODUCTION RUN FAILED"; exit 1; fi
FA_SM100_MODE=pair FA_SM100_LIB=$GUARD timeout 300 python tools/quick_bench.py --shapes "1,128,2;2,200,3" --reps 2 --warmup 1 --check 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "GUARDED PRODUCTION RUN FAILED"; exit 1; fi
timeout 900 python tools/sweep_variants.py --shapes "${SWEEP_SHAPES:-4,4096,32;16,1024,16}" --modes "${SWEEP_MODES:-single,pair}" --reps 20 --out gpurun_out/sweep_${TAG:-g6}.json 2>&1 | tail -40
timeout 200 python tools/gpu_trace.py > gpurun_out/trace_${TAG:-g6}.txt 2>&1; tail -3 gpurun_out/trace_${TAG:-g6}.txt | cut -c1-900
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_${TAG:-g6}.txt 2>&1; tail -4 gpurun_out/pytest_gpu_${TAG:-g6}.txt
