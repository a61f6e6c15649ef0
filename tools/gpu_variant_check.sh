#!/bin/bash
# Validates one library variant end to end before it is trusted: guarded bring-up ladder in both machine
# mappings, A/B sweep against base, then the whole GPU test suite running ON the variant.
#   V=splits bash tools/gpu_variant_check.sh      (needs csrc/variants/libfa_$V.so and libfa_guard_$V.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${V:?variant name}
VDIR=$PWD/flash_attention_from_scratch_b200/csrc/variants
for M in pair single; do
  FA_SM100_MODE=$M FA_GUARD_LIB=$VDIR/libfa_guard_$V.so timeout 400 python tools/gpu_bringup.py --quick --out gpurun_out/bringup_${V}_$M.json > gpurun_out/bringup_${V}_$M.log 2>&1
  echo "bringup $M rc=$?"
  python - <<PY
import json, sys
log = json.load(open('gpurun_out/bringup_${V}_$M.json'))
res = [r for r in log if r['name'] == 'RESULT'][0]
bad = [(r['name'], r.get('rc'), r.get('full_maxerr'), r.get('S_maxerr')) for r in log if (r['name'].startswith('shape') or r['name'].startswith('level')) and not (r.get('rc') == 0 and (r.get('full_maxerr', 0) < 2e-2))]
print('$M passed_level', res['passed_level'], 'bad', bad)
sys.exit(0 if res['passed_level'] == 4 and not bad else 1)
PY
  if [ $? -ne 0 ]; then echo "GATE FAILED ($M)"; tail -5 gpurun_out/bringup_${V}_$M.log | cut -c1-400; exit 1; fi
done
timeout 600 python tools/sweep_variants.py --only base,$V --shapes "4,4096,32;16,1024,16;16,512,16" --modes pair,single --reps 20 --out gpurun_out/sweep_$V.json 2>&1 | tail -14
FA_SM100_LIB=$VDIR/libfa_$V.so timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_$V.txt 2>&1; tail -3 gpurun_out/pytest_gpu_$V.txt
