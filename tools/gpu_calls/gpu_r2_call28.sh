#!/bin/bash
# the chain's tcgen05.ld leg: S fetched with one wait (arrive right behind it, row max afterwards) vs the split fetch
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/sweep_variants.py --timeout 100 --only base,noldsplit --shapes "4,4096,32;16,2048,16;8,8192,16;16,4096,16" --modes pair --reps 20 --out gpurun_out/r02_g15_sweep_ldsplit.json 2>&1 | tail -9
timeout 300 python tools/sweep_variants.py --timeout 100 --only base,noldsplit --shapes "4,4096,32;16,1024,16" --modes single,pp --reps 15 --out gpurun_out/r02_g15_sweep_ldsplit_single.json 2>&1 | tail -9
