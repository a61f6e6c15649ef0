#!/bin/bash
# power-capped regime: does the amount of FMA-pipe exp2 emulation change FLOP per joule?  (>= 2.5 s back to back)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
for SH in 4,16384,16 4,4096,32; do
  for X in base emu0 emu2 emu6; do
    FA_SM100_LIB=$V/libfa_$X.so timeout 100 python tools/sustained_bench.py --seconds 2.5 --chunk 10 --shape $SH --what fa --out gpurun_out/r02_g15_sustained_${X}_${SH//,/_}.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('$X', r['shape'], 'region', round(r['tflops_region'],1), 'first', round(r['tflops_first_chunk'],1), 'last', round(r['tflops_last_chunks'],1), r['sm_mhz_median'], 'MHz', round(r['power_w_mean'],0), 'W', round(r['gflop_per_joule'],1), 'GF/J')
"
  done
  timeout 100 python tools/sustained_bench.py --seconds 2.5 --chunk 10 --shape $SH --what cudnn --out gpurun_out/r02_g15_sustained_cudnn_${SH//,/_}.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('cudnn', r['shape'], 'region', round(r['tflops_region'],1), 'first', round(r['tflops_first_chunk'],1), 'last', round(r['tflops_last_chunks'],1), r['sm_mhz_median'], 'MHz', round(r['power_w_mean'],0), 'W', round(r['gflop_per_joule'],1), 'GF/J')
"
done
