#!/bin/bash
# cycle trace of the pair kernel with and without the S-before-PV issue order
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
for X in base noqkf; do
  FA_SM100_MODE=pair FA_SM100_LIB=$V/libfa_$X.so FA_TRACE_OUT=r02_g16_trace_$X.json timeout 120 python tools/gpu_trace2.py > gpurun_out/r02_g16_trace_$X.txt 2>&1
  echo "== $X"; grep MEDIANS gpurun_out/r02_g16_trace_$X.txt | cut -c1-600
done
