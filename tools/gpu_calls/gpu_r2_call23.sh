#!/bin/bash
# cycle trace of the PRODUCTION pair kernel (FA_TRACE builds) with and without the S-before-PV issue order
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
for X in trace trace_noqkf; do
  FA_TRACE_LEVEL=40 FA_SM100_MODE=pair FA_SM100_LIB=$V/libfa_$X.so FA_TRACE_OUT=r02_g16_prod_$X.json timeout 120 python tools/gpu_trace2.py > gpurun_out/r02_g16_prod_$X.txt 2>&1
  echo "== $X"; grep MEDIANS gpurun_out/r02_g16_prod_$X.txt | cut -c1-600
done
FA_TRACE_LEVEL=40 FA_SM100_MODE=single FA_SM100_LIB=$V/libfa_trace_noqkf.so FA_TRACE_OUT=r02_g16_prod_single.json timeout 120 python tools/gpu_trace2.py > gpurun_out/r02_g16_prod_single.txt 2>&1
echo "== single noqkf"; grep MEDIANS gpurun_out/r02_g16_prod_single.txt | cut -c1-600
