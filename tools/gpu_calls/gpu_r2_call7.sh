#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/flash_attention_from_scratch_b200/csrc/variants
echo "== token ahead4 (production)"; FA_SM100_MODE=pp timeout 120 python tools/gpu_pp_trace.py --out gpurun_out/pp_trace_tok4.json > gpurun_out/pp_trace_tok4.txt 2>&1; tail -4 gpurun_out/pp_trace_tok4.txt
for X in ppnotoken3 ppnotoken ppahead3; do echo "== $X"; FA_SM100_LIB=$V/libfa_$X.so FA_SM100_MODE=pp timeout 120 python tools/gpu_pp_trace.py --out gpurun_out/pp_trace_$X.json > gpurun_out/pp_trace_$X.txt 2>&1; tail -4 gpurun_out/pp_trace_$X.txt; done
