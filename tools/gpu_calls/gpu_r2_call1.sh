#!/bin/bash
# Round 2, GPU call 1: validate the P-through-smem candidates, sanitizers on what ships, power / sustained
# comparison with cuDNN and cuBLAS, wait-loop back-off variants, the reference kernel rebuilt for sm_100.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VDIR=$PWD/flash_attention_from_scratch_b200/csrc/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv | tail -1
echo "=== 1. psmem bring-up (pair)"
FA_SM100_MODE=pair FA_GUARD_LIB=$VDIR/libfa_guard_psmem.so timeout 300 python tools/gpu_bringup.py --quick --out gpurun_out/bringup_psmem_pair.json > gpurun_out/bringup_psmem_pair.log 2>&1
echo "bringup rc=$?"; tail -3 gpurun_out/bringup_psmem_pair.log | cut -c1-300
echo "=== 2. sweep"
timeout 900 python tools/sweep_variants.py --only base,psmem,psmem2,psmemr,hint,sleep32 --shapes "4,4096,32;16,1024,16;4,16384,16" --modes pair --reps 15 --out gpurun_out/r02_sweep1.json 2>&1 | tail -24
echo "=== 3. sustained"
timeout 300 python tools/sustained_bench.py --seconds 3 --what fa,cudnn,gemm --out gpurun_out/r02_sustained_base.json 2>&1 | tail -4
for V in hint sleep32 psmem; do
  FA_SM100_LIB=$VDIR/libfa_$V.so timeout 200 python tools/sustained_bench.py --seconds 3 --what fa --out gpurun_out/r02_sustained_$V.json 2>&1 | tail -1
done
timeout 300 python tools/sustained_bench.py --seconds 3 --shape 4,16384,16 --what fa,cudnn --out gpurun_out/r02_sustained_16k.json 2>&1 | tail -3
echo "=== 4. reference kernel 16 rebuilt for sm_100"
if [ -f oracle/_ref/flash_attention_kernels.so ]; then
  timeout 600 python tools/ref_kernel_bench.py --out gpurun_out/r02_ref_kernel.json > gpurun_out/r02_ref_kernel.txt 2>&1; tail -4 gpurun_out/r02_ref_kernel.txt | cut -c1-400
else echo "no oracle/_ref"; fi
echo "=== 5. sanitizers (production library)"
for M in single pair; do
  for TOOL in memcheck racecheck synccheck; do
    FA_SM100_MODE=$M timeout 400 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 640 --batch 1 --n_heads 2 --n_runs 1 > gpurun_out/r02_${TOOL}_${M}.txt 2>&1
    echo "$TOOL $M rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_${TOOL}_${M}.txt | tail -1)"
  done
done
FA_SM100_MODE=pair timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/benchmark/run_kernels.py --seq_len 2048 --batch 1 --n_heads 1 --n_runs 1 > gpurun_out/r02_racecheck_pair2048.txt 2>&1
echo "racecheck pair 2048 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_racecheck_pair2048.txt | tail -1)"
echo "=== 6. GPU suite on psmem (only if bring-up passed)"
if grep -q '"passed_level": 4' gpurun_out/bringup_psmem_pair.json 2>/dev/null; then
  FA_SM100_LIB=$VDIR/libfa_psmem.so timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_psmem.txt 2>&1; tail -3 gpurun_out/pytest_gpu_psmem.txt
fi
