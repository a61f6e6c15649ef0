#!/bin/sh
# Writes the SASS listing + resource usage of the production kernel to profiles/ (the reference keeps
# per-iteration listings under kernel_sass/; tools/build/extract_sass.py there).
set -e
cd "$(dirname "$0")/.."
LIB=flash_attention_from_scratch_b200/csrc/libfa_sm100.so
TAG=${1:-r01}
# production instantiation: bf16, no debug hooks, seq_len % 128 == 0  (fa_fwd_kernel<true,false,false>)
cuobjdump -sass "$LIB" | awk '/Function :/{on = ($0 ~ /fa_fwd_kernelILb1ELb0ELb0E/)} on' > profiles/${TAG}_fa_fwd_kernel_bf16.sass
cuobjdump -res-usage "$LIB" 2>/dev/null | grep -A1 "fa_fwd_kernelILb1ELb0ELb0" > profiles/${TAG}_resource_usage.txt || true
{
  echo "# SASS mnemonic histogram of fa::fa_fwd_kernel<bf16, production> (sm_100a)"
  grep -E "^ +/\*[0-9a-f]{4}\*/" profiles/${TAG}_fa_fwd_kernel_bf16.sass | awk '{m=$2; if (m ~ /^@/) m=$3; print m}' | sort | uniq -c | sort -rn | head -45
} > profiles/${TAG}_sass_histogram.txt
wc -l profiles/${TAG}_fa_fwd_kernel_bf16.sass
grep -c "UTCHMMA" profiles/${TAG}_fa_fwd_kernel_bf16.sass
