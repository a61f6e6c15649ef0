#!/bin/sh
# Writes the SASS listings + resource usage of the production kernels to profiles/ (the reference keeps
# per-iteration listings under kernel_sass/; tools/build/extract_sass.py there).
set -e
cd "$(dirname "$0")/.."
LIB=flash_attention_from_scratch_b200/csrc/libfa_sm100.so
TAG=${1:-r01}
# production instantiations: bf16, no debug hooks, seq_len % 128 == 0
#   fa_fwd_kernel_pair<true,false,false>  (CTA pairs, AUTO for seq_len > 1024)      -> *_pair_bf16.sass
#   pp::fa_fwd_kernel_pp<true,false,false> (ping-pong, AUTO for seq_len <= 1024)   -> *_pp_bf16.sass
#   fa_fwd_kernel<true,false,false>       (single CTA, explicit mode only)         -> *_bf16.sass
cuobjdump -sass "$LIB" | awk '/Function :/{on = ($0 ~ /fa_fwd_kernel_pairILb1ELb0ELb0E/)} on' > profiles/${TAG}_fa_fwd_kernel_pair_bf16.sass
cuobjdump -sass "$LIB" | awk '/Function :/{on = ($0 ~ /fa_fwd_kernelILb1ELb0ELb0E/)} on' > profiles/${TAG}_fa_fwd_kernel_bf16.sass
cuobjdump -sass "$LIB" | awk '/Function :/{on = ($0 ~ /fa_fwd_kernel_ppILb1ELb0ELb0E/)} on' > profiles/${TAG}_fa_fwd_kernel_pp_bf16.sass
cuobjdump -res-usage "$LIB" 2>/dev/null | grep -A1 "fa_fwd_kernel\(_pair\|_pp\)\?ILb1ELb0ELb0" > profiles/${TAG}_resource_usage.txt || true
for K in fa_fwd_kernel_pair fa_fwd_kernel_pp fa_fwd_kernel; do
  {
    echo "# SASS mnemonic histogram of fa::${K}<bf16, production> (sm_100a)"
    grep -E "^ +/\*[0-9a-f]{4}\*/" profiles/${TAG}_${K}_bf16.sass | awk '{m=$2; if (m ~ /^@/) m=$3; print m}' | sort | uniq -c | sort -rn | head -45
  } > profiles/${TAG}_${K}_sass_histogram.txt
  wc -l profiles/${TAG}_${K}_bf16.sass
  grep -c "UTCHMMA" profiles/${TAG}_${K}_bf16.sass
done
