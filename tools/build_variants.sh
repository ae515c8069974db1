#!/usr/bin/env bash
# Builds the tuning variants of libofspmm_b200.so that are waiting for GPU time (see ROUND_NOTES.md)
# into of-spmm_b200/lib_variants/<name>/ and prints the gpurun command that compares them.
# Every variant is a -D switch over the same sources; the default build is untouched.
set -euo pipefail
cd "$(dirname "$0")/../of-spmm_b200/csrc"
variants=(
  "base:"
  "fence:-DOFSPMM_PROXY_FENCE"
  "unroll8_c7:-DOFSPMM_CHUNKS_PER_ITER=2 -DOFSPMM_MIN_CTAS=7"
  "unroll8_c6:-DOFSPMM_CHUNKS_PER_ITER=2 -DOFSPMM_MIN_CTAS=6"
  "c8:-DOFSPMM_MIN_CTAS=8"
  "c10:-DOFSPMM_MIN_CTAS=10"
)
for v in "${variants[@]}"; do
  name=${v%%:*}; extra=${v#*:}
  make -j8 VARIANT="$name" EXTRA="$extra" > "/tmp/mk_$name.log" 2>&1 &
done
wait
ls ../lib_variants/
cat <<'MSG'

Run (1 GPU, ~1 min):
  gpurun --timeout 600 -- 'for w in cfg2_reddit_n128_fp32 cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32; do
      timeout 180 python tools/sweep_fwd.py --workload $w --reps 5 of-spmm_b200/lib_variants/*/libofspmm_b200.so; done'
Remove of-spmm_b200/lib_variants afterwards (each .so is ~35 MB of snapshot).
MSG
