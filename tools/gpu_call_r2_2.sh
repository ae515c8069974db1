#!/usr/bin/env bash
# Round-2 GPU call #2 (1 GPU): full GPU test suite on the reworked library + option sweeps.
set -u
O=gpurun_out/r2c2
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee $O/pytest_gpu.log
echo "== sweeps"
for w in cfg1_uniform4096_n64_fp32 cfg2_reddit_n128_fp32 cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32; do
  timeout 300 python tools/sweep_opts.py --workload $w 2>&1 | tee -a $O/sweep_opts.log
done
for g in "rmat:20 --n 32" "rmat:20 --n 64" "products:16 --n 64 --dtype bf16" "reddit:16 --n 32" "reddit:16 --n 64" "products:16 --n 32"; do
  timeout 200 python tools/sweep_opts.py --graph $g 2>&1 | tee -a $O/sweep_small_n.log
done
for w in cfg2_reddit_n128_fp32 cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32; do
  timeout 300 python tools/sweep_opts.py --workload $w --lib of-spmm_b200/lib_variants/bload3/libofspmm_b200.so 2>&1 | tee -a $O/sweep_bload3.log
done
timeout 200 python tools/opbench.py --workload cfg1_uniform4096_n64_fp32 --reps 9 2>&1 | tee $O/opbench_cfg1.log
