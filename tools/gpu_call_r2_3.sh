#!/usr/bin/env bash
# Round-2 GPU call #3 (1 GPU): full GPU suite on the new defaults, plan probe, bench at N=1.
set -u
O=gpurun_out/r2c3
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee $O/pytest_gpu.log
echo "== plan probe"
for w in cfg2_reddit_n128_fp32 cfg3_products_n256_bf16; do timeout 300 python tools/plan_probe.py --workload $w 2>&1 | tee -a $O/plan_probe.log; done
echo "== chooser"
for g in "rmat:20 --n 32" "rmat:20 --n 64" "products:16 --n 64 --dtype bf16"; do
  timeout 200 python tools/sweep_opts.py --graph $g 2>&1 | grep -E "workload|dynamic" | tee -a $O/sweep_small_n.log
done
echo "== bench"
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 3000 $O/bench_cfg2.json; tail -5 $O/bench_cfg2.err
for w in cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32 cfg1_uniform4096_n64_fp32; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 10 > $O/bench_$w.json 2> $O/bench_$w.err; tail -c 1500 $O/bench_$w.json; tail -5 $O/bench_$w.err
done
