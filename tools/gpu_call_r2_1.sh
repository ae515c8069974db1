#!/usr/bin/env bash
# Round-2 GPU call #1 (1 GPU): validate what round 1 left unverified, re-measure the shipped build,
# capture ncu evidence of the shipped kernels.  Everything lands in gpurun_out/r2c1/.
set -u
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1

echo "== pending tests"; timeout 400 python -m pytest tests/pending/round2_candidates.py -q -m gpu -x 2>&1 | tail -15 | tee $O/pending.log
echo "== bench base";    timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_base.json 2> $O/bench_base.err; tail -c 1500 $O/bench_base.json
echo "== bench pipelined e2e"; timeout 300 python bench.py --steps 10 --warmup 3 --e2e-mode pipelined --no-cpu-baseline > $O/bench_pipe.json 2> $O/bench_pipe.err; tail -c 600 $O/bench_pipe.json; tail -5 $O/bench_pipe.err

echo "== variant sweep"
for w in cfg2_reddit_n128_fp32 cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32; do
  timeout 240 python tools/sweep_fwd.py --workload $w --reps 7 of-spmm_b200/lib_variants/{base,fence,unroll8_c7,unroll8_c6}/libofspmm_b200.so 2>&1 | tee -a $O/sweep.log
done

echo "== opbench"
for w in cfg1_uniform4096_n64_fp32 cfg2_reddit_n128_fp32 cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32; do
  timeout 240 python tools/opbench.py --workload $w --reps 7 2>&1 | tee -a $O/opbench.log
done

echo "== ncu launch list"
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/plain_launch.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
echo "launch list rc=$?"

cap() {  # workload op kernel-regex name
  timeout 200 python tools/ncu_target.py --workload $1 --op $2 > $O/plain_$4.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o $O/$4 \
      python tools/ncu_target.py --workload $1 --op $2 > $O/ncu_$4.log 2>&1
  echo "cap $4 rc=$?"; cat $O/plain_$4.log | tail -3
}
cap cfg2_reddit_n128_fp32 fwd spmm_merge_kernel full_cfg2_fwd
cap cfg3_products_n256_bf16 fwd spmm_merge_kernel full_cfg3_fwd
cap cfg4_rmat24_n128_fp32 fwd spmm_merge_kernel full_cfg4_fwd
cap cfg2_reddit_n128_fp32 sddmm sddmm_merge_kernel full_cfg2_sddmm
cap cfg2_reddit_n128_fp32 atomic bwd_atomic_kernel full_cfg2_atomic
ls -la $O
