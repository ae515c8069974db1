#!/usr/bin/env bash
# Round-2 GPU call #10 (1 GPU): the final build — full GPU test suite, smoke, default bench + reference arm,
# ncu --set full of the dominant kernels (each only after the same command exited 0 without ncu),
# the launch list of a bench step, cfg1 launch latency.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c10
mkdir -p $O
echo "== gpu tests"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/smoke.log
echo "== bench default"; timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo rc=$?; cut -c1-1500 $O/bench_default.json
echo "== bench reference arm"; timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo rc=$?; cut -c1-600 $O/bench_reference.json
echo "== cfg1 latency"; timeout 120 python tools/cfg1_latency.py --json $O/cfg1_latency.json 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
cap() {  # name kernel-regex ncu_target args...
  local name=$1 rx=$2; shift 2
  timeout 150 python tools/ncu_target.py "$@" > $O/$name.plain.log 2>&1 && \
  timeout 240 $NCU -k regex:$rx -o $O/full_$name python tools/ncu_target.py "$@" > $O/$name.ncu.log 2>&1
  echo "== ncu $name rc=$? $(tail -1 $O/$name.plain.log)"
  ncu -i $O/full_$name.ncu-rep --page raw --csv > $O/full_$name.raw.csv 2>/dev/null
  ncu -i $O/full_$name.ncu-rep --page source --csv > $O/full_$name.source.csv 2>/dev/null
  case $name in cfg2_fwd|cfg4_fwd) ;; *) rm -f $O/full_$name.ncu-rep ;; esac   # gpurun_out comes back capped at 64 MiB
}
cap cfg2_fwd spmm_merge --workload cfg2_reddit_n128_fp32 --op fwd --plan
cap cfg2_bwd_t spmm_merge --workload cfg2_reddit_n128_fp32 --op bwd_t
cap cfg3_fwd spmm_merge --workload cfg3_products_n256_bf16 --op fwd --plan
cap cfg4_fwd spmm_merge --workload cfg4_rmat24_n128_fp32 --op fwd --plan
cap cfg2_sddmm sddmm --workload cfg2_reddit_n128_fp32 --op sddmm
echo "== launch list"
BA="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 python bench.py $BA > $O/bench_short.json 2> $O/bench_short.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_|sddmm_|partition|gather_vals|transpose_' -c 400 --csv \
   --log-file $O/launches_bench_steps3.csv python bench.py $BA > $O/bench_short_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $O/launches_bench_steps3.csv 2>/dev/null)"
for w in cfg3_products_n256_bf16 cfg4_rmat24_n128_fp32 cfg1_uniform4096_n64_fp32; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "== $w rc=$?"; cut -c1-400 $O/bench_$w.json
done
ls -la $O | head -40
