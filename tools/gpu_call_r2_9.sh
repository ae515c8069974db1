#!/usr/bin/env bash
# Round-2 GPU call #9 (8 GPUs): fused exchange at N=8 — parity at world 4, cfg2/3/4/5.
set -u
O=gpurun_out/r2c9
mkdir -p $O
run() {
  local name=$1 np=$2; shift 2
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err
  echo "== $name rc=$?"; tail -c 700 $O/$name.json | cut -c1-700; tail -2 $O/$name.err | cut -c1-300
}
echo "== multi-rank tests (world 4)"; timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "4" 2>&1 | tail -6 | tee $O/pytest_multi.log
run cfg2_n8_pull 8 --steps 20 --warmup 5
run cfg4_n8_pull 8 --steps 10 --warmup 3 --workload cfg4_rmat24_n128_fp32 --no-e2e
run cfg3_n8_pull 8 --steps 10 --warmup 3 --workload cfg3_products_n256_bf16 --no-e2e
run cfg5_n8      8 --steps 10 --warmup 3 --workload cfg5_gcn_reddit_h256
run cfg4_n8_pull_b2 8 --steps 10 --warmup 3 --workload cfg4_rmat24_n128_fp32 --no-e2e --buckets 2 --pull-ctas 148
run cfg2_n4_pull 4 --steps 20 --warmup 5 --no-e2e
