#!/usr/bin/env bash
# Round-2 GPU call #11 (8 GPUs, the last one): interleaved step, combine overlap and block-cyclic
# ownership.  Phase A uses disjoint GPU sets side by side (one rank per GPU throughout): the
# world-4 parity tests on GPUs 4-7 while cfg2 runs at N=4 and N=2 on GPUs 0-3.  Phase B: all 8.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c11
mkdir -p $O
run() {   # name nproc port bench-args...
  local name=$1 np=$2 port=$3; shift 3
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err
  echo "== $name rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    x = d.get("impl_detail", {})
    print("ms/step", round(d["ms_per_step"], 4), "fwd_ms", round(d.get("fwd_ms") or 0, 4), "value", round(d["value"], 1), "verified", d.get("verified"),
          "layout", x.get("shard_layout"), "order", x.get("step_order"), "e2e_ms", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("no json:", e)
PY
)"
  tail -2 $O/$name.err | cut -c1-300
}
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 420 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "spmm and 4" > $O/pytest_multi.log 2>&1; echo "== multi-rank tests (world 4) rc=$?"; tail -4 $O/pytest_multi.log ) &
TESTS=$!
CUDA_VISIBLE_DEVICES=0,1,2,3 run cfg2_n4 4 29581 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline
CUDA_VISIBLE_DEVICES=0,1 run cfg2_n2 2 29582 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline
CUDA_VISIBLE_DEVICES=0,1,2,3 run cfg4_n4 4 29583 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload cfg4_rmat24_n128_fp32
wait $TESTS
B="--no-e2e --no-cpu-baseline"
run cfg2_n8 8 29584 --steps 20 --warmup 5 --no-cpu-baseline
run cfg4_n8 8 29586 --steps 10 --warmup 3 $B --workload cfg4_rmat24_n128_fp32
run cfg2_n8_b2 8 29591 --steps 20 --warmup 5 $B --buckets 2
run cfg3_n8 8 29588 --steps 10 --warmup 3 $B --workload cfg3_products_n256_bf16
run cfg5_n8 8 29589 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5_gcn_reddit_h256
run cfg2_n8_comb 8 29585 --steps 20 --warmup 5 $B --combine-ctas 148
run cfg2_n8_ag 8 29592 --steps 20 --warmup 5 $B --scheme allgather
run cfg4_n8_block 8 29593 --steps 10 --warmup 3 $B --workload cfg4_rmat24_n128_fp32 --layout block
run cfg3_n8_comb 8 29590 --steps 10 --warmup 3 $B --workload cfg3_products_n256_bf16 --combine-ctas 148
run cfg4_n8_comb 8 29587 --steps 10 --warmup 3 $B --workload cfg4_rmat24_n128_fp32 --combine-ctas 148
