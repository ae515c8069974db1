#!/usr/bin/env bash
# Round-2 GPU call #7 (2 GPUs): TMA pull kernel — kernel tests, multi-rank parity, N=2 bench.
set -u
O=gpurun_out/r2c7
mkdir -p $O
echo "== kernel tests (1 GPU)"; CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_formats.py -q -m gpu -x -k "multi_peer or device or acc or fp32_accumulator or scatter" 2>&1 | tail -8 | tee $O/pytest_kernels.log
echo "== multi-rank tests"; timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15 | tee $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
for args in "--steps 20 --warmup 5 --no-e2e" "--steps 10 --warmup 3 --workload cfg3_products_n256_bf16 --no-e2e" "--steps 10 --warmup 3 --workload cfg4_rmat24_n128_fp32 --no-e2e"; do
  name=$(echo $args | tr -c 'a-zA-Z0-9' '_' | cut -c1-60)
  timeout 200 $TR bench.py --gpus 2 $args > $O/n2_$name.json 2> $O/n2_$name.err
  echo "== N=2 $args rc=$?"; tail -c 1000 $O/n2_$name.json | cut -c1-1000; tail -3 $O/n2_$name.err | cut -c1-300
done
