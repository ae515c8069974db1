#!/usr/bin/env bash
# Round-2 GPU call #6 (2 GPUs): fused exchange — multi-rank parity tests, N=2 bench, single-GPU regression.
set -u
O=gpurun_out/r2c6
mkdir -p $O
echo "== multi-rank tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -30 | tee $O/pytest_multi.log
echo "== single-GPU suite"; CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -8 | tee $O/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for args in "--steps 20 --warmup 5" "--steps 10 --warmup 3 --workload cfg3_products_n256_bf16 --no-e2e" "--steps 10 --warmup 3 --workload cfg4_rmat24_n128_fp32 --no-e2e" "--steps 10 --warmup 3 --workload twin_gcn_reddit16_h256"; do
  name=$(echo $args | tr -c 'a-zA-Z0-9' '_' | cut -c1-60)
  timeout 240 $TR bench.py --gpus 2 $args > $O/n2_$name.json 2> $O/n2_$name.err
  echo "== N=2 $args rc=$?"; tail -c 1200 $O/n2_$name.json | cut -c1-1200; tail -3 $O/n2_$name.err | cut -c1-300
done
