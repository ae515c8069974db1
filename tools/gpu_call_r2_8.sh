#!/usr/bin/env bash
set -u
O=gpurun_out/r2c8
mkdir -p $O
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "multi_peer" 2>&1 | tail -3 | tee $O/pytest_kernels.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
for w in cfg2_reddit_n128_fp32 cfg4_rmat24_n128_fp32; do
  timeout 200 $TR tools/dist_probe.py --workload $w 2>$O/probe_$w.err | tee -a $O/probe.log; tail -3 $O/probe_$w.err | cut -c1-300
done
