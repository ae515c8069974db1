#!/usr/bin/env bash
# Round-2 GPU call #4 (2 GPUs): multi-rank parity tests on real kernels + peer memory, N=2 bench, plan probe.
set -u
O=gpurun_out/r2c4
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== multi-rank tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -30 | tee $O/pytest_multi.log
echo "== quick single-GPU regression"; timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -x -k "variants or plan or epilogue or sddmm or reddit or golden" 2>&1 | tail -5 | tee $O/pytest_quick.log
echo "== plan probe"
for w in cfg2_reddit_n128_fp32 cfg3_products_n256_bf16; do CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/plan_probe.py --workload $w 2>&1 | tee -a $O/plan_probe.log; done
echo "== bench N=2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_cfg2_n2_pull.json 2> $O/bench_cfg2_n2_pull.err; tail -c 2500 $O/bench_cfg2_n2_pull.json; tail -3 $O/bench_cfg2_n2_pull.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scheme allgather --no-e2e > $O/bench_cfg2_n2_ag.json 2> $O/bench_cfg2_n2_ag.err; tail -c 1500 $O/bench_cfg2_n2_ag.json; tail -3 $O/bench_cfg2_n2_ag.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload twin_gcn_reddit16_h256 > $O/bench_gcn_twin_n2.json 2> $O/bench_gcn_twin_n2.err; tail -c 1500 $O/bench_gcn_twin_n2.json; tail -3 $O/bench_gcn_twin_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2_n1.json 2> $O/bench_cfg2_n1.err; tail -c 2500 $O/bench_cfg2_n1.json; tail -3 $O/bench_cfg2_n1.err
