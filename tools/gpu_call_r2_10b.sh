#!/usr/bin/env bash
# Round-2 GPU call #10b (1 GPU): the whole-row family for small problems — tests, cfg1 latency, cfg1 bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c10b
mkdir -p $O
echo "== gpu tests"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
echo "== cfg1 latency"; timeout 120 python tools/cfg1_latency.py --json $O/cfg1_latency.json 2>&1 | tail -2
timeout 100 python bench.py --workload cfg1_uniform4096_n64_fp32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "== cfg1 bench rc=$?"; cut -c1-300 $O/bench_cfg1.json
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
