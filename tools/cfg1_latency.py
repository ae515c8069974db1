#!/usr/bin/env python
"""Launch latency of the smallest configuration (cfg1: 4096 x 4096, 1 % dense, n = 64) — the
reference's own CPU-runnable case, where the product is launch bound, not bandwidth bound.

Three numbers per op, one GPU:
  eager      python -> ofspmm_* call per launch, back to back (what an eager framework pays)
  prepared   the plan's pre-bound launcher, back to back (no argument marshalling per call)
  graph      K launches captured in ONE CUDA graph and replayed: the device-side cost of a launch,
             which is what the op costs inside OneFlow's graph / stream-ordered executor

    python tools/cfg1_latency.py [--workload cfg1_uniform4096_n64_fp32] [--k 64] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ofspmm_b200 as ofs  # noqa: E402

ops = __import__("importlib").import_module("of-spmm_b200.ops")


def b2b(fn, reps):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def graph_us(fn, k, replays=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (replays * k) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1_uniform4096_n64_fp32")
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    spec = bench.WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    A = bench._make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    dY = ofs.graphs.upstream_grad(A.rows, n, 12, dev, dtype)
    C = torch.empty((A.rows, n), dtype=dtype, device=dev)
    dB = torch.empty((A.cols, n), dtype=dtype, device=dev)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, n, dtype, transpose=True)
    res = {"workload": args.workload, "rows": A.rows, "cols": A.cols, "nnz": A.nnz, "n": n,
           "variant": plan.variant_name(), "variant_T": plan.variant_name(True), "unit": "us per launch", "graph_k": args.k}
    fwd_plain = lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C)
    fwd_plan = lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan)
    bwd_plan = lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, out=dB, plan=plan)
    launcher = plan.prepared()
    prep = lambda: launcher(A.val, B, C)
    res["fwd_eager_unplanned"] = b2b(fwd_plain, 200)
    res["fwd_eager_planned"] = b2b(fwd_plan, 200)
    res["fwd_prepared"] = b2b(prep, 200)
    res["fwd_graph"] = graph_us(prep, args.k)
    res["bwd_eager_planned"] = b2b(bwd_plan, 200)
    res["bwd_graph"] = graph_us(bwd_plan, args.k)
    res["step_graph"] = graph_us(lambda: (fwd_plan(), bwd_plan()), args.k // 2)
    # the static-order launch has no counter memset in front of it
    fwd_static = lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan, order="static")
    res["fwd_graph_static_order"] = graph_us(fwd_static, args.k)
    print(json.dumps(res))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
