#!/usr/bin/env python
"""One op on one bench workload, a few launches, nothing else — the command that is wrapped in
`ncu --set full -k regex:<kernel> -s 2 -c 1` (development tool, GPU box).

    python tools/ncu_target.py --workload cfg2_reddit_n128_fp32 --op fwd|bwd_t|atomic|sddmm [--reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ofspmm_b200 as ofs  # noqa: E402

ops = __import__("importlib").import_module("of-spmm_b200.ops")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_reddit_n128_fp32")
    ap.add_argument("--op", default="fwd", choices=["fwd", "bwd_t", "atomic", "sddmm"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--plan", action="store_true", help="forward with a cached plan (no partition kernel in the call)")
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
    if args.op == "fwd":
        plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, n, dtype) if args.plan else None
        fn = lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan)
    elif args.op == "bwd_t":
        tr = ops.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols)
        fn = lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, transposed=tr, out=dB)
    elif args.op == "atomic":
        fn = lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, out=dB, atomic=True)
    else:
        dv = torch.empty(A.nnz, dtype=torch.float32, device=dev)
        fn = lambda: ops.sddmm_csr_compute(A.crow, A.col, dY, B, A.rows, A.cols, out=dv)
    torch.cuda.synchronize()
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        print(f"{args.workload} {args.op}: {a.elapsed_time(b):.3f} ms", flush=True)


if __name__ == "__main__":
    main()
