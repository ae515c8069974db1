#!/usr/bin/env python
"""Why is a planned dynamic-order forward sometimes slower than an unplanned one?  (development
probe, one GPU).  Times the same product (a) unplanned, (b) planned, (c) planned with an unrelated
tiny kernel in front, (d) planned after touching crow — each per-launch (sync between launches) and
back to back (no sync), dynamic and static order.

    python tools/plan_probe.py --workload cfg3_products_n256_bf16
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


def per_launch(fn, reps=9):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


def back_to_back(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3_products_n256_bf16")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    spec = bench.WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    A = bench._make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    C = torch.empty((A.rows, n), dtype=dtype, device=dev)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, n, dtype)
    tiny = torch.zeros(1024, device=dev)
    for order in ("dynamic", "static"):
        cases = {
            "unplanned": lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, variant=plan.variant, order=order),
            "planned": lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan, order=order),
            "partition kernel + planned": lambda: (ofs.merge_path_partition(A.crow, A.nnz, (A.rows + A.nnz + 255) // 256), ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan, order=order)),
            "row_hist(crow) + planned": lambda: (ops.row_hist(A.crow), ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, plan=plan, order=order)),
        }
        for name, fn in cases.items():
            med, mn, mx = per_launch(fn)
            bb = back_to_back(fn)
            print(json.dumps({"workload": args.workload, "order": order, "case": name, "per_launch_med": round(med, 4),
                              "per_launch_min": round(mn, 4), "per_launch_max": round(mx, 4), "back_to_back": round(bb, 4)}), flush=True)


if __name__ == "__main__":
    main()
