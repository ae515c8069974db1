#!/usr/bin/env python
"""Per-op timing on one GPU (development tool): forward, A^T·dY (both routes), SDDMM, the one-off
transpose and the partitioner, for any bench workload.

    python tools/opbench.py --workload cfg4_rmat24_n128_fp32 [--reps 10]
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


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_reddit_n128_fp32")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    spec = bench.WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    s = 4 if dtype == torch.float32 else 2
    A = bench._make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    dY = ofs.graphs.upstream_grad(A.rows, n, 12, dev, dtype)
    alg = ofs.graphs.expected_alg_bytes(A.rows, A.cols, A.nnz, n, s)
    hist = ofs.row_hist(A.crow).cpu().tolist()
    print(json.dumps({"workload": args.workload, "rows": A.rows, "nnz": A.nnz, "n": n, "dtype": spec["dtype"],
                      "max_row": int(A.row_lengths().max()), "row_hist_log2": hist[:24],
                      "m2_gb": alg["m2"] / 1e9, "m1_gb": alg["m1"] / 1e9, "gflop": alg["flop"] / 1e9}), flush=True)

    def report(name, med, mn, bytes_=alg["m2"], flop=alg["flop"]):
        print(json.dumps({"op": name, "ms_med": round(med, 4), "ms_min": round(mn, 4),
                          "m2_gbs": round(bytes_ / med / 1e6, 1), "gflops": round(flop / med / 1e6, 1)}), flush=True)

    C = torch.empty((A.rows, n), dtype=dtype, device=dev)
    dB = torch.empty((A.cols, n), dtype=dtype, device=dev)
    report("fwd", *timed(lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C), args.reps))
    if "transpose" not in args.skip:
        med, mn = timed(lambda: ops.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols), 3, warm=1)
        print(json.dumps({"op": "csr_transpose (one-off)", "ms_med": round(med, 3), "ms_min": round(mn, 3)}), flush=True)
        tr = ops.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols)
        report("bwd_b transpose-route", *timed(lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, transposed=tr, out=dB), args.reps))
        del tr
    if "atomic" not in args.skip:
        report("bwd_b atomic-route", *timed(lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, out=dB, atomic=True), args.reps))
        report("bwd_b transient-route", *timed(lambda: ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, out=dB), args.reps))
    if "sddmm" not in args.skip:
        dv = torch.empty(A.nnz, dtype=torch.float32, device=dev)
        report("sddmm", *timed(lambda: ops.sddmm_csr_compute(A.crow, A.col, dY, B, A.rows, A.cols, out=dv), args.reps))
    med, mn = timed(lambda: ofs.merge_path_partition(A.crow, A.nnz, 8), args.reps)
    print(json.dumps({"op": "partition(parts=8)", "ms_med": round(med, 4)}), flush=True)
    med, mn = timed(lambda: ofs.row_hist(A.crow), args.reps)
    print(json.dumps({"op": "row_hist", "ms_med": round(med, 4)}), flush=True)


if __name__ == "__main__":
    main()
