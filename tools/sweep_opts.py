#!/usr/bin/env python
"""Forward-kernel option sweep on one GPU (development tool): task order x kernel family x plan,
same inputs, bitwise cross-check, one JSON line per configuration.

    python tools/sweep_opts.py --workload cfg2_reddit_n128_fp32 [--n 64] [--reps 7] [--lib path.so]
    python tools/sweep_opts.py --graph rmat:20 --n 32 --dtype fp32
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
_lib = ofs._lib
X = _lib.VARIANT_EXPLICIT


def timed(fn, reps):
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
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=None)
    ap.add_argument("--graph", default=None, help="rmat:<scale> | reddit:<div> | products:<div> | uniform")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--dtype", default=None, choices=[None, "fp32", "bf16"])
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--lib", default=None, help="alternative build of the library to load")
    args = ap.parse_args()
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)
    dev = torch.device("cuda:0")
    if args.workload:
        spec = dict(bench.WORKLOADS[args.workload])
    else:
        kind, _, arg = args.graph.partition(":")
        spec = dict(kind=kind, n=128, dtype="fp32")
        if kind == "rmat":
            spec["scale"] = int(arg or 20)
        else:
            spec["scale_div"] = int(arg or 1)
    if args.n:
        spec["n"] = args.n
    if args.dtype:
        spec["dtype"] = args.dtype
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    A = bench._make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    C = torch.empty((A.rows, n), dtype=dtype, device=dev)
    alg = ofs.graphs.expected_alg_bytes(A.rows, A.cols, A.nnz, n, 4 if dtype == torch.float32 else 2)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, n, dtype)
    tag = args.workload or f"{args.graph}/n{n}/{spec['dtype']}"
    print(json.dumps({"workload": tag, "rows": A.rows, "nnz": A.nnz, "n": n, "lib": args.lib or "default",
                      "hist_variant": plan.variant_name(), "auto_variant":
                      _lib.lib().ofspmm_fwd_variant(A.rows, A.nnz, n, 2 if dtype == torch.float32 else 11).decode()}), flush=True)
    ref = {}
    fams = [("base", X), ("rowpar", X | _lib.VARIANT_ROWPAR),
            ("items64", X | _lib.VARIANT_ITEMS64)]
    for fam, v in fams:
        name = _lib.lib().ofspmm_variant_name(v, A.rows, A.nnz, n, 2 if dtype == torch.float32 else 11).decode()
        if fam != "base" and name == _lib.lib().ofspmm_variant_name(X, A.rows, A.nnz, n, 2 if dtype == torch.float32 else 11).decode():
            continue   # family has no kernel for this width: would just repeat the base family
        if fam == "items64" and A.nnz > 40_000_000:
            continue   # 4x the carry workspace for nothing on big graphs
        p_v = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, n, dtype, variant=v)
        for order in ("static", "dynamic"):
            for use_plan in (False, True):
                fn = lambda: ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C, variant=v, order=order,
                                                  plan=p_v if use_plan else None)
                med, mn = timed(fn, args.reps)
                same = None
                if fam in ref:
                    same = bool(torch.equal(ref[fam], C))
                else:
                    ref[fam] = C.clone()
                print(json.dumps({"family": fam, "order": order, "plan": use_plan, "ms_med": round(med, 4),
                                  "ms_min": round(mn, 4), "m2_gbs": round(alg["m2"] / med / 1e6, 1),
                                  "gflops": round(alg["flop"] / med / 1e6, 1), "bitwise_same_in_family": same}), flush=True)


if __name__ == "__main__":
    main()
