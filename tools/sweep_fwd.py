#!/usr/bin/env python
"""Kernel-tuning sweep (development tool, GPU box): times ofspmm_fwd of several builds of the
library on one workload, same inputs, and checks that every build returns the same bits.

    python tools/sweep_fwd.py --workload cfg2_reddit_n128_fp32 lib_a.so lib_b.so ...
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ofspmm_b200 as ofs  # noqa: E402

_lib = ofs._lib


def bind(path):
    L = ctypes.CDLL(path)
    i64, i32, vp, sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    L.ofspmm_fwd_workspace_bytes.argtypes = [i64, i64, i64, i64, i32]
    L.ofspmm_fwd_workspace_bytes.restype = sz
    L.ofspmm_fwd.argtypes = [ctypes.POINTER(_lib.CsrStruct), vp, vp, i64, i32, vp, sz, vp]
    L.ofspmm_fwd.restype = i32
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_reddit_n128_fp32")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--transpose", action="store_true", help="time the forward kernel on A^T (the bwd route)")
    ap.add_argument("libs", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    spec = bench.WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    dcode = 2 if dtype == torch.float32 else 11
    A = bench._make_graph(spec, dev)
    if args.transpose:
        tc, tcol, tv = ofs.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols)
        A = ofs.graphs.CsrMatrix(tc, tcol, tv, A.cols, A.rows)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    alg = ofs.graphs.expected_alg_bytes(A.rows, A.cols, A.nnz, n, 4 if dtype == torch.float32 else 2)
    ref = None
    stream = torch.cuda.current_stream().cuda_stream
    for path in args.libs:
        L = bind(path)
        cs = _lib.CsrStruct(A.rows, A.cols, A.nnz, A.crow.data_ptr(), A.col.data_ptr(), A.val.data_ptr(), 5, 2)
        nbytes = L.ofspmm_fwd_workspace_bytes(A.rows, A.cols, A.nnz, n, dcode)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        C = torch.empty((A.rows, n), dtype=dtype, device=dev)

        def run():
            rc = L.ofspmm_fwd(ctypes.byref(cs), B.data_ptr(), C.data_ptr(), n, dcode, ws.data_ptr(), nbytes, stream)
            assert rc == 0, rc
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        if ref is None:
            ref = C.clone()
            same = True
        else:
            same = bool(torch.equal(ref, C))
        med = ts[len(ts) // 2]
        print(json.dumps({"lib": os.path.relpath(path, ROOT), "ms_med": round(med, 4), "ms_min": round(ts[0], 4),
                          "m2_gbs": round(alg["m2"] / med / 1e6, 1), "gflops": round(alg["flop"] / med / 1e6, 1),
                          "bitwise_same_as_first": same}), flush=True)


if __name__ == "__main__":
    main()
