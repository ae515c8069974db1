#!/usr/bin/env python
"""CPU context table of BASELINE.md §4 (no GPU needed): the reference-style loop (1 thread / all
cores) next to scipy's and torch's CSR·dense, on configs[0] and a Reddit-shaped twin.

    python tools/cpu_context.py
"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ofspmm_b200 as ofs  # noqa: E402
from oracle import oracle as O  # noqa: E402

warnings.filterwarnings("ignore")


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]


def main():
    cores = os.cpu_count() or 1
    cases = [("cfg1 uniform 4096^2 1%, N=64", ofs.graphs.uniform_csr(4096, 4096, 0.01, seed=1), 64),
             ("Reddit-shaped twin /32, N=128", ofs.graphs.reddit_like(32, seed=2), 128)]
    for name, A, n in cases:
        B = ofs.graphs.dense_operand(A.cols, n, 11)
        crow, col, val, Bn = A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy()
        S = A.scipy()
        T = torch.sparse_csr_tensor(A.crow.long(), A.col.long(), A.val, size=(A.rows, A.cols))
        flop = 2.0 * A.nnz * n
        rows = {
            "ref-style loop, 1 thread": best(lambda: O.spmm_f32(crow, col, val, Bn, native=True)),
            f"ref-style loop, {cores} threads": best(lambda: O.spmm_f32(crow, col, val, Bn, threads=cores, native=True)),
            "scipy csr @ dense (1 thread)": best(lambda: S @ Bn),
            f"torch sparse_csr @ dense (MKL, {torch.get_num_threads()} threads)": best(lambda: T @ B),
        }
        print(json.dumps({"case": name, "rows": A.rows, "nnz": A.nnz, "n": n,
                          "forward_gflops": {k: round(flop / v / 1e9, 2) for k, v in rows.items()}}))


if __name__ == "__main__":
    main()
