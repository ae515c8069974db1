"""Generates tests/golden/*.npz — small input/output vectors for the SpMM path.

The mounted reference has no SpMM op, test or golden vector (parity
unpinned), and OneFlow cannot be imported here, so the expected outputs come from two
*independent* implementations available in this container:
  * scipy.sparse.csr_matrix (float64 inputs → float64 result) — stored as the expected values;
  * torch.sparse_csr_tensor @ dense (MKL, fp32) — cross-checked against scipy at generation time
    and stored for the record.
Neither touches oracle/ or the CUDA library.  Run:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def random_csr(rng, M, K, density, empty_frac=0.0, hub=None):
    A = sp.random(M, K, density=density, format="csr", random_state=rng, dtype=np.float64,
                  data_rvs=lambda s: rng.uniform(-1, 1, s))
    A = A.tolil()
    if empty_frac > 0:
        for r in rng.choice(M, int(M * empty_frac), replace=False):
            A.rows[r], A.data[r] = [], []
    if hub is not None:  # one hub row touching `hub` columns
        cols = np.sort(rng.choice(K, hub, replace=False))
        A.rows[M // 3] = list(cols)
        A.data[M // 3] = list(rng.uniform(-1, 1, hub))
    A = A.tocsr()
    A.sort_indices()
    A.data = A.data.astype(np.float32)
    return A


def emit(name, A, N, rng):
    M, K = A.shape
    B = rng.standard_normal((K, N)).astype(np.float32)
    dY = rng.uniform(0, 1, (M, N)).astype(np.float32)
    A64 = A.astype(np.float64)
    C = A64 @ B.astype(np.float64)
    dB = A64.T @ dY.astype(np.float64)
    rows = np.repeat(np.arange(M), np.diff(A.indptr))
    dval = np.einsum("ij,ij->i", dY[rows].astype(np.float64), B[A.indices].astype(np.float64))
    # independent cross-check with torch's CSR kernel
    At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)),
                                 torch.from_numpy(A.indices.astype(np.int64)),
                                 torch.from_numpy(A.data), size=(M, K))
    C_torch = (At @ torch.from_numpy(B)).numpy()
    scale = np.abs(C).max() + 1e-30
    assert np.abs(C_torch - C).max() <= 1e-5 * scale * max(1, np.diff(A.indptr).max()), name
    AT = A.T.tocsr()
    AT.sort_indices()
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        crow=A.indptr.astype(np.int32), col=A.indices.astype(np.int32), val=A.data,
                        rows=M, cols=K, B=B, dY=dY, C=C, C_torch=C_torch, dB=dB, dval=dval,
                        t_crow=AT.indptr.astype(np.int32), t_col=AT.indices.astype(np.int32),
                        t_val=AT.data.astype(np.float32))
    print(name, "M", M, "K", K, "nnz", A.nnz, "N", N)


def main():
    import warnings
    warnings.filterwarnings("ignore")
    rng = np.random.default_rng(20261018)
    emit("uniform_96x80_n64", random_csr(rng, 96, 80, 0.08), 64, rng)
    emit("ragged_empty_200x150_n128", random_csr(rng, 200, 150, 0.05, empty_frac=0.4), 128, rng)
    emit("hub_300x1500_n32", random_csr(rng, 300, 1500, 0.004, empty_frac=0.2, hub=1200), 32, rng)
    emit("tall_600x64_n20", random_csr(rng, 600, 64, 0.1), 20, rng)         # n not a multiple of 4
    emit("wide_8x700_n256", random_csr(rng, 8, 700, 0.6), 256, rng)        # few long rows
    emit("allzero_40x40_n16", sp.csr_matrix((40, 40), dtype=np.float32), 16, rng)


if __name__ == "__main__":
    main()
