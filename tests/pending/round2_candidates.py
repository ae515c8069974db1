"""GPU checks for code that was written after round 1's GPU budget ran out.  NOT collected by the
default `pytest tests` run (file name has no test_ prefix); run explicitly on a B200:

    python -m pytest tests/pending/round2_candidates.py -q -m gpu

Promote each test into tests/test_gpu_parity.py once it has passed on hardware."""
import importlib

import numpy as np
import pytest
import torch

import ofspmm_b200 as ofs
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ops = importlib.import_module("of-spmm_b200.ops")
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_bwd_b_transient_route(dtype, idx):
    A = ofs.graphs.rmat_csr(12, 16, seed=4)
    N = 128
    dY = ofs.graphs.upstream_grad(A.rows, N, 6).to(dtype)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    ref = O.spmm_t_f64(crow, col, val, dY.float().numpy(), A.cols)
    amax, cnt = O.spmm_t_absmax(crow, col, val, dY.float().numpy(), A.cols)
    got = ops.spmm_csr_grad_b_transient_compute(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), dY.to(DEV),
                                                A.rows, A.cols)
    again = ops.spmm_csr_grad_b_transient_compute(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), dY.to(DEV),
                                                  A.rows, A.cols)
    assert torch.equal(got, again)                       # deterministic
    err = np.abs(got.float().cpu().numpy().astype(np.float64) - ref)
    if dtype == torch.float32:
        assert (err <= O.fp32_tolerance(ref, amax, cnt) + 1e-30).all()
        tr = ofs.csr_transpose(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), A.rows, A.cols)
        cached = ofs.spmm_csr_grad_b(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), dY.to(DEV), A.rows, A.cols,
                                     transposed=tr)
        assert torch.equal(got, cached)                  # same kernel, same transposed CSR -> same bits
    else:
        assert (err <= 1e-2 * np.abs(ref) + 2.0 ** -8 * amax + 1e-30).all()


def test_fwd_host_entry_point():
    """ofspmm_fwd_host: host pointers in, host pointers out, staging carved from the workspace."""
    import ctypes
    L = ofs._lib.lib()
    A = ofs.graphs.uniform_csr(2000, 1500, 0.01, seed=3)
    N = 64
    B = ofs.graphs.dense_operand(A.cols, N, 3)
    crow, col, val = (t.pin_memory() for t in (A.crow, A.col, A.val))
    Bp = B.pin_memory()
    C = torch.empty((A.rows, N)).pin_memory()
    cs = ofs._lib.CsrStruct(A.rows, A.cols, A.nnz, crow.data_ptr(), col.data_ptr(), val.data_ptr(), 5, 2)
    nbytes = L.ofspmm_fwd_host_workspace_bytes(A.rows, A.cols, A.nnz, N, 2, 5, 2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    rc = L.ofspmm_fwd_host(ctypes.byref(cs), Bp.data_ptr(), C.data_ptr(), N, 2, ws.data_ptr(), nbytes,
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = O.spmm_f64(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy())
    amax = O.spmm_absmax(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy())
    assert (np.abs(C.numpy() - ref) <= O.fp32_tolerance(ref, amax, np.diff(A.crow.numpy())) + 1e-30).all()
