"""python/oneflow/test/modules/test_global_spmm_csr.py — global-tensor (SBP) drop-in test in the house
style of test_global_matmul.py:25-58: loop over all placements and the SBPs the op's GetSbp admits,
create the inputs as broadcast tensors, re-shard with to_global, and compare the (broadcast-gathered)
result and gradients with a single-device scipy oracle.  PyTorch has no CSR x dense op with this
schema, so the oracle is explicit (like test_gather.py:27-38) instead of the @autotest dual object.

Signatures exercised (of-spmm_b200/oneflow_glue/spmm_op.cpp:BilinearSbp): CSR structure always B;
  b: S(1) -> out: S(1);   b: P -> out: P;   a_val: P -> out: P;   all B.
Runs inside a OneFlow tree that carries the glue in this directory."""
import unittest

import numpy as np
import scipy.sparse as sp

import oneflow as flow
import oneflow.unittest
from oneflow.test_utils.automated_test_util import all_placement, globaltest


def _csr(rows, cols, density, seed):
    a = sp.random(rows, cols, density=density, format="csr", random_state=np.random.RandomState(seed), dtype=np.float32)
    a.sort_indices()
    return a


def _to_global(x, placement, sbp, requires_grad=False):
    t = flow.tensor(x).to_global(placement=placement, sbp=flow.sbp.broadcast)   # same data on every rank
    t = t.to_global(placement=placement, sbp=sbp)
    t.requires_grad = requires_grad
    return t


def _test_global_spmm_csr(test_case, placement, val_sbp, b_sbp, n=16):
    a = _csr(48, 40, 0.2, 3)
    rng = np.random.RandomState(5)
    b_np = rng.randn(40, n).astype(np.float32)
    dy_np = rng.rand(48, n).astype(np.float32)
    B = flow.sbp.broadcast
    crow = _to_global(a.indptr.astype(np.int32), placement, B)
    col = _to_global(a.indices.astype(np.int32), placement, B)
    val = _to_global(a.data, placement, val_sbp, requires_grad=True)
    b = _to_global(b_np, placement, b_sbp, requires_grad=True)
    out = flow._C.spmm_csr(crow, col, val, b, 48, 40)
    want = a @ b_np
    got = out.to_global(placement=placement, sbp=B).to_local().numpy()
    test_case.assertTrue(np.allclose(got, want, rtol=1e-4, atol=1e-5))
    out.backward(_to_global(dy_np, placement, B))
    rows = np.repeat(np.arange(48), np.diff(a.indptr))
    test_case.assertTrue(np.allclose(b.grad.to_global(placement=placement, sbp=B).to_local().numpy(), a.T @ dy_np,
                                     rtol=1e-4, atol=1e-5))
    test_case.assertTrue(np.allclose(val.grad.to_global(placement=placement, sbp=B).to_local().numpy(),
                                     np.einsum("ij,ij->i", dy_np[rows], b_np[a.indices]), rtol=1e-4, atol=1e-5))


class TestGlobalSpmmCsr(flow.unittest.TestCase):
    @globaltest
    def test_spmm_csr(test_case):
        B, S1 = flow.sbp.broadcast, flow.sbp.split(1)
        for placement in all_placement():
            if placement.type != "cuda":          # no CPU kernel is registered on this path
                continue
            for val_sbp, b_sbp in ((B, B), (B, S1)):
                _test_global_spmm_csr(test_case, placement, [val_sbp] * len(placement.ranks.shape),
                                      [b_sbp] * len(placement.ranks.shape))

    @globaltest
    def test_cached_transpose_structure(test_case):
        a = _csr(48, 40, 0.2, 3)
        B = flow.sbp.broadcast
        for placement in all_placement():
            if placement.type != "cuda":
                continue
            crow = _to_global(a.indptr.astype(np.int32), placement, B)
            col = _to_global(a.indices.astype(np.int32), placement, B)
            t_crow, t_col, t_perm = flow._C.csr_transpose_structure(crow, col, 48, 40)
            at = a.T.tocsr()
            at.sort_indices()
            test_case.assertTrue(np.array_equal(t_crow.to_local().numpy(), at.indptr))
            test_case.assertTrue(np.array_equal(t_col.to_local().numpy(), at.indices))
            val = _to_global(a.data, placement, B, requires_grad=False)
            b = _to_global(np.ones((40, 8), np.float32), placement, B, requires_grad=True)
            flow._C.spmm_csr(crow, col, val, b, 48, 40, t_crow, t_col, t_perm).sum().backward()
            test_case.assertTrue(np.allclose(b.grad.to_local().numpy(), a.T @ np.ones((48, 8), np.float32), rtol=1e-4, atol=1e-5))


if __name__ == "__main__":
    unittest.main()
