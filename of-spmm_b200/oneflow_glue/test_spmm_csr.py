"""python/oneflow/test/modules/test_spmm_csr.py — drop-in test in the reference's house style:
numpy/scipy oracle + explicit backward oracle like test_gather.py:27-38,87-101, device loop via
GenArgList, @flow.unittest.skip_unless_1n1d().  Runs inside a OneFlow tree that carries the glue
in this directory; the same cases run against the C ABI in this repo's tests/test_gpu_parity.py."""
import unittest
from collections import OrderedDict

import numpy as np
import scipy.sparse as sp

import oneflow as flow
import oneflow.unittest
from oneflow.test_utils.test_util import GenArgList


def _random_csr(rows, cols, density, rng):
    a = sp.random(rows, cols, density=density, format="csr", random_state=rng, dtype=np.float32)
    a.sort_indices()
    return a


def _test_spmm_csr_forward_backward(test_case, device, n, index_dtype):
    rng = np.random.RandomState(7)
    a = _random_csr(300, 200, 0.05, rng)
    b_np = rng.randn(200, n).astype(np.float32)
    dy_np = rng.rand(300, n).astype(np.float32)
    crow = flow.tensor(a.indptr.astype(index_dtype), device=device)
    col = flow.tensor(a.indices.astype(index_dtype), device=device)
    val = flow.tensor(a.data, device=device, requires_grad=True)
    b = flow.tensor(b_np, device=device, requires_grad=True)
    out = flow._C.spmm_csr(crow, col, val, b, 300, 200)
    test_case.assertTrue(np.allclose(out.numpy(), a @ b_np, rtol=1e-4, atol=1e-5))
    out.backward(flow.tensor(dy_np, device=device))
    test_case.assertTrue(np.allclose(b.grad.numpy(), a.T @ dy_np, rtol=1e-4, atol=1e-5))
    rows = np.repeat(np.arange(300), np.diff(a.indptr))
    dval = np.einsum("ij,ij->i", dy_np[rows], b_np[a.indices])
    test_case.assertTrue(np.allclose(val.grad.numpy(), dval, rtol=1e-4, atol=1e-5))
    test_case.assertTrue(crow.grad is None and col.grad is None)


def _test_fused_spmm_csr_bias_act(test_case, device, n, relu):
    rng = np.random.RandomState(11)
    a = _random_csr(300, 200, 0.05, rng)
    b_np = rng.randn(200, n).astype(np.float32)
    bias_np = np.linspace(-0.5, 0.5, n).astype(np.float32)
    dy_np = rng.rand(300, n).astype(np.float32)
    crow = flow.tensor(a.indptr.astype(np.int32), device=device)
    col = flow.tensor(a.indices.astype(np.int32), device=device)
    val = flow.tensor(a.data, device=device, requires_grad=True)
    b = flow.tensor(b_np, device=device, requires_grad=True)
    bias = flow.tensor(bias_np, device=device, requires_grad=True)
    out = flow._C.fused_spmm_csr_bias_act(crow, col, val, b, bias, 300, 200, relu)
    pre = a @ b_np + bias_np
    want = np.maximum(pre, 0) if relu else pre
    test_case.assertTrue(np.allclose(out.numpy(), want, rtol=1e-4, atol=1e-5))
    out.backward(flow.tensor(dy_np, device=device))
    dz = dy_np * (pre > 0) if relu else dy_np
    rows = np.repeat(np.arange(300), np.diff(a.indptr))
    test_case.assertTrue(np.allclose(bias.grad.numpy(), dz.sum(0), rtol=1e-4, atol=1e-4))
    test_case.assertTrue(np.allclose(b.grad.numpy(), a.T @ dz, rtol=1e-4, atol=1e-5))
    test_case.assertTrue(np.allclose(val.grad.numpy(), np.einsum("ij,ij->i", dz[rows], b_np[a.indices]), rtol=1e-4, atol=1e-5))


@flow.unittest.skip_unless_1n1d()
class TestSpmmCsr(flow.unittest.TestCase):
    def test_spmm_csr(test_case):
        arg_dict = OrderedDict()
        arg_dict["device"] = ["cuda"]           # no CPU kernel is registered on this path
        arg_dict["n"] = [1, 20, 64, 128, 256]
        arg_dict["index_dtype"] = [np.int32, np.int64]
        for arg in GenArgList(arg_dict):
            _test_spmm_csr_forward_backward(test_case, *arg)

    def test_fused_spmm_csr_bias_act(test_case):
        arg_dict = OrderedDict()
        arg_dict["device"] = ["cuda"]
        arg_dict["n"] = [16, 64, 136]
        arg_dict["relu"] = [False, True]
        for arg in GenArgList(arg_dict):
            _test_fused_spmm_csr_bias_act(test_case, *arg)

    def test_spmm_csr_shape_error(test_case):
        with test_case.assertRaises(Exception) as ctx:
            flow._C.spmm_csr(flow.zeros(3, dtype=flow.int32, device="cuda"), flow.zeros(0, dtype=flow.int32, device="cuda"),
                             flow.zeros(0, device="cuda"), flow.ones(5, 4, device="cuda"), 2, 7)
        test_case.assertTrue("a_cols" in str(ctx.exception))


if __name__ == "__main__":
    unittest.main()
