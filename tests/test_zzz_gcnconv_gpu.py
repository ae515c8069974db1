"""GCNConv on the real kernels: bias + ReLU fused into the SpMM store, the
ordering-by-width rule, a shared op state across layers — against dense float64 autograd."""
import importlib

import pytest
import torch

import ofspmm_b200 as ofs
from test_gcn import _conv_vs_dense, _setup

gcn = importlib.import_module("of-spmm_b200.gcn")
F = importlib.import_module("of-spmm_b200.functional")
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_gcnconv_gpu_matches_dense_fp64():
    A, _, _ = _setup(DEV)
    X = ofs.graphs.dense_operand(A.rows, 128, 3).to(DEV)          # CPU generator: same numbers on every box
    l1 = gcn.GCNConv(128, 64, bias=True, activation="relu", seed=3, device=DEV)      # multiply first: fused bias + ReLU at width 64
    l2 = gcn.GCNConv(64, 128, bias=True, activation=None, seed=4, device=DEV)        # aggregate first at width 64
    assert l1.multiply_first and not l2.multiply_first
    with torch.no_grad():
        l1.bias.copy_(torch.linspace(-0.3, 0.3, 64))
        l2.bias.copy_(torch.linspace(0.2, -0.2, 128))
    val = A.val.clone().requires_grad_(True)
    state = F.SpmmOpKernelState()
    before = ofs.launch_count()
    _conv_vs_dense([l1, l2], A, X, val, DEV, state)
    assert ofs.launch_count() > before                       # the C ABI did the work
    # fused epilogue == unfused composition of the same op, to rounding
    Z = ofs.graphs.dense_operand(A.cols, 64, 8).to(DEV)
    bias = torch.linspace(-1, 1, 64, device=DEV)
    fused = F.spmm_csr_bias_act(A.crow, A.col, A.val, Z, A.rows, A.cols, bias=bias, relu=True)
    plain = torch.relu(ofs.spmm_csr(A.crow, A.col, A.val, Z, A.rows, A.cols) + bias)
    assert torch.allclose(fused, plain, rtol=1e-5, atol=1e-6)
