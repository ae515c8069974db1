"""Late additions (run last): GCNConv on the real kernels — bias + ReLU fused into
the SpMM store, the ordering-by-width rule, a shared op state across layers, against dense float64
autograd — and the launch policy of the collective multi-GPU scheme (static task order with CTAs
that retire after 2 tasks per warp) against the default policy, bit for bit."""
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


@pytest.mark.parametrize("N", [64, 128])
@pytest.mark.parametrize("graph", ["rmat", "reddit"])
def test_static_order_short_lived_ctas_same_bits(N, graph):
    """dist.AllGatherSpmm launches the products that run beside NCCL with OFSPMM_ORDER_STATIC and
    tasks_per_warp = 2 (round 1's measured policy): same arithmetic as the default launch."""
    ops = importlib.import_module("of-spmm_b200.ops")
    # large enough that the policy really changes the grid: more than 2 tasks per resident warp
    # (148 SMs x 9 CTAs x 4 warps), otherwise tasks_per_warp is a no-op
    A = ofs.graphs.rmat_csr(18, 16, seed=4, device=DEV) if graph == "rmat" else ofs.graphs.reddit_like(16, seed=2, device=DEV)
    assert (A.rows + A.nnz) // 256 > 2 * 148 * 9 * 4
    B = ofs.graphs.dense_operand(A.cols, N, 11).to(DEV)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, N, torch.float32, transpose=True)
    ref = ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, plan=plan)
    got = ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, plan=plan, tasks_per_warp=2, order="static")
    assert torch.equal(ref, got)
    dY = ofs.graphs.upstream_grad(A.rows, N, 12).to(DEV)
    ref_t = ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, plan=plan)
    launcher = plan.prepared(True)                    # the path dist.CudaCompute.spmm_t takes
    out = torch.empty_like(ref_t)
    launcher(plan.transposed_values(A.val), dY, out, importlib.import_module("of-spmm_b200._lib").ORDER_STATIC, None, None, 0, 2)
    assert torch.equal(ref_t, out)
