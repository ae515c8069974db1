"""2-layer GCN harness (BASELINE configs[4]): the composition is checked on
the CPU with an injected dense stand-in for the SpMM op, and on the GPU with the real op against a
dense float64 torch model."""
import importlib

import pytest
import torch

import ofspmm_b200 as ofs

gcn = importlib.import_module("of-spmm_b200.gcn")


def _dense_model(A_dense, val_like, W1, W2, X, labels):
    h1 = torch.relu(A_dense @ (X @ W1))
    out = (A_dense @ h1) @ W2
    return torch.nn.functional.cross_entropy(out, labels)


def _setup(device, dtype=torch.float32):
    A = ofs.graphs.gcn_normalize(ofs.graphs.add_self_loops(ofs.graphs.uniform_csr(300, 300, 0.03, seed=9)))
    X = ofs.graphs.dense_operand(300, 24, 3).to(device)
    labels = torch.randint(0, 7, (300,), generator=torch.Generator().manual_seed(1)).to(device)
    return A.to(device), X, labels


def _check_against_dense(model, A, X, labels):
    loss = model.train_step(X, labels)
    rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths().cpu())
    val64 = model.val.detach().double().cpu().requires_grad_(True)
    dense = torch.zeros(A.rows, A.cols, dtype=torch.float64).index_put((rows, A.col.cpu().long()), val64)
    W1 = model.W1.detach().double().cpu().requires_grad_(True)
    W2 = model.W2.detach().double().cpu().requires_grad_(True)
    ref = _dense_model(dense, None, W1, W2, X.double().cpu(), labels.cpu())
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert torch.allclose(model.W1.grad.double().cpu(), W1.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(model.W2.grad.double().cpu(), W2.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(model.val.grad.double().cpu(), val64.grad, rtol=1e-4, atol=1e-6)   # SDDMM path


def test_gcn_composition_cpu_with_injected_spmm():
    A, X, labels = _setup("cpu")
    rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths())

    def dense_spmm(crow, col, val, b, m, k):      # differentiable CPU stand-in for the op
        return torch.zeros(m, k).index_put((rows, col.long()), val) @ b
    model = gcn.GCN2(A, in_dim=24, hidden=16, out_dim=7, seed=5, spmm=dense_spmm)
    _check_against_dense(model, A, X, labels)
    assert model.spmm_flops_per_step() == 2 * 3 * 2 * A.nnz * 16


@pytest.mark.gpu
def test_gcn_forward_backward_gpu_matches_dense():
    A, X, labels = _setup("cuda:0")
    model = gcn.GCN2(A, in_dim=24, hidden=16, out_dim=7, seed=5)
    before = ofs.launch_count()
    _check_against_dense(model, A, X, labels)
    assert ofs.launch_count() - before >= 2 * 3 + 2 * 3 + 2   # 2 fwd, 2 bwd_b(+transpose once), 2 sddmm
    # hidden width of configs[4] on a Reddit twin: runs and stays finite
    A2 = ofs.graphs.gcn_normalize(ofs.graphs.add_self_loops(ofs.graphs.reddit_like(256, seed=2))).to("cuda:0")
    X2 = ofs.graphs.dense_operand(A2.rows, 602, 4, "cuda:0")
    y2 = torch.randint(0, 41, (A2.rows,), generator=torch.Generator().manual_seed(2)).to("cuda:0")
    m2 = gcn.GCN2(A2)
    l0 = float(m2.train_step(X2, y2, lr=0.1))
    l1 = float(m2.train_step(X2, y2, lr=0.1))
    assert l0 == l0 and l1 < l0                    # finite, and SGD on the weights reduces the loss


# ------------------------------------------------------------------ GCNConv: fused epilogue, ordering by width

class _DenseKernels:
    """CPU stand-in for functional.OpsKernels (same three methods, dense torch arithmetic): lets the
    autograd wiring of the fused-epilogue function run without a GPU."""

    def __init__(self, rows_of):
        self.rows_of = rows_of
        self.calls = []

    def _dense(self, a_col, a_val, m, k):
        return torch.zeros(m, k, dtype=a_val.dtype).index_put((self.rows_of, a_col.long()), a_val)

    def fwd(self, a_crow, a_col, a_val, b, m, k, bias, relu, plan):
        self.calls.append(("fwd", b.shape[1], bias is not None, relu))
        out = self._dense(a_col, a_val, m, k) @ b
        if bias is not None:
            out = out + bias
        return torch.relu(out) if relu else out

    def grad_b(self, a_crow, a_col, a_val, dy, m, k, plan):
        self.calls.append(("grad_b", dy.shape[1]))
        return self._dense(a_col, a_val, m, k).t() @ dy

    def sddmm(self, a_crow, a_col, dy, b, m, k, val_dtype, plan):
        self.calls.append(("sddmm", dy.shape[1]))
        return (dy[self.rows_of] * b[a_col.long()]).sum(1).to(val_dtype)


def _conv_vs_dense(conv_layers, A, X, val, dev="cpu", state=None):
    """Two stacked GCNConv layers against the same model in dense float64 autograd."""
    rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths().cpu())
    Xg = X.clone().requires_grad_(True)
    h = Xg
    for layer in conv_layers:
        h = layer(A, h, val=val, state=state)
    loss = (h * torch.linspace(-1, 1, h.shape[1], device=h.device)).sum() / h.shape[0]
    loss.backward()
    v64 = val.detach().double().cpu().requires_grad_(True)
    dense = torch.zeros(A.rows, A.cols, dtype=torch.float64).index_put((rows, A.col.cpu().long()), v64)
    X64 = X.detach().double().cpu().requires_grad_(True)
    ps = []
    h64 = X64
    for layer in conv_layers:
        W = layer.weight.detach().double().cpu().requires_grad_(True)
        bb = layer.bias.detach().double().cpu().requires_grad_(True) if layer.bias is not None else None
        ps.append((W, bb))
        h64 = dense @ (h64 @ W)                       # the order of the products does not change the maths
        if bb is not None:
            h64 = h64 + bb
        if layer.activation == "relu":
            h64 = torch.relu(h64)
    ref = (h64 * torch.linspace(-1, 1, h64.shape[1], dtype=torch.float64)).sum() / h64.shape[0]
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    close = lambda a, b: torch.allclose(a.double().cpu(), b, rtol=2e-4, atol=2e-6)
    assert close(Xg.grad, X64.grad)
    assert close(val.grad, v64.grad)
    for layer, (W, bb) in zip(conv_layers, ps):
        assert close(layer.weight.grad, W.grad)
        if bb is not None:
            assert close(layer.bias.grad, bb.grad)


def test_gcnconv_cpu_wiring_fused_epilogue_and_ordering():
    A, X, _ = _setup("cpu")
    rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths())
    K = _DenseKernels(rows)
    l1 = gcn.GCNConv(24, 16, bias=True, activation="relu", seed=3, device="cpu", kernels=K)     # 24 -> 16: multiply first
    l2 = gcn.GCNConv(16, 40, bias=True, activation=None, seed=4, device="cpu", kernels=K)       # 16 -> 40: aggregate first
    assert l1.multiply_first and not l2.multiply_first and l1.aggregation_width() == 16 == l2.aggregation_width()
    with torch.no_grad():
        l1.bias.copy_(torch.linspace(-0.3, 0.3, 16))
        l2.bias.copy_(torch.linspace(0.2, -0.2, 40))
    val = A.val.clone().requires_grad_(True)
    _conv_vs_dense([l1, l2], A, X, val)
    fwd = [c for c in K.calls if c[0] == "fwd"]
    assert fwd == [("fwd", 16, True, True), ("fwd", 16, False, False)]    # layer 1: bias + ReLU inside the product
    assert sorted(c[0] for c in K.calls if c[0] != "fwd") == ["grad_b", "grad_b", "sddmm", "sddmm"]
    # forced orders give the same numbers
    K2 = _DenseKernels(rows)
    l3 = gcn.GCNConv(24, 16, seed=3, device="cpu", order="aggregate_first", kernels=K2)
    with torch.no_grad():
        l3.bias.copy_(l1.bias)
    assert torch.allclose(l3(A, X), l1(A, X), rtol=1e-5, atol=1e-6)
    assert [c for c in K2.calls if c[0] == "fwd"] == [("fwd", 24, False, False)]
    # dropout acts on the input and is off in eval mode
    l4 = gcn.GCNConv(24, 16, seed=3, device="cpu", dropout=0.5, kernels=K)
    l4.training = False
    with torch.no_grad():
        l4.bias.copy_(l1.bias)
    assert torch.equal(l4(A, X), l1(A, X))
    # the product path refuses CPU tensors (no CPU fallback): only injected kernels run here
    with pytest.raises(Exception):
        gcn.GCNConv(24, 16, device="cpu")(A, X)
    # bias shape / dtype is checked like the op's other inputs
    from importlib import import_module
    F = import_module("of-spmm_b200.functional")
    with pytest.raises(ofs.OpInferError):
        F.spmm_csr_bias_act(A.crow, A.col, A.val, X, A.rows, A.cols, bias=torch.zeros(5), kernels=K)
