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
