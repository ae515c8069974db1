"""CPU tests of the format / construction utilities: COO→CSR with
coalescing, self loops, normalisations, scipy-compatible .npz round trips and edge lists, each
checked against scipy.sparse as the independent implementation."""
import importlib

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import ofspmm_b200 as ofs

fmt = importlib.import_module("of-spmm_b200.formats")


def _dense(A):
    return A.scipy().toarray()


def test_coo_to_csr_matches_scipy_with_duplicates():
    rng = np.random.default_rng(0)
    M, K, E = 50, 40, 2000                      # many duplicates
    r, c = rng.integers(0, M, E), rng.integers(0, K, E)
    v = rng.uniform(-1, 1, E).astype(np.float32)
    A = fmt.coo_to_csr(torch.from_numpy(r), torch.from_numpy(c), torch.from_numpy(v), (M, K))
    ref = sp.coo_matrix((v.astype(np.float64), (r, c)), shape=(M, K)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(A.crow.numpy(), ref.indptr) and np.array_equal(A.col.numpy(), ref.indices)
    np.testing.assert_allclose(A.val.numpy(), ref.data, rtol=1e-5, atol=1e-6)
    # sorted + unique within rows, int32 indices
    rows = torch.repeat_interleave(torch.arange(M), A.row_lengths())
    key = rows * K + A.col.long()
    assert (key[1:] > key[:-1]).all() and A.col.dtype == torch.int32
    # other coalesce modes
    Amax = fmt.coo_to_csr(torch.from_numpy(r), torch.from_numpy(c), torch.from_numpy(v), (M, K), coalesce="max")
    dense_max = np.full((M, K), -np.inf)
    np.maximum.at(dense_max, (r, c), v)
    got = _dense(Amax)
    mask = np.isfinite(dense_max)
    np.testing.assert_allclose(got[mask], dense_max[mask], rtol=1e-6)
    with pytest.raises(ValueError, match="outside"):
        fmt.coo_to_csr(torch.tensor([0, 60]), torch.tensor([0, 1]), None, (M, K))


def test_empty_and_int64():
    A = fmt.coo_to_csr(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), None, (5, 7),
                       index_dtype=torch.int64)
    assert A.nnz == 0 and A.crow.tolist() == [0] * 6 and A.crow.dtype == torch.int64


def test_self_loops_and_normalisations():
    A = ofs.graphs.uniform_csr(200, 200, 0.03, seed=4)
    A.val = A.val.abs() + 0.1
    L = fmt.add_self_loops(A)
    d = _dense(L)
    assert (np.diag(d) != 0).all()
    off = ~np.eye(200, dtype=bool)
    np.testing.assert_allclose(d[off], _dense(A)[off])
    # existing diagonal entries keep their value, missing ones get 1
    da = np.diag(_dense(A))
    np.testing.assert_allclose(np.diag(d), np.where(da != 0, da, 1.0), rtol=1e-6)
    S = fmt.sym_normalize(L)
    deg = d.sum(1)
    np.testing.assert_allclose(_dense(S), d / np.sqrt(deg)[:, None] / np.sqrt(deg)[None, :], rtol=1e-5, atol=1e-7)
    R = fmt.row_normalize(L)
    np.testing.assert_allclose(_dense(R).sum(1), np.ones(200), rtol=1e-5)


def test_npz_round_trip_is_scipy_compatible(tmp_path):
    A = ofs.graphs.rmat_csr(9, 8, seed=4)
    p = str(tmp_path / "a.npz")
    fmt.save_csr_npz(p, A)
    S = sp.load_npz(p)                                   # scipy reads our file
    assert S.shape == (A.rows, A.cols) and S.nnz == A.nnz
    assert np.array_equal(S.indptr, A.crow.numpy()) and np.array_equal(S.indices, A.col.numpy())
    B = fmt.load_csr_npz(p)
    assert torch.equal(B.crow, A.crow) and torch.equal(B.col, A.col) and torch.equal(B.val, A.val)
    # we read scipy's files, in any of its formats, unsorted indices included
    rng = np.random.default_rng(1)
    M = sp.random(60, 45, density=0.1, format="coo", random_state=rng, dtype=np.float32)
    for f in ("csr", "csc", "coo"):
        q = str(tmp_path / f"s_{f}.npz")
        sp.save_npz(q, M.asformat(f))
        C = fmt.load_csr_npz(q)
        np.testing.assert_allclose(_dense(C), M.toarray(), rtol=1e-6)


def test_edge_list(tmp_path):
    p = tmp_path / "edges.txt"
    p.write_text("# src dst w\n0 1 0.5\n1 2 2.0\n0 1 0.25\n3 0 1.0\n")
    A = fmt.load_edge_list(str(p))
    assert A.rows == A.cols == 4 and A.nnz == 3
    np.testing.assert_allclose(_dense(A)[0, 1], 0.75)
    U = fmt.load_edge_list(str(p), symmetric=True, num_nodes=6)
    d = _dense(U)
    assert U.rows == 6 and np.allclose(d, d.T) and d[1, 0] == 0.5 and d[0, 3] == 1.0


# ------------------------------------------------------------------ the same utilities on the device (csrc/build.cu)

@pytest.mark.gpu
@pytest.mark.parametrize("index_dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("coalesce", ["sum", "max", "first"])
def test_device_coo_to_csr_bit_exact_vs_host_path(index_dtype, coalesce):
    """ofspmm_coo_to_csr on the GPU == the torch host path on the same edge list: structure bit-exact;
    values bit-exact for max / first, and for sum too (duplicates are added in edge-list order on
    both sides)."""
    rng = np.random.default_rng(1)
    M, K, E = 3000, 2500, 400_000                 # ~5 % duplicates
    r = torch.from_numpy(rng.integers(0, M, E))
    c = torch.from_numpy(rng.integers(0, K, E))
    v = torch.from_numpy(rng.uniform(-1, 1, E).astype(np.float32))
    host = fmt.coo_to_csr(r, c, v, (M, K), coalesce=coalesce, index_dtype=index_dtype)
    before = ofs.launch_count()
    dev = fmt.coo_to_csr(r.cuda(), c.cuda(), v.cuda(), (M, K), coalesce=coalesce, index_dtype=index_dtype)
    assert ofs.launch_count() > before             # the library's kernels ran, not torch ops
    assert dev.crow.dtype == index_dtype and dev.nnz == host.nnz
    assert torch.equal(dev.crow.cpu(), host.crow) and torch.equal(dev.col.cpu(), host.col)
    if coalesce == "sum":
        ref = sp.coo_matrix((v.numpy().astype(np.float64), (r.numpy(), c.numpy())), shape=(M, K)).tocsr()
        ref.sum_duplicates()
        ref.sort_indices()
        np.testing.assert_allclose(dev.val.cpu().numpy(), ref.data, rtol=1e-5, atol=1e-6)
    else:
        assert torch.equal(dev.val.cpu(), host.val)
    # out-of-range edges are counted on the device and reported
    with pytest.raises(ValueError, match="outside"):
        fmt.coo_to_csr(torch.tensor([0, M]).cuda(), torch.tensor([0, 1]).cuda(), None, (M, K))
    # empty edge list
    E0 = fmt.coo_to_csr(torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), None, (5, 7))
    assert E0.nnz == 0 and E0.crow.cpu().tolist() == [0] * 6


@pytest.mark.gpu
def test_device_self_loops_and_normalisations_match_host_path():
    A = ofs.graphs.rmat_csr(11, 8, seed=4)
    A.val = A.val.abs() + 0.1
    Ad = A.to("cuda:0")
    Lh, Ld = fmt.add_self_loops(A), fmt.add_self_loops(Ad)
    assert torch.equal(Ld.crow.cpu(), Lh.crow) and torch.equal(Ld.col.cpu(), Lh.col) and torch.equal(Ld.val.cpu(), Lh.val)
    rows_d, _, _ = fmt.csr_to_coo(Ad)
    assert torch.equal(rows_d.cpu(), fmt.csr_to_coo(A)[0])
    for fn in (fmt.sym_normalize, fmt.row_normalize):
        h, d = fn(Lh), fn(Ld)
        assert torch.equal(d.col.cpu(), h.col)
        np.testing.assert_allclose(d.val.cpu().numpy(), h.val.numpy(), rtol=2e-6, atol=1e-7)
    # the GCN propagation matrix built on the device feeds the op
    S = fmt.sym_normalize(Ld)
    B = ofs.graphs.dense_operand(S.cols, 32, 1, "cuda:0")
    out = ofs.spmm_csr(S.crow, S.col, S.val, B, S.rows, S.cols)
    np.testing.assert_allclose(out.cpu().numpy(), S.scipy() @ B.cpu().numpy(), rtol=1e-4, atol=1e-5)
