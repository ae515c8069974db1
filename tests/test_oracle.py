"""CPU tests of the oracle (oracle/): pinned against the committed golden vectors (scipy fp64 +
torch CSR, tests/golden/make_golden.py) and against scipy / torch live on seeded inputs."""
import numpy as np
import pytest
import torch

import ofspmm_b200 as ofs
from oracle import oracle as O

graphs = ofs.graphs


def _tol(crow, col, val, B, C64):
    return O.fp32_tolerance(C64, O.spmm_absmax(crow, col, val, B), np.diff(crow))


def test_oracle_matches_golden_forward(golden):
    g = golden
    C64 = O.spmm_f64(g["crow"], g["col"], g["val"], g["B"], int(g["cols"]))
    np.testing.assert_allclose(C64, g["C"], rtol=1e-12, atol=1e-12)      # fp64 oracle == scipy fp64
    C32 = O.spmm_f32(g["crow"], g["col"], g["val"], g["B"], int(g["cols"]))
    tol = _tol(g["crow"], g["col"], g["val"], g["B"], g["C"])
    assert (np.abs(C32 - g["C"]) <= tol + 1e-30).all()                   # oracle-A within §8c tolerance
    assert (np.abs(g["C_torch"] - g["C"]) <= tol + 1e-30).all()          # and so is torch's kernel


def test_oracle_matches_golden_backward_and_sddmm(golden):
    g = golden
    K = int(g["cols"])
    dB64 = O.spmm_t_f64(g["crow"], g["col"], g["val"], g["dY"], K)
    np.testing.assert_allclose(dB64, g["dB"], rtol=1e-12, atol=1e-12)
    dB32 = O.spmm_t_f32(g["crow"], g["col"], g["val"], g["dY"], K)
    amax, cnt = O.spmm_t_absmax(g["crow"], g["col"], g["val"], g["dY"], K)
    assert (np.abs(dB32 - g["dB"]) <= O.fp32_tolerance(g["dB"], amax, cnt) + 1e-30).all()
    dv64, aabs = O.sddmm_f64(g["crow"], g["col"], g["dY"], g["B"])
    np.testing.assert_allclose(dv64, g["dval"], rtol=1e-12, atol=1e-12)
    dv32 = O.sddmm_f32(g["crow"], g["col"], g["dY"], g["B"])
    assert (np.abs(dv32 - g["dval"]) <= 1e-5 * np.abs(g["dval"]) + 2.0 ** -23 * g["B"].shape[1] * aabs + 1e-30).all()


def test_oracle_transpose_matches_golden(golden):
    g = golden
    t_crow, t_col, t_val, t_perm = O.csr_transpose(g["crow"], g["col"], g["val"], int(g["cols"]))
    assert np.array_equal(t_crow, g["t_crow"]) and np.array_equal(t_col, g["t_col"])
    assert np.array_equal(t_val, g["t_val"])
    assert np.array_equal(g["val"][t_perm], t_val)


@pytest.mark.parametrize("idx", [np.int32, np.int64])
def test_oracle_vs_scipy_and_torch_cfg1(idx):
    A = graphs.uniform_csr(1024, 1024, 0.01, seed=1)
    B = graphs.dense_operand(1024, 64, 1).numpy()
    crow, col, val = A.crow.numpy().astype(idx), A.col.numpy().astype(idx), A.val.numpy()
    C64 = O.spmm_f64(crow, col, val, B)
    Cs = A.scipy().astype(np.float64) @ B.astype(np.float64)
    np.testing.assert_allclose(C64, Cs, rtol=1e-12, atol=1e-13)
    Ct = (torch.sparse_csr_tensor(A.crow.long(), A.col.long(), A.val, size=(1024, 1024)) @ torch.from_numpy(B)).numpy()
    tol = _tol(crow, col, val, B, C64)
    assert (np.abs(O.spmm_f32(crow, col, val, B) - C64) <= tol).all()
    assert (np.abs(Ct - C64) <= tol).all()


def test_oracle_multithread_is_bitwise_single_thread():
    A = graphs.community_csr(3000, 90000, communities=4, seed=7)
    B = graphs.dense_operand(3000, 32, 7).numpy()
    a = O.spmm_f32(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B)
    for t in (2, 3, 8):
        assert np.array_equal(a, O.spmm_f32(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B, threads=t))


def test_balanced_split_matches_reference_semantics():
    # BalancedSplitter: first (total % parts) ranges get one extra element
    # (reference: oneflow/core/common/balanced_splitter.cpp:20-39)
    total, parts = 103, 8
    ranges = [O.balanced_split(total, parts, i) for i in range(parts)]
    assert ranges[0] == (0, 13) and ranges[6] == (78, 91) and ranges[7] == (91, 103)
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(parts - 1))


def test_out_of_range_columns_are_skipped():
    crow = np.array([0, 3], np.int32)
    col = np.array([0, 7, -1], np.int32)   # 7 and -1 are outside [0, 2)
    val = np.array([2.0, 5.0, 9.0], np.float32)
    B = np.array([[1.0, 2.0], [3.0, 4.0]], np.float32)
    np.testing.assert_array_equal(O.spmm_f32(crow, col, val, B), [[2.0, 4.0]])


@pytest.mark.parametrize("P", [1, 2, 3, 7, 64, 1000])
def test_merge_path_partition_properties(P):
    A = graphs.rmat_csr(10, 8, seed=4)
    crow = A.crow.numpy()
    M, nnz = A.rows, A.nnz
    rows, nz = O.merge_path_partition(crow, P)
    total = M + nnz
    ipw = -(-total // P)
    assert np.array_equal(rows + nz, np.minimum(np.arange(P + 1) * ipw, total))
    assert rows[0] == 0 and nz[0] == 0 and rows[-1] == M and nz[-1] == nnz
    assert (np.diff(rows) >= 0).all() and (np.diff(nz) >= 0).all()
    # a split point never sits before the end of a row it claims finished, nor past the next row end
    for r, z in zip(rows, nz):
        assert crow[r] <= z or r == 0 or crow[r] <= z
        if r < M:
            assert z <= crow[r + 1]
    # brute force the merge order on a tiny case
    if P <= 7:
        order = []
        i = j = 0
        while i < M or j < nnz:
            if i < M and crow[i + 1] <= j:
                i += 1
            else:
                j += 1
            order.append((i, j))
        for k in range(1, P + 1):
            d = min(k * ipw, total)
            assert (rows[k], nz[k]) == order[d - 1]


def test_row_hist_and_blocks():
    crow = np.array([0, 0, 1, 3, 7, 7, 1031], np.int32)
    h = O.row_hist(crow)
    assert h[0] == 2 and h[1] == 1 and h[2] == 1 and h[3] == 1 and h[11] == 1 and h.sum() == 6
    A = graphs.community_csr(5000, 200000, communities=5, seed=3)
    b = O.row_blocks(A.crow.numpy(), 4)
    assert b[0] == 0 and b[-1] == A.rows and (np.diff(b) > 0).all()
    per = np.diff(A.crow.numpy()[b])
    assert per.max() / per.mean() < 1.05   # nnz-balanced


def test_bf16_helpers_match_torch():
    x = torch.randn(4096) * 100
    x[::97] = float("inf")
    x[5] = float("nan")
    ours = O.f32_to_bf16(x.numpy())
    theirs = x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    nan = np.isnan(x.numpy())
    assert np.array_equal(ours[~nan], theirs[~nan])
    back = O.bf16_to_f32(ours)
    assert np.array_equal(back[~nan], x.to(torch.bfloat16).float().numpy()[~nan])
    assert np.isnan(back[nan]).all()


def test_sbp_signatures_of_the_glue_are_semantically_valid():
    """The GetSbp signatures in oneflow_glue/spmm_op.cpp, checked on the maths with the oracle:
    B×4 ⊗ S(1)(b) → S(1)(out) for spmm_csr / spmm_csr_grad_b (a column split of the dense side
    commutes with the row-wise linear map), and S(1)(dy), S(1)(b) → P(dval) for sddmm_csr (a split
    of the contraction axis makes the result a partial sum)."""
    A = graphs.uniform_csr(300, 200, 0.05, seed=12)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    B = graphs.dense_operand(200, 48, 12).numpy()
    dY = graphs.upstream_grad(300, 48, 12).numpy()
    halves = [slice(0, 20), slice(20, 48)]          # uneven split on purpose
    C = O.spmm_f32(crow, col, val, B)
    assert np.array_equal(np.concatenate([O.spmm_f32(crow, col, val, np.ascontiguousarray(B[:, h])) for h in halves], 1), C)
    dB = O.spmm_t_f32(crow, col, val, dY, 200)
    assert np.array_equal(np.concatenate([O.spmm_t_f32(crow, col, val, np.ascontiguousarray(dY[:, h]), 200) for h in halves], 1), dB)
    dv64, _ = O.sddmm_f64(crow, col, dY, B)
    parts = [O.sddmm_f64(crow, col, np.ascontiguousarray(dY[:, h]), np.ascontiguousarray(B[:, h]))[0] for h in halves]
    np.testing.assert_allclose(parts[0] + parts[1], dv64, rtol=1e-12, atol=1e-12)
