"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI
(include/ofspmm.h via of-spmm_b200/_lib.py) and is compared with the CPU oracle on the same
seeded inputs, with the committed golden vectors, and — at BASELINE.json's full cfg2 size —
through size-independent properties.

Tolerances (BASELINE.json north_star, made concrete in SURVEY.md §8c):
  fp32:  |got - oracleB| <= 1e-5*|oracleB| + 2^-23 * len_i * max_p|val[p]*B[col[p],j]|
  bf16:  rtol 1e-2 vs oracle-B computed from the bf16-rounded inputs (+ one bf16 ulp of the
         largest term, since the output itself is rounded to bf16)
  integer / index work (partition, histogram, transpose): bit-exact.
"""
import numpy as np
import pytest
import torch

import ofspmm_b200 as ofs
from oracle import oracle as O

pytestmark = pytest.mark.gpu
graphs = ofs.graphs
ops = __import__("importlib").import_module("of-spmm_b200.ops")
_lib = ofs._lib
DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


def _assert_fp32(got, ref64, amax, lens, what):
    tol = O.fp32_tolerance(ref64, amax, lens, rtol=1e-5) + 1e-30
    err = np.abs(got.astype(np.float64) - ref64)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance, worst excess {float((err - tol).max()):.3e}"


def _assert_bf16(got_bf16, ref64, amax, what):
    got = got_bf16.float().cpu().numpy().astype(np.float64)
    tol = 1e-2 * np.abs(ref64) + 2.0 ** -8 * amax.astype(np.float64) + 1e-30
    err = np.abs(got - ref64)
    assert (err <= tol).all(), f"{what}: worst excess {float((err - tol).max()):.3e}"


def _fwd_case(A, N, idx=torch.int32, seed=11):
    B = graphs.dense_operand(A.cols, N, seed)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    C64 = O.spmm_f64(crow, col, val, B.numpy(), A.cols)
    amax = O.spmm_absmax(crow, col, val, B.numpy(), A.cols)
    got = ofs.spmm_csr(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), B.to(DEV), A.rows, A.cols)
    torch.cuda.synchronize()
    _assert_fp32(_np(got), C64, amax, np.diff(crow), f"fwd N={N}")
    return got


# ------------------------------------------------------------------ golden vectors

def test_golden_forward_backward_sddmm(golden):
    g = golden
    M, K = int(g["rows"]), int(g["cols"])
    crow, col, val = (torch.from_numpy(g[k]).to(DEV) for k in ("crow", "col", "val"))
    B, dY = torch.from_numpy(g["B"]).to(DEV), torch.from_numpy(g["dY"]).to(DEV)
    lens = np.diff(g["crow"])
    C = ofs.spmm_csr(crow, col, val, B, M, K)
    _assert_fp32(_np(C), g["C"], O.spmm_absmax(g["crow"], g["col"], g["val"], g["B"], K), lens, "golden fwd")
    amax_t, cnt = O.spmm_t_absmax(g["crow"], g["col"], g["val"], g["dY"], K)
    dB_atomic = ofs.spmm_csr_grad_b(crow, col, val, dY, M, K, atomic=True)
    _assert_fp32(_np(dB_atomic), g["dB"], amax_t, cnt, "golden bwd atomic")
    dB_default = ofs.spmm_csr_grad_b(crow, col, val, dY, M, K)            # default = transient transpose
    _assert_fp32(_np(dB_default), g["dB"], amax_t, cnt, "golden bwd transient")
    tr = ofs.csr_transpose(crow, col, val, M, K)
    assert np.array_equal(_np(tr[0]), g["t_crow"]) and np.array_equal(_np(tr[1]), g["t_col"])
    assert np.array_equal(_np(tr[2]), g["t_val"])
    dB_t = ofs.spmm_csr_grad_b(crow, col, val, dY, M, K, transposed=tr)
    _assert_fp32(_np(dB_t), g["dB"], amax_t, cnt, "golden bwd transpose")
    dv = ofs.sddmm_csr(crow, col, dY, B, M, K)
    _, aabs = O.sddmm_f64(g["crow"], g["col"], g["dY"], g["B"])
    err = np.abs(_np(dv).astype(np.float64) - g["dval"])
    assert (err <= 1e-5 * np.abs(g["dval"]) + 2.0 ** -23 * g["B"].shape[1] * aabs + 1e-30).all()


# ------------------------------------------------------------------ forward, seeded vs oracle

@pytest.mark.parametrize("N", [1, 3, 4, 8, 20, 32, 64, 100, 128, 192, 256, 512, 640])
def test_forward_fp32_widths(N):
    _fwd_case(graphs.uniform_csr(700, 900, 0.02, seed=5), N)


@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_forward_cfg1_uniform_4096(idx):
    _fwd_case(graphs.uniform_csr(4096, 4096, 0.01, seed=1), 64, idx)      # BASELINE configs[0]


def test_forward_reddit_twin_n128():
    _fwd_case(graphs.reddit_like(64, seed=2), 128)                         # configs[1] twin


def test_forward_rmat_hub_rows_n128():
    A = graphs.rmat_csr(14, 16, seed=4)                                    # configs[3] twin, hub skew
    assert int(A.row_lengths().max()) > 1500
    _fwd_case(A, 128)


def test_forward_edge_cases():
    # all rows empty
    M, K, N = 50, 40, 16
    crow = torch.zeros(M + 1, dtype=torch.int32, device=DEV)
    e = torch.empty(0, dtype=torch.int32, device=DEV)
    C = ofs.spmm_csr(crow, e, torch.empty(0, device=DEV), torch.randn(K, N, device=DEV), M, K)
    assert C.shape == (M, N) and (C == 0).all()
    # zero rows / zero width
    assert ofs.spmm_csr(torch.zeros(1, dtype=torch.int32, device=DEV), e, torch.empty(0, device=DEV),
                        torch.randn(K, N, device=DEV), 0, K).shape == (0, N)
    assert ofs.spmm_csr(crow, e, torch.empty(0, device=DEV), torch.randn(K, 0, device=DEV), M, K).shape == (M, 0)
    # one hub row far longer than a task, surrounded by empty rows (segmented fix-up path)
    M, K, N = 9, 5000, 128
    lens = torch.tensor([0, 0, 0, 4097, 0, 1, 0, 0, 255])
    cols = torch.cat([torch.randperm(K)[:l].sort().values for l in lens.tolist()]).int()
    crow = torch.zeros(M + 1, dtype=torch.int32)
    crow[1:] = lens.cumsum(0)
    A = graphs.CsrMatrix(crow, cols, torch.rand(cols.numel()) * 2 - 1, M, K)
    _fwd_case(A, N)
    # exactly task-sized and task+1 rows, empty rows at both ends
    for L in (255, 256, 257, 511, 512):
        lens = torch.tensor([0, L, 0, 0, L, L, 0])
        cols = torch.cat([torch.randperm(K)[:l].sort().values for l in lens.tolist()]).int()
        crow = torch.zeros(8, dtype=torch.int32)
        crow[1:] = lens.cumsum(0)
        _fwd_case(graphs.CsrMatrix(crow, cols, torch.rand(cols.numel()) * 2 - 1, 7, K), 64)


def test_forward_out_of_range_columns_skipped():
    crow = torch.tensor([0, 3, 4], dtype=torch.int32, device=DEV)
    col = torch.tensor([0, 7, -1, 1], dtype=torch.int32, device=DEV)
    val = torch.tensor([2.0, 5.0, 9.0, 1.0], device=DEV)
    B = torch.tensor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]], device=DEV)
    C = ofs.spmm_csr(crow, col, val, B, 2, 2)
    assert torch.equal(C.cpu(), torch.tensor([[2.0, 4.0, 6.0, 8.0], [5.0, 6.0, 7.0, 8.0]]))


def test_forward_unaligned_csr_slices():
    """CSR arrays that start at odd element offsets (views into bigger buffers): the TMA staging
    must take its ragged-edge path."""
    A = graphs.uniform_csr(300, 400, 0.05, seed=9)
    B = graphs.dense_operand(400, 64, 9)
    want = O.spmm_f64(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy())
    amax = O.spmm_absmax(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy())
    for off_i, off_v in ((1, 3), (2, 1), (3, 2)):
        crow = torch.zeros(A.rows + 1 + off_i, dtype=torch.int32, device=DEV)[off_i:].copy_(A.crow)
        col = torch.zeros(A.nnz + off_i, dtype=torch.int32, device=DEV)[off_i:].copy_(A.col)
        val = torch.zeros(A.nnz + off_v, device=DEV)[off_v:].copy_(A.val)
        assert crow.data_ptr() % 16 != 0 and val.data_ptr() % 16 != 0
        got = ofs.spmm_csr(crow, col, val, B.to(DEV), A.rows, A.cols)
        _assert_fp32(_np(got), want, amax, np.diff(A.crow.numpy()), "unaligned slices")


def test_forward_is_bitwise_deterministic():
    A = graphs.rmat_csr(13, 16, seed=4).to(DEV)
    B = graphs.dense_operand(A.cols, 128, 3, DEV)
    a = ofs.spmm_csr(A.crow, A.col, A.val, B, A.rows, A.cols)
    for _ in range(3):
        assert torch.equal(a, ofs.spmm_csr(A.crow, A.col, A.val, B, A.rows, A.cols))


@pytest.mark.parametrize("N", [8, 64, 256, 264])
@pytest.mark.parametrize("bf16_vals", [False, True])
def test_forward_bf16(N, bf16_vals):
    A = graphs.products_like(256, seed=3)                                 # configs[2] twin
    B = graphs.dense_operand(A.cols, N, 4).to(torch.bfloat16)
    val = A.val.to(torch.bfloat16) if bf16_vals else A.val
    crow, col = A.crow.numpy(), A.col.numpy()
    B32, v32 = B.float().numpy(), val.float().numpy()
    C64 = O.spmm_f64(crow, col, v32, B32, A.cols)
    amax = O.spmm_absmax(crow, col, v32, B32, A.cols)
    got = ofs.spmm_csr(A.crow.to(DEV), A.col.to(DEV), val.to(DEV), B.to(DEV), A.rows, A.cols)
    assert got.dtype == torch.bfloat16
    _assert_bf16(got, C64, amax, f"bf16 fwd N={N}")
    # oracle-A (fp32 accumulate, one rounding) agrees to <= 1 bf16 ulp almost everywhere
    ref_bits = O.spmm_bf16(crow, col, v32, B.view(torch.int16).numpy().view(np.uint16), A.cols)
    ref = torch.from_numpy(O.bf16_to_f32(ref_bits))
    diff = (got.float().cpu() - ref).abs()
    assert (diff <= 2.0 ** -7 * ref.abs() + 2.0 ** -8 * torch.from_numpy(amax)).all()


# ------------------------------------------------------------------ backward wrt B, SDDMM

@pytest.mark.parametrize("N", [4, 64, 128, 130])
@pytest.mark.parametrize("route", ["atomic", "transpose", "transient", "plan"])
def test_backward_b_fp32(N, route):
    A = graphs.rmat_csr(12, 16, seed=4)
    dY = graphs.upstream_grad(A.rows, N, 6)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    dB64 = O.spmm_t_f64(crow, col, val, dY.numpy(), A.cols)
    amax, cnt = O.spmm_t_absmax(crow, col, val, dY.numpy(), A.cols)
    Ad = A.to(DEV)
    tr = ofs.csr_transpose(Ad.crow, Ad.col, Ad.val, A.rows, A.cols) if route == "transpose" else None
    plan = ops.SpmmPlan(Ad.crow, Ad.col, A.rows, A.cols, N, torch.float32, transpose=True) if route == "plan" else None
    kw = dict(transposed=tr, plan=plan, atomic=route == "atomic")
    got = ofs.spmm_csr_grad_b(Ad.crow, Ad.col, Ad.val, dY.to(DEV), A.rows, A.cols, **kw)
    _assert_fp32(_np(got), dB64, amax, cnt, f"bwd_b {route} N={N}")
    if route != "atomic":   # deterministic routes: bitwise repeatable
        assert torch.equal(got, ofs.spmm_csr_grad_b(Ad.crow, Ad.col, Ad.val, dY.to(DEV), A.rows, A.cols, **kw))


@pytest.mark.parametrize("route", ["atomic", "transpose", "transient", "plan"])
def test_backward_b_bf16(route):
    A = graphs.products_like(512, seed=3)
    N = 256
    dY = graphs.upstream_grad(A.rows, N, 6).to(torch.bfloat16)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    dB64 = O.spmm_t_f64(crow, col, val, dY.float().numpy(), A.cols)
    amax, _ = O.spmm_t_absmax(crow, col, val, dY.float().numpy(), A.cols)
    Ad = A.to(DEV)
    tr = ofs.csr_transpose(Ad.crow, Ad.col, Ad.val, A.rows, A.cols) if route == "transpose" else None
    plan = ops.SpmmPlan(Ad.crow, Ad.col, A.rows, A.cols, N, torch.bfloat16, transpose=True) if route == "plan" else None
    got = ofs.spmm_csr_grad_b(Ad.crow, Ad.col, Ad.val, dY.to(DEV), A.rows, A.cols, transposed=tr, plan=plan,
                              atomic=route == "atomic")
    _assert_bf16(got, dB64, amax, f"bwd_b bf16 {route}")


@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_transpose_bit_exact_vs_oracle(idx):
    A = graphs.rmat_csr(12, 8, seed=4)
    t_crow, t_col, t_val, t_perm = O.csr_transpose(A.crow.numpy(), A.col.numpy(), A.val.numpy(), A.cols)
    got = ofs.csr_transpose(A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), A.rows, A.cols, want_perm=True)
    assert got[0].dtype == idx
    assert np.array_equal(_np(got[0]), t_crow) and np.array_equal(_np(got[1]), t_col)
    assert np.array_equal(_np(got[2]), t_val) and np.array_equal(_np(got[3]), t_perm)


@pytest.mark.parametrize("N", [4, 64, 128, 100, 256, 1024, 2052])
def test_sddmm_fp32(N):
    A = graphs.rmat_csr(11, 16, seed=4)
    B = graphs.dense_operand(A.cols, N, 2)
    dY = graphs.upstream_grad(A.rows, N, 2)
    ref, aabs = O.sddmm_f64(A.crow.numpy(), A.col.numpy(), dY.numpy(), B.numpy())
    Ad = A.to(DEV)
    got = ofs.sddmm_csr(Ad.crow, Ad.col, dY.to(DEV), B.to(DEV), A.rows, A.cols)
    err = np.abs(_np(got).astype(np.float64) - ref)
    assert (err <= 1e-5 * np.abs(ref) + 2.0 ** -23 * N * aabs + 1e-30).all()


def test_sddmm_bf16():
    A = graphs.products_like(512, seed=3)
    N = 256
    B = graphs.dense_operand(A.cols, N, 2).to(torch.bfloat16)
    dY = graphs.upstream_grad(A.rows, N, 2).to(torch.bfloat16)
    ref, aabs = O.sddmm_f64(A.crow.numpy(), A.col.numpy(), dY.float().numpy(), B.float().numpy())
    Ad = A.to(DEV)
    for vdt in (torch.float32, torch.bfloat16):
        got = ofs.sddmm_csr(Ad.crow, Ad.col, dY.to(DEV), B.to(DEV), A.rows, A.cols, vdt)
        assert got.dtype == vdt
        err = np.abs(got.float().cpu().numpy().astype(np.float64) - ref)
        assert (err <= 1e-2 * np.abs(ref) + 2.0 ** -8 * aabs + 1e-30).all()


# ------------------------------------------------------------------ integer work: bit-exact

@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
@pytest.mark.parametrize("parts", [1, 2, 8, 148, 4736, 100003])
def test_device_partition_bit_exact_vs_host(idx, parts):
    A = graphs.rmat_csr(14, 16, seed=4)
    crow = A.crow.to(idx)
    dr, dz = ofs.merge_path_partition(crow.to(DEV), A.nnz, parts)
    hr, hz = ofs.merge_path_partition_host(crow, A.nnz, parts)
    orr, oz = O.merge_path_partition(crow.numpy(), parts)
    assert torch.equal(dr.cpu(), hr) and torch.equal(dz.cpu(), hz)
    assert np.array_equal(hr.numpy(), orr) and np.array_equal(hz.numpy(), oz)
    assert torch.equal(ofs.row_blocks(crow.to(DEV), A.nnz, min(parts, 8)).cpu(), ofs.row_blocks(crow, A.nnz, min(parts, 8)))


def test_row_hist_bit_exact():
    for A in (graphs.rmat_csr(14, 16, seed=4), graphs.reddit_like(64, seed=2)):
        h = ofs.row_hist(A.crow.to(DEV))
        assert np.array_equal(_np(h), O.row_hist(A.crow.numpy()))
        assert int(h.sum()) == A.rows


# ------------------------------------------------------------------ autograd mirror of the grad function

def test_autograd_matches_oracles_and_torch():
    A = graphs.uniform_csr(500, 300, 0.03, seed=8)
    N = 64
    B = graphs.dense_operand(A.cols, N, 8)
    dY = graphs.upstream_grad(A.rows, N, 8)
    Ad = A.to(DEV)
    val = Ad.val.clone().requires_grad_(True)
    Bd = B.to(DEV).requires_grad_(True)
    state = ofs.SpmmOpKernelState()
    out = ofs.spmm_csr(Ad.crow, Ad.col, val, Bd, A.rows, A.cols, state)
    out.backward(dY.to(DEV))
    ref_dv, aabs = O.sddmm_f64(A.crow.numpy(), A.col.numpy(), dY.numpy(), B.numpy())
    assert (np.abs(_np(val.grad) - ref_dv) <= 1e-5 * np.abs(ref_dv) + 2.0 ** -23 * N * aabs).all()
    ref_db = O.spmm_t_f64(A.crow.numpy(), A.col.numpy(), A.val.numpy(), dY.numpy(), A.cols)
    amax, cnt = O.spmm_t_absmax(A.crow.numpy(), A.col.numpy(), A.val.numpy(), dY.numpy(), A.cols)
    _assert_fp32(_np(Bd.grad), ref_db, amax, cnt, "autograd dB")
    # torch dense autograd as an independent witness (house style: PyTorch is the oracle,
    # python/oneflow/test_utils/automated_test_util/torch_flow_dual_object.py:1066-1107)
    Bt = B.clone().requires_grad_(True)
    dense = torch.tensor(A.scipy().toarray())
    (dense @ Bt).backward(dY)
    assert torch.allclose(Bd.grad.cpu(), Bt.grad, rtol=1e-4, atol=1e-5)
    # only b needs grad -> no SDDMM; index inputs never get grads
    Bd2 = B.to(DEV).requires_grad_(True)
    ofs.spmm_csr(Ad.crow, Ad.col, Ad.val, Bd2, A.rows, A.cols).sum().backward()
    assert Bd2.grad is not None and Ad.val.grad is None


def test_cuda_graph_capture_and_replay():
    """The kernels are capturable (user_op::CudaGraphSupport contract): no sync, no allocation,
    no host-dependent control flow inside the C-ABI call."""
    A = graphs.reddit_like(128, seed=2).to(DEV)
    B = graphs.dense_operand(A.cols, 128, 3, DEV)
    want = ofs.spmm_csr(A.crow, A.col, A.val, B, A.rows, A.cols)
    out = torch.empty_like(want)
    ops = __import__("importlib").import_module("of-spmm_b200.ops")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=out)   # warm-up
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    out.zero_()
    with torch.cuda.graph(g):
        ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=out)
    B.mul_(2.0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ofs.spmm_csr(A.crow, A.col, A.val, B, A.rows, A.cols))


# ------------------------------------------------------------------ full cfg2 size: properties

def test_full_size_reddit_properties():
    """BASELINE configs[1] at full size (232 965 nodes, ~114.6 M nnz, N=128 fp32): the CPU oracle
    would take minutes, so check size-independent properties instead —
      (1) A·1 = row sums of val (checked against a torch segment reduce, rtol/atol per §8c);
      (2) linearity: A·(2B1 - 3B2) == 2·A·B1 - 3·A·B2 within fp32 rounding;
      (3) a row sample compared against the oracle on exactly those rows;
      (4) <dY, A·B> == <A^T·dY, B> (adjoint identity ties forward and both backward routes);
      (5) bitwise determinism."""
    A = graphs.reddit_like(1, seed=2, device=DEV)
    assert A.rows == 232965 and abs(A.nnz - 114615892) <= 1000
    N = 128
    lens = A.row_lengths()
    ones = torch.ones(A.cols, N, device=DEV)
    C1 = ofs.spmm_csr(A.crow, A.col, A.val, ones, A.rows, A.cols)
    csum = torch.zeros(A.nnz + 1, dtype=torch.float64, device=DEV)
    csum[1:] = torch.cumsum(A.val.double(), 0)
    rowsum = csum[A.crow[1:].long()] - csum[A.crow[:-1].long()]
    absmax = torch.segment_reduce(A.val.abs(), "max", lengths=lens, unsafe=True, initial=0.0)
    tol = 1e-5 * rowsum.abs() + 2.0 ** -23 * lens.double() * absmax.double() + 1e-30
    assert ((C1[:, 0].double() - rowsum).abs() <= tol).all()
    assert torch.equal(C1[:, :1].expand(-1, N), C1)
    B1 = graphs.dense_operand(A.cols, N, 21, DEV)
    B2 = graphs.dense_operand(A.cols, N, 22, DEV)
    Ca = ofs.spmm_csr(A.crow, A.col, A.val, B1, A.rows, A.cols)
    Cb = ofs.spmm_csr(A.crow, A.col, A.val, B2, A.rows, A.cols)
    Cc = ofs.spmm_csr(A.crow, A.col, A.val, 2 * B1 - 3 * B2, A.rows, A.cols)
    scale = (lens.float().sqrt() * 4 * absmax)[:, None] + 1e-6
    assert (((2 * Ca - 3 * Cb) - Cc).abs() / scale).max() < 2e-4
    assert torch.equal(Ca, ofs.spmm_csr(A.crow, A.col, A.val, B1, A.rows, A.cols))
    # row sample vs oracle
    rows = torch.randint(0, A.rows, (64,), generator=torch.Generator().manual_seed(5)).tolist()
    B1h = B1.cpu().numpy()
    for r in rows:
        sub = A.row_slice(r, r + 1)
        crow, col, val = _np(sub.crow), _np(sub.col), _np(sub.val)
        want = O.spmm_f64(crow, col, val, B1h, A.cols)
        amax = O.spmm_absmax(crow, col, val, B1h, A.cols)
        _assert_fp32(_np(Ca[r:r + 1]), want, amax, np.diff(crow), f"cfg2 row {r}")
    # adjoint identity
    dY = graphs.upstream_grad(A.rows, N, 23, DEV)
    lhs = (dY.double() * Ca.double()).sum()
    dB_atomic = ofs.spmm_csr_grad_b(A.crow, A.col, A.val, dY, A.rows, A.cols, atomic=True)
    tr = ofs.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols)
    dB_t = ofs.spmm_csr_grad_b(A.crow, A.col, A.val, dY, A.rows, A.cols, transposed=tr)
    dB_tr = ofs.spmm_csr_grad_b(A.crow, A.col, A.val, dY, A.rows, A.cols)          # transient route
    assert torch.equal(dB_t, dB_tr)
    for dB in (dB_atomic, dB_t):
        rhs = (dB.double() * B1.double()).sum()
        assert abs(float(lhs - rhs)) <= 1e-6 * float((dY.double() * Ca.double()).abs().sum())
    # SDDMM adjoint: <dval, val> == <dY, A·B>
    dv = ofs.sddmm_csr(A.crow, A.col, dY, B1, A.rows, A.cols)
    assert abs(float((dv.double() * A.val.double()).sum() - lhs)) <= 1e-6 * float((dY.double() * Ca.double()).abs().sum())


# ------------------------------------------------------------------ strided dense operands (column panels)

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_strided_column_panels(dtype):
    """ofspmm_fwd_strided: B and C as column slices of wider row-major buffers (what the multi-GPU
    path uses to pipeline column panels); must equal the contiguous call bit for bit."""
    A = graphs.rmat_csr(12, 16, seed=4).to(DEV)
    N, w = 128, 32
    B = graphs.dense_operand(A.cols, N, 3, DEV, dtype)
    want = ofs.spmm_csr(A.crow, A.col, A.val, B, A.rows, A.cols)
    ops = __import__("importlib").import_module("of-spmm_b200.ops")
    C = torch.full((A.rows, N + 32), 7.0, dtype=dtype, device=DEV)        # wider output buffer
    for j in range(N // w):
        ops.spmm_csr_compute(A.crow, A.col, A.val, B[:, j * w:(j + 1) * w], A.rows, A.cols,
                             out=C[:, j * w:(j + 1) * w])
    # a panel result equals the same columns of the full-width product up to summation order
    # within a row (different lane grouping) → compare with tolerance, and exactly against a
    # contiguous call of the same width
    for j in range(N // w):
        Bj = B[:, j * w:(j + 1) * w].contiguous()
        assert torch.equal(C[:, j * w:(j + 1) * w], ofs.spmm_csr(A.crow, A.col, A.val, Bj, A.rows, A.cols))
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert torch.allclose(C[:, :N].float(), want.float(), rtol=tol, atol=tol * 10)
    assert (C[:, N:] == 7.0).all()                                          # nothing written outside
