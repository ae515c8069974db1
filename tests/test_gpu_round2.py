"""GPU parity tests added in round 2 (run with -m gpu on the B200 box): kernel variants chosen by the
row-length histogram, cached plans, the fused epilogue / accumulate mode, skipped (out-of-range)
columns on every route, the structure-cached backward, the host-buffer entry, the row-exchange
kernels of the multi-GPU path, and full-size property tests for BASELINE configs[2] and [3].
Every call goes through the C ABI; the checker is the CPU oracle (oracle/), tolerances as in
tests/test_gpu_parity.py."""
import ctypes

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
X = _lib.VARIANT_EXPLICIT


def _np(t):
    return t.detach().cpu().numpy()


def _fp32_ok(got, ref64, amax, lens, what, extra=0.0):
    tol = O.fp32_tolerance(ref64, amax, lens, rtol=1e-5) + extra + 1e-30
    err = np.abs(np.asarray(got, dtype=np.float64) - ref64)
    assert (err <= tol).all(), f"{what}: worst excess {float((err - tol).max()):.3e}"


def _oracle_fwd(A, B):
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    return (O.spmm_f64(crow, col, val, B.numpy(), A.cols), O.spmm_absmax(crow, col, val, B.numpy(), A.cols),
            np.diff(crow))


# ------------------------------------------------------------------ variants, plans, task order

@pytest.mark.parametrize("N", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("variant", [X, X | _lib.VARIANT_ITEMS64, X | _lib.VARIANT_ROWPAR, X | _lib.VARIANT_ROWS])
@pytest.mark.parametrize("graph", ["rmat", "reddit"])
def test_forward_explicit_variants(N, variant, graph):
    """Every kernel family on every lane layout against the fp64 oracle (a family that has no
    kernel for a width falls back to the base family — still has to be right)."""
    A = graphs.rmat_csr(12, 16, seed=4) if graph == "rmat" else graphs.reddit_like(128, seed=2)
    B = graphs.dense_operand(A.cols, N, 11)
    C64, amax, lens = _oracle_fwd(A, B)
    Ad = A.to(DEV)
    got = ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols, variant=variant)
    _fp32_ok(_np(got), C64, amax, lens, f"variant {variant:#x} N={N}")
    # dynamic and static task order are the same arithmetic: bitwise equal
    for order in ("dynamic", "static"):
        assert torch.equal(got, ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols,
                                                     variant=variant, order=order))
    # short-lived CTAs (multi-GPU launch policy): same bits again
    assert torch.equal(got, ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols,
                                                 variant=variant, tasks_per_warp=2))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_plan_matches_unplanned_bitwise(dtype):
    """A cached plan (histogram-chosen variant + task partition) changes no bit of the result of
    the same variant computed without a plan; forward, SDDMM and the cached-structure backward."""
    A = graphs.products_like(256, seed=3).to(DEV)
    N = 256
    B = graphs.dense_operand(A.cols, N, 4, DEV, dtype)
    dY = graphs.upstream_grad(A.rows, N, 5, DEV, dtype)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, N, dtype, transpose=True)
    assert plan.variant & X and plan.t_variant & X
    a = ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, plan=plan)
    b = ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, variant=plan.variant)
    assert torch.equal(a, b)
    assert torch.equal(ops.sddmm_csr_compute(A.crow, A.col, dY, B, A.rows, A.cols, plan=plan),
                       ops.sddmm_csr_compute(A.crow, A.col, dY, B, A.rows, A.cols))
    tr = ofs.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols)
    d_plan = ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, plan=plan)
    d_tr = ops.spmm_csr_compute(tr[0], tr[1], tr[2], dY, A.cols, A.rows, variant=plan.t_variant)
    assert torch.equal(d_plan, d_tr)
    d_glue = ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, plan=plan, regather=True)
    assert torch.equal(d_plan, d_glue)                      # ofspmm_bwd_b_cached: values gathered inside the call
    # a plan of another structure is rejected, not silently used
    other = graphs.products_like(512, seed=3).to(DEV)
    with pytest.raises(ofs.OpInferError):
        ops.spmm_csr_compute(other.crow, other.col, other.val, B[:other.cols], other.rows, other.cols, plan=plan)


def test_cached_structure_backward_sees_value_updates():
    """a_val updated in place must reach dB — only the structure of A^T is cached."""
    A = graphs.rmat_csr(11, 16, seed=4)
    N = 64
    dY = graphs.upstream_grad(A.rows, N, 6)
    Ad = A.to(DEV)
    val = Ad.val.clone()
    plan = ops.SpmmPlan(Ad.crow, Ad.col, A.rows, A.cols, N, torch.float32, transpose=True)
    ops.spmm_csr_grad_b_compute(Ad.crow, Ad.col, val, dY.to(DEV), A.rows, A.cols, plan=plan)
    val.mul_(-3.0).add_(0.25)                                        # in place: same pointer
    got = ops.spmm_csr_grad_b_compute(Ad.crow, Ad.col, val, dY.to(DEV), A.rows, A.cols, plan=plan)
    assert torch.equal(got, ops.spmm_csr_grad_b_compute(Ad.crow, Ad.col, val, dY.to(DEV), A.rows, A.cols, plan=plan,
                                                        regather=True))
    v2 = _np(val)
    ref = O.spmm_t_f64(A.crow.numpy(), A.col.numpy(), v2, dY.numpy(), A.cols)
    amax, cnt = O.spmm_t_absmax(A.crow.numpy(), A.col.numpy(), v2, dY.numpy(), A.cols)
    _fp32_ok(_np(got), ref, amax, cnt, "cached-structure dB after in-place value update")


def test_small_problem_variants():
    """BASELINE configs[0] (4096^2, 1 %, N=64): without the histogram AUTO picks 64-item tasks; the
    plan (histogram: no row reaches 512 non-zeros) picks the one-launch whole-row kernel, for the
    product and for the product on A^T; results vs the oracle, launches counted."""
    A = graphs.uniform_csr(4096, 4096, 0.01, seed=1)
    name = _lib.lib().ofspmm_fwd_variant(A.rows, A.nnz, 64, 2).decode()
    assert "64-item" in name, name
    B = graphs.dense_operand(A.cols, 64, 1)
    dY = graphs.upstream_grad(A.rows, 64, 2)
    C64, amax, lens = _oracle_fwd(A, B)
    Ad = A.to(DEV)
    plan = ops.SpmmPlan(Ad.crow, Ad.col, A.rows, A.cols, 64, transpose=True)
    assert plan.variant == X | _lib.VARIANT_ROWS and plan.t_variant == X | _lib.VARIANT_ROWS, (plan.variant, plan.t_variant)
    assert "whole_rows" in plan.variant_name()
    n0 = _lib.lib().ofspmm_launch_count()
    got = ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols, plan=plan)
    assert _lib.lib().ofspmm_launch_count() - n0 == 1            # no partition, no fix-up
    _fp32_ok(_np(got), C64, amax, lens, "cfg1 planned (whole rows)")
    small = ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols, variant=X | _lib.VARIANT_ITEMS64)
    _fp32_ok(_np(small), C64, amax, lens, "cfg1 64-item tasks")
    dB64 = O.spmm_t_f64(A.crow.numpy(), A.col.numpy(), A.val.numpy(), dY.numpy(), A.cols)
    amax_t, cnt = O.spmm_t_absmax(A.crow.numpy(), A.col.numpy(), A.val.numpy(), dY.numpy(), A.cols)
    dB = ops.spmm_csr_grad_b_compute(Ad.crow, Ad.col, Ad.val, dY.to(DEV), A.rows, A.cols, plan=plan)
    _fp32_ok(_np(dB), dB64, amax_t, cnt, "cfg1 dB (whole rows on A^T)")
    # a hub row sends the plan back to the merge path: rows of >= 512 non-zeros need the task split
    hub = graphs.rmat_csr(12, 16, seed=4).to(DEV)
    assert not (ops.SpmmPlan(hub.crow, hub.col, hub.rows, hub.cols, 64).variant & _lib.VARIANT_ROWS)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N", [24, 64, 128, 256])
def test_whole_row_kernel_epilogue_skips_strides(dtype, N):
    """The whole-row family on its own: ragged rows (empty ones included), masked lanes (N not a
    power of two), skipped column indices with NaN bait, strided operands, accumulate + bias + relu,
    and the fp32 running-sum passes of 16-bit products."""
    g = torch.Generator().manual_seed(9)
    M, K = 1500, 700
    A = graphs.uniform_csr(M, K, 0.03, seed=21)
    col = A.col.clone()
    bad = torch.rand(col.numel(), generator=g) < 0.05
    col[bad] = torch.tensor([-1, K, 2 ** 30], dtype=torch.int32)[torch.randint(0, 3, (int(bad.sum()),), generator=g)]
    if dtype == torch.float32 and N > 128:
        pytest.skip("fp32 rows wider than 512 bytes run the 64-item family (covered by test_forward_explicit_variants)")
    V = X | _lib.VARIANT_ROWS
    B = graphs.dense_operand(K, N, 5).to(dtype)
    C_old = (graphs.dense_operand(M, N, 6) * 2).to(dtype)
    bias = (torch.arange(N, dtype=torch.float32) * 0.02 - 0.4).to(dtype)
    crow, coln, val = A.crow.numpy(), col.numpy(), A.val.numpy()
    Bf = B.float().numpy()
    ref = O.spmm_f64(crow, coln, val, Bf, K)
    amax = O.spmm_absmax(crow, coln, val, Bf, K)
    lens = np.diff(crow)

    def ok(got, want, what, mag=0.0):
        gg = got.float().cpu().numpy().astype(np.float64)
        if dtype == torch.float32:
            _fp32_ok(gg, want, amax, lens, what, extra=2.0 ** -22 * mag)
        else:
            assert (np.abs(gg - want) <= 1e-2 * np.abs(want) + 2.0 ** -8 * (amax + mag) + 1e-30).all(), what

    d = dict(crow=A.crow.to(DEV), col=col.to(DEV), val=A.val.to(DEV))
    n0 = _lib.lib().ofspmm_launch_count()
    got = ops.spmm_csr_compute(d["crow"], d["col"], d["val"], B.to(DEV), M, K, variant=V)
    assert _lib.lib().ofspmm_launch_count() - n0 == 1
    ok(got, ref, "plain")
    # NaN bait in B row 0: skipped entries must not touch it
    Bn = B.clone()
    Bn[0] = float("nan")
    keep0 = np.ones(M, dtype=bool)
    keep0[np.unique(np.repeat(np.arange(M), lens)[coln == 0])] = False
    gn = ops.spmm_csr_compute(d["crow"], d["col"], d["val"], Bn.to(DEV), M, K, variant=V).float().cpu().numpy()
    assert np.isfinite(gn[keep0]).all()
    # accumulate + bias + relu
    out = C_old.to(DEV).clone()
    ops.spmm_csr_compute(d["crow"], d["col"], d["val"], B.to(DEV), M, K, out=out, accumulate=True, bias=bias.to(DEV),
                         relu=True, variant=V)
    want = np.maximum(ref + C_old.float().numpy().astype(np.float64) + bias.float().numpy().astype(np.float64)[None, :], 0)
    ok(out, want, "acc+bias+relu", mag=np.abs(C_old.float().numpy()) + np.abs(bias.float().numpy())[None, :])
    # strided B and C (row stride > n, a multiple of the 16-byte vector)
    pad = 16 // B.element_size()
    Bw = torch.zeros((K, N + pad), dtype=dtype, device=DEV)
    Bw[:, :N] = B.to(DEV)
    Cw = torch.full((M, N + 2 * pad), float("nan"), dtype=dtype, device=DEV)
    ops.spmm_csr_compute(d["crow"], d["col"], d["val"], Bw[:, :N], M, K, out=Cw[:, :N], variant=V)
    assert torch.equal(Cw[:, :N], got) and bool(torch.isnan(Cw[:, N:]).all())
    # int64 indices: no whole-row kernel for them — the 64-item family answers, same tolerance
    g64 = ops.spmm_csr_compute(d["crow"].long(), d["col"].long(), d["val"], B.to(DEV), M, K, variant=V)
    ok(g64, ref, "int64 fallback")
    if dtype != torch.float32:
        # two passes with fp32 running sums == one pass, bit for bit (one rounding)
        half = K // 2
        rows_of = torch.repeat_interleave(torch.arange(M), torch.from_numpy(lens))
        acc32 = torch.zeros((M, N), dtype=torch.float32, device=DEV)
        outs = torch.empty((M, N), dtype=dtype, device=DEV)
        masks = ((col < half) | bad, (col >= half) & ~bad)
        for i, mask in enumerate(masks):
            cnt = torch.bincount(rows_of[mask], minlength=M)
            c = torch.zeros(M + 1, dtype=torch.int32)
            c[1:] = cnt.cumsum(0)
            ops.spmm_csr_compute(c.to(DEV), col[mask].to(DEV), A.val[mask].to(DEV), B.to(DEV), M, K, out=outs, variant=V,
                                 acc32=acc32, acc32_in=i > 0, acc32_out=i == 0)
        ok(outs, ref, "two passes, fp32 running sums")


# ------------------------------------------------------------------ fused epilogue / accumulate

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["acc", "bias", "relu", "acc+bias+relu"])
@pytest.mark.parametrize("N", [64, 128, 136])
def test_forward_epilogue(dtype, mode, N):
    """C = relu(A·B [+ C_old] [+ bias]) with hub rows that span many tasks (the epilogue of a
    stitched row runs in the fix-up kernel, exactly once)."""
    A = graphs.rmat_csr(12, 16, seed=4)
    assert int(A.row_lengths().max()) > 600
    B = graphs.dense_operand(A.cols, N, 3).to(dtype)
    C_old = (graphs.dense_operand(A.rows, N, 4) * 3).to(dtype)
    bias = (torch.arange(N, dtype=torch.float32) * 0.01 - 0.5).to(dtype)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    ref = O.spmm_f64(crow, col, val, B.float().numpy(), A.cols)
    amax = O.spmm_absmax(crow, col, val, B.float().numpy(), A.cols)
    acc, bi, relu = "acc" in mode, "bias" in mode, "relu" in mode
    if acc:
        ref = ref + C_old.float().numpy().astype(np.float64)
    if bi:
        ref = ref + bias.float().numpy().astype(np.float64)[None, :]
    if relu:
        ref = np.maximum(ref, 0.0)
    Ad = A.to(DEV)
    out = C_old.to(DEV).clone() if acc else torch.full((A.rows, N), float("nan"), dtype=dtype, device=DEV)
    got = ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols, out=out, accumulate=acc,
                               bias=bias.to(DEV) if bi else None, relu=relu)
    g = got.float().cpu().numpy().astype(np.float64)
    mag = np.abs(C_old.float().numpy()) * acc + np.abs(bias.float().numpy())[None, :] * bi
    if dtype == torch.float32:
        _fp32_ok(g, ref, amax, np.diff(crow), f"epilogue {mode}", extra=2.0 ** -22 * mag)
    else:
        tol = 1e-2 * np.abs(ref) + 2.0 ** -8 * (amax + mag) + 1e-30
        assert (np.abs(g - ref) <= tol).all()
    again = C_old.to(DEV).clone() if acc else torch.empty_like(out)
    ops.spmm_csr_compute(Ad.crow, Ad.col, Ad.val, B.to(DEV), A.rows, A.cols, out=again, accumulate=acc,
                         bias=bias.to(DEV) if bi else None, relu=relu, order="static")
    assert torch.equal(got, again)


def test_two_pass_column_buckets_equal_one_pass_within_tolerance():
    """The multi-GPU forward computes C = A_local·B_local, then C += A_remote·B_remote.  Splitting
    the columns in two buckets and accumulating must match the fp64 oracle of the whole product."""
    A = graphs.reddit_like(64, seed=2)
    N = 128
    B = graphs.dense_operand(A.cols, N, 7)
    C64, amax, lens = _oracle_fwd(A, B)
    split = A.cols // 3
    crow = A.crow.long()
    rows_of = torch.repeat_interleave(torch.arange(A.rows), crow[1:] - crow[:-1])
    parts = []
    for mask in (A.col < split, A.col >= split):
        cnt = torch.bincount(rows_of[mask], minlength=A.rows)
        c = torch.zeros(A.rows + 1, dtype=torch.int32)
        c[1:] = cnt.cumsum(0)
        parts.append((c.to(DEV), A.col[mask].to(DEV), A.val[mask].to(DEV)))
    Bd = B.to(DEV)
    out = ops.spmm_csr_compute(*parts[0], Bd, A.rows, A.cols)
    ops.spmm_csr_compute(*parts[1], Bd, A.rows, A.cols, out=out, accumulate=True)
    _fp32_ok(_np(out), C64, amax, lens, "two-pass column buckets", extra=2.0 ** -22 * np.abs(C64))


# ------------------------------------------------------------------ skipped columns on every route

def test_out_of_range_columns_skipped_everywhere():
    """Indices outside [0, cols) add nothing — forward, SDDMM, atomic, transpose / transient /
    cached backward — and nothing is loaded for them: NaN in an unrelated B row cannot leak
    (the old code rewrote them to (col 0, val 0), so 0 * NaN leaked)."""
    g = torch.Generator().manual_seed(3)
    M, K, N = 300, 256, 64                                           # K a power of two (the aliasing case)
    A = graphs.uniform_csr(M, K, 0.08, seed=12)
    col = A.col.clone()
    bad = torch.rand(col.numel(), generator=g) < 0.07
    junk = torch.tensor([-1, K, K + 5, 2 ** 30], dtype=torch.int32)
    col[bad] = junk[torch.randint(0, 4, (int(bad.sum()),), generator=g)]
    B = graphs.dense_operand(K, N, 5)
    dY = graphs.upstream_grad(M, N, 6)
    crow, coln, val = A.crow.numpy(), col.numpy(), A.val.numpy()
    C64 = O.spmm_f64(crow, coln, val, B.numpy(), K)
    amax = O.spmm_absmax(crow, coln, val, B.numpy(), K)
    Bnan = B.clone()
    Bnan[0] = float("nan")                                           # row 0 is where skipped entries used to point
    keep0 = np.ones(M, dtype=bool)                                   # rows without a legitimate column-0 entry
    keep0[np.unique(np.repeat(np.arange(M), np.diff(crow))[coln == 0])] = False
    d = dict(crow=A.crow.to(DEV), col=col.to(DEV), val=A.val.to(DEV))
    got = ofs.spmm_csr(d["crow"], d["col"], d["val"], B.to(DEV), M, K)
    _fp32_ok(_np(got), C64, amax, np.diff(crow), "fwd with skipped columns")
    got_nan = _np(ofs.spmm_csr(d["crow"], d["col"], d["val"], Bnan.to(DEV), M, K))
    assert np.isfinite(got_nan[keep0]).all(), "NaN leaked through a skipped entry"
    # A^T·dY on all four routes
    dB64 = O.spmm_t_f64(crow, coln, val, dY.numpy(), K)
    amax_t, cnt = O.spmm_t_absmax(crow, coln, val, dY.numpy(), K)
    plan = ops.SpmmPlan(d["crow"], d["col"], M, K, N, transpose=True)
    tr = ofs.csr_transpose(d["crow"], d["col"], d["val"], M, K)
    assert int(tr[0][-1]) == A.nnz and int((tr[1] < 0).sum()) == int(bad.sum())
    want_t = O.csr_transpose(crow, coln, val, K)                      # skipped entries trail A^T, column -1
    got_t = ofs.csr_transpose(d["crow"], d["col"], d["val"], M, K, want_perm=True)
    for w, h in zip(want_t, got_t):
        assert np.array_equal(w, _np(h))
    for name, kw in (("atomic", dict(atomic=True)), ("transient", {}), ("transpose", dict(transposed=tr)),
                     ("plan", dict(plan=plan))):
        got = ofs.spmm_csr_grad_b(d["crow"], d["col"], d["val"], dY.to(DEV), M, K, **kw)
        _fp32_ok(_np(got), dB64, amax_t, cnt + 1, f"bwd {name} with skipped columns")
    # SDDMM: skipped entries give exactly 0
    ref, aabs = O.sddmm_f64(crow, coln, dY.numpy(), B.numpy())
    dv = _np(ofs.sddmm_csr(d["crow"], d["col"], dY.to(DEV), B.to(DEV), M, K))
    assert (dv[bad.numpy()] == 0).all()
    assert (np.abs(dv - ref) <= 1e-5 * np.abs(ref) + 2.0 ** -23 * N * aabs + 1e-30).all()


def test_packed_operands_required_for_grad_b_and_sddmm():
    """a column-slice view would be read with the wrong row stride."""
    A = graphs.uniform_csr(64, 48, 0.1, seed=2).to(DEV)
    wide = torch.randn(A.rows, 96, device=DEV)
    with pytest.raises(ofs.OpInferError):
        ofs.spmm_csr_grad_b(A.crow, A.col, A.val, wide[:, :32], A.rows, A.cols)
    with pytest.raises(ofs.OpInferError):
        ofs.sddmm_csr(A.crow, A.col, wide[:, :32], torch.randn(A.cols, 32, device=DEV), A.rows, A.cols)
    # the forward takes the view: its row stride travels as the leading dimension
    Bw = torch.randn(A.cols, 96, device=DEV)
    assert torch.equal(ofs.spmm_csr(A.crow, A.col, A.val, Bw[:, 32:64], A.rows, A.cols),
                       ofs.spmm_csr(A.crow, A.col, A.val, Bw[:, 32:64].contiguous(), A.rows, A.cols))


# ------------------------------------------------------------------ promoted from tests/pending (round 1)

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_bwd_b_transient_route(dtype, idx):
    A = graphs.rmat_csr(12, 16, seed=4)
    N = 128
    dY = graphs.upstream_grad(A.rows, N, 6).to(dtype)
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    ref = O.spmm_t_f64(crow, col, val, dY.float().numpy(), A.cols)
    amax, cnt = O.spmm_t_absmax(crow, col, val, dY.float().numpy(), A.cols)
    a = (A.crow.to(DEV, idx), A.col.to(DEV, idx), A.val.to(DEV), dY.to(DEV), A.rows, A.cols)
    got = ops.spmm_csr_grad_b_transient_compute(*a)
    assert torch.equal(got, ops.spmm_csr_grad_b_transient_compute(*a))          # deterministic
    err = np.abs(got.float().cpu().numpy().astype(np.float64) - ref)
    if dtype == torch.float32:
        assert (err <= O.fp32_tolerance(ref, amax, cnt) + 1e-30).all()
        tr = ofs.csr_transpose(a[0], a[1], a[2], A.rows, A.cols)
        assert torch.equal(got, ofs.spmm_csr_grad_b(*a, transposed=tr))         # same kernel, same A^T
    else:
        assert (err <= 1e-2 * np.abs(ref) + 2.0 ** -8 * amax + 1e-30).all()


def test_fwd_host_entry_point():
    """ofspmm_fwd_host: host pointers in, host pointers out, staging carved from the workspace."""
    L = _lib.lib()
    A = graphs.uniform_csr(2000, 1500, 0.01, seed=3)
    N = 64
    B = graphs.dense_operand(A.cols, N, 3)
    crow, col, val = (t.pin_memory() for t in (A.crow, A.col, A.val))
    Bp = B.pin_memory()
    C = torch.empty((A.rows, N)).pin_memory()
    cs = _lib.CsrStruct(A.rows, A.cols, A.nnz, crow.data_ptr(), col.data_ptr(), val.data_ptr(), 5, 2)
    nbytes = L.ofspmm_fwd_host_workspace_bytes(A.rows, A.cols, A.nnz, N, 2, 5, 2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    rc = L.ofspmm_fwd_host(ctypes.byref(cs), Bp.data_ptr(), C.data_ptr(), N, 2, ws.data_ptr(), nbytes,
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    C64, amax, lens = _oracle_fwd(A, B)
    _fp32_ok(C.numpy(), C64, amax, lens, "fwd_host")


# ------------------------------------------------------------------ row exchange kernels (multi-GPU path)

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [128, 256, 20])
@pytest.mark.parametrize("idx", [torch.int32, torch.int64])
def test_gather_and_scatter_add_rows_bit_exact(dtype, n, idx):
    g = torch.Generator().manual_seed(n)
    K, cnt, off = 5000, 1777, 1000
    src = torch.randn(K, n, generator=g).to(dtype).to(DEV)
    index = (torch.randperm(K - off, generator=g)[:cnt].sort().values + off).to(idx).to(DEV)
    dst = torch.full((cnt + 3, n), 5.0, dtype=dtype, device=DEV)
    ops.gather_rows(dst, src[off:], index, index_offset=off, max_ctas=8)
    assert torch.equal(dst[:cnt], src[index.long()]) and (dst[cnt:] == 5.0).all()
    # identity list
    d2 = torch.empty((100, n), dtype=dtype, device=DEV)
    ops.gather_rows(d2, src[40:], None, count=100)
    assert torch.equal(d2, src[40:140])
    # scatter-add of distinct rows == index_add in fp32 then one rounding (bf16)
    acc = torch.randn(K - off, n, generator=g).to(dtype).to(DEV)
    want = acc.clone()
    want[(index - off).long()] = (want[(index - off).long()].float() + dst[:cnt].float()).to(dtype)
    ops.scatter_add_rows(acc, dst[:cnt], index, index_offset=off)
    assert torch.equal(acc, want)


# ------------------------------------------------------------------ full-size properties: configs[2], configs[3]

def _full_size_properties(A, N, dtype, name):
    """Size-independent checks at BASELINE size (the CPU oracle would take minutes to hours):
    A·1 = row sums, sampled rows vs the oracle, the adjoint identity tying forward / A^T·dY / SDDMM
    together, bitwise determinism, planned == unplanned."""
    lens = A.row_lengths()
    f32 = dtype == torch.float32
    ones = torch.ones(A.cols, N, dtype=dtype, device=DEV)
    C1 = ofs.spmm_csr(A.crow, A.col, A.val, ones, A.rows, A.cols)
    csum = torch.zeros(A.nnz + 1, dtype=torch.float64, device=DEV)
    csum[1:] = torch.cumsum(A.val.double(), 0)
    rowsum = csum[A.crow[1:].long()] - csum[A.crow[:-1].long()]
    absmax = torch.segment_reduce(A.val.abs(), "max", lengths=lens, unsafe=True, initial=0.0)
    if f32:
        tol = 1e-5 * rowsum.abs() + 2.0 ** -23 * lens.double() * absmax.double() + 1e-30
    else:
        tol = 1e-2 * rowsum.abs() + 2.0 ** -8 * absmax.double() * lens.double().sqrt().clamp(min=1) + 1e-30
    assert ((C1[:, 0].double() - rowsum).abs() <= tol).all(), f"{name}: A·1 != row sums"
    assert torch.equal(C1[:, :1].expand(-1, N), C1)
    del ones, C1, csum
    B1 = graphs.dense_operand(A.cols, N, 21, DEV, dtype)
    plan = ops.SpmmPlan(A.crow, A.col, A.rows, A.cols, N, dtype, transpose=True)
    Ca = ops.spmm_csr_compute(A.crow, A.col, A.val, B1, A.rows, A.cols, plan=plan)
    assert torch.equal(Ca, ops.spmm_csr_compute(A.crow, A.col, A.val, B1, A.rows, A.cols, variant=plan.variant))
    # sampled rows (long, short and empty ones) vs the oracle on exactly those rows
    gen = torch.Generator().manual_seed(5)
    rows = torch.randint(0, A.rows, (48,), generator=gen).tolist() + torch.topk(lens, 4).indices.tolist()
    B1h = B1.float().cpu().numpy()
    for r in rows:
        sub = A.row_slice(r, r + 1)
        crow, col, val = _np(sub.crow), _np(sub.col), _np(sub.val)
        want = O.spmm_f64(crow, col, val, B1h, A.cols)
        amax = O.spmm_absmax(crow, col, val, B1h, A.cols)
        got = Ca[r:r + 1].float().cpu().numpy().astype(np.float64)
        if f32:
            _fp32_ok(got, want, amax, np.diff(crow), f"{name} row {r}")
        else:
            assert (np.abs(got - want) <= 1e-2 * np.abs(want) + 2.0 ** -8 * amax + 1e-30).all(), f"{name} row {r}"
    # adjoint identities
    dY = graphs.upstream_grad(A.rows, N, 23, DEV, dtype)
    prod = dY.double() * Ca.double()
    lhs, scale = float(prod.sum()), float(prod.abs().sum())
    del prod
    rtol = 1e-6 if f32 else 2e-3
    dB = ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, plan=plan)
    assert abs(lhs - float((dB.double() * B1.double()).sum())) <= rtol * scale, f"{name}: <dY,AB> != <A^T dY,B>"
    assert torch.equal(dB, ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, plan=plan))
    del dB
    dv = ops.sddmm_csr_compute(A.crow, A.col, dY, B1, A.rows, A.cols, plan=plan)
    assert abs(lhs - float((dv.double() * A.val.double()).sum())) <= rtol * scale, f"{name}: SDDMM adjoint"


def test_full_size_products_bf16_properties():
    """BASELINE configs[2]: ogbn-products-shaped, 2 449 029 nodes, ~123.7 M nnz, N=256 bf16."""
    A = graphs.products_like(1, seed=3, device=DEV)
    assert A.rows == 2449029 and abs(A.nnz - 123718280) <= 2000
    _full_size_properties(A, 256, torch.bfloat16, "cfg3")


def test_full_size_rmat24_properties():
    """BASELINE configs[3]: R-MAT scale 24, edge factor 16 (duplicates coalesced), N=128 fp32 —
    hub rows of > 200 000 non-zeros spread over ~1000 tasks each, 56 % empty rows."""
    A = graphs.rmat_csr(24, 16, seed=4, device=DEV)
    assert A.rows == 1 << 24 and A.nnz > 250_000_000
    assert int(A.row_lengths().max()) > 100_000
    _full_size_properties(A, 128, torch.float32, "cfg4")


# ------------------------------------------------------------------ fp32 running sums for multi-pass bf16 products

@pytest.mark.parametrize("N", [64, 256, 264])
def test_bf16_multi_pass_with_fp32_accumulator_rounds_once(N):
    """Three column buckets, bf16 in / bf16 out, running sums in the fp32 buffer: the result must meet
    the SINGLE-pass bf16 tolerance (one rounding), hub rows stitched by the fix-up kernel included."""
    A = graphs.rmat_csr(12, 16, seed=4)
    B = graphs.dense_operand(A.cols, N, 7).to(torch.bfloat16)
    crow_np, col_np, val_np = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    C64 = O.spmm_f64(crow_np, col_np, val_np, B.float().numpy(), A.cols)
    amax = O.spmm_absmax(crow_np, col_np, val_np, B.float().numpy(), A.cols)
    crow = A.crow.long()
    rows_of = torch.repeat_interleave(torch.arange(A.rows), crow[1:] - crow[:-1])
    cuts = [0, A.cols // 4, A.cols // 2, A.cols]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        mask = (A.col >= lo) & (A.col < hi)
        c = torch.zeros(A.rows + 1, dtype=torch.int32)
        c[1:] = torch.bincount(rows_of[mask], minlength=A.rows).cumsum(0)
        parts.append((c.to(DEV), A.col[mask].to(DEV), A.val[mask].to(DEV)))
    Bd = B.to(DEV)
    acc = torch.full((A.rows, N), float("nan"), dtype=torch.float32, device=DEV)
    out = torch.full((A.rows, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    bias = (torch.arange(N, dtype=torch.float32) * 0.01 - 0.3).to(torch.bfloat16).to(DEV)
    ops.spmm_csr_compute(*parts[0], Bd, A.rows, A.cols, out=out, acc32=acc, acc32_out=True)
    ops.spmm_csr_compute(*parts[1], Bd, A.rows, A.cols, out=out, acc32=acc, acc32_in=True, acc32_out=True)
    ops.spmm_csr_compute(*parts[2], Bd, A.rows, A.cols, out=out, acc32=acc, acc32_in=True, bias=bias, relu=True)
    ref = np.maximum(C64 + bias.float().cpu().numpy().astype(np.float64)[None, :], 0.0)
    got = out.float().cpu().numpy().astype(np.float64)
    tol = 1e-2 * np.abs(ref) + 2.0 ** -8 * (amax + np.abs(bias.float().cpu().numpy())[None, :]) + 1e-30
    assert (np.abs(got - ref) <= tol).all(), float((np.abs(got - ref) - tol).max())
    # fp32 products must not take the acc32 route
    with pytest.raises(ofs.OpInferError):
        ops.spmm_csr_compute(*parts[0], Bd.float(), A.rows, A.cols, acc32=acc, acc32_out=True)


def test_scatter_add_rows_into_fp32_then_cast_once():
    g = torch.Generator().manual_seed(3)
    K, n, cnt = 3000, 256, 900
    acc = torch.randn(K, n, generator=g).to(DEV)
    want = acc.clone()
    for r in range(5):                                    # five "ranks" contribute bf16 partial rows
        idx = torch.randperm(K, generator=g)[:cnt].sort().values.int().to(DEV)
        part = torch.randn(cnt, n, generator=g).to(torch.bfloat16).to(DEV)
        ops.scatter_add_rows_f32(acc, part, idx)
        want[idx.long()] += part.float()
    assert torch.equal(acc, want)                         # fp32 adds in the same order: exact
    out = torch.empty((K, n), dtype=torch.bfloat16, device=DEV)
    ops.cast_from_f32(out, acc)
    assert torch.equal(out, want.to(torch.bfloat16))


# ------------------------------------------------------------------ flag-synchronised multi-peer kernels, one GPU

@pytest.mark.parametrize("dtype,n", [(torch.float32, 128), (torch.bfloat16, 256), (torch.float32, 20)])
def test_multi_peer_pull_and_combine_kernels_bit_exact(dtype, n):
    """ofspmm_signal_peers / ofspmm_pull_rows_multi (TMA ring when rows are packed) /
    ofspmm_combine_rows_multi with the 'peers' being buffers on the same GPU: the flags are written by
    the signal kernel earlier on the same stream, so no kernel ever waits on a concurrently running
    one.  Checked against torch indexing, bit for bit; the cross-GPU use is in tests/test_gpu_multi.py."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(n)
    peers, shard = 3, 4000
    if (n * (4 if dtype == torch.float32 else 2)) % 16 != 0:
        pytest.skip("multi-peer kernels move whole 16-byte units")
    pad = torch.zeros((4, peers), dtype=torch.int64, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    epoch = 7
    slots = (ctypes.c_void_p * peers)(*[pad.data_ptr() + (0 * peers + r) * 8 for r in range(peers)])
    assert L.ofspmm_signal_peers(slots, peers, epoch, stream) == 0
    srcs = [torch.randn(shard, n, generator=g).to(dtype).to(DEV) for _ in range(peers)]
    lists = [torch.randperm(shard, generator=g)[: 700 + 311 * r].sort().values.int().to(DEV) for r in range(peers)]
    total = sum(int(l.numel()) for l in lists)
    dst = torch.full((total, n), 3.0, dtype=dtype, device=DEV)
    segs = (_lib.PullSeg * peers)()
    off = 0
    for r in range(peers):
        segs[r] = _lib.PullSeg(srcs[r].data_ptr(), lists[r].data_ptr(), pad.data_ptr() + r * 8, int(lists[r].numel()), off)
        off += int(lists[r].numel())
    dd = 2 if dtype == torch.float32 else 11
    for ctas in (0, 3):
        dst.fill_(3.0)
        assert L.ofspmm_pull_rows_multi(dst.data_ptr(), n, n, segs, peers, epoch, n, dd, 5, ctas, stream) == 0
        torch.cuda.synchronize()
        assert torch.equal(dst, torch.cat([srcs[r][lists[r].long()] for r in range(peers)]))
    # combine: acc[row] += partial rows of every "peer", in segment order, fp32 accumulator
    acc = torch.randn(shard, n, generator=g).to(DEV)
    want = acc.clone()
    parts, invs = [], []
    csegs = (_lib.CombineSeg * peers)()
    for r in range(peers):
        part = torch.randn(int(lists[r].numel()), n, generator=g).to(dtype).to(DEV)
        inv = torch.full((shard,), -1, dtype=torch.int32, device=DEV)
        inv[lists[r].long()] = torch.arange(lists[r].numel(), dtype=torch.int32, device=DEV)
        parts.append(part)
        invs.append(inv)
        csegs[r] = _lib.CombineSeg(part.data_ptr(), inv.data_ptr(), pad.data_ptr() + r * 8)
        want[lists[r].long()] += part.float()
    assert L.ofspmm_combine_rows_multi(acc.data_ptr(), n, n, csegs, peers, epoch, shard, n, dd, 0, stream) == 0
    torch.cuda.synchronize()
    assert torch.equal(acc, want)
