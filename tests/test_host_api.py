"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/ofspmm.h
declares, size queries and argument validation work without a GPU, the host partitioner is
bit-exact with the oracle, and the op mirror raises the reference-style inference errors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import ofspmm_b200 as ofs
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = __import__("importlib").import_module("of-spmm_b200._lib")


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ofspmm.h")).read()
    return sorted(set(re.findall(r"OFSPMM_API[^;]*?\b(ofspmm_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 17
    L = ctypes.CDLL(ofs.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ofspmm.h but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_library_has_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", ofs.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_strerror_and_queries():
    L = _lib.lib()
    assert L.ofspmm_version() == 100
    assert L.ofspmm_strerror(0) == b"ok"
    assert b"workspace" in L.ofspmm_strerror(3)
    # workspace grows with the task count and the dense width; bf16 adds the fp32 head buffer
    w32 = L.ofspmm_fwd_workspace_bytes(1000, 1000, 100000, 128, 2)
    wbf = L.ofspmm_fwd_workspace_bytes(1000, 1000, 100000, 128, 11)
    assert 0 < w32 < wbf
    tasks = -(-(1000 + 100000) // 256)
    assert w32 >= (tasks + 1) * 8 + tasks * 128 * 4
    assert L.ofspmm_sddmm_workspace_bytes(1000, 1000, 100000, 128, 2) >= (tasks + 1) * 8
    assert b"warp-per-row" in L.ofspmm_fwd_variant(1, 1, 128, 2)
    assert b"vector-per-row" in L.ofspmm_fwd_variant(1, 1, 64, 2)


def test_argument_validation_without_gpu():
    L = _lib.lib()
    crow = np.zeros(3, np.int32)
    A = _lib.CsrStruct(2, 2, 0, crow.ctypes.data, None, None, 5, 2)
    assert L.ofspmm_fwd(None, None, None, 4, 2, None, 0, None) == 1            # null csr
    A.rows = -1
    assert L.ofspmm_fwd(ctypes.byref(A), None, None, 4, 2, None, 0, None) == 1  # negative size
    A.rows, A.idx_dtype = 2, 9
    assert L.ofspmm_fwd(ctypes.byref(A), None, None, 4, 2, None, 0, None) == 2  # kFloat16 index
    A.idx_dtype, A.val_dtype = 5, 11
    assert L.ofspmm_fwd(ctypes.byref(A), None, None, 4, 2, None, 0, None) == 2  # bf16 val + fp32 dense
    A.val_dtype, A.nnz = 2, 2 ** 31
    assert L.ofspmm_fwd(ctypes.byref(A), None, None, 4, 2, None, 0, None) == 5  # too large
    assert L.ofspmm_partition_host(None, 5, 1, 1, 1, None, None) == 1


@pytest.mark.parametrize("idx", [np.int32, np.int64])
@pytest.mark.parametrize("parts", [1, 2, 5, 8, 333])
def test_partition_host_bit_exact_vs_oracle(idx, parts):
    A = ofs.graphs.rmat_csr(11, 8, seed=4)
    crow = A.crow.numpy().astype(idx)
    r0, z0 = O.merge_path_partition(crow, parts)
    r1, z1 = ofs.merge_path_partition_host(torch.from_numpy(crow), A.nnz, parts)
    assert np.array_equal(r0, r1.numpy()) and np.array_equal(z0, z1.numpy())
    assert np.array_equal(O.row_blocks(crow, parts), ofs.row_blocks(torch.from_numpy(crow), A.nnz, parts).numpy())


def test_infer_errors_mirror_reference_checks():
    crow = torch.tensor([0, 1, 2], dtype=torch.int32)
    col = torch.tensor([0, 1], dtype=torch.int32)
    val = torch.ones(2)
    b = torch.ones(2, 4)
    ops = __import__("importlib").import_module("of-spmm_b200.ops")
    assert ops.infer_spmm_csr(crow, col, val, b, 2, 2) == ((2, 4), torch.float32)
    with pytest.raises(ofs.OpInferError, match="a_rows\\+1"):
        ops.infer_spmm_csr(crow, col, val, b, 3, 2)
    with pytest.raises(ofs.OpInferError, match="a_cols"):
        ops.infer_spmm_csr(crow, col, val, b, 2, 5)
    with pytest.raises(ofs.OpInferError, match="index dtype"):
        ops.infer_spmm_csr(crow.float(), col, val, b, 2, 2)
    with pytest.raises(ofs.OpInferError, match="no registered kernel"):
        ops.infer_spmm_csr(crow, col, val, b.double(), 2, 2)
    with pytest.raises(ofs.OpInferError, match="bfloat16 a_val"):
        ops.infer_spmm_csr(crow, col, val.bfloat16(), b, 2, 2)
    # exactly one kernel is registered and it is the CUDA one: CPU tensors do not fall back
    with pytest.raises(ofs.OpInferError, match="no kernel registered for device type cpu"):
        ofs.spmm_csr(crow, col, val, b, 2, 2)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libofspmm_b200.so")
    with pytest.raises(ofs.OfspmmLibraryError, match="no CPU fallback"):
        _lib.lib()


def test_generators_are_seeded_and_well_formed():
    for A in (ofs.graphs.uniform_csr(512, 512, 0.02, seed=1), ofs.graphs.reddit_like(256, seed=2),
              ofs.graphs.products_like(512, seed=3), ofs.graphs.rmat_csr(10, 16, seed=4)):
        crow, col = A.crow.long(), A.col.long()
        assert crow[0] == 0 and crow[-1] == A.nnz and (crow[1:] >= crow[:-1]).all()
        rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths())
        key = rows * A.cols + col
        assert (key[1:] > key[:-1]).all()                # sorted + unique within rows
        assert col.min() >= 0 and col.max() < A.cols
    a = ofs.graphs.reddit_like(256, seed=2)
    b = ofs.graphs.reddit_like(256, seed=2)
    assert torch.equal(a.col, b.col) and torch.equal(a.val, b.val)
    full = ofs.graphs.expected_alg_bytes(232965, 232965, 114615892, 128, 4)
    assert abs(full["m2"] / 1e9 - 59.72) < 0.01 and abs(full["m1"] / 1e9 - 1.156) < 0.001


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/ofspmm.h compiles as C11 (-pedantic) and a C program can link the library and call
    the host-side entry points — the boundary carries no C++ / CUDA / torch types."""
    import subprocess
    exe = tmp_path / "abi_host_check"
    libdir = os.path.dirname(ofs.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_host_check.c"), "-o", str(exe),
                    "-L", libdir, "-lofspmm_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_host_check ok" in out.stdout


def test_partition_host_property_based():
    """Hypothesis: for arbitrary row-length profiles (empty rows, hub rows, empty matrices) and
    part counts, the C-ABI host partitioner equals the oracle bit for bit, equals a brute-force
    walk of the merge list, and its split points are monotone and consistent with crow."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(lens=st.lists(st.one_of(st.just(0), st.integers(0, 6), st.integers(0, 300)), min_size=0, max_size=60),
           parts=st.integers(1, 40), wide=st.booleans())
    def check(lens, parts, wide):
        dt = np.int64 if wide else np.int32
        crow = np.concatenate([[0], np.cumsum(lens)]).astype(dt)
        M, nnz = len(lens), int(crow[-1])
        r, z = ofs.merge_path_partition_host(torch.from_numpy(crow), nnz, parts)
        r, z = r.numpy(), z.numpy()
        ro, zo = O.merge_path_partition(crow, parts)
        assert np.array_equal(r, ro) and np.array_equal(z, zo)
        total = M + nnz
        ipw = -(-total // parts) if total else 0
        assert np.array_equal(r + z, np.minimum(np.arange(parts + 1) * ipw, total))
        assert (np.diff(r) >= 0).all() and (np.diff(z) >= 0).all()
        assert r[-1] == M and z[-1] == nnz
        # brute-force merge order: a row-end is consumed as soon as all its non-zeros are
        i = j = 0
        order = [(0, 0)]
        while i < M or j < nnz:
            if i < M and crow[i + 1] <= j:
                i += 1
            else:
                j += 1
            order.append((i, j))
        for k in range(parts + 1):
            assert (r[k], z[k]) == order[min(k * ipw, total)]
    check()


# ------------------------------------------------------------------ histogram -> kernel variant (host function)

def _hist_of(lengths):
    """ofspmm_row_hist's bucketing on the host: bucket 0 = empty rows, bucket b = lengths in [2^(b-1), 2^b)."""
    h = np.zeros(32, dtype=np.int64)
    lengths = np.asarray(lengths, dtype=np.int64)
    b = np.where(lengths <= 0, 0, np.floor(np.log2(np.maximum(lengths, 1))).astype(np.int64) + 1)
    np.add.at(h, np.minimum(b, 31), 1)
    return h


@pytest.mark.parametrize("name,rows,lengths,n,dd,want", [
    # BASELINE configs[0]: 4096 rows of ~41 non-zeros, n = 64 fp32 -> one launch, whole rows
    ("cfg1", 4096, lambda r: r.binomial(4096, 0.01, 4096), 64, "f32", "ROWS"),
    # same rows but one hub of 600 non-zeros: back to the merge path (64-item tasks: still a small problem)
    ("cfg1+hub", 4096, lambda r: np.concatenate([r.binomial(4096, 0.01, 4095), [600]]), 64, "f32", "ITEMS64"),
    # too few rows to fill the machine with one lane group per row
    ("few rows", 300, lambda r: r.binomial(4096, 0.05, 300), 64, "f32", "ITEMS64"),
    # fp32 rows wider than 512 bytes have no whole-row kernel
    ("wide", 4096, lambda r: r.binomial(4096, 0.01, 4096), 256, "f32", "ITEMS64"),
    # Reddit-shaped (median ~ 400), n = 128: base family
    ("cfg2-like", 200_000, lambda r: r.integers(100, 900, 200_000), 128, "f32", "BASE"),
    # R-MAT-like (most rows under 16 non-zeros), narrow operand: row-parallel groups
    ("rmat n32", 1_000_000, lambda r: r.geometric(0.2, 1_000_000) - 1, 32, "f32", "ROWPAR"),
    # same rows, warp-wide operand: the row-parallel family has no kernel -> base
    ("rmat n128", 1_000_000, lambda r: r.geometric(0.2, 1_000_000) - 1, 128, "f32", "BASE"),
    # products-shaped (median ~ 50), n = 64 bf16: base wins (profiles/r2_variant_sweeps.md)
    ("products n64 bf16", 150_000, lambda r: r.integers(20, 90, 150_000), 64, "bf16", "BASE"),
])
def test_choose_variant_from_histogram(name, rows, lengths, n, dd, want):
    from importlib import import_module
    _lib = import_module("of-spmm_b200._lib")
    L = _lib.lib()
    lens = lengths(np.random.default_rng(5))
    assert len(lens) == rows
    hist = _hist_of(lens)
    arr = (ctypes.c_int64 * 32)(*hist.tolist())
    dense = _lib.DTYPE_FLOAT if dd == "f32" else _lib.DTYPE_BFLOAT16
    code = L.ofspmm_choose_variant(arr, rows, int(lens.sum()), n, dense)
    assert code & _lib.VARIANT_EXPLICIT
    fam = ("ROWS" if code & _lib.VARIANT_ROWS else "ITEMS64" if code & _lib.VARIANT_ITEMS64
           else "ROWPAR" if code & _lib.VARIANT_ROWPAR else "BASE")
    assert fam == want, (name, hex(code), L.ofspmm_variant_name(code, rows, int(lens.sum()), n, dense).decode())
    # without a histogram only the problem size decides
    auto = L.ofspmm_choose_variant(None, rows, int(lens.sum()), n, dense)
    assert not (auto & (_lib.VARIANT_ROWS | _lib.VARIANT_ROWPAR))
