"""Host-side mirror of the OneFlow user-op pieces of SURVEY.md §8a (a2, a7): shape / dtype
inference with the reference's error behaviour, and the kernel-compute bodies that pull raw
pointers out of tensors and call the C ABI on the current stream.

Tensors are ``torch`` CUDA tensors — torch is only the memory / stream plumbing here; the maths is
all in libofspmm_b200.so.  Names follow the provisional op schema (`spmm_csr`,
`spmm_csr_grad_b`, `sddmm_csr`; inputs `a_crow, a_col, a_val, b`; attrs `a_rows, a_cols`).

Reference conventions mirrored:
  * InferLogicalTensorDesc / InferDataType with CHECK_*_OR_RETURN → Python exception
    (oneflow/user/ops/matmul_op.cpp:23-75, oneflow/user/ops/unsorted_segment_sum_op.cpp:66-78);
  * kernel selection by (device, dense dtype, index dtype); exactly one kernel may match
    (oneflow/core/framework/user_op_registry_manager.cpp:93-117) — here: CUDA only, no CPU kernel;
  * tmp_buffer sized by an InferTmpSizeFn (oneflow/user/kernels/unsorted_segment_sum_kernel.cpp:191-202)
    → the ``*_workspace_bytes`` queries, allocated from torch's stream-ordered caching allocator;
  * outputs are allocated uninitialised and fully overwritten by the kernel.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CsrStruct, check

_DENSE = {torch.float32: _lib.DTYPE_FLOAT, torch.bfloat16: _lib.DTYPE_BFLOAT16}
_INDEX = {torch.int32: _lib.DTYPE_INT32, torch.int64: _lib.DTYPE_INT64}


class OpInferError(RuntimeError):
    """What a failed CHECK_*_OR_RETURN in an op's Infer* function surfaces as in Python."""


def _chk(cond: bool, msg: str) -> None:
    if not cond:
        raise OpInferError(msg)


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None if t is None else (t.data_ptr() or None)
    return t.data_ptr()


def _workspace(nbytes: int, device) -> Tuple[Optional[torch.Tensor], Optional[int]]:
    if nbytes == 0:
        return None, None
    w = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return w, w.data_ptr()


def infer_spmm_csr(a_crow, a_col, a_val, b, a_rows: int, a_cols: int) -> Tuple[Tuple[int, int], torch.dtype]:
    """SpmmCsrOp::InferLogicalTensorDesc + InferDataType."""
    _chk(a_crow.dim() == 1 and a_col.dim() == 1, "a_crow and a_col must be 1-D")
    _chk(b.dim() == 2, f"b must be 2-D (cols x n), got {b.dim()}-D")
    _chk(a_rows >= 0 and a_cols >= 0, "a_rows / a_cols must be non-negative")
    _chk(a_crow.numel() == a_rows + 1, f"a_crow must have a_rows+1 = {a_rows + 1} entries, got {a_crow.numel()}")
    _chk(b.shape[0] == a_cols, f"b has {b.shape[0]} rows but a_cols = {a_cols}")
    _chk(a_crow.dtype in _INDEX, f"a_crow must be an index dtype (int32/int64), got {a_crow.dtype}")
    _chk(a_col.dtype == a_crow.dtype, "a_col and a_crow must share one index dtype")
    _chk(b.dtype in _DENSE, f"b dtype {b.dtype} has no registered kernel (float32 / bfloat16)")
    if a_val is not None:
        _chk(a_val.dim() == 1 and a_val.numel() == a_col.numel(), "a_val and a_col must have nnz entries each")
        _chk(a_val.dtype in _DENSE, f"a_val dtype {a_val.dtype} unsupported")
        _chk(a_val.dtype == torch.float32 or b.dtype == torch.bfloat16,
             "bfloat16 a_val requires a bfloat16 dense operand")
    return (a_rows, int(b.shape[1])), b.dtype


def _check_device(*tensors) -> None:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            # exactly one kernel is registered for this op and it is the CUDA one (no CPU kernel):
            raise OpInferError("spmm_csr: no kernel registered for device type cpu — tensors must be on a CUDA device")
        # 1-D tensors contiguous; dense 2-D operands row-major with unit inner stride (a column
        # slice of a wider matrix is fine: its row stride is passed on as the leading dimension)
        ok = t.is_contiguous() or (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1])
        _chk(ok, "spmm_csr kernels take contiguous tensors (2-D operands: unit inner stride)")
        dev = dev or t.device
        _chk(t.device == dev, "all tensors must live on one device")


def _csr_struct(a_crow, a_col, a_val, rows: int, cols: int, val_dtype=None) -> CsrStruct:
    vd = _DENSE[a_val.dtype] if a_val is not None else _DENSE[val_dtype or torch.float32]
    return CsrStruct(rows, cols, int(a_col.numel()), _ptr(a_crow), _ptr(a_col), _ptr(a_val),
                     _INDEX[a_crow.dtype], vd)


def spmm_csr_compute(a_crow, a_col, a_val, b, a_rows: int, a_cols: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SpmmCsrKernel::Compute — out[a_rows, n] = A · b."""
    _check_device(a_crow, a_col, a_val, b)
    (m, n), dt = infer_spmm_csr(a_crow, a_col, a_val, b, a_rows, a_cols)
    if out is None:
        out = torch.empty((m, n), dtype=dt, device=b.device)
    else:
        _chk(out.shape == (m, n) and out.dtype == dt and out.device == b.device,
             "out has the wrong shape / dtype / device")
        _check_device(out)
    L = _lib.lib()
    with torch.cuda.device(b.device):
        A = _csr_struct(a_crow, a_col, a_val, a_rows, a_cols)
        nbytes = L.ofspmm_fwd_workspace_bytes(a_rows, a_cols, A.nnz, n, _DENSE[dt])
        ws, wsp = _workspace(nbytes, b.device)
        ldb = b.stride(0) if b.shape[0] > 1 else max(n, 1)
        ldc = out.stride(0) if out.shape[0] > 1 else max(n, 1)
        if ldb == n and ldc == n:
            rc = L.ofspmm_fwd(ctypes.byref(A), _ptr(b), _ptr(out), n, _DENSE[dt], wsp, nbytes, _stream_ptr(b))
        else:
            rc = L.ofspmm_fwd_strided(ctypes.byref(A), _ptr(b), ldb, _ptr(out), ldc, n, _DENSE[dt], wsp, nbytes,
                                      _stream_ptr(b))
        check(rc, "spmm_csr")
    return out


def spmm_csr_grad_b_compute(a_crow, a_col, a_val, dy, a_rows: int, a_cols: int,
                            transposed: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SpmmCsrGradBKernel::Compute — db[a_cols, n] = A^T · dy.  ``transposed`` = (t_crow, t_col,
    t_val) from csr_transpose (kept in the op state) selects the deterministic route."""
    _check_device(a_crow, a_col, a_val, dy)
    _chk(dy.dim() == 2 and dy.shape[0] == a_rows, f"dy must be (a_rows={a_rows}) x n")
    infer_spmm_csr(a_crow, a_col, a_val, dy.new_empty((a_cols, dy.shape[1])), a_rows, a_cols)
    n, dt = int(dy.shape[1]), dy.dtype
    if out is None:
        out = torch.empty((a_cols, n), dtype=dt, device=dy.device)
    L = _lib.lib()
    with torch.cuda.device(dy.device):
        A = _csr_struct(a_crow, a_col, a_val, a_rows, a_cols)
        At_ref = None
        if transposed is not None:
            t_crow, t_col, t_val = transposed
            _check_device(t_crow, t_col, t_val)
            At = _csr_struct(t_crow, t_col, t_val, a_cols, a_rows)
            At_ref = ctypes.byref(At)
        nbytes = L.ofspmm_bwd_b_workspace_bytes(a_rows, a_cols, A.nnz, n, _DENSE[dt], 1 if transposed is not None else 0)
        ws, wsp = _workspace(nbytes, dy.device)
        check(L.ofspmm_bwd_b(ctypes.byref(A), At_ref, _ptr(dy), _ptr(out), n, _DENSE[dt], wsp, nbytes,
                             _stream_ptr(dy)), "spmm_csr_grad_b")
    return out


def sddmm_csr_compute(a_crow, a_col, dy, b, a_rows: int, a_cols: int,
                      val_dtype: torch.dtype = torch.float32,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SddmmCsrKernel::Compute — dval[p] = <dy[i,:], b[col[p],:]>."""
    _check_device(a_crow, a_col, dy, b)
    infer_spmm_csr(a_crow, a_col, None, b, a_rows, a_cols)
    _chk(dy.dim() == 2 and dy.shape[0] == a_rows and dy.shape[1] == b.shape[1] and dy.dtype == b.dtype,
         "dy must be (a_rows x n) with b's dtype")
    _chk(val_dtype == torch.float32 or b.dtype == torch.bfloat16, "bfloat16 values require a bfloat16 dense operand")
    n, dt = int(b.shape[1]), b.dtype
    nnz = int(a_col.numel())
    if out is None:
        out = torch.empty((nnz,), dtype=val_dtype, device=b.device)
    L = _lib.lib()
    with torch.cuda.device(b.device):
        A = _csr_struct(a_crow, a_col, None, a_rows, a_cols, val_dtype)
        nbytes = L.ofspmm_sddmm_workspace_bytes(a_rows, a_cols, nnz, n, _DENSE[dt])
        ws, wsp = _workspace(nbytes, b.device)
        check(L.ofspmm_sddmm(ctypes.byref(A), _ptr(dy), _ptr(b), _ptr(out), n, _DENSE[dt], wsp, nbytes,
                             _stream_ptr(b)), "sddmm_csr")
    return out


def merge_path_partition(a_crow: torch.Tensor, nnz: int, parts: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device merge-path partitioner: parts+1 (row, nz) split points as int64 tensors."""
    _check_device(a_crow)
    _chk(a_crow.dtype in _INDEX and a_crow.dim() == 1 and a_crow.numel() >= 1, "a_crow must be a 1-D index tensor")
    _chk(parts >= 1, "parts must be >= 1")
    rows = a_crow.numel() - 1
    out_row = torch.empty(parts + 1, dtype=torch.int64, device=a_crow.device)
    out_nz = torch.empty(parts + 1, dtype=torch.int64, device=a_crow.device)
    with torch.cuda.device(a_crow.device):
        check(_lib.lib().ofspmm_partition(_ptr(a_crow), _INDEX[a_crow.dtype], rows, nnz, parts,
                                          out_row.data_ptr(), out_nz.data_ptr(), _stream_ptr(a_crow)),
              "merge_path_partition")
    return out_row, out_nz


def merge_path_partition_host(a_crow: torch.Tensor, nnz: int, parts: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Host twin of the device partitioner (C ABI ofspmm_partition_host); CPU tensors."""
    _chk(not a_crow.is_cuda and a_crow.dtype in _INDEX and a_crow.is_contiguous(), "a_crow must be a contiguous CPU index tensor")
    rows = a_crow.numel() - 1
    out_row = torch.empty(parts + 1, dtype=torch.int64)
    out_nz = torch.empty(parts + 1, dtype=torch.int64)
    check(_lib.lib().ofspmm_partition_host(a_crow.data_ptr(), _INDEX[a_crow.dtype], rows, nnz, parts,
                                           out_row.data_ptr(), out_nz.data_ptr()), "merge_path_partition_host")
    return out_row, out_nz


def row_blocks(a_crow: torch.Tensor, nnz: int, parts: int) -> torch.Tensor:
    """Whole-row nnz-balanced row blocks for ``parts`` devices: bounds[k] = the
    merge-path split row of diagonal k, bounds[0] = 0, bounds[parts] = rows."""
    rows, _ = (merge_path_partition if a_crow.is_cuda else merge_path_partition_host)(a_crow, nnz, parts)
    rows = rows.clone()
    rows[0] = 0
    rows[-1] = a_crow.numel() - 1
    return rows


def row_hist(a_crow: torch.Tensor) -> torch.Tensor:
    """32 log2 buckets of the row lengths (device)."""
    _check_device(a_crow)
    hist = torch.empty(32, dtype=torch.int64, device=a_crow.device)
    with torch.cuda.device(a_crow.device):
        check(_lib.lib().ofspmm_row_hist(_ptr(a_crow), _INDEX[a_crow.dtype], a_crow.numel() - 1,
                                         hist.data_ptr(), _stream_ptr(a_crow)), "row_hist")
    return hist


def csr_transpose(a_crow, a_col, a_val, a_rows: int, a_cols: int, want_perm: bool = False):
    """Device CSR → CSR of A^T (entries of a column in ascending row order).  Returns
    (t_crow, t_col, t_val[, t_perm])."""
    _check_device(a_crow, a_col, a_val)
    nnz = int(a_col.numel())
    dev = a_crow.device
    t_crow = torch.empty(a_cols + 1, dtype=a_crow.dtype, device=dev)
    t_col = torch.empty(nnz, dtype=a_crow.dtype, device=dev)
    t_val = torch.empty(nnz, dtype=a_val.dtype, device=dev) if a_val is not None else None
    t_perm = torch.empty(nnz, dtype=a_crow.dtype, device=dev) if want_perm else None
    L = _lib.lib()
    with torch.cuda.device(dev):
        A = _csr_struct(a_crow, a_col, a_val, a_rows, a_cols)
        nbytes = L.ofspmm_csr_transpose_workspace_bytes(a_rows, a_cols, nnz, _INDEX[a_crow.dtype])
        ws, wsp = _workspace(nbytes, dev)
        check(L.ofspmm_csr_transpose(ctypes.byref(A), _ptr(t_crow), _ptr(t_col), _ptr(t_val), _ptr(t_perm),
                                     wsp, nbytes, _stream_ptr(a_crow)), "csr_transpose")
    return (t_crow, t_col, t_val, t_perm) if want_perm else (t_crow, t_col, t_val)


def spmm_csr_grad_b_transient_compute(a_crow, a_col, a_val, dy, a_rows: int, a_cols: int,
                                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """db = A^T · dy through route (3) of the C ABI (``ofspmm_bwd_b_transient``): A^T is built
    inside the workspace for this call only, then the forward kernel runs on it — deterministic,
    for callers that hold no op state.  NOTE (round 1): compiled and exported, not yet exercised on
    a GPU (the round's GPU budget was spent); ``spmm_csr_grad_b_compute`` remains the default."""
    _check_device(a_crow, a_col, a_val, dy)
    _chk(dy.dim() == 2 and dy.shape[0] == a_rows and dy.is_contiguous(), f"dy must be contiguous (a_rows={a_rows}) x n")
    infer_spmm_csr(a_crow, a_col, a_val, dy.new_empty((a_cols, dy.shape[1])), a_rows, a_cols)
    n, dt = int(dy.shape[1]), dy.dtype
    if out is None:
        out = torch.empty((a_cols, n), dtype=dt, device=dy.device)
    L = _lib.lib()
    with torch.cuda.device(dy.device):
        A = _csr_struct(a_crow, a_col, a_val, a_rows, a_cols)
        nbytes = L.ofspmm_bwd_b_transient_workspace_bytes(a_rows, a_cols, A.nnz, n, _DENSE[dt], A.idx_dtype, A.val_dtype)
        ws, wsp = _workspace(nbytes, dy.device)
        check(L.ofspmm_bwd_b_transient(ctypes.byref(A), _ptr(dy), _ptr(out), n, _DENSE[dt], wsp, nbytes,
                                       _stream_ptr(dy)), "spmm_csr_grad_b(transient)")
    return out
