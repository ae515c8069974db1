"""Functional layer + autograd wiring — the mirror of SURVEY.md §8 rows a8 / a9.

``spmm_csr(a_crow, a_col, a_val, b, a_rows, a_cols)`` is the call a OneFlow user would make as
``flow._C.spmm_csr(...)`` (pattern: oneflow/core/functional/functional_api.yaml:1011-1015,
oneflow/core/functional/impl/nn_functor.cpp:290-323).  Gradients follow the grad-function pattern of
oneflow/core/autograd/gradient_funcs/matmul.cpp:36-104 and gather.cpp:29-72:

    Capture: remember which inputs need grad, save crow/col/val/b as needed
    Apply:   in_grads[a_val] = sddmm_csr(crow, col, dy, b)        iff a_val.requires_grad
             in_grads[b]     = spmm_csr_grad_b(crow, col, val, dy) iff b.requires_grad
             index inputs never get a gradient (ModifyInputArg: set_requires_grad(false),
             oneflow/user/ops/unsorted_segment_sum_op.cpp:71-78)
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops


class SpmmOpKernelState:
    """Per-op persistent data — the role OpKernelState plays in the reference
    (oneflow/user/kernels/stateful_opkernel.cpp:919-928): a CSR of A^T built once on the device and
    keyed by the identity of the CSR arrays, so `spmm_csr_grad_b` runs the deterministic,
    atomic-free route."""

    def __init__(self) -> None:
        self._cache: Dict[Tuple[int, int, int, int, int, int], Tuple[torch.Tensor, ...]] = {}
        self._keep = {}

    @staticmethod
    def _key(a_crow, a_col, a_val, a_rows, a_cols):
        return (a_crow.data_ptr(), a_col.data_ptr(), a_val.data_ptr(), a_val._version, a_rows, a_cols)

    def transposed(self, a_crow, a_col, a_val, a_rows: int, a_cols: int):
        k = self._key(a_crow, a_col, a_val, a_rows, a_cols)
        hit = self._cache.get(k)
        if hit is None:
            hit = ops.csr_transpose(a_crow, a_col, a_val.detach(), a_rows, a_cols)
            self._cache = {k: hit}                      # one entry: the op sees one graph at a time
            self._keep = (a_crow, a_col, a_val)         # pin the key tensors so pointers stay unique
        return hit

    def clear(self) -> None:
        self._cache.clear()
        self._keep = {}


class _SpmmCsrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_crow, a_col, a_val, b, a_rows, a_cols, state):
        out = ops.spmm_csr_compute(a_crow, a_col, a_val.detach(), b.detach(), a_rows, a_cols)
        # Capture (matmul.cpp:48-73): save only what Apply will read
        ctx.val_requires_grad = a_val.requires_grad
        ctx.b_requires_grad = b.requires_grad
        ctx.a_rows, ctx.a_cols, ctx.state = a_rows, a_cols, state
        ctx.save_for_backward(a_crow, a_col, a_val, b if a_val.requires_grad else None)
        ctx.mark_non_differentiable(a_crow, a_col)
        return out

    @staticmethod
    def backward(ctx, dy):
        a_crow, a_col, a_val, b = ctx.saved_tensors
        dy = dy.contiguous()
        d_val = d_b = None
        if ctx.val_requires_grad:
            d_val = ops.sddmm_csr_compute(a_crow, a_col, dy, b, ctx.a_rows, ctx.a_cols, a_val.dtype)
        if ctx.b_requires_grad:
            tr = None
            if ctx.state is not None:
                tr = ctx.state.transposed(a_crow, a_col, a_val, ctx.a_rows, ctx.a_cols)
            d_b = ops.spmm_csr_grad_b_compute(a_crow, a_col, a_val.detach(), dy, ctx.a_rows, ctx.a_cols, tr)
        return None, None, d_val, d_b, None, None, None


def spmm_csr(a_crow: torch.Tensor, a_col: torch.Tensor, a_val: torch.Tensor, b: torch.Tensor,
             a_rows: int, a_cols: int, state: Optional[SpmmOpKernelState] = None) -> torch.Tensor:
    """out[a_rows, n] = CSR(a_crow, a_col, a_val; a_rows x a_cols) @ b[a_cols, n].

    Differentiable wrt ``a_val`` (SDDMM) and ``b`` (A^T·dy).  ``state`` (optional) caches the device
    transpose for the deterministic backward; without it the backward uses the atomic route."""
    if not (torch.is_grad_enabled() and (a_val.requires_grad or b.requires_grad)):
        return ops.spmm_csr_compute(a_crow, a_col, a_val, b, a_rows, a_cols)
    ops._check_device(a_crow, a_col, a_val, b)
    ops.infer_spmm_csr(a_crow, a_col, a_val, b, a_rows, a_cols)
    return _SpmmCsrFn.apply(a_crow, a_col, a_val, b, a_rows, a_cols, state)


def spmm_csr_grad_b(a_crow, a_col, a_val, dy, a_rows: int, a_cols: int, transposed=None) -> torch.Tensor:
    """db[a_cols, n] = A^T @ dy (the op the grad function dispatches for ``b``)."""
    return ops.spmm_csr_grad_b_compute(a_crow, a_col, a_val, dy, a_rows, a_cols, transposed)


def sddmm_csr(a_crow, a_col, dy, b, a_rows: int, a_cols: int, val_dtype=torch.float32) -> torch.Tensor:
    """dval[p] = <dy[row(p), :], b[a_col[p], :]> (the op the grad function dispatches for ``a_val``)."""
    return ops.sddmm_csr_compute(a_crow, a_col, dy, b, a_rows, a_cols, val_dtype)
