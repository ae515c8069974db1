"""2-layer GCN on top of the SpMM op — the caller of the hot path in BASELINE.json configs[4]
("2-layer GCN forward+backward (SpMM + SDDMM value grad, hidden 256) on a Reddit-shaped synthetic
graph, random-init weights").  SURVEY.md §8f rank 2 ("next" row): a thin harness, not a framework —
dense X·W products are left to torch (cuBLAS, a plain library GEMM), every aggregation Â·H goes
through ``spmm_csr`` and therefore through the C ABI: forward ``ofspmm_fwd``, backward
``ofspmm_bwd_b`` on the cached transpose, and ``ofspmm_sddmm`` when the edge weights need a
gradient.

    H1  = relu( Â · (X · W1) )          aggregate at width `hidden`   (602 → 256: multiply first)
    out =       (Â · H1) · W2           aggregate at width `hidden`   (256 → 41: aggregate first)

so both aggregations run at the dense width BASELINE.md quotes for cfg5 (N = 256).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

from .functional import SpmmOpKernelState, spmm_csr
from .graphs import CsrMatrix


def glorot(fan_in: int, fan_out: int, seed: int, device, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    w = (torch.rand(fan_in, fan_out, generator=g, dtype=torch.float32) * 2 - 1) * bound
    return w.to(device=device, dtype=dtype)


class GCN2:
    """Two GCNConv layers sharing one normalised adjacency Â = (crow, col, val)."""

    def __init__(self, A: CsrMatrix, in_dim: int = 602, hidden: int = 256, out_dim: int = 41, seed: int = 5,
                 edge_weight_grad: bool = True, spmm: Optional[Callable] = None):
        dev = A.crow.device
        self.A = A
        self.val = A.val.detach().clone().requires_grad_(edge_weight_grad)
        self.W1 = glorot(in_dim, hidden, seed, dev).requires_grad_(True)
        self.W2 = glorot(hidden, out_dim, seed + 1, dev).requires_grad_(True)
        self.state = SpmmOpKernelState() if spmm is None else None   # cached Aᵀ for the backward
        self._spmm = spmm

    def parameters(self):
        return [p for p in (self.W1, self.W2, self.val) if p.requires_grad]

    def aggregate(self, H: torch.Tensor) -> torch.Tensor:
        A = self.A
        if self._spmm is not None:
            return self._spmm(A.crow, A.col, self.val, H, A.rows, A.cols)
        return spmm_csr(A.crow, A.col, self.val, H.contiguous(), A.rows, A.cols, self.state)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        h1 = torch.relu(self.aggregate(X @ self.W1))
        return self.aggregate(h1) @ self.W2

    def train_step(self, X: torch.Tensor, labels: torch.Tensor, lr: float = 0.0) -> torch.Tensor:
        """One forward + backward (cross-entropy on all nodes); optional plain SGD update."""
        for p in self.parameters():
            p.grad = None
        loss = torch.nn.functional.cross_entropy(self.forward(X), labels)
        loss.backward()
        if lr > 0:
            with torch.no_grad():
                for p in (self.W1, self.W2):
                    p -= lr * p.grad
        return loss.detach()

    def spmm_flops_per_step(self) -> int:
        """FLOPs of the sparse products in one train_step: 2 aggregations × (forward + Aᵀ·dY)
        [+ SDDMM when edge weights need a gradient], each 2·nnz·hidden."""
        per = 2 * self.A.nnz * self.W1.shape[1]
        return 2 * (2 + (1 if self.val.requires_grad else 0)) * per
