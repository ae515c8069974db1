"""Multi-GPU SpMM inside one NVSwitch box: one process per GPU over
``torch.distributed`` (NCCL over NVLink 5; gloo on CPU for the host-logic tests).

Partitioning
  * A (M×K) → ``world`` contiguous, **nnz-balanced whole-row blocks** from the device merge-path
    partitioner (``row_blocks``); rank r keeps only its block's CSR.  Output rows stay local.
    Contrast: the reference can only split rows in equal counts (BalancedSplitter,
    oneflow/core/common/balanced_splitter.cpp:20-39) and its S(0)→B boxing needs dim0 % world == 0
    (oneflow/core/boxing/ccl_boxing_function.cpp:115), which is why this lives here and not in SBP.
  * B (K×n) is row-sharded in equal blocks of ceil(K/world) rows (zero-padded), which is what an
    all-gather needs.

Collectives and overlap (measured on 8 B200s, profiles/r1_multigpu.md)
  * ``forward``: all-gather of the B shards, then ``C_blk = A_blk · B``.  ``backward``: partial
    ``A_blkᵀ·dY_blk`` (K×n), then reduce-scatter — the dual collective.
  * ``step`` (both products of a training step) hides the collectives of one product behind the
    compute of the other: all-gather(B) ‖ A_blkᵀ·dY, then reduce-scatter(dB) ‖ A_blk·B.  The
    reference instead finishes a blocking, unfused all-gather before the op is even issued
    (oneflow/core/framework/op_interpreter/eager_global_op_interpreter.cpp:156-181).
  * ``panels`` > 1 additionally pipelines a collective with its *own* product over column panels of
    the dense operand (``ofspmm_fwd_strided`` writes panel j straight into C_blk's columns).
    Measured slower than whole-width products on cfg2, so the default is 1.
  * NCCL kernels need SMs; a persistent compute grid owns all of them, so by default the compute
    kernels run as short-lived CTAs while collectives are in flight (``tasks_per_warp``), or the
    exchange uses copy engines over symmetric peer memory (``comm="peer"``).

The compute callbacks default to the CUDA ops; tests inject CPU stand-ins to exercise the
partition / shard / pipeline logic under gloo with world_size 2.
"""
from __future__ import annotations

from typing import Callable, List, Optional  # noqa: F401

import torch
import torch.distributed as dist

from . import ops
from .graphs import CsrMatrix


def _default_spmm(crow, col, val, b, rows, cols, out):
    return ops.spmm_csr_compute(crow, col, val, b, rows, cols, out=out)


def _default_transpose(crow, col, val, rows, cols):
    return ops.csr_transpose(crow, col, val, rows, cols)


def shard_rows_count(k: int, world: int) -> int:
    return (k + world - 1) // world


class _PeerExchange:
    """Collectives over NVSwitch peer memory without SMs: every rank exposes its send buffers
    through CUDA symmetric memory (``torch.distributed._symmetric_memory``), synchronises with a
    stream-ordered device barrier and *pulls* the peers' blocks with plain device-to-device copies
    on a side stream — those run on the copy engines, so the persistent SpMM kernel keeps all 148
    SMs while the next panel is in flight.  (NCCL's all-gather kernel needs SMs of its own; when
    the persistent compute grid already owns them the two serialise instead of overlapping.)"""

    def __init__(self, shard: int, kp: int, w: int, panels: int, dtype, device, rank: int, world: int, group):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world, self.shard, self.w, self.panels = rank, world, shard, w, panels
        gname = (group or dist.group.WORLD).group_name
        # outgoing B panels (forward) and partial dB panels (backward), visible to every peer
        self.send = symm.empty((panels, shard, w), dtype=dtype, device=device)
        self.part = symm.empty((panels, kp, w), dtype=dtype, device=device)
        self.h_send = symm.rendezvous(self.send, gname)
        self.h_part = symm.rendezvous(self.part, gname)
        self.peer_send = [self.h_send.get_buffer(r, (panels, shard, w), dtype) for r in range(world)]
        self.peer_part = [self.h_part.get_buffer(r, (panels, kp, w), dtype) for r in range(world)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.stage = torch.empty((world, shard, w), dtype=dtype, device=device)

    def begin_forward(self):
        # every peer has finished pulling the previous step's panels before they are overwritten
        self.h_send.barrier(channel=0)

    def all_gather_panel(self, j: int, out_full: torch.Tensor) -> torch.cuda.Event:
        """send[j] was filled on the current stream; returns the event after which out_full
        (kp x w) holds every rank's panel j."""
        self.h_send.barrier(channel=1 + j)          # all ranks have filled their panel j
        cur = torch.cuda.current_stream()
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            for step in range(self.world):
                r = (self.rank + step) % self.world   # own block first, then a ring of peers
                out_full[r * self.shard:(r + 1) * self.shard].copy_(self.peer_send[r][j], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev

    def begin_backward(self):
        self.h_part.barrier(channel=0)

    def reduce_scatter_panel(self, j: int, out_shard: torch.Tensor) -> torch.cuda.Event:
        """part[j] (this rank's partial kp x w) was written on the current stream; out_shard
        (shard x w) = sum over ranks of their partial rows of this rank's shard, fixed order."""
        self.h_part.barrier(channel=1 + j)          # every rank's partial panel j is complete
        cur = torch.cuda.current_stream()
        self.copy_stream.wait_stream(cur)
        lo, hi = self.rank * self.shard, (self.rank + 1) * self.shard
        with torch.cuda.stream(self.copy_stream):
            for step in range(self.world):
                r = (self.rank + step) % self.world
                self.stage[r].copy_(self.peer_part[r][j, lo:hi], non_blocking=True)
            torch.sum(self.stage, dim=0, out=out_shard)   # rank order 0..world-1: deterministic
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev


class ShardedSpmm:
    """Row-block-partitioned SpMM operator bound to one rank."""

    def __init__(self, A: CsrMatrix, n: int, dtype: torch.dtype, rank: int, world: int, device,
                 bwd: str = "transpose", panels: int = 1,
                 spmm_fn: Callable = _default_spmm, transpose_fn: Callable = _default_transpose,
                 group=None, comm: str = "nccl", tasks_per_warp: int = 2):
        assert A.rows >= world, "fewer rows than ranks"
        # Launch policy of the compute kernels while collectives are in flight: CTAs that retire
        # after ~tasks_per_warp tasks per warp instead of one persistent wave, so the NCCL kernels
        # (higher-priority stream) get SMs as soon as they are ready.  8 B200s, cfg2: 1.19 ms →
        # 0.98 ms per step (profiles/r1_multigpu.md).  The library reads the knob at every launch.
        if world > 1 and tasks_per_warp > 0 and comm == "nccl":
            import os
            os.environ.setdefault("OFSPMM_TASKS_PER_WARP", str(tasks_per_warp))
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.n, self.dtype = n, dtype
        self.rows, self.cols = A.rows, A.cols
        self.spmm_fn, self.bwd_mode = spmm_fn, bwd
        # column panels must keep 16-byte alignment of every panel start (8 bf16 / 4 fp32)
        vec = 8 if dtype == torch.bfloat16 else 4
        panels = max(1, min(panels, n // vec if n >= vec else 1))
        while n % (panels * vec) != 0 and panels > 1:
            panels -= 1
        self.panels = panels
        self.w = n // panels
        # nnz-balanced whole-row blocks (device partitioner when the graph is on the GPU)
        self.bounds = ops.row_blocks(A.crow, A.nnz, world).cpu().tolist()
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.A_blk = A.row_slice(self.r0, self.r1)
        self.shard = shard_rows_count(A.cols, world)
        self.kp = self.shard * world  # padded K
        self.At_blk = None
        if bwd == "transpose":
            t = transpose_fn(self.A_blk.crow, self.A_blk.col, self.A_blk.val, self.A_blk.rows, self.A_blk.cols)
            self.At_blk = CsrMatrix(t[0], t[1], t[2], self.A_blk.cols, self.A_blk.rows)
        m = self.A_blk.rows
        # persistent buffers (allocated once: the op state of the multi-GPU path)
        self._b_send = [torch.empty((self.shard, self.w), dtype=dtype, device=device) for _ in range(panels)]
        self._b_full = [torch.empty((self.kp, self.w), dtype=dtype, device=device) for _ in range(panels)]
        self._c = torch.empty((m, n), dtype=dtype, device=device)
        self._db_part = [torch.empty((self.kp, self.w), dtype=dtype, device=device) for _ in range(panels)]
        self._db_out = [torch.empty((self.shard, self.w), dtype=dtype, device=device) for _ in range(panels)]
        self._db = torch.empty((self.shard, n), dtype=dtype, device=device)
        self._nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        # comm = "peer": copy-engine pulls over symmetric memory instead of NCCL kernels
        self.comm, self._peer = "nccl" if self._nccl else "gloo", None
        if comm == "peer" and self._nccl:
            try:
                self._peer = _PeerExchange(self.shard, self.kp, self.w, panels, dtype, device, rank, world, group)
                self.comm = "peer"
            except Exception as e:  # pragma: no cover - needs NVLink peers
                import warnings
                warnings.warn(f"symmetric-memory peer exchange unavailable ({e}); using NCCL collectives")

    # ------------------------------------------------------------------ sharding helpers
    def shard_rows(self, B_full: torch.Tensor) -> torch.Tensor:
        """This rank's equal-size row shard of a K×n dense operand (zero-padded past K)."""
        s0 = self.rank * self.shard
        out = torch.zeros((self.shard, self.n), dtype=B_full.dtype, device=B_full.device)
        hi = min(self.cols, s0 + self.shard)
        if hi > s0:
            out[: hi - s0] = B_full[s0:hi]
        return out

    def shard_rows_out(self, dY_full: torch.Tensor) -> torch.Tensor:
        """The rows of an M×n tensor that belong to this rank's row block of A."""
        return dY_full[self.r0:self.r1].contiguous()

    # ------------------------------------------------------------------ collectives
    def _all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        return dist.all_gather_into_tensor(out, inp, group=self.group, async_op=True)

    def _reduce_scatter(self, out: torch.Tensor, inp: torch.Tensor):
        if self._nccl:
            return dist.reduce_scatter_tensor(out, inp, group=self.group, async_op=True)
        # gloo has no reduce-scatter: all-reduce then keep the local shard (host-logic tests only)
        dist.all_reduce(inp, group=self.group)
        out.copy_(inp[self.rank * self.shard:(self.rank + 1) * self.shard])
        return None

    # ------------------------------------------------------------------ forward
    def forward(self, B_shard: torch.Tensor) -> torch.Tensor:
        """C_blk[m, n] = A_blk · allgather(B_shard), panel-pipelined."""
        A, w = self.A_blk, self.w
        if self._peer is not None:
            px = self._peer
            px.begin_forward()
            events = []
            for j in range(self.panels):
                px.send[j].copy_(B_shard[:, j * w:(j + 1) * w])
                events.append(px.all_gather_panel(j, self._b_full[j]))
            for j in range(self.panels):
                torch.cuda.current_stream().wait_event(events[j])
                self.spmm_fn(A.crow, A.col, A.val, self._b_full[j][: self.cols], A.rows, A.cols,
                             self._c[:, j * w:(j + 1) * w])
            return self._c
        works = []
        for j in range(self.panels):
            self._b_send[j].copy_(B_shard[:, j * w:(j + 1) * w])
            works.append(self._all_gather(self._b_full[j], self._b_send[j]))
        for j in range(self.panels):
            works[j].wait()  # compute stream waits for panel j only; later panels keep flowing
            self.spmm_fn(A.crow, A.col, A.val, self._b_full[j][: self.cols], A.rows, A.cols,
                         self._c[:, j * w:(j + 1) * w])
        return self._c

    # ------------------------------------------------------------------ backward wrt B
    def backward(self, dY_blk: torch.Tensor) -> torch.Tensor:
        """dB_shard[shard, n] = reduce_scatter(A_blkᵀ · dY_blk), panel-pipelined."""
        w = self.w
        if self._peer is not None and self.At_blk is not None:
            px, At = self._peer, self.At_blk
            px.begin_backward()
            events = []
            for j in range(self.panels):
                part = px.part[j]
                if self.kp > self.cols:
                    part[self.cols:].zero_()
                self.spmm_fn(At.crow, At.col, At.val, dY_blk[:, j * w:(j + 1) * w], At.rows, At.cols, part[: self.cols])
                events.append(px.reduce_scatter_panel(j, self._db_out[j]))
            for j in range(self.panels):
                torch.cuda.current_stream().wait_event(events[j])
                self._db[:, j * w:(j + 1) * w].copy_(self._db_out[j])
            return self._db
        works: List[Optional[object]] = []
        for j in range(self.panels):
            part = self._db_part[j]
            if self.kp > self.cols:
                part[self.cols:].zero_()
            dyj = dY_blk[:, j * w:(j + 1) * w]
            if self.At_blk is not None:
                At = self.At_blk
                self.spmm_fn(At.crow, At.col, At.val, dyj, At.rows, At.cols, part[: self.cols])
            else:
                A = self.A_blk
                part[: self.cols].copy_(ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dyj.contiguous(),
                                                                   A.rows, A.cols))
            works.append(self._reduce_scatter(self._db_out[j], part))
        for j in range(self.panels):
            if works[j] is not None:
                works[j].wait()
            self._db[:, j * w:(j + 1) * w].copy_(self._db_out[j])
        return self._db

    def step(self, B_shard: torch.Tensor, dY_blk: torch.Tensor, overlap: bool = True):
        """One benchmark step: C_blk = A_blk·allgather(B) and dB_shard = reduce_scatter(A_blkᵀ·dY).
        The two halves are independent, so with ``overlap`` the collectives of one hide behind the
        compute of the other:  all-gather(B) ‖ A_blkᵀ·dY,  then  reduce-scatter(dB) ‖ A_blk·B.
        (Measured on 8 B200s, cfg2: splitting the dense width into column panels to overlap inside
        one product costs more in narrower gathers than the overlap returns; whole-width products
        with cross-product overlap are faster — profiles/r1_multigpu.md.)"""
        if not overlap or self.At_blk is None or self.panels != 1:
            return self.forward(B_shard), self.backward(dY_blk)
        A, At, n = self.A_blk, self.At_blk, self.n
        if self._peer is not None:
            px = self._peer
            cur = torch.cuda.current_stream()
            px.begin_forward()
            px.send[0].copy_(B_shard)
            ev_ag = px.all_gather_panel(0, self._b_full[0])           # copy engines, side stream
            px.begin_backward()
            part = px.part[0]
            if self.kp > self.cols:
                part[self.cols:].zero_()
            self.spmm_fn(At.crow, At.col, At.val, dY_blk, At.rows, At.cols, part[: self.cols])
            ev_rs = px.reduce_scatter_panel(0, self._db_out[0])      # pulls + ordered sum, side stream
            cur.wait_event(ev_ag)
            self.spmm_fn(A.crow, A.col, A.val, self._b_full[0][: self.cols], A.rows, A.cols, self._c)
            cur.wait_event(ev_rs)
            self._db.copy_(self._db_out[0])
            return self._c, self._db
        self._b_send[0].copy_(B_shard)
        w_ag = self._all_gather(self._b_full[0], self._b_send[0])     # NCCL stream
        part = self._db_part[0]
        if self.kp > self.cols:
            part[self.cols:].zero_()
        self.spmm_fn(At.crow, At.col, At.val, dY_blk, At.rows, At.cols, part[: self.cols])
        w_rs = self._reduce_scatter(self._db_out[0], part)
        w_ag.wait()
        self.spmm_fn(A.crow, A.col, A.val, self._b_full[0][: self.cols], A.rows, A.cols, self._c)
        if w_rs is not None:
            w_rs.wait()
        self._db.copy_(self._db_out[0])
        return self._c, self._db

    def capture(self, fn: Callable[[], object], warmup: int = 3):
        """Capture ``fn`` (a closure over static input tensors, e.g. ``lambda: self.step(B, dY)``)
        into a CUDA graph — kernels, copies and the NCCL collectives — and return the replay
        callable.  A step at 8 GPUs is ~30 short launches; replaying one graph removes the host
        launch gaps between them (the reference's lazy mode does the same per kernel,
        oneflow/core/kernel/user_kernel.cpp:689-714).  Falls back to eager ``fn`` if capture is
        not possible (e.g. gloo / CPU)."""
        if not (torch.cuda.is_available() and str(self.device).startswith("cuda")):
            return fn
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fn()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
            torch.cuda.synchronize(self.device)
            return graph.replay
        except Exception as e:  # pragma: no cover - depends on the NCCL / driver combination
            import warnings
            warnings.warn(f"CUDA-graph capture of the sharded step failed ({e}); running eagerly")
            torch.cuda.synchronize(self.device)
            return fn
