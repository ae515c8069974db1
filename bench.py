#!/usr/bin/env python
"""bench.py — headline benchmark of the SpMM path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A *step* is one pass of the hot path over the workload: SpMM forward C = A·B plus the A^T·dY
backward (BASELINE.json configs[1]: "forward + A^T·dY backward").  Workload at N=1 = configs[1]
(Reddit-shaped synthetic graph, 232 965 nodes, ~114.6 M nnz, dense N=128 fp32).  For N>1 the same
graph is partitioned into nnz-balanced row blocks (one per rank), B / dY row-sharded and
all-gathered with NCCL overlapped with local compute (strong scaling: total work fixed).

Prints ONE JSON line (rank 0).  `value` = GFLOP/s = flops of the whole job / max-over-ranks device
time, inputs resident in HBM.  `e2e` = the same metric through the public op API with pinned HOST
buffers, host↔device copies inside the timed region.  `roofline` is for the dominant kernel
(spmm_merge_kernel, forward) under the gather model M2 of SURVEY.md §8d; `cpu_baseline` is the
oracle's OneFlow-style CPU loop timed on a bounded row sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SpMM GFLOP/s (2*nnz*N/t), fwd + A^T*dY"
UNIT = "GFLOP/s"

WORKLOADS = {
    # name: (generator kwargs, dense width, dtype)
    "cfg2_reddit_n128_fp32": dict(kind="reddit", scale_div=1, n=128, dtype="fp32"),
    "cfg3_products_n256_bf16": dict(kind="products", scale_div=1, n=256, dtype="bf16"),
    "cfg4_rmat24_n128_fp32": dict(kind="rmat", scale=24, n=128, dtype="fp32"),
    "cfg1_uniform4096_n64_fp32": dict(kind="uniform", n=64, dtype="fp32"),
    "twin_reddit16_n128_fp32": dict(kind="reddit", scale_div=16, n=128, dtype="fp32"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _make_graph(spec, device):
    import ofspmm_b200 as ofs
    g = ofs.graphs
    if spec["kind"] == "reddit":
        return g.reddit_like(spec["scale_div"], seed=2, device=device)
    if spec["kind"] == "products":
        return g.products_like(spec["scale_div"], seed=3, device=device)
    if spec["kind"] == "rmat":
        return g.rmat_csr(spec["scale"], 16, seed=4, device=device)
    return g.uniform_csr(4096, 4096, 0.01, seed=1, device=device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def _cpu_baseline(A, B, dY, n, seconds_target=12.0):
    """Oracle (kind "port": the reference has no CPU SpMM kernel to compile) timed on
    a bounded row sample of the same workload: forward with rows split equally over all host cores
    (MultiThreadLoop + BalancedSplitter idiom) + A^T·dY with per-thread accumulators."""
    import numpy as np
    from oracle import oracle as O
    try:
        O.lib(native=True)
        native = True
    except Exception:
        native = False
    cores = os.cpu_count() or 1
    crow_all = A.crow.cpu().numpy()
    M = A.rows
    Bh = B.float().cpu().numpy()
    # probe a small sample to size the timed one
    def run(rows):
        p1 = int(crow_all[rows])
        crow = crow_all[:rows + 1]
        col = A.col[:p1].cpu().numpy()
        val = A.val[:p1].float().cpu().numpy()
        dYh = dY[:rows].float().cpu().numpy()
        t0 = time.perf_counter()
        O.spmm_f32(crow, col, val, Bh, A.cols, threads=cores, native=native)
        O.spmm_t_f32(crow, col, val, dYh, A.cols, threads=cores, native=native)
        return time.perf_counter() - t0, p1
    probe_rows = max(1, min(M, M // 64))
    t_probe, nnz_probe = run(probe_rows)
    rate = nnz_probe / max(t_probe, 1e-6)
    want_nnz = min(int(crow_all[-1]), int(rate * seconds_target))
    rows = int(np.searchsorted(crow_all, want_nnz, side="right")) - 1
    rows = max(probe_rows, min(M, rows))
    t, nnz_s = run(rows)
    flops = 2.0 * 2.0 * nnz_s * n
    # the reference's default CPU_THREADING_RUNTIME is SEQ (CMakeLists.txt:54): also time the plain
    # single-thread loop, on a 1/16 sample so the whole baseline stays within seconds
    rows1 = max(1, rows // 16)
    p1 = int(crow_all[rows1])
    t0 = time.perf_counter()
    O.spmm_f32(crow_all[:rows1 + 1], A.col[:p1].cpu().numpy(), A.val[:p1].float().cpu().numpy(), Bh, A.cols,
               threads=1, native=native)
    O.spmm_t_f32(crow_all[:rows1 + 1], A.col[:p1].cpu().numpy(), A.val[:p1].float().cpu().numpy(),
                 dY[:rows1].float().cpu().numpy(), A.cols, threads=1, native=native)
    t1 = max(time.perf_counter() - t0, 1e-9)
    return {"value": flops / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": 2.0 * 2.0 * p1 * n / t1 / 1e9,
            "sample": f"first {rows} of {M} rows ({nnz_s} nnz), fwd (rows split over {cores} threads) + A^T*dY "
                      f"(per-thread accumulators), {t:.2f} s, oracle built {'-march=native' if native else 'x86-64-v3'}"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference has
    no SpMM kernel and cannot be built offline, so this arm times the oracle
    port (OneFlow CPU-kernel idiom) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import ofspmm_b200 as ofs
    spec = WORKLOADS[args.workload]
    n = spec["n"]
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    A = _make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev)
    dY = ofs.graphs.upstream_grad(A.rows, n, 12, dev)
    per_step = max(2.0, min(12.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = _cpu_baseline(A, B, dY, n, seconds_target=per_step)
        if i >= args.warmup:
            vals.append(base["value"])
    v = statistics.mean(vals)
    base["value"] = v
    nnz = A.nnz
    ms = 2.0 * 2.0 * nnz * n / (v * 1e9) * 1e3
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "rows": A.rows, "cols": A.cols, "nnz": nnz, "n": n,
                   "step": "forward C=A*B + backward dB=A^T*dY", "parallelism": f"{base['cores']} host threads",
                   "note": "ms_per_step extrapolated from the sampled rate to the full workload"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _e2e_pipelined(args, ofs, ops, A, B, dY, n, dtype, dev, flops_step, blocks=8):
    """EXPERIMENTAL (--e2e-mode pipelined; not yet validated on a GPU in round 1): stream the CSR to
    the device in nnz-balanced row blocks so the H2D copy of block i+1 overlaps the forward of block
    i and the D2H of block i-1; A^T·dY runs once the whole CSR and dY have arrived (transient
    transpose route).  Same bytes per step as the simple mode."""
    import torch
    host = {k: v.cpu().pin_memory() for k, v in dict(crow=A.crow, col=A.col, val=A.val, B=B, dY=dY).items()}
    C_h = torch.empty((A.rows, n), dtype=dtype).pin_memory()
    dB_h = torch.empty((A.cols, n), dtype=dtype).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = C_h.numel() * C_h.element_size() + dB_h.numel() * dB_h.element_size()
    d = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    C_d = torch.empty((A.rows, n), dtype=dtype, device=dev)
    dB_d = torch.empty((A.cols, n), dtype=dtype, device=dev)
    bounds = ofs.row_blocks(host["crow"], A.nnz, blocks).tolist()      # host twin of the partitioner
    offs = [int(host["crow"][r]) for r in bounds]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    def e2e_step():
        cur = torch.cuda.current_stream()
        s_in.wait_stream(cur)
        ev_in, ev_c = [], []
        with torch.cuda.stream(s_in):
            d["B"].copy_(host["B"], non_blocking=True)
            for i in range(blocks):
                r0, r1, p0, p1 = bounds[i], bounds[i + 1], offs[i], offs[i + 1]
                d["crow"][r0:r1 + 1].copy_(host["crow"][r0:r1 + 1], non_blocking=True)
                d["col"][p0:p1].copy_(host["col"][p0:p1], non_blocking=True)
                d["val"][p0:p1].copy_(host["val"][p0:p1], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                ev_in.append(e)
            d["dY"].copy_(host["dY"], non_blocking=True)
            ev_dy = torch.cuda.Event()
            ev_dy.record(s_in)
        for i in range(blocks):
            r0, r1, p0, p1 = bounds[i], bounds[i + 1], offs[i], offs[i + 1]
            cur.wait_event(ev_in[i])
            if r1 > r0:
                crow_blk = d["crow"][r0:r1 + 1] - d["crow"][r0:r0 + 1]
                ops.spmm_csr_compute(crow_blk, d["col"][p0:p1], d["val"][p0:p1], d["B"], r1 - r0, A.cols, out=C_d[r0:r1])
            e = torch.cuda.Event()
            e.record(cur)
            ev_c.append(e)
        cur.wait_event(ev_dy)
        ops.spmm_csr_grad_b_transient_compute(d["crow"], d["col"], d["val"], d["dY"], A.rows, A.cols, out=dB_d)
        ev_b = torch.cuda.Event()
        ev_b.record(cur)
        with torch.cuda.stream(s_out):
            for i in range(blocks):
                r0, r1 = bounds[i], bounds[i + 1]
                s_out.wait_event(ev_c[i])
                C_h[r0:r1].copy_(C_d[r0:r1], non_blocking=True)
            s_out.wait_event(ev_b)
            dB_h.copy_(dB_d, non_blocking=True)
        cur.wait_stream(s_out)

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    k = max(3, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        e2e_step()
    b.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b) / k
    return {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
            "path": f"pinned host -> device in {blocks} nnz-balanced row blocks on a copy stream, forward per block as "
                    "it lands, A^T*dY via transient transpose, C blocks / dB -> pinned host on a third stream"}


def _teardown(world):
    """Leave the process group without ever hanging the launcher: a watchdog force-exits if NCCL
    teardown does not return (seen after CUDA-graph capture of collectives)."""
    if world <= 1:
        return
    import torch.distributed as dist

    def _bail():
        sys.stdout.flush()
        os._exit(0)
    t = threading.Timer(20.0, _bail)
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    except Exception:
        pass
    t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_reddit_n128_fp32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--bwd", default="transpose", choices=["transpose", "atomic"])
    ap.add_argument("--e2e-mode", default="simple", choices=["simple", "pipelined"],
                    help="pipelined = experimental row-block streaming of the CSR (see _e2e_pipelined)")
    ap.add_argument("--graph", action="store_true",
                    help="N>1: replay the sharded step from a CUDA graph (experimental: measured no faster than eager "
                         "on 8 B200s, and NCCL teardown after capture can hang — off by default)")
    ap.add_argument("--comm", default="nccl", choices=["nccl", "peer"],
                    help="N>1: NCCL all-gather / reduce-scatter kernels, or copy-engine pulls over symmetric (peer) memory")
    ap.add_argument("--panels", type=int, default=1,
                    help="N>1: column panels used to pipeline a collective with its own product (1 = whole width; the "
                         "collectives then overlap with the other product of the step)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import ofspmm_b200 as ofs
    ops = __import__("importlib").import_module("of-spmm_b200.ops")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a multi-rank run that stops making progress must not hold the box: hard exit after 5 min
        # (a healthy run takes well under one)
        wd = threading.Timer(300.0, lambda: os._exit(2))
        wd.daemon = True
        wd.start()
        dist.init_process_group("nccl", device_id=dev)

    spec = WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    A = _make_graph(spec, dev)          # identical on every rank (seeded)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    dY = ofs.graphs.upstream_grad(A.rows, n, 12, dev, dtype)
    s_dense = 4 if dtype == torch.float32 else 2
    alg = ofs.graphs.expected_alg_bytes(A.rows, A.cols, A.nnz, n, s_dense)
    flops_step = 2.0 * alg["flop"]      # forward + A^T·dY

    if world > 1:
        dmod = __import__("importlib").import_module("of-spmm_b200.dist")
        runner = dmod.ShardedSpmm(A, n, dtype, rank, world, dev, bwd=args.bwd, panels=args.panels, comm=args.comm)
        B_in, dY_in = runner.shard_rows(B), runner.shard_rows_out(dY)
        step_eager = lambda: runner.step(B_in, dY_in)
        l0 = ofs.launch_count()
        step_eager()
        graph_launches_per_step = ofs.launch_count() - l0
        if not args.graph:
            step, fwd_only = step_eager, (lambda: runner.forward(B_in))
        else:  # one CUDA graph per step: kernels + copies + NCCL collectives, no host launch gaps
            step = runner.capture(step_eager)
            fwd_only = runner.capture(lambda: runner.forward(B_in))
        parallelism = (f"row-block x{world} (nnz-balanced whole rows), B/dB row-sharded, comm={runner.comm}, "
                       f"{runner.panels} column panel(s), all-gather(B) overlapped with A^T*dY and reduce-scatter(dB) "
                       f"overlapped with A*B, " + ("CUDA-graph replay" if args.graph else "eager launches"))
    else:
        t0 = time.perf_counter()
        tr = ops.csr_transpose(A.crow, A.col, A.val, A.rows, A.cols) if args.bwd == "transpose" else None
        torch.cuda.synchronize()
        plan_ms = (time.perf_counter() - t0) * 1e3
        C = torch.empty((A.rows, n), dtype=dtype, device=dev)
        dB = torch.empty((A.cols, n), dtype=dtype, device=dev)

        def fwd_only():
            ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C)

        def step():
            ops.spmm_csr_compute(A.crow, A.col, A.val, B, A.rows, A.cols, out=C)
            ops.spmm_csr_grad_b_compute(A.crow, A.col, A.val, dY, A.rows, A.cols, transposed=tr, out=dB)
        parallelism = "single GPU"

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = ofs.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
    ev[1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    launches = ofs.launch_count() - launches0
    if world > 1 and args.graph and launches == 0:
        launches = graph_launches_per_step * args.steps   # replayed from the captured graph
    # dominant kernel (forward): its own CUDA-event loop right after, same residency / clocks
    for i in range(args.steps):
        fwd_ev[i][0].record()
        fwd_only()
        fwd_ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ev)

    t = torch.tensor([total_ms, fwd_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = flops_step / (ms_per_step * 1e-3) / 1e9

    # ---- cold-L2 forward (flush by writing 256 MB) for the record
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cold = []
    for _ in range(min(5, args.steps)):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fwd_only(); b.record()
        torch.cuda.synchronize()
        cold.append(a.elapsed_time(b))
    del flush

    # ---- e2e through the public op API with pinned host buffers (N=1 path; per rank for N>1)
    e2e = None
    if not args.no_e2e and world == 1 and args.e2e_mode == "pipelined":
        e2e = _e2e_pipelined(args, ofs, ops, A, B, dY, n, dtype, dev, flops_step)
    elif not args.no_e2e and world == 1:
        host = {k: v.cpu().pin_memory() for k, v in dict(crow=A.crow, col=A.col, val=A.val, B=B, dY=dY).items()}
        C_h = torch.empty((A.rows, n), dtype=dtype).pin_memory()
        dB_h = torch.empty((A.cols, n), dtype=dtype).pin_memory()
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = C_h.numel() * C_h.element_size() + dB_h.numel() * dB_h.element_size()

        def e2e_step():
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            c = ofs.spmm_csr(d["crow"], d["col"], d["val"], d["B"], A.rows, A.cols)
            g = ofs.spmm_csr_grad_b(d["crow"], d["col"], d["val"], d["dY"], A.rows, A.cols)
            C_h.copy_(c, non_blocking=True)
            dB_h.copy_(g, non_blocking=True)
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        k = max(3, min(args.steps, 10))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            e2e_step()
        b.record()
        torch.cuda.synchronize()
        e2e_ms = a.elapsed_time(b) / k
        e2e = {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
               "path": "pinned host -> device copies, ofs.spmm_csr + ofs.spmm_csr_grad_b (atomic route: no cached "
                       "transpose for fresh inputs), device -> pinned host"}

    if not args.no_e2e and world > 1:
        # per-rank e2e: this rank's CSR block, B shard and dY block come from pinned host memory every
        # step and its C block / dB shard go back; the sharded step itself is the one timed above.
        # Never let a failure here cost the scaling numbers.
        try:
            import torch.distributed as dist
            blk = runner.A_blk
            host = {k: v.cpu().pin_memory() for k, v in dict(crow=blk.crow, col=blk.col, val=blk.val, B=B_in, dY=dY_in).items()}
            dst = dict(crow=blk.crow, col=blk.col, val=blk.val, B=B_in, dY=dY_in)
            C_h = torch.empty((blk.rows, n), dtype=dtype).pin_memory()
            dB_h = torch.empty((runner.shard, n), dtype=dtype).pin_memory()
            h2d = sum(v.numel() * v.element_size() for v in host.values())
            d2h = C_h.numel() * C_h.element_size() + dB_h.numel() * dB_h.element_size()

            def e2e_step():
                for key, src in host.items():
                    dst[key].copy_(src, non_blocking=True)
                c, g = runner.step(B_in, dY_in)
                C_h.copy_(c, non_blocking=True)
                dB_h.copy_(g, non_blocking=True)
            # no extra collectives in this section (only the ones inside runner.step, issued the
            # same number of times by every rank): a rank-local failure can then never deadlock.
            for _ in range(3):
                e2e_step()
            torch.cuda.synchronize()
            k = max(3, min(args.steps, 10))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                e2e_step()
            b.record()
            torch.cuda.synchronize()
            e2e_ms = a.elapsed_time(b) / k
            tt = [e2e_ms, float(h2d) * world, float(d2h) * world]
            e2e = {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(tt[1]),
                   "d2h_bytes_per_step": int(tt[2]), "ms_per_step": e2e_ms,
                   "path": "per rank: pinned host -> device copies of its CSR row block, B shard and dY block, "
                           "ShardedSpmm.step (cached transpose of the block: the graph is static across steps), "
                           "C block and dB shard -> pinned host; time = rank 0's device time of the collective step (the "
                           "collectives inside it synchronise the ranks), bytes = rank 0's x world (nnz-balanced blocks)"}
        except Exception as exc:  # pragma: no cover
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "error": repr(exc)[:200]}

    if rank != 0:
        _teardown(world)
        return

    peak, peak_src = _peaks()
    fwd_bytes = alg["m2"] if world == 1 else alg["m2"] / world
    achieved = fwd_bytes / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by hand from the ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if dtype == torch.float32 else "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "rows": A.rows, "cols": A.cols, "nnz": A.nnz, "n": n,
                   "step": "forward C=A*B + backward dB=A^T*dY (" + args.bwd + " route)", "parallelism": parallelism,
                   "l2": "inputs larger than L2 (CSR stream 0.9 GB/step >> 126 MB); no flush between steps; "
                         "cold-L2 forward reported as fwd_ms_cold_l2",
                   "variant": ofs._lib.lib().ofspmm_fwd_variant(A.rows, A.nnz, n, 2 if dtype == torch.float32 else 11).decode()},
        "fwd_ms": fwd_ms, "fwd_ms_cold_l2": statistics.mean(cold) if cold else None,
        "fwd_gflops": alg["flop"] / (fwd_ms * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "spmm_merge_kernel (forward)", "peak_source": peak_src,
                     "model": "M2 gather model: nnz*(idx+val) + (M+1)*idx + nnz*N*s + M*N*s bytes per launch "
                              "(SURVEY.md 8d); B-row gathers are mostly L2 hits, so achieved may exceed the HBM copy "
                              "peak - see traffic (ncu dram bytes) and DESIGN.md",
                     "frac_of_nominal_8tbs": achieved / 8000.0,
                     "alg_bytes": fwd_bytes, "m1_compulsory_bytes": alg["m1"],
                     "m1_gbs": alg["m1"] / (fwd_ms * 1e-3) / 1e9 / max(1, world)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "reference_cuda_path": "n/a: the reference snapshot has no SpMM kernel and OneFlow does not build offline "
                               "(SURVEY.md 0.1-0.2); the CPU arm is `bench.py --impl reference`",
    }
    if world == 1:
        out["plan_ms_one_off"] = plan_ms
    if e2e is not None:
        out["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = _cpu_baseline(A, B, dY, n)
    print(json.dumps(out), flush=True)
    _teardown(world)


if __name__ == "__main__":
    main()
