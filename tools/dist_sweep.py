#!/usr/bin/env python
"""Multi-GPU tuning sweep (development tool): one torchrun launch, one graph, several
communication / grid configurations of dist.ShardedSpmm timed back to back.

    torchrun --nproc-per-node 8 tools/dist_sweep.py [--workload cfg2_reddit_n128_fp32]
"""
import argparse
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import ofspmm_b200 as ofs  # noqa: E402

dmod = __import__("importlib").import_module("of-spmm_b200.dist")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_reddit_n128_fp32")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    # never hang the launcher
    killer = threading.Timer(100.0, lambda: os._exit(3))
    killer.daemon = True
    killer.start()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    spec = bench.WORKLOADS[args.workload]
    n = spec["n"]
    dtype = torch.float32 if spec["dtype"] == "fp32" else torch.bfloat16
    A = bench._make_graph(spec, dev)
    B = ofs.graphs.dense_operand(A.cols, n, 11, dev, dtype)
    dY = ofs.graphs.upstream_grad(A.rows, n, 12, dev, dtype)
    flops_step = 2.0 * 2.0 * A.nnz * n
    # (comm, column panels, OFSPMM_TASKS_PER_WARP, cross-product overlap in step())
    configs = [("nccl", 1, 0, False), ("nccl", 1, 0, True), ("nccl", 1, 2, True), ("peer", 1, 0, True), ("peer", 1, 0, False)]
    ref_c = None
    for comm, panels, tpw, ov in configs:
        try:
            if tpw:
                os.environ["OFSPMM_TASKS_PER_WARP"] = str(tpw)
            else:
                os.environ.pop("OFSPMM_TASKS_PER_WARP", None)
            r = dmod.ShardedSpmm(A, n, dtype, rank, world, dev, panels=panels, comm=comm)
            Bs, dYs = r.shard_rows(B), r.shard_rows_out(dY)
            for _ in range(3):
                r.step(Bs, dYs, overlap=ov)
            dist.barrier()
            torch.cuda.synchronize()

            def timed(fn):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                torch.cuda.synchronize()
                a.record()
                for _ in range(args.steps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t[0])
            step_ms = timed(lambda: r.step(Bs, dYs, overlap=ov))
            fwd_ms = timed(lambda: r.forward(Bs))
            bwd_ms = timed(lambda: r.backward(dYs))
            c, db = r.step(Bs, dYs, overlap=ov)
            torch.cuda.synchronize()
            chk = torch.tensor([float(c.double().abs().sum()), float(db.double().abs().sum())], device=dev, dtype=torch.float64)
            dist.all_reduce(chk)
            if rank == 0:
                same = None
                if ref_c is None:
                    ref_c = chk.clone()
                else:
                    same = bool(torch.allclose(chk, ref_c, rtol=1e-6))
                print(json.dumps({"comm": r.comm, "panels": r.panels, "tasks_per_warp": tpw, "overlap": ov, "step_ms": round(step_ms, 4),
                                  "fwd_ms": round(fwd_ms, 4), "bwd_ms": round(bwd_ms, 4),
                                  "tflops": round(flops_step / step_ms / 1e9, 2), "checksum_matches_first": same}), flush=True)
            del r
        except Exception as e:  # keep going: one failing configuration must not waste the box
            if rank == 0:
                print(json.dumps({"comm": comm, "panels": panels, "tasks_per_warp": tpw, "error": repr(e)[:300]}), flush=True)
    bench._teardown(world)


if __name__ == "__main__":
    main()
