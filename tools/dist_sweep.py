#!/usr/bin/env python
"""Multi-GPU tuning sweep (development tool): one torchrun launch, one graph, several
exchange-scheme / launch-policy configurations of the sharded products (dist.make_sharded) timed back to back.

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
    killer = threading.Timer(240.0, lambda: os._exit(3))
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
    # (name, scheme, constructor keywords): every launch policy is an argument of the C ABI now
    # (round 1 switched them through an environment variable)
    configs = [
        ("needed rows", "pull", {}),
        ("needed rows, forward(); backward()", "pull", dict(interleave=False)),
        ("needed rows, combine beside the last forward pass", "pull", dict(combine_ctas=148)),
        ("needed rows, 2 buckets", "pull", dict(buckets=2)),
        ("needed rows, block ownership", "pull", dict(shard_layout="block")),
        ("needed rows, cyclic ownership", "pull", dict(shard_layout="cyclic")),
        ("all-gather, static order, 2 tasks/warp", "allgather", dict(allgather_kw=dict(tasks_per_warp=2, static_order=True))),
        ("all-gather, dynamic order, 2 tasks/warp", "allgather", dict(allgather_kw=dict(tasks_per_warp=2, static_order=False))),
        ("all-gather, dynamic order, 4 tasks/warp", "allgather", dict(allgather_kw=dict(tasks_per_warp=4, static_order=False))),
        ("all-gather, persistent grids", "allgather", dict(allgather_kw=dict(tasks_per_warp=0))),
    ]
    ref_c = None
    for name, scheme, kw in configs:
        try:
            r, _, _ = dmod.make_sharded(A, n, dtype, rank, world, dev, scheme=scheme, **kw)
            Bs, dYs = r.shard_rows(B), r.shard_rows_out(dY)
            for _ in range(3):
                r.step(Bs, dYs)
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
            step_ms = timed(lambda: r.step(Bs, dYs))
            fwd_ms = timed(lambda: r.forward(Bs))
            bwd_ms = timed(lambda: r.backward(dYs))
            c, db = r.step(Bs, dYs)
            torch.cuda.synchronize()
            # layout-independent checksums: |C| over all row blocks, |dB| over all owned rows
            chk = torch.tensor([float(c.double().abs().sum()), float(db[: r.own].double().abs().sum())], device=dev, dtype=torch.float64)
            dist.all_reduce(chk)
            if rank == 0:
                same = None
                if ref_c is None:
                    ref_c = chk.clone()
                else:
                    same = bool(torch.allclose(chk, ref_c, rtol=1e-5))
                print(json.dumps({"config": name, "comm": r.comm, "step_ms": round(step_ms, 4), "fwd_ms": round(fwd_ms, 4),
                                  "bwd_ms": round(bwd_ms, 4), "tflops": round(flops_step / step_ms / 1e9, 2),
                                  "checksum_matches_first": same}), flush=True)
            del r
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        except Exception as e:  # keep going: one failing configuration must not waste the box
            if rank == 0:
                print(json.dumps({"config": name, "error": repr(e)[:300]}), flush=True)
    bench._teardown(world)


if __name__ == "__main__":
    main()
