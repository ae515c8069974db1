#!/usr/bin/env python
"""Where does a sharded forward spend its time?  (development probe, torchrun, N >= 2)

    torchrun --nproc-per-node 2 tools/dist_probe.py --workload cfg2_reddit_n128_fp32

Times, with CUDA events on the streams involved and max over ranks: the pull kernel alone, the local
and remote products alone, forward-only loops and forward+backward loops, for a few launch policies."""
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
    threading.Timer(150.0, lambda: os._exit(3)).start()
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

    def timed(fn, sync_each=False):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
            if sync_each:
                torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t[0]), 4)

    for pull_ctas in (64, 148):
        r = dmod.ShardedSpmm(A, n, dtype, rank, world, dev, pull_ctas=pull_ctas)
        Bs, dYs = r.shard_rows(B), r.shard_rows_out(dY)
        cp, s0, s1 = r.cp, r.sub[0], r.sub[1]
        C = r._c
        out = {"workload": args.workload, "world": world, "pull_ctas": pull_ctas,
               "local_nnz": s0.A.nnz, "remote_nnz": s1.A.nnz, "pulled_rows": r.pulled_rows,
               "local_spmm": timed(lambda: cp.spmm(s0.A, r.B_pub[: s0.A.cols], C, plan=s0.plan)),
               "local_spmm_reserve1": timed(lambda: cp.spmm(s0.A, r.B_pub[: s0.A.cols], C, plan=s0.plan, reserve_ctas=1)),
               "remote_spmm_acc": timed(lambda: cp.spmm(s1.A, s1.Bc[0], C, plan=s1.plan, accumulate=(dtype == torch.float32))),
               "forward_sync_each": timed(lambda: r.forward(Bs), sync_each=True),
               "forward_back_to_back": timed(lambda: r.forward(Bs)),
               "backward_back_to_back": timed(lambda: r.backward(dYs)),
               "step_back_to_back": timed(lambda: r.step(Bs, dYs))}
        # local product on a plain (non-symmetric) copy of the shard: is symmetric memory slower to read locally?
        plain = r.B_pub.clone()
        out["local_spmm_plain_memory"] = timed(lambda: cp.spmm(s0.A, plain[: s0.A.cols], C, plan=s0.plan))
        if rank == 0:
            print(json.dumps(out), flush=True)
        del r
    bench._teardown(world)


if __name__ == "__main__":
    main()
