"""The reference arm of bench.py (`--impl reference`) needs no GPU: check its one-line JSON contract
here, on BASELINE configs[0] (the reference's own CPU-runnable case)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1_uniform4096_n64_fp32", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_json_contract():
    lines = _run()
    assert len(lines) == 1                                   # exactly ONE JSON line
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GFLOP/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("SpMM GFLOP/s") and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["config"]["workload"] == "cfg1_uniform4096_n64_fp32" and j["config"]["n"] == 64
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "rows" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None and j["data"] == "synthetic"


def test_reference_arm_only_rank0_works_under_torchrun_env():
    # ranks != 0 exit 0 without output (the driver launches the arm with torchrun for N > 1)
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


# ------------------------------------------------------------------ N > 1: which sharded product the bench runs

def _pick_worker(rank, world, port, scheme, no_autotune, graph, q):
    import argparse
    import importlib

    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        import ofspmm_b200 as ofs
        from test_dist_cpu import OracleCompute
        dmod = importlib.import_module("of-spmm_b200.dist")
        n = 8
        # "dense": every rank touches every column block (nothing saved); "banded": almost nothing remote
        if graph == "dense":
            A = ofs.graphs.uniform_csr(96, 96, 0.5, seed=3)
        else:
            idx = torch.arange(96)
            A = ofs.formats.coo_to_csr(idx, idx, torch.ones(96), (96, 96))
        B = ofs.graphs.dense_operand(A.cols, n, 5)
        dY = ofs.graphs.upstream_grad(A.rows, n, 6)
        args = argparse.Namespace(scheme=scheme, no_autotune=no_autotune, tasks_per_warp=0, ag_dynamic_order=False, buckets=1,
                                  pull_ctas=64, layout="auto", no_interleave=False, combine_ctas=0, tune_combine=True)
        runner, used, saving, tuned = bench.pick_runner(args, dmod, A, B, dY, n, torch.float32, rank, world, "cpu",
                                                        compute=OracleCompute(), steps=2)
        C, dB = runner.step(runner.shard_rows(B), runner.shard_rows_out(dY))
        q.put((rank, used, saving, tuned, type(runner).__name__, runner.r0, runner.r1, C.clone().numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put(("error", rank, traceback.format_exc()))


def _pick(world, scheme, no_autotune, graph):
    import socket

    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pick_worker, args=(r, world, port, scheme, no_autotune, graph, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[0] != "error", r[2]
    return sorted(res)


def test_pick_runner_dense_exchange_is_autotuned_and_sparse_is_not():
    import numpy as np
    sys.path.insert(0, ROOT)
    import ofspmm_b200 as ofs
    from oracle import oracle as O
    # dense exchange, fp32: both schemes are timed, they agree, every rank takes the same one
    res = _pick(2, "auto", False, "dense")
    used = {r[1] for r in res}
    assert len(used) == 1 and used <= {"pull", "allgather"}
    for r in res:
        assert r[2] < 0.25 and r[3]["schemes_agree"] is True
        assert set(r[3]) >= {"pull", "pull_fwd_only", "allgather[static order, 2 tasks/warp]",
                             "allgather[dynamic order, 2 tasks/warp]", "allgather[dynamic order, 4 tasks/warp]"}
        assert (r[1] == "allgather") == ("allgather_policy" in r[3]) and "chosen" in r[3]
        assert "pull[combine beside the last forward pass]" in r[3]
        assert not any(k.endswith("error") for k in r[3])
        assert r[4] == ("AllGatherSpmm" if r[1] == "allgather" else "ShardedSpmm")
    A = ofs.graphs.uniform_csr(96, 96, 0.5, seed=3)
    B = ofs.graphs.dense_operand(A.cols, 8, 5)
    ref = O.spmm_f64(A.crow.numpy(), A.col.numpy(), A.val.numpy(), B.numpy(), A.cols)
    got = np.concatenate([r[7] for r in res], axis=0)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)
    # the rule alone (no timing): dense -> all-gather
    assert {r[1] for r in _pick(2, "auto", True, "dense")} == {"allgather"}
    # sparse exchange (a diagonal: no remote row at all): needed-rows exchange, nothing timed
    for r in _pick(2, "auto", False, "banded"):
        assert r[1] == "pull" and r[2] > 0.9 and r[3] is None
    # an explicit request is honoured
    assert {r[1] for r in _pick(2, "pull", False, "dense")} == {"pull"}
