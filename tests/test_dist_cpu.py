"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic in of-spmm_b200/dist.py: nnz-balanced
row blocks, equal B shards with padding, panel-pipelined all-gather / reduce-scatter and the
reassembly of C and dB.  The per-rank compute callbacks are CPU stand-ins built on the oracle
(tests may use it as the checker's arithmetic); the CUDA kernels themselves are covered by the
-m gpu tests and the N>1 bench."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_spmm(crow, col, val, b, rows, cols, out):
    from oracle import oracle as O
    res = O.spmm_f32(crow.numpy(), col.numpy(), val.numpy(), b.contiguous().numpy(), cols)
    out.copy_(torch.from_numpy(res))
    return out


def _cpu_transpose(crow, col, val, rows, cols):
    from oracle import oracle as O
    tc, tcol, tv, _ = O.csr_transpose(crow.numpy(), col.numpy(), val.numpy(), cols)
    return (torch.from_numpy(tc.astype(np.int32)), torch.from_numpy(tcol.astype(np.int32)), torch.from_numpy(tv))


def _worker(rank, world, port, n, panels, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import ofspmm_b200 as ofs
        dmod = importlib.import_module("of-spmm_b200.dist")
        A = ofs.graphs.rmat_csr(10, 12, seed=4)        # skewed rows: equal-count blocks would be unbalanced
        A = ofs.graphs.CsrMatrix(A.crow, A.col, A.val, A.rows, A.cols)
        K = A.cols
        B = ofs.graphs.dense_operand(K, n, 5)
        dY = ofs.graphs.upstream_grad(A.rows, n, 6)
        sh = dmod.ShardedSpmm(A, n, torch.float32, rank, world, "cpu", bwd="transpose", panels=panels,
                              spmm_fn=_cpu_spmm, transpose_fn=_cpu_transpose)
        C_blk, dB_shard = sh.step(sh.shard_rows(B), sh.shard_rows_out(dY))
        q.put((rank, sh.bounds, sh.r0, sh.r1, sh.shard, C_blk.clone().numpy(), dB_shard.clone().numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:   # report instead of leaving the parent (and the peer rank) waiting
        import traceback
        q.put(("error", rank, traceback.format_exc()))
        os._exit(1)


@pytest.mark.parametrize("world,n,panels", [(2, 16, 2), (2, 12, 1), (3, 32, 4)])
def test_sharded_spmm_gloo(world, n, panels):
    sys.path.insert(0, ROOT)
    import ofspmm_b200 as ofs
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, panels, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(world):
        r = q.get(timeout=120)
        if r[0] == "error":
            for p in procs:
                p.kill()
            raise AssertionError(f"rank {r[1]} failed:\n{r[2]}")
        results.append(r)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda t: t[0])

    A = ofs.graphs.rmat_csr(10, 12, seed=4)
    B = ofs.graphs.dense_operand(A.cols, n, 5).numpy()
    dY = ofs.graphs.upstream_grad(A.rows, n, 6).numpy()
    crow, col, val = A.crow.numpy(), A.col.numpy(), A.val.numpy()
    C_ref = O.spmm_f32(crow, col, val, B)
    dB_ref = O.spmm_t_f64(crow, col, val, dY, A.cols)

    bounds = results[0][1]
    assert bounds == list(O.row_blocks(crow, world))          # device/host partitioner == oracle
    assert bounds[0] == 0 and bounds[-1] == A.rows
    per_rank_nnz = np.diff(crow[np.array(bounds)])
    assert per_rank_nnz.max() <= 1.25 * per_rank_nnz.mean()   # nnz-balanced (equal-count would not be)
    C = np.concatenate([r[5] for r in results], axis=0)
    assert C.shape == C_ref.shape
    assert np.array_equal(C, C_ref)                           # same arithmetic per row → bitwise
    shard = results[0][4]
    dB = np.concatenate([r[6] for r in results], axis=0)
    assert dB.shape[0] == shard * world >= A.cols
    np.testing.assert_allclose(dB[: A.cols], dB_ref, rtol=1e-4, atol=1e-4)
    assert np.all(dB[A.cols:] == 0)                            # padding rows stay zero
