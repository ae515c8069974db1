"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic in of-spmm_b200/dist.py: nnz-balanced
row blocks, equal B shards with padding, the split of a block into a local and compact remote
sub-CSRs, the needed-rows pull / ordered scatter-add exchange (and round 1's all-gather /
reduce-scatter scheme), the SDDMM on the pulled rows and the reassembly of C, dB and dval.  The
per-rank compute is a CPU stand-in built on the oracle (tests may use it as the checker's
arithmetic) and the transport is the all-gather emulation of peer memory; the CUDA kernels and the
symmetric-memory transport are covered by tests/test_gpu_multi.py and the N>1 bench."""
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


class OracleCompute:
    """CPU stand-in for dist.CudaCompute: same interface, oracle arithmetic."""
    is_cuda = False

    def plan(self, *a, **k):
        return None

    def spmm(self, A, b, out, plan=None, accumulate=False, tasks_per_warp=0, bias=None, relu=False, static_order=False,
             acc32=None, acc32_in=False, acc32_out=False, reserve_ctas=0):
        from oracle import oracle as O
        assert not (acc32_in or acc32_out), "fp32 running sums are a 16-bit feature; the CPU stand-in is fp32"
        res = torch.from_numpy(O.spmm_f32(A.crow.numpy(), A.col.numpy(), A.val.numpy(), b.contiguous().numpy(), A.cols))
        if accumulate:
            res = res + out
        if bias is not None:
            res = res + bias
        if relu:
            res = res.clamp(min=0)
        out.copy_(res)
        return out

    def spmm_t(self, A, dy, out, plan=None, tasks_per_warp=0, acc32_out=None, reserve_ctas=0, static_order=False):
        from oracle import oracle as O
        out.copy_(torch.from_numpy(O.spmm_t_f32(A.crow.numpy(), A.col.numpy(), A.val.numpy(), dy.contiguous().numpy(), A.cols)))
        return out

    def sddmm(self, A, dy, b, plan=None):
        from oracle import oracle as O
        return torch.from_numpy(O.sddmm_f32(A.crow.numpy(), A.col.numpy(), dy.contiguous().numpy(), b.contiguous().numpy()))

    def gather_rows(self, dst, src, index, max_ctas=0):
        dst.copy_(src[index.long()])
        return dst

    def scatter_add_rows(self, dst, src, index, max_ctas=0):
        dst[index.long()] += src
        return dst


def _cache_worker(rank, world, port, n, cache_dir, q):
    """Every rank loads ONLY its row block from a partition cache; the result must equal the
    whole-graph construction bit for bit."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import ofspmm_b200 as ofs
        dmod = importlib.import_module("of-spmm_b200.dist")
        A = ofs.graphs.rmat_csr(10, 12, seed=4)
        B = ofs.graphs.dense_operand(A.cols, n, 5)
        dY = ofs.graphs.upstream_grad(A.rows, n, 6)
        G = ofs.formats.load_partition(cache_dir, rank, world)
        assert G.block.nnz < A.nnz and G.rows == A.rows and G.cols == A.cols
        out = {}
        for name, graph in (("cache", G), ("full", A)):
            sh = dmod.ShardedSpmm(graph, n, torch.float32, rank, world, "cpu", compute=OracleCompute(), shard_layout="block")
            C, dB = sh.step(sh.shard_rows(B), sh.shard_rows_out(dY))
            out[name] = (C.clone(), dB.clone(), sh.bounds)
            del sh
        ag = dmod.make_sharded(G, n, torch.float32, rank, world, "cpu", scheme="allgather", compute=OracleCompute())[0]
        C3, dB3 = ag.step(ag.shard_rows(B), ag.shard_rows_out(dY))
        saving = dmod.needed_rows_saving(G, rank, world)
        q.put((rank, bool(torch.equal(out["cache"][0], out["full"][0])), bool(torch.equal(out["cache"][1], out["full"][1])),
               out["cache"][2] == out["full"][2], bool(torch.allclose(C3, out["full"][0], rtol=1e-5, atol=1e-5)), saving))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put(("error", rank, traceback.format_exc()))


def test_partition_cache_round_trip_gloo(tmp_path):
    import ofspmm_b200 as ofs
    world, n = 3, 8
    A = ofs.graphs.rmat_csr(10, 12, seed=4)
    m = ofs.formats.save_partition(str(tmp_path), A, world)
    assert m["bounds"][0] == 0 and m["bounds"][-1] == A.rows and sum(m["block_nnz"]) == A.nnz
    items = [m["bounds"][r + 1] - m["bounds"][r] + m["block_nnz"][r] for r in range(world)]     # merge items: rows + non-zeros
    assert max(items) - min(items) <= int(A.row_lengths().max()) + 1                   # balanced up to one whole row
    import scipy.sparse as sp                                                          # blocks are plain scipy files
    blk1 = sp.load_npz(os.path.join(str(tmp_path), "block_1.npz"))
    assert blk1.shape == (m["bounds"][2] - m["bounds"][1], A.cols) and blk1.nnz == m["block_nnz"][1]
    with pytest.raises(ValueError):
        ofs.formats.load_partition(str(tmp_path), 0, world + 1)                        # written for another world size
    G0 = ofs.formats.load_partition(str(tmp_path), 0, world)
    with pytest.raises(ValueError):
        G0.row_slice(m["bounds"][1], m["bounds"][2])                                   # a rank holds only its own block
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cache_worker, args=(r, world, port, n, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[0] != "error", r[2]
        assert r[1] and r[2] and r[3] and r[4], r
        assert 0.0 <= r[5] <= 1.0


def _worker(rank, world, port, n, scheme, buckets, layout, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import ofspmm_b200 as ofs
        dmod = importlib.import_module("of-spmm_b200.dist")
        A = ofs.graphs.rmat_csr(10, 12, seed=4)        # skewed rows: equal-count blocks would be unbalanced
        B = ofs.graphs.dense_operand(A.cols, n, 5)
        dY = ofs.graphs.upstream_grad(A.rows, n, 6)
        bias = torch.linspace(-1, 1, n)
        if scheme == "pull":
            sh = dmod.ShardedSpmm(A, n, torch.float32, rank, world, "cpu", buckets=buckets, compute=OracleCompute(),
                                  shard_layout=layout, cyclic_block=16)
        elif scheme.startswith("auto"):
            # the factory's collective decision: threshold above / below any possible saving
            thr = 2.0 if scheme == "auto_ag" else -1.0
            sh, used, saving = dmod.make_sharded(A, n, torch.float32, rank, world, "cpu", scheme="auto", saving_threshold=thr,
                                                 compute=OracleCompute(), buckets=buckets, shard_layout=layout, cyclic_block=16)
            assert used == ("allgather" if scheme == "auto_ag" else "pull") and 0.0 <= saving <= 1.0
            assert isinstance(sh, dmod.AllGatherSpmm if scheme == "auto_ag" else dmod.ShardedSpmm)
            # 16-bit products never take the collective scheme (their partial sums are combined in fp32)
            assert dmod.make_sharded(A, n, torch.bfloat16, rank, world, "cpu", scheme="auto", saving_threshold=2.0,
                                     compute=OracleCompute())[1] == "pull"
            scheme = "pull" if used == "pull" else "allgather"
        else:
            sh = dmod.AllGatherSpmm(A, n, torch.float32, rank, world, "cpu", compute=OracleCompute())
        out = {}
        for it in range(2):                             # twice: the publish / consume hand-shake repeats
            C_blk, dB_shard = sh.step(sh.shard_rows(B), sh.shard_rows_out(dY))
        out["C"], out["dB"] = C_blk.clone().numpy(), dB_shard.clone().numpy()
        # forward and backward called separately must agree with the interleaved step
        C2, dB2 = sh.forward(sh.shard_rows(B)).clone(), sh.backward(sh.shard_rows_out(dY)).clone()
        out["sep_equals_step"] = bool(torch.equal(C2, torch.from_numpy(out["C"]))) and bool(torch.equal(dB2, torch.from_numpy(out["dB"])))
        if scheme == "pull":
            out["dval"] = sh.sddmm(sh.shard_rows_out(dY)).numpy()
            out["C_ep"] = sh.forward(sh.shard_rows(B), bias=bias, relu=True).clone().numpy()
            out["stats"] = sh.exchange_bytes()
            out["nsub"] = len(sh.sub)
            # new edge values, same structure
            sh.update_values(sh.A_blk.val * -2.0)
            out["C_scaled"] = sh.forward(sh.shard_rows(B)).clone().numpy()
        out["shard_ids"], out["layout"] = sh.shard_ids.numpy(), getattr(sh, "layout", "block")
        q.put((rank, sh.bounds, sh.r0, sh.r1, sh.shard, out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:   # report instead of leaving the parent (and the peer rank) waiting
        import traceback
        q.put(("error", rank, traceback.format_exc()))
        os._exit(1)


@pytest.mark.parametrize("world,n,scheme,buckets,layout", [(2, 16, "pull", 1, "auto"), (3, 12, "pull", 1, "block"),
                                                           (3, 8, "pull", 2, "cyclic"), (4, 8, "pull", 1, "auto"),
                                                           (2, 12, "allgather", 1, "block"), (2, 8, "auto_ag", 1, "block"),
                                                           (3, 8, "auto_pull", 1, "auto")])
def test_sharded_spmm_gloo(world, n, scheme, buckets, layout):
    sys.path.insert(0, ROOT)
    import ofspmm_b200 as ofs
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, scheme, buckets, layout, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(world):
        r = q.get(timeout=180)
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
    C_ref = O.spmm_f64(crow, col, val, B)
    dB_ref = O.spmm_t_f64(crow, col, val, dY, A.cols)

    bounds = results[0][1]
    assert bounds == list(O.row_blocks(crow, world))          # device/host partitioner == oracle
    assert bounds[0] == 0 and bounds[-1] == A.rows
    per_rank_nnz = np.diff(crow[np.array(bounds)])
    assert per_rank_nnz.max() <= 1.25 * per_rank_nnz.mean()   # nnz-balanced (equal-count would not be)
    C = np.concatenate([r[5]["C"] for r in results], axis=0)
    assert C.shape == C_ref.shape
    np.testing.assert_allclose(C, C_ref, rtol=1e-4, atol=1e-4)
    shard = results[0][4]
    dB = np.full((A.cols, n), np.nan)
    owned = 0
    for r in results:                                         # row i of a shard is row shard_ids[i] of dB
        ids = r[5]["shard_ids"]
        assert r[5]["dB"].shape[0] == shard and len(ids) <= shard
        dB[ids] = r[5]["dB"][: len(ids)]
        assert np.all(r[5]["dB"][len(ids):] == 0)            # padding rows stay zero
        owned += len(ids)
    assert owned == A.cols and not np.isnan(dB).any()         # every row of dB has exactly one owner
    np.testing.assert_allclose(dB, dB_ref, rtol=1e-4, atol=1e-4)
    if layout == "cyclic":
        assert all(r[5]["layout"] == "cyclic" for r in results)
        assert results[0][5]["shard_ids"][:17].tolist() == list(range(16)) + [16 * world]   # blocks of 16 dealt round-robin
    assert all(r[5]["sep_equals_step"] for r in results)      # forward(); backward() == the interleaved step(), bit for bit
    assert len({r[5]["layout"] for r in results}) == 1        # "auto" is a collective decision: all ranks agree
    if scheme not in ("pull", "auto_pull"):
        return
    dval = np.concatenate([r[5]["dval"] for r in results])
    dv_ref, _ = O.sddmm_f64(crow, col, dY, B)
    np.testing.assert_allclose(dval, dv_ref, rtol=1e-4, atol=1e-4)
    bias = np.linspace(-1, 1, n)
    C_ep = np.concatenate([r[5]["C_ep"] for r in results], axis=0)
    np.testing.assert_allclose(C_ep, np.maximum(C_ref + bias[None, :], 0), rtol=1e-4, atol=1e-4)
    C_scaled = np.concatenate([r[5]["C_scaled"] for r in results], axis=0)
    np.testing.assert_allclose(C_scaled, -2.0 * C_ref, rtol=1e-4, atol=2e-4)
    for r in results:
        st = r[5]["stats"]
        assert r[5]["nsub"] == 1 + min(buckets, world - 1)
        assert 0 < st["pulled"] <= st["all_gather"]           # never more than an all-gather moves
        assert 0.0 < st["local_nnz_fraction"] < 1.0


# ------------------------------------------------------------------ sharded 2-layer GCN (configs[4] host logic)

def _gcn_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        import ofspmm_b200 as ofs
        gcn = importlib.import_module("of-spmm_b200.gcn")
        A = ofs.graphs.gcn_normalize(ofs.graphs.add_self_loops(ofs.graphs.uniform_csr(300, 300, 0.03, seed=9)))
        X = ofs.graphs.dense_operand(300, 24, 3)
        labels = torch.randint(0, 7, (300,), generator=torch.Generator().manual_seed(1))
        model = gcn.ShardedGCN2(A, rank, world, "cpu", in_dim=24, hidden=16, out_dim=7, seed=5, compute=OracleCompute())
        Xr, yr = model.local_rows(X), model.local_rows(labels)
        loss = model.train_step(Xr, yr)
        loss2 = model.train_step(Xr, yr)                  # second step: slots / hand-shakes are reusable
        q.put((rank, model.sh.r0, model.sh.r1, float(loss), float(loss2), model.grads["W1"].clone().numpy(),
               model.grads["W2"].clone().numpy(), model.grads["val"].clone().numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put(("error", rank, traceback.format_exc()))
        os._exit(1)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gcn_gloo_matches_dense_fp64(world):
    sys.path.insert(0, ROOT)
    import ofspmm_b200 as ofs
    import importlib
    gcn = importlib.import_module("of-spmm_b200.gcn")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gcn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(world):
        r = q.get(timeout=180)
        if r[0] == "error":
            for p in procs:
                p.kill()
            raise AssertionError(f"rank {r[1]} failed:\n{r[2]}")
        results.append(r)
    for p in procs:
        p.join(timeout=60)
    results.sort(key=lambda t: t[0])
    A = ofs.graphs.gcn_normalize(ofs.graphs.add_self_loops(ofs.graphs.uniform_csr(300, 300, 0.03, seed=9)))
    X = ofs.graphs.dense_operand(300, 24, 3).double()
    labels = torch.randint(0, 7, (300,), generator=torch.Generator().manual_seed(1))
    rows = torch.repeat_interleave(torch.arange(A.rows), A.row_lengths())
    val = A.val.double().requires_grad_(True)
    dense = torch.zeros(A.rows, A.cols, dtype=torch.float64).index_put((rows, A.col.long()), val)
    W1 = gcn.glorot(24, 16, 5, "cpu").double().requires_grad_(True)
    W2 = gcn.glorot(16, 7, 6, "cpu").double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy((dense @ torch.relu(dense @ (X @ W1))) @ W2, labels)
    ref.backward()
    for r in results:
        assert abs(r[3] - float(ref)) < 1e-5 and r[4] == pytest.approx(r[3], abs=1e-6)
        assert np.allclose(r[5], W1.grad.numpy(), rtol=1e-4, atol=1e-6)
        assert np.allclose(r[6], W2.grad.numpy(), rtol=1e-4, atol=1e-6)
    dval = np.concatenate([r[7] for r in results])
    assert np.allclose(dval, val.grad.numpy(), rtol=1e-4, atol=1e-6)
