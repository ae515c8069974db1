"""Executable model of the flag-synchronised exchange of ``dist.ShardedSpmm`` (fused mode) —
checks the claim the code relies on: *published buffers double buffered by epoch parity need no
"done reading" barrier*, for every interleaving of the ranks' streams.

What is modelled (of-spmm_b200/dist.py, ``_fwd_begin / _bwd_remote / _fwd_local / _bwd_local /
_fwd_remote / _bwd_combine`` and ``step``): per rank a main stream and a communication stream, each
executing its operations in order; cross-stream event waits; ``signal`` = write my epoch into every
peer's pad (after everything earlier on the stream); ``pull`` / ``combine`` = for every peer, wait
until its flag reaches the epoch, then read its published buffer of that epoch's parity over a time
interval (begin … end) during which other operations of other ranks may run.  The scheduler picks
any enabled stream head at random, so ranks drift apart as far as the dependencies allow.

Safety: a read of peer p's buffer at epoch e must see epoch e from begin to end (never the
overwrite of epoch e + 2, never a stale e − 2).  Liveness: every schedule runs to completion.
The model has teeth: with ONE buffer instead of two (parity ignored) the same scheduler finds the
overwrite within a few hundred schedules, and dropping the forward signal dead-locks.

Model, not product code: it documents and checks the protocol; the kernels themselves are tested
on GPUs in tests/test_gpu_multi.py and tests/test_gpu_round2.py.
"""
import random

import pytest


class Violation(Exception):
    pass


class Op:
    __slots__ = ("kind", "rank", "epoch", "buf", "peer", "deps", "done", "name")

    def __init__(self, kind, rank, epoch, buf=None, peer=None, name=""):
        self.kind, self.rank, self.epoch, self.buf, self.peer = kind, rank, epoch, buf, peer
        self.deps, self.done, self.name = [], False, name


def build_program(world, steps, variant="step", double_buffered=True, signal_forward=True, combine_on_comm=False):
    """Per rank two op lists (main, comm).  Kinds: write(buf), signal(kind), wait_flag+read_begin,
    read_end, local (no shared state), each with explicit extra deps (event waits)."""
    prog = {r: {"main": [], "comm": []} for r in range(world)}
    for r in range(world):
        main, comm = prog[r]["main"], prog[r]["comm"]
        prev_step_last_main = None
        for k in range(steps):
            e = k + 1
            par = (e & 1) if double_buffered else 0
            peers = [p for p in range(world) if p != r]

            def add(lst, op, *deps):
                op.deps.extend(d for d in deps if d is not None)
                lst.append(op)
                return op

            ev_in = add(main, Op("local", r, e, name="ev_in"))
            wB = add(main, Op("write", r, e, buf=("B", r, par), name="publish B"))
            sF = add(main, Op("signal", r, e, buf="F", name="signal fwd")) if signal_forward else wB
            # pull on the comm stream: waits ev_in (the previous step's readers of the pulled rows are done)
            pulls = []
            for p in peers:
                b = add(comm, Op("read_begin", r, e, buf=("B", p, par), peer=p, name="pull begin"), ev_in)
                b.kind = "read_begin_F"
                pulls.append(add(comm, Op("read_end", r, e, buf=("B", p, par), peer=p, name="pull end")))
            ev_g = pulls[-1] if pulls else None

            def bwd_remote():
                w = add(main, Op("write", r, e, buf=("D", r, par), name="publish partials"))
                add(main, Op("signal", r, e, buf="Bk", name="signal bwd"))
                return w

            def combine(lst, *deps):
                last = None
                for p in peers:
                    b = add(lst, Op("read_begin", r, e, buf=("D", p, par), peer=p, name="combine begin"), *deps)
                    b.kind = "read_begin_Bk"
                    last = add(lst, Op("read_end", r, e, buf=("D", p, par), peer=p, name="combine end"))
                return last

            if variant == "step":                      # ShardedSpmm.step, interleaved
                bwd_remote()
                add(main, Op("local", r, e, name="fwd local"))
                loc = add(main, Op("local", r, e, name="bwd local"))
                if combine_on_comm:
                    done = combine(comm, loc)
                    add(main, Op("local", r, e, name="fwd remote"), ev_g)
                    add(main, Op("local", r, e, name="wait combine"), done)
                else:
                    add(main, Op("local", r, e, name="fwd remote"), ev_g)
                    combine(main)
            else:                                      # forward(); backward()
                add(main, Op("local", r, e, name="fwd local"))
                add(main, Op("local", r, e, name="fwd remote"), ev_g)
                bwd_remote()
                add(main, Op("local", r, e, name="bwd local"))
                combine(main)
    return prog


def run_schedule(world, steps, rng, **kw):
    prog = build_program(world, steps, **kw)
    flags = {("F", dst, src): 0 for dst in range(world) for src in range(world)}
    flags.update({("Bk", dst, src): 0 for dst in range(world) for src in range(world)})
    content = {}                                        # buffer -> epoch it holds
    readers = {}                                        # buffer -> set of (rank, epoch) currently reading
    heads = {(r, s): 0 for r in range(world) for s in ("main", "comm")}
    total = sum(len(prog[r][s]) for r in range(world) for s in ("main", "comm"))
    executed = 0
    while executed < total:
        enabled = []
        for (r, s), i in heads.items():
            lst = prog[r][s]
            if i >= len(lst):
                continue
            op = lst[i]
            if not all(d.done for d in op.deps):
                continue
            if op.kind == "read_begin_F" and flags[("F", r, op.peer)] < op.epoch:
                continue
            if op.kind == "read_begin_Bk" and flags[("Bk", r, op.peer)] < op.epoch:
                continue
            enabled.append((r, s))
        if not enabled:
            raise Violation("deadlock")
        r, s = rng.choice(enabled)
        op = prog[r][s][heads[(r, s)]]
        if op.kind == "write":
            if readers.get(op.buf):
                raise Violation(f"rank {op.rank} overwrites {op.buf} with epoch {op.epoch} under readers {readers[op.buf]}")
            content[op.buf] = op.epoch
        elif op.kind == "signal":
            for p in range(world):
                if p != r:
                    flags[(op.buf, p, r)] = op.epoch    # the release: everything earlier on this stream is visible
        elif op.kind.startswith("read_begin"):
            if content.get(op.buf) != op.epoch:
                raise Violation(f"rank {op.rank} reads {op.buf} at epoch {op.epoch}, buffer holds {content.get(op.buf)}")
            readers.setdefault(op.buf, set()).add((op.rank, op.epoch))
        elif op.kind == "read_end":
            if content.get(op.buf) != op.epoch:
                raise Violation(f"{op.buf} changed under rank {op.rank}'s read of epoch {op.epoch}")
            readers[op.buf].discard((op.rank, op.epoch))
        op.done = True
        heads[(r, s)] += 1
        executed += 1
    return True


@pytest.mark.parametrize("variant,combine_on_comm", [("step", False), ("step", True), ("separate", False)])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_double_buffered_exchange_is_safe_and_live_under_any_interleaving(world, variant, combine_on_comm):
    rng = random.Random(1000 * world + len(variant) + combine_on_comm)
    for _ in range(400):
        assert run_schedule(world, 5, rng, variant=variant, combine_on_comm=combine_on_comm)


def test_model_has_teeth_single_buffer_is_unsafe_and_missing_signal_deadlocks():
    rng = random.Random(7)
    seen = 0
    for _ in range(600):
        try:
            run_schedule(3, 5, rng, double_buffered=False)
        except Violation as v:
            assert "overwrites" in str(v) or "changed under" in str(v) or "holds" in str(v)
            seen += 1
    assert seen > 0, "a single published buffer must be caught being overwritten under a reader"
    with pytest.raises(Violation, match="deadlock"):
        run_schedule(2, 2, random.Random(1), signal_forward=False)
