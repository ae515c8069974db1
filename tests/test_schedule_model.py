"""Executable model of the forward kernel's scheduling (csrc/spmm_kernels.cuh), checked against the
oracle with hypothesis on the CPU: merge-path tasks of ITEMS items, per-row 16-byte index chunks with
masked edge chunks, chunk → lane-group ownership, carry / head partials and the fix-up search
(4 linear steps, then binary search) that stitches rows spanning several tasks.

The model uses small-integer data so that float64 sums are exact: the comparison with the oracle is
then an equality that holds iff every non-zero is counted exactly once and every partial row lands
in the right output row.  It mirrors the *algorithm*; the CUDA implementation itself is covered by
tests/test_gpu_parity.py.  Keep the two in sync when the kernel's scheduling changes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O


def merge_path_search(crow, M, nnz, d):
    lo, hi = max(d - nnz, 0), min(d, M)
    while lo < hi:
        mid = (lo + hi) >> 1
        if crow[mid + 1] <= d - mid - 1:
            lo = mid + 1
        else:
            hi = mid
    return lo


def model_forward(crow, col, val, B, items, G, base_off):
    """Returns C computed the way the kernel schedules it.  ``base_off`` emulates the alignment
    phase of the col / val arrays (slot of element p in shared memory = (p + base_off) % 4 + ...)."""
    M, nnz = len(crow) - 1, int(crow[-1])
    N = B.shape[1]
    total = M + nnz
    P = -(-total // items) if total else 0
    part = []
    for k in range(P + 1):
        d = min(k * items, total)
        r = merge_path_search(crow, M, nnz, d)
        part.append((r, d - r))
    C = np.zeros((M, N))
    carry = np.full((max(P, 1), N), np.nan)      # NaN = never written: reading it poisons the result
    written = np.zeros(M, dtype=bool)
    for k in range(P):
        (rs, ns), (re, ne) = part[k], part[k + 1]
        cnt_nz = ne - ns
        assert 0 <= cnt_nz <= items and re - rs + 1 <= items + 1
        pre_c = (ns + base_off) % 4              # alignment phase of col + ns
        for r in range(rs, re + 1):
            e0 = 0 if r == rs else int(crow[r]) - ns
            e1 = int(crow[r + 1]) - ns if r < re else cnt_nz
            if r == re and e1 <= e0:
                break
            acc = np.zeros((G, N))
            s0, s1 = pre_c + e0, pre_c + e1
            cbase = s0 & ~3
            nchunks = ((s1 + 3) >> 2) - (s0 >> 2) if s1 > s0 else 0
            seen = 0
            for g in range(G):                   # each lane group walks its own chunks
                owned = []
                if nchunks > 0:
                    if g == 0:
                        owned.append(0)
                    q = G if g == 0 else g
                    while q < nchunks - 1:       # interior loop
                        owned.append(q)
                        q += G
                    if nchunks > 1 and q == nchunks - 1:
                        owned.append(nchunks - 1)
                for q in owned:
                    sa = cbase + 4 * q
                    edge = q == 0 or q == nchunks - 1
                    for u in range(4):
                        s = sa + u
                        if edge and not (s0 <= s < s1):
                            continue             # masked slot
                        assert s0 <= s < s1, "interior chunk must lie inside the row"
                        p = ns + (s - pre_c)
                        acc[g] += val[p] * B[col[p]]
                        seen += 1
            assert seen == e1 - e0               # every element of the segment exactly once
            row_sum = acc.sum(0)
            if r == re:
                carry[k] = row_sum               # trailing partial row
            else:
                assert not written[r]
                C[r] = row_sum                   # (head partial for fp32 outputs lives in C too)
                written[r] = True
    assert written.all()
    # fix-up: task k ends a row that started earlier -> add carries of tasks j..k-1
    for k in range(1, P):
        (rs, ns), (re, _) = part[k], part[k + 1]
        if re <= rs:
            continue
        cr = int(crow[rs])
        if cr >= ns:
            continue
        j, steps = k - 1, 0
        while j > 0 and steps < 4 and part[j][1] > cr:
            j -= 1
            steps += 1
        if j > 0 and part[j][1] > cr:
            lo, hi = 0, j
            while lo < hi:
                mid = (lo + hi) >> 1
                if part[mid + 1][1] > cr:
                    hi = mid
                else:
                    lo = mid + 1
            j = lo
        C[rs] += carry[j:k].sum(0)
    return C


@settings(max_examples=150, deadline=None)
@given(lens=st.lists(st.one_of(st.just(0), st.integers(0, 9), st.integers(0, 140)), min_size=1, max_size=40),
       items=st.sampled_from([8, 16, 32, 256]), G=st.sampled_from([1, 2, 4]), base_off=st.integers(0, 3),
       seed=st.integers(0, 10))
def test_schedule_model_counts_every_nonzero_once(lens, items, G, base_off, seed):
    rng = np.random.default_rng(seed)
    crow = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    M, nnz, K, N = len(lens), int(crow[-1]), 23, 3
    col = rng.integers(0, K, nnz).astype(np.int32)
    val = rng.integers(-3, 4, nnz).astype(np.float32)
    B = rng.integers(-4, 5, (K, N)).astype(np.float32)
    want = O.spmm_f64(crow, col, val, B, K)
    got = model_forward(crow, col, val.astype(np.float64), B.astype(np.float64), items, G, base_off)
    assert not np.isnan(got).any()
    assert np.array_equal(got, want)


def test_schedule_model_hub_row_uses_binary_search_path():
    lens = [0, 3, 0, 1000, 2, 0, 0, 77]            # 1000 non-zeros / 16-item tasks: > 60 tasks in one row
    crow = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rng = np.random.default_rng(1)
    nnz = int(crow[-1])
    col = rng.integers(0, 50, nnz).astype(np.int32)
    val = rng.integers(-3, 4, nnz).astype(np.float32)
    B = rng.integers(-4, 5, (50, 4)).astype(np.float32)
    for G in (1, 2, 4):
        got = model_forward(crow, col, val.astype(np.float64), B.astype(np.float64), 16, G, 1)
        assert np.array_equal(got, O.spmm_f64(crow, col, val, B, 50))
