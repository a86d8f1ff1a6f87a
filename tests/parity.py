"""Shared helpers for the GPU parity tests: compare the CUDA path (through the C ABI) with the
CPU oracle on the same inputs.  Integer outputs (IDs, positions, counts) and float32 scores are
compared BIT-EXACTLY: both sides order by (score, scan position)."""
import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same_results(g_ids, g_sc, g_cnt, o_ids, o_sc, what=""):
    """g_*: one query's row of the GPU output; o_*: oracle arrays."""
    n = int(g_cnt)
    assert n == len(o_ids), f"{what}: count {n} != oracle {len(o_ids)}"
    gs, os_ = bits(g_sc[:n]), bits(o_sc)
    if not np.array_equal(gs, os_):
        bad = np.nonzero(gs != os_)[0][:5]
        raise AssertionError(f"{what}: scores differ at ranks {bad.tolist()}: gpu {g_sc[bad]} oracle {o_sc[bad]}")
    if not np.array_equal(g_ids[:n], o_ids):
        bad = np.nonzero(g_ids[:n] != o_ids)[0][:5]
        raise AssertionError(f"{what}: ids differ at ranks {bad.tolist()}: gpu {g_ids[bad]} oracle {o_ids[bad]}")


def assert_same_up_to_ties(g_ids, g_sc, g_cnt, o_ids, o_sc, what=""):
    """The reference's own guarantee (unstable sort): equal score sequence, and equal ID sets inside
    every run of bit-equal scores except a run cut by the k boundary."""
    n = int(g_cnt)
    assert n == len(o_ids), f"{what}: count {n} != oracle {len(o_ids)}"
    assert np.array_equal(bits(g_sc[:n]), bits(o_sc)), f"{what}: score sequences differ"
    i = 0
    while i < n:
        j = i
        while j + 1 < n and bits(o_sc[j + 1:j + 2])[0] == bits(o_sc[i:i + 1])[0]:
            j += 1
        if j < n - 1:   # run not cut by the boundary
            assert set(g_ids[i:j + 1].tolist()) == set(o_ids[i:j + 1].tolist()), f"{what}: tie group {i}..{j} differs"
        i = j + 1
