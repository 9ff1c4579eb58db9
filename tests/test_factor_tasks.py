"""
Host-side check of the task list the persistent factorisation kernel executes (no GPU needed: the list is built by the
library on the host, ``gpso_debug_factor_tasks``).  The kernel hands out tasks in queue order and a CTA waits for the
counters its task names, so the schedule is deadlock-free exactly when the queue is a topological order of those
dependencies: replaying it sequentially, every dependency must already be satisfied when its task comes up.  The replay
also checks the per-tile operation counts of the blocked Cholesky and that every pair of the inverse recursion is complete.
"""
import ctypes

import numpy as np
import pytest

from pygpso_b200 import backend

OPS = {0: "DIAG", 1: "PANEL", 2: "UPD", 3: "TRANSPOSE", 4: "XT", 5: "Y"}


def task_list(nb, nsm=148):
    lib = backend.load_library()
    nt, nc = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.gpso_debug_factor_tasks(nb, nsm, None, 0, ctypes.byref(nt), ctypes.byref(nc)) == 0
    buf = np.zeros(nt.value * 16, dtype=np.int32)
    rc = lib.gpso_debug_factor_tasks(nb, nsm, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), buf.size, ctypes.byref(nt), ctypes.byref(nc))
    assert rc == 0
    return buf.reshape(-1, 16), nc.value


@pytest.mark.parametrize("nb", [2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 47, 48, 49, 64, 65])
def test_queue_is_topological_and_complete(nb):
    tasks, ncounters = task_list(nb)
    W = 4 if nb >= 48 else 1
    ops = lambda j: j // W + j % W
    ctr = np.zeros(ncounters, dtype=np.int64)
    seen = {name: 0 for name in OPS.values()}
    for t in tasks:
        op = OPS[int(t[0])]
        seen[op] += 1
        for k in range(3):
            if t[6 + k] >= 0:
                assert ctr[t[6 + k]] >= t[9 + k], (op, t.tolist(), int(ctr[t[6 + k]]))
        if t[13] > 0:
            # tile counters only ever move forward by one operation (updates are applied in order) -- except the diagonal
            # tile, whose last update is fused into its DIAG task
            step = t[13] - ctr[t[12]]
            assert step == (2 if op == "DIAG" and t[1] > 0 else 1), (op, t.tolist(), int(ctr[t[12]]))
            ctr[t[12]] = t[13]
        else:
            ctr[t[12]] += 1
    # every tile of the lower triangle is final, every transpose done; the tile below each diagonal block is solved in four
    # row strips that count up their own counter (everything that reads the tile waits for that one)
    strip = lambda j: nb * nb + 1 + 2 * 8 * nb + j
    for i in range(nb):
        for j in range(i + 1):
            if i == j + 1:
                assert ctr[i * nb + j] == ops(j) and ctr[strip(j)] == 4, (i, j, int(ctr[i * nb + j]), int(ctr[strip(j)]))
            else:
                assert ctr[i * nb + j] == ops(j) + 1, (i, j, int(ctr[i * nb + j]))
    trc = lambda p: nb * nb + 1 + 2 * 8 * nb + 2 * nb + p  # transposes are chained: counter p = tiles 0 .. p transposed
    assert all(ctr[trc(p)] == 1 for p in range(nb))
    assert seen["DIAG"] == nb and seen["TRANSPOSE"] == nb and seen["PANEL"] == nb * (nb - 1) // 2 + 3 * (nb - 1)
    for t in tasks:
        if OPS[int(t[0])] == "PANEL" and t[4] > 0:
            assert t[2] == t[1] + 1 and 1 <= t[4] <= 4 and t[12] == strip(t[1]) and t[13] == 0, t.tolist()
        for k in range(3):  # nobody waits for the tile counter of a strip-solved tile to become final
            c, v = int(t[6 + k]), int(t[9 + k])
            if 0 <= c < nb * nb and c // nb == c % nb + 1:
                assert v <= ops(c % nb), t.tolist()
    # the inverse recursion: XT and Y tiles of every level, s * nv per pair
    expect = 0
    s = 1
    while s < nb:
        q = 0
        while 2 * q * s < nb:
            rem = nb - (2 * q * s + s)
            expect += s * max(0, min(rem, s))
            q += 1
        s *= 2
    assert seen["XT"] == expect and seen["Y"] == expect


@pytest.mark.parametrize("cap", [0, 1, 2, 4, 8])
@pytest.mark.parametrize("nb", [4, 5, 8, 9, 16, 24, 32])
def test_capped_inverse_levels_keep_the_queue_topological(nb, cap):
    """What the library runs from N = 512 up: the inverse-factor merges of the levels s < cap tiles inside the queue (entered
    as soon as their inputs are), the levels above left to the int8 engine."""
    lib = backend.load_library()
    nt, nc = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.gpso_debug_factor_tasks_cap(nb, 148, cap, None, 0, ctypes.byref(nt), ctypes.byref(nc)) == 0
    buf = np.zeros(nt.value * 16, dtype=np.int32)
    assert lib.gpso_debug_factor_tasks_cap(nb, 148, cap, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), buf.size, ctypes.byref(nt), ctypes.byref(nc)) == 0
    tasks = buf.reshape(-1, 16)
    ctr = np.zeros(nc.value, dtype=np.int64)
    levels_seen = set()
    for t in tasks:
        for k in range(3):
            if t[6 + k] >= 0:
                assert ctr[t[6 + k]] >= t[9 + k], (OPS[int(t[0])], t.tolist(), int(ctr[t[6 + k]]))
        if t[13] > 0:
            ctr[t[12]] = t[13]
        else:
            ctr[t[12]] += 1
        if OPS[int(t[0])] in ("XT", "Y"):
            levels_seen.add(int(t[4]))
    want = {s for s in (1, 2, 4, 8, 16) if s < cap and s < nb}
    assert levels_seen == want, (levels_seen, want)
    # XT and Y tasks of the levels present: s * nv per pair, both kinds
    expect = sum(s * max(0, min(nb - (2 * q * s + s), s)) for s in want for q in range(nb) if 2 * q * s < nb)
    kinds = [OPS[int(t[0])] for t in tasks]
    assert kinds.count("XT") == expect and kinds.count("Y") == expect


def test_single_panel_matrix_has_no_scheduler_tasks_beyond_the_block():
    tasks, _ = task_list(1)
    assert [OPS[int(t[0])] for t in tasks] == ["DIAG", "TRANSPOSE"]


@pytest.mark.parametrize("nb,chain_slack,work_slack", [(16, 1.20, None), (32, 1.20, None), (64, None, 1.30)])
def test_schedule_quality_in_the_discrete_event_replay(nb, chain_slack, work_slack):
    """tools/factor_sim.py replays the queue with the measured task durations on 148 workers.  Up to N = 4096 the
    factorisation is bound by its chains of dependent tile tasks -- DIAG -> four PANEL strips -> DIAG on the diagonal, PANEL ->
    UPDATE -> PANEL on the rows below it -- and the queue must not add more than 20 % to the longer one; at N = 8192 the makespan has to stay
    within 30 % of the work bound (its chain-bound tail costs ~20 %).  Guards the look-ahead ordering of the host builder."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import factor_sim

    r = factor_sim.simulate(nb)
    assert r["makespan_us"] >= max(r["chain_us"], r["work_bound_us"]) * 0.999
    if chain_slack:
        assert r["makespan_us"] <= chain_slack * max(r["chain_us"], r["update_chain_us"]), r
    if work_slack:
        assert r["makespan_us"] <= work_slack * r["work_bound_us"], r
