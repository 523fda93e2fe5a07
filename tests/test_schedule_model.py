"""The periodic visit schedule of the split pipeline's solver, checked on the host model of it
(profiles/sched_sim.py): a schedule that satisfies the ring constraints executes every body's visits in pool order,
one at a time -- which is all the sequential Gauss-Seidel loop (reference src/px_world.c:249-260) needs for identical
bits.  The device's schedules themselves are covered by the bit-parity tests (-m gpu)."""
from __future__ import annotations

import importlib.util
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402
from pdphyzx_b200 import batch as pxb, scenes  # noqa: E402


def _sched_sim():
    spec = importlib.util.spec_from_file_location("sched_sim", os.path.join(ROOT, "profiles", "sched_sim.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["sched_sim", "--static", "--modulo"]  # import only: no main() runs
    try:
        spec.loader.exec_module(mod)
    except SystemExit:
        pass
    finally:
        sys.argv = argv
    return mod


def _manifold_lists(n_worlds: int, n_dynamic: int, substeps: int, mix: str):
    sc = [scenes.pile_scene(w, seed=3, mix=mix, n_dynamic=n_dynamic, cols=8) for w in range(n_worlds)]
    worlds = pxb.HostWorlds(sc)
    L, mut, imm = pxb.pack_worlds(worlds)
    ora = parity.OracleBatch(L, mut, imm, parity.sin_table(worlds))
    ora.step(substeps)
    tr = ora.trace()
    out = [tr.world_manifolds(w, int(ora.blocks.header[w]["statManifolds"]))[1] for w in range(n_worlds)]
    worlds.close()
    return out


@pytest.mark.parametrize("mix", ["mixed", "circles"])
def test_periodic_schedule_keeps_every_ring_in_order(mix):
    ss = _sched_sim()
    iterations = 6
    for idm in _manifold_lists(2, 70, 350, mix):
        assert len(idm) > 40
        lists, *_ = ss.rings(idm)
        d = idm[:, 3].astype(np.int64)
        ring = max(int(sum(d[m] for m in lst)) for lst in lists.values())
        # the period ladder of the device: the first feasible one
        t = None
        for II in (ring + 3, ring + 5, ring + 8, 2 * ring + 8):
            t, passes = ss.relax_in_rank_order(idm, II, max_passes=7)
            if t is not None:
                break
        assert t is not None, "no period of the ladder was feasible (the device would fall back to sequential iterations)"
        assert ring <= II  # a period can never be shorter than the longest ring
        # execute: visit k of manifold m occupies rounds [t(m) + k * II, t(m) + k * II + contacts(m))
        for body, lst in lists.items():
            busy_until, expect = -1, 0
            events = sorted((int(t[m]) + k * II, i, k, m) for k in range(iterations) for i, m in enumerate(lst))
            for start, i, k, m in events:
                assert start > busy_until, f"body {body}: two of its visits overlap"
                # pool order inside an iteration, iterations in order
                assert (k, i) == divmod(expect, len(lst)), f"body {body}: visits out of the sequential loop's order"
                expect += 1
                busy_until = start + int(d[m]) - 1


def test_too_short_a_period_is_detected():
    ss = _sched_sim()
    idm = _manifold_lists(1, 70, 350, "mixed")[0]
    lists, *_ = ss.rings(idm)
    d = idm[:, 3].astype(np.int64)
    ring = max(int(sum(d[m] for m in lst)) for lst in lists.values())
    t, _ = ss.relax_in_rank_order(idm, ring - 1, max_passes=7)
    assert t is None  # shorter than a ring: never converges, the device moves on to the next period after 7 passes


# ---- the replay kernel's chunk iterator (pxReplayKernel, px_step.cu: nextChunk) --------------------------------------
#
# Visits are sorted by key = (round mod II) * QN + (round div II); round R of the replay is slot R mod II, period index
# Q = R div II, and the visits alive in it are the phases q with 0 <= Q - q < iterations: ONE contiguous segment
# seg[row + qlo] .. seg[row + qhi1] of the sorted order (row = slot * QN).  The kernel slides the window [qlo, qhi1) by
# increments when the period wraps instead of evaluating max / min; this is the same arithmetic on the host.

def _iterator_chunks(seg, II, QN, depth, iterations):
    n_rounds = depth + (iterations - 1) * II
    out = []
    R = slot = row = Q = 0
    qlo, qhi1 = 0, 1
    beg, end = seg[0], seg[1]
    while True:
        while beg >= end:
            R += 1
            if R >= n_rounds:
                return out
            row += QN
            slot += 1
            if slot == II:
                slot = row = 0
                if Q + 1 >= iterations:
                    qlo += 1
                if Q + 1 < QN:
                    qhi1 += 1
                Q += 1
                assert qlo == max(0, Q + 1 - iterations) and qhi1 == min(Q, QN - 1) + 1
            if qlo < qhi1:
                beg, end = seg[row + qlo], seg[row + qhi1]
        out.append((beg, min(beg + 32, end)))  # a chunk: [base, min(base + 32, limit))
        beg += 32


def test_replay_iterator_visits_every_contact_once_per_iteration():
    rng = random.Random(7)
    for _ in range(3000):
        II, QN, iterations = rng.randint(1, 9), rng.randint(1, 6), rng.randint(1, 7)
        depth = rng.randint((QN - 1) * II + 1, QN * II)
        count = [0] * (II * QN)
        for slot in range(II):
            for q in range(QN):
                if slot + II * q < depth:
                    count[slot * QN + q] = rng.choice([0, 0, 1, 2, 33, 70])
        last = depth - 1
        count[(last % II) * QN + last // II] = max(1, count[(last % II) * QN + last // II])
        seg = [0]
        for c in count:
            seg.append(seg[-1] + c)
        seen = [0] * seg[-1]
        order = []
        for base, lim in _iterator_chunks(seg, II, QN, depth, iterations):
            assert 0 < lim - base <= 32
            for v in range(base, lim):
                seen[v] += 1
            order.append(base)
        assert all(s == iterations for s in seen)
        # the phases alive in a round share its chunks: never more chunks than one per (key, iteration) and 32 visits
        assert -(-seg[-1] * iterations // 32) <= len(order) <= sum(-(-c // 32) for c in count) * iterations
