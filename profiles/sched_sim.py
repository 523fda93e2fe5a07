#!/usr/bin/env python
"""Host simulation of the solver's visit schedule (design aid, not product code).

Steps a few bench worlds with the CPU oracle until the pile has settled, takes the ordered
manifold list of the last substep and measures, per world:
  * ASAP rounds per substep of the dataflow schedule (32-wide and unlimited rounds),
  * L = rounds of one iteration when iterations are NOT pipelined into each other,
  * the spread W of one iteration's visits in the ASAP timeline and the period
    max_m (T(m,k+1) - T(m,k)) a replayed (modulo) schedule would need,
  * the size of the ready set.
Run: python profiles/sched_sim.py [--worlds 4] [--settle 1000] [--iterations 10] [--mix mixed]
     python profiles/sched_sim.py --modulo     the periodic schedule pxReplayKernel derives: longest ring, least
                                               feasible period, relaxation passes in rank order, replay chunks
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def manifold_list(args):
    import parity
    from pdphyzx_b200 import batch as pxb, scenes
    sc = [scenes.pile_scene(w, seed=1, mix=args.mix, iterations=args.iterations, n_dynamic=252) for w in range(args.worlds)]
    worlds = pxb.HostWorlds(sc)
    L, mut, imm = pxb.pack_worlds(worlds)
    ora = parity.OracleBatch(L, mut, imm, parity.sin_table(worlds))
    ora.step(args.settle)
    tr = ora.trace()
    out = []
    for w in range(args.worlds):
        nm = int(ora.blocks.header[w]["statManifolds"])
        _, idm = tr.world_manifolds(w, nm)
        out.append(idm)  # columns: slotA, slotB, bIsDynamic, contactCount
    return out


def simulate(idm, iterations, cap, merge_contacts=False):
    """ASAP dataflow schedule in lock-step rounds of at most `cap` contact visits (FIFO ready queue).
    Returns (rounds, T[m][k] round of contact 0 of visit k, ready-set sizes)."""
    M = len(idm)
    lists = {}
    for m, (a, b, bd, cnt) in enumerate(idm):
        lists.setdefault(("d", a), []).append(m)
        if bd:
            lists.setdefault(("d", b), []).append(m)
    # ring position of each manifold in each of its bodies' lists
    pos = {}
    for key, lst in lists.items():
        for i, m in enumerate(lst):
            pos[(key, m)] = i
    bodies = [[("d", a)] + ([("d", b)] if bd else []) for (a, b, bd, cnt) in idm]
    head = {key: 0 for key in lists}      # index into list of the body's next visit
    iters = {key: 0 for key in lists}     # iteration of the body's next visit
    T = np.full((M, iterations), -1, np.int64)
    queue = []                            # (m, contactIndex)

    def ready(m):
        k = None
        for key in bodies[m]:
            if lists[key][head[key]] != m:
                return None
            if k is None:
                k = iters[key]
            elif k != iters[key]:
                return None
        return k if k is not None and k < iterations else None

    for m in range(M):
        if ready(m) is not None:
            queue.append((m, 0))
    rounds = 0
    sizes = []
    while queue:
        sizes.append(len(queue))
        now, queue = queue[:cap], queue[cap:]
        pushed = []
        for (m, ci) in now:
            cnt = 1 if merge_contacts else int(idm[m][3])
            if ci == 0:
                T[m, iters[bodies[m][0]]] = rounds
            if ci + 1 < cnt:
                pushed.append((m, ci + 1))
                continue
            for key in bodies[m]:
                head[key] += 1
                if head[key] == len(lists[key]):
                    head[key] = 0
                    iters[key] += 1
            for key in bodies[m]:
                nxt = lists[key][head[key]]
                if ready(nxt) is not None and (nxt, 0) not in pushed:
                    pushed.append((nxt, 0))
        queue += pushed
        rounds += 1
    assert (T >= 0).all(), "schedule did not visit everything"
    return rounds, T, sizes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=4)
    ap.add_argument("--settle", type=int, default=1000)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--mix", default="mixed")
    args = ap.parse_args()
    for w, idm in enumerate(manifold_list(args)):
        M = len(idm)
        contacts = int(idm[:, 3].sum())
        r32, T32, sz = simulate(idm, args.iterations, 32)
        rinf, Tinf, szi = simulate(idm, args.iterations, 1 << 30)
        r1, _, _ = simulate(idm, 1, 1 << 30)
        r1c, _, _ = simulate(idm, 1, 32)
        rm, _, _ = simulate(idm, args.iterations, 32, merge_contacts=True)
        k = min(3, args.iterations - 1)
        period32 = int((T32[:, k] - T32[:, k - 1]).max())
        periodinf = int((Tinf[:, k] - Tinf[:, k - 1]).max())
        spread32 = int(T32[:, k].max() - T32[:, k].min() + 1)
        spreadinf = int(Tinf[:, k].max() - Tinf[:, k].min() + 1)
        print(f"world {w}: M={M} contact visits/iter={contacts} | ASAP rounds cap32 {r32} ({r32 / args.iterations:.1f}/it) "
              f"unlimited {rinf} ({rinf / args.iterations:.1f}/it) | one iteration alone: {r1} (cap32 {r1c}) | "
              f"period cap32 {period32} unlimited {periodinf} | spread of iteration {k}: cap32 {spread32} unlimited {spreadinf} | "
              f"mean ready {np.mean(sz):.1f} (unlimited {np.mean(szi):.1f}) | both contacts per visit: {rm} rounds")


if __name__ == "__main__" and "--static" not in sys.argv and "--modulo" not in sys.argv:
    main()


# ---------------------------------------------------------------- static (replayed) schedule
def static_levels(idm):
    """lev[m] = longest-path level of manifold m inside ONE iteration (contact visits as unit of
    time: a 2-contact manifold occupies lev and lev + 1); span per body; lists."""
    M = len(idm)
    last = {}
    lev = np.zeros(M, np.int64)
    first_of, last_of = {}, {}
    for m, (a, b, bd, cnt) in enumerate(idm):
        l = 0
        for key in [a] + ([b] if bd else []):
            if key in last:
                p = last[key]
                l = max(l, lev[p] + int(idm[p][3]))
        lev[m] = l
        for key in [a] + ([b] if bd else []):
            last[key] = m
            first_of.setdefault(key, m)
            last_of[key] = m
    return lev, first_of, last_of


def alap_refine(idm, lev, first_of, last_of):
    """push every manifold as late as its successors inside the iteration allow (sinks stay)."""
    M = len(idm)
    nxt = {}
    out = lev.copy()
    for m in range(M - 1, -1, -1):
        a, b, bd, cnt = idm[m]
        lim = None
        for key in [a] + ([b] if bd else []):
            if key in nxt:
                s = nxt[key]
                v = out[s] - int(cnt)
                lim = v if lim is None else min(lim, v)
        if lim is not None:
            out[m] = lim
        for key in [a] + ([b] if bd else []):
            nxt[key] = m
    return out


def replay_rounds(idm, lev, first_of, last_of, iterations, period=None, cap=32):
    d = idm[:, 3].astype(np.int64)
    span = max(int(lev[last_of[k]] + d[last_of[k]] - lev[first_of[k]]) for k in first_of)
    P = max(span, period or 0)
    # contact visits: (level, manifold)
    lv = np.concatenate([lev, lev[d > 1] + 1])
    tmax = int(lv.max()) + (iterations - 1) * P + 1
    counts = np.zeros(tmax, np.int64)
    for k in range(iterations):
        np.add.at(counts, lv + k * P, 1)
    rounds = int(np.ceil(counts / cap).sum())
    return span, P, tmax, rounds, float(counts.sum() / max(rounds, 1))


def main_static():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=4)
    ap.add_argument("--settle", type=int, default=1000)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--mix", default="mixed")
    ap.add_argument("--static", action="store_true")
    args = ap.parse_args()
    for w, idm in enumerate(manifold_list(args)):
        lev, fo, lo = static_levels(idm)
        r32, _, _ = simulate(idm, args.iterations, 32)
        line = f"world {w}: M={len(idm)} ASAP dynamic {r32} rounds | levels {int(lev.max()) + 1}"
        for name, lv in (("asap", lev), ("alap", alap_refine(idm, lev, fo, lo))):
            span, P, tmax, rounds, lanes = replay_rounds(idm, lv, fo, lo, args.iterations)
            line += f" | {name}: period {span} steps {tmax} rounds {rounds} lanes {lanes:.1f}"
            for extra in (2, 4, 8):
                _, P2, tmax2, rounds2, lanes2 = replay_rounds(idm, lv, fo, lo, args.iterations, period=span + extra)
                line += f" [P+{extra}: {rounds2}]"
        print(line)


if __name__ == "__main__" and "--static" in sys.argv:
    main_static()


# ---------------------------------------------------------------- periodic (modulo) schedule: pxReplayKernel's
def rings(idm):
    """Per manifold: predecessor in its A body's ring and in its B body's ring (-1: static B), and whether that edge
    closes the ring (carries the iteration boundary)."""
    M = len(idm)
    lists = {}
    for m, (a, b, bd, cnt) in enumerate(idm):
        lists.setdefault(int(a), []).append(m)
        if bd:
            lists.setdefault(int(b), []).append(m)
    pA = np.zeros(M, np.int64); wA = np.zeros(M, np.int64)
    pB = np.full(M, -1, np.int64); wB = np.zeros(M, np.int64)
    for body, lst in lists.items():
        for i, m in enumerate(lst):
            p, w = lst[i - 1], (1 if i == 0 else 0)
            if int(idm[m][0]) == body:
                pA[m], wA[m] = p, w
            else:
                pB[m], wB[m] = p, w
    return lists, pA, wA, pB, wB


def relax_in_rank_order(idm, II, max_passes=40):
    """Least start rounds t(m) with t(m) >= t(p) + contacts(p) - [closing edge] * II, relaxed in pool (rank) order as
    the device does.  Returns (t, passes) or (None, passes) if the period is infeasible."""
    M = len(idm)
    d = idm[:, 3].astype(np.int64)
    _, pA, wA, pB, wB = rings(idm)
    t = np.zeros(M, np.int64)
    for p in range(1, max_passes + 1):
        moved = False
        for m in range(M):
            v = max(t[m], t[pA[m]] + d[pA[m]] - wA[m] * II)
            if pB[m] >= 0:
                v = max(v, t[pB[m]] + d[pB[m]] - wB[m] * II)
            if v != t[m]:
                t[m] = v
                moved = True
        if not moved:
            return t, p
        if t.max() > 4000:
            break
    return None, max_passes


def replay_chunks(idm, t, II, iterations, cap=32):
    d = idm[:, 3].astype(np.int64)
    lv = np.concatenate([t, t[d > 1] + 1])
    rounds = int(lv.max()) + (iterations - 1) * II + 1
    counts = np.zeros(rounds, np.int64)
    for k in range(iterations):
        np.add.at(counts, lv + k * II, 1)
    chunks = int(np.ceil(counts / cap).sum())
    return rounds, chunks, float(counts.sum() / max(chunks, 1))


def main_modulo():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=6)
    ap.add_argument("--settle", type=int, default=1000)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--mix", default="mixed")
    ap.add_argument("--modulo", action="store_true")
    args = ap.parse_args()
    for w, idm in enumerate(manifold_list(args)):
        lists, *_ = rings(idm)
        d = idm[:, 3].astype(np.int64)
        ring = max(sum(d[m] for m in l) for l in lists.values())
        dyn, _, _ = simulate(idm, args.iterations, 32)
        least = next(II for II in range(max(ring, 1), ring + 200) if relax_in_rank_order(idm, II)[0] is not None)
        line = f"world {w}: M={len(idm)} visits/it={int(d.sum())} longest ring={ring} dataflow rounds={dyn} | least period={least}"
        for II in (least, ring + 3, ring + 5):
            t, passes = relax_in_rank_order(idm, II)
            if t is None:
                line += f" | II={II}: infeasible"
                continue
            rounds, chunks, lanes = replay_chunks(idm, t, II, args.iterations)
            line += f" | II={II}: passes {passes} depth {int(t.max()) + 1} rounds {rounds} chunks {chunks} lanes {lanes:.1f}"
        print(line, flush=True)


if __name__ == "__main__" and "--modulo" in sys.argv:
    main_modulo()
