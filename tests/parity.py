"""Shared helpers for the parity tests: run the device and the CPU oracle side by side on the
same packed worlds and compare everything bit for bit."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pdphyzx_b200 import batch as pxb  # noqa: E402
from oracle import oraclebind  # noqa: E402


def sin_table(worlds: pxb.HostWorlds) -> np.ndarray:
    return np.ctypeslib.as_array(worlds.lib.pxh_sin_table(), shape=(256,)).copy()


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint32)


class OracleBatch:
    """The CPU oracle stepping a private copy of a batch's blocks."""

    def __init__(self, layout, mut: np.ndarray, imm: np.ndarray, table: np.ndarray):
        self.L = layout
        self.mut = mut.copy()
        self.imm = imm.copy()
        self.table = table
        self.blocks = pxb.Blocks(layout, self.mut, self.imm)
        self.trace_buf = np.zeros(layout.nWorlds * layout.traceStride, np.uint8)

    def step(self, k: int = 1):
        oraclebind.substeps(self.L, self.mut, self.imm, self.table, k, trace=self.trace_buf)

    def trace(self) -> pxb.Trace:
        return pxb.Trace(self.L, self.trace_buf)


def compare_state(dev: pxb.Blocks, ora: pxb.Blocks, what: str = "") -> str | None:
    """None if the mutable state of every world matches bitwise, else a description."""
    n = dev.n
    for w in range(n):
        hd, ho = dev.header[w], ora.header[w]
        nd = int(ho["nDynamic"])
        for key in ("statManifolds", "statPairs", "statSleeping", "statOverflow", "substeps"):
            if int(hd[key]) != int(ho[key]):
                return f"{what} world {w}: header.{key} device {int(hd[key])} oracle {int(ho[key])}"
        od, oo = dev.order[w, :nd], ora.order[w, :nd]
        if not np.array_equal(od, oo):
            k = int(np.argmax(od != oo))
            return f"{what} world {w}: order differs first at sorted position {k}: device {od[k]} oracle {oo[k]}"
        slots = oo.astype(np.int64)
        a, b = bits(dev.dyn[w][:, slots]), bits(ora.dyn[w][:, slots])
        if not np.array_equal(a, b):
            r, c = np.argwhere(a != b)[0]
            return (f"{what} world {w}: row {pxb.DYN_ROWS[r]} slot {slots[c]} device {dev.dyn[w][r, slots[c]]!r} "
                    f"oracle {ora.dyn[w][r, slots[c]]!r} ({int((a != b).sum())} words differ)")
        if not np.array_equal(dev.flags[w, slots], ora.flags[w, slots]):
            k = int(np.argmax(dev.flags[w, slots] != ora.flags[w, slots]))
            return f"{what} world {w}: flags differ at slot {slots[k]}"
    return None


def compare_trace(dev_t: pxb.Trace, ora_t: pxb.Trace, ora_blocks: pxb.Blocks, what: str = "") -> str | None:
    for w in range(ora_blocks.n):
        h = ora_blocks.header[w]
        npairs, nm = int(h["statPairs"]), int(h["statManifolds"])
        for part, (x, y) in enumerate(zip(dev_t.world_pairs(w, npairs), ora_t.world_pairs(w, npairs))):
            if not np.array_equal(x, y):
                k = int(np.argmax(x != y))
                return f"{what} world {w}: pair list differs (component {part}) first at pair {k}: device {x[k]} oracle {y[k]}"
        dr, di = dev_t.world_manifolds(w, nm)
        orr, oi = ora_t.world_manifolds(w, nm)
        if not np.array_equal(di, oi):
            k = int(np.argwhere(di != oi)[0][0])
            return f"{what} world {w}: manifold ids differ first at manifold {k}: device {di[k]} oracle {oi[k]}"
        if not np.array_equal(bits(dr), bits(orr)):
            k, c = np.argwhere(bits(dr) != bits(orr))[0]
            return f"{what} world {w}: manifold {k} column {c}: device {dr[k, c]!r} oracle {orr[k, c]!r} ids {oi[k]}"
    return None


def run_side_by_side(scenes, substeps: int, k: int = 1, flags: int = 0, threads: int = 0, check_trace: bool = True,
                     check_every: int = 1, **caps):
    """Build `scenes` on the host, step device and oracle `substeps` substeps in launches of k
    and compare.  Returns (None | first mismatch description, stats dict)."""
    worlds = pxb.HostWorlds(scenes)
    b = pxb.Batch(worlds, flags=flags | (pxb.FLAG_TRACE if check_trace else 0), threads=threads, **caps)
    try:
        dev0 = b.download_blocks()
        ora = OracleBatch(b.L, dev0.mut, dev0.imm, sin_table(worlds))
        done = 0
        launches = 0
        while done < substeps:
            kk = min(k, substeps - done)
            b.step(kk)
            ora.step(kk)
            done += kk
            launches += 1
            if launches % check_every == 0 or done >= substeps:
                dev = b.download_blocks()
                msg = compare_state(dev, ora.blocks, f"substep {done}")
                if msg is None and check_trace:
                    msg = compare_trace(b.trace(), ora.trace(), ora.blocks, f"substep {done}")
                if msg:
                    return msg, dict(done=done)
        h = ora.blocks.header
        return None, dict(done=done, manifolds=h["statManifolds"].copy(), pairs=h["statPairs"].copy(),
                          sleeping=h["statSleeping"].copy(), info=b.kernel_info())
    finally:
        b.close()
        worlds.close()


def oracle_substeps_threaded(L, mut: np.ndarray, imm: np.ndarray, table: np.ndarray, k: int, trace: np.ndarray | None = None,
                             threads: int | None = None):
    """pxOracleSubsteps over every world of a block pair, disjoint world ranges on `threads` host
    threads (worlds are independent; ctypes releases the GIL while the C oracle runs)."""
    import threading
    n = L.nWorlds
    threads = max(1, min(threads or (os.cpu_count() or 1), n))
    bounds = [n * t // threads for t in range(threads + 1)]
    errs = []

    def run(a, b):
        try:
            if b > a:
                tr = None if trace is None else trace[a * L.traceStride:]  # the oracle indexes its trace from `first`
                oraclebind.substeps(L, mut, imm, table, k, first=a, n=b - a, trace=tr)
        except Exception as exc:  # pragma: no cover
            errs.append(exc)

    ts = [threading.Thread(target=run, args=(bounds[t], bounds[t + 1])) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
