"""The CPU restatement oracle (oracle/px_oracle.c) pinned against the compiled reference
(oracle/_ref): ordered AABB-pass pair lists, ordered manifolds and the full body state, bit for
bit, substep by substep (parity tiers T1/T2 of SURVEY 8c), plus arithmetic (T0)."""
import numpy as np
import pytest

import parity
from oracle import oraclebind, refbind
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes

bits = parity.bits


def run_vs_reference(sc, substeps, variant="wide", every=1):
    ref = refbind.RefWorld(sc, variant)
    worlds = pxb.HostWorlds([sc])
    table = parity.sin_table(worlds)
    L, mut, imm = pxb.pack_worlds(worlds)
    blocks = pxb.Blocks(L, mut, imm)
    trace = np.zeros(L.traceStride, np.uint8)
    seen_manifolds = 0
    for s in range(substeps):
        a, b, d = ref.substep_traced()
        oraclebind.substeps(L, mut, imm, table, 1, trace=trace)
        if s % every:
            continue
        tr = pxb.Trace(L, trace)
        h = blocks.header[0]
        pa, pb, pd = tr.world_pairs(0, int(h["statPairs"]))
        assert len(a) == len(pa), f"substep {s}: pair count"
        np.testing.assert_array_equal(a, pa)
        np.testing.assert_array_equal(b, pb)
        np.testing.assert_array_equal(d, pd)
        rrows, rids = ref.manifolds()
        orows, oids = tr.world_manifolds(0, int(h["statManifolds"]))
        np.testing.assert_array_equal(rids.astype(np.int64), oids, err_msg=f"substep {s}: manifold ids")
        rrows = rrows.copy()
        rrows[rids[:, 3] < 2, 8:10] = 0  # second contact slot is garbage when unused
        np.testing.assert_array_equal(bits(rrows), bits(orows), err_msg=f"substep {s}: manifold values")
        rb, rfl, ro = ref.bodies(True)
        n = len(ro)
        np.testing.assert_array_equal(ro, blocks.order[0, :n], err_msg=f"substep {s}: order")
        mine = blocks.dyn[0].T
        np.testing.assert_array_equal(bits(rb[ro][:, :14]), bits(mine[ro]), err_msg=f"substep {s}: state")
        np.testing.assert_array_equal(rfl[ro], blocks.flags[0, ro], err_msg=f"substep {s}: flags")
        seen_manifolds = max(seen_manifolds, len(rids))
    worlds.close()
    return seen_manifolds, blocks


@pytest.mark.parametrize("unit", [1.0, 100.0, 1000.0])
def test_circle_pile_free_running(reflib_available, unit):
    m, _ = run_vs_reference(scenes.pile_scene(1, mix="circles", unit=unit), 700, every=7)
    assert m > 300


@pytest.mark.parametrize("unit", [1.0, 1000.0])
def test_mixed_pile_free_running(reflib_available, unit):
    m, _ = run_vs_reference(scenes.pile_scene(2, mix="mixed", unit=unit), 500, every=5)
    assert m > 250


def test_polygon_stack(reflib_available):
    m, _ = run_vs_reference(scenes.stack_scene(3), 400, every=4)
    assert m > 50


def test_every_substep_small_world_unmodified_reference(reflib_available):
    """<= 64 bodies / manifolds: checked against the UNMODIFIED (narrow) reference build."""
    run_vs_reference(scenes.drop_scene(4, n_dynamic=18), 600, variant="narrow", every=1)


def test_sleeping_bodies_and_index_mixup(reflib_available):
    """Long enough for bodies to fall asleep and be woken again (px_world.c:141-219)."""
    sc = scenes.pile_scene(9, mix="circles", n_dynamic=30, cols=6, unit=1000.0)
    _, blocks = run_vs_reference(sc, 2500, every=25)
    assert int(blocks.header[0]["statSleeping"]) > 0


def test_oracle_worlds_are_independent():
    """Stepping a batch equals stepping its worlds one by one (no cross-world state)."""
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=40, cols=6) for w in range(4)]
    worlds = pxb.HostWorlds(sc)
    table = parity.sin_table(worlds)
    L, mut, imm = pxb.pack_worlds(worlds)
    together = mut.copy()
    oraclebind.substeps(L, together, imm, table, 120)
    apart = mut.copy()
    for w in (2, 0, 3, 1):
        oraclebind.substeps(L, apart, imm, table, 120, first=w, n=1)
    np.testing.assert_array_equal(together, apart)
    worlds.close()


def test_oracle_overflow_flags():
    sc = [scenes.pile_scene(0, mix="circles", n_dynamic=60, cols=8)]
    worlds = pxb.HostWorlds(sc)
    table = parity.sin_table(worlds)
    for caps, want in ((dict(maxManifolds=32), pxb.OVERFLOW_MANIFOLDS), (dict(maxPairs=64), pxb.OVERFLOW_PAIRS)):
        L, mut, imm = pxb.pack_worlds(worlds, pxb.make_desc(1.0, **caps))
        oraclebind.substeps(L, mut, imm, table, 400)
        assert int(pxb.Blocks(L, mut, imm).header[0]["statOverflow"]) & want
    L, mut, imm = pxb.pack_worlds(worlds)
    oraclebind.substeps(L, mut, imm, table, 400)
    assert int(pxb.Blocks(L, mut, imm).header[0]["statOverflow"]) == 0
    worlds.close()


@pytest.mark.parametrize("op", range(6))
def test_oracle_math_matches_reference(reflib_available, op):
    rng = np.random.default_rng(100 + op)
    x = np.concatenate([rng.uniform(1e-6, 1e6, 50000), rng.uniform(-100, 100, 50000)]).astype(np.float32)
    if op < 4:
        x = np.abs(x)
    y = rng.uniform(1e-3, 1e3, x.size).astype(np.float32)
    ref = refbind.load("wide", 1.0)
    table = np.ctypeslib.as_array(ref.ref_sin_table(), shape=(256,)).copy()
    out_ref = np.zeros_like(x)
    if op == 3:
        ref.ref_math_div(x, y, out_ref, x.size)
    else:
        [ref.ref_math_rsqrt, ref.ref_math_rcp, ref.ref_math_sqrt, None, ref.ref_math_sin, ref.ref_math_cos][op](x, out_ref, x.size)
    np.testing.assert_array_equal(bits(oraclebind.math(op, x, y, table)), bits(out_ref))


def test_oracle_rsqrt_sweep_matches_reference_on_a_slice(reflib_available):
    """Block sums used by the exhaustive device test, checked against the reference's own
    pxFastRsqrt on two 2^20-pattern blocks."""
    ref = refbind.load("wide", 1.0)
    for block in (1012, 2040, 3051):  # ordinary values, +inf/NaN inputs, inputs that yield NaN
        pats = (np.arange(1 << 20, dtype=np.uint64) + (np.uint64(block) << np.uint64(20))).astype(np.uint32)
        x = pats.view(np.float32)
        out = np.zeros_like(x)
        ref.ref_math_rsqrt(x, out, x.size)
        words = np.where(np.isnan(out), np.uint32(0x7FC00000), out.view(np.uint32))
        assert int(words.astype(np.uint64).sum()) == int(oraclebind.rsqrt_sweep(block, 1)[0])


@pytest.mark.parametrize("unit", [1.0, 1000.0])
def test_edge_scene_big_polygons_and_many_statics(reflib_available, unit):
    """Polygons with 9..32 vertices, 40 static boxes with a non-monotone max.x along the min.x
    order (the reference's static lower-bound search is literal, px_body_array.c:104-131)."""
    sc = scenes.edge_scene(5, unit=unit)
    assert int(sc.nverts.max()) > 20 and sc.n_static == 40
    m, _ = run_vs_reference(sc, 450, every=3)
    assert m > 40
