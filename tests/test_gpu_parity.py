"""GPU parity tests: the CUDA world step through the C ABI against the CPU oracle and against
the compiled reference -- ordered AABB-pass pair lists, ordered manifolds and body state, bit
for bit (0 ulp; the kernel is built with -fmad=false).  Tolerance stated by the north_star for
positions/velocities is therefore met with margin: the test asserts exact equality."""
import numpy as np
import pytest

import parity
from oracle import oraclebind, refbind
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes, sharding

pytestmark = pytest.mark.gpu
bits = parity.bits


def side_by_side(sc, **kw):
    msg, stats = parity.run_side_by_side(sc, **kw)
    assert msg is None, msg
    return stats


@pytest.mark.parametrize("mix,unit", [("circles", 1.0), ("mixed", 1.0), ("mixed", 100.0), ("mixed", 1000.0), ("polygons", 1.0)])
def test_pile_every_substep(mix, unit):
    """T1/T2: 252 dynamic + 4 static bodies, compared after every substep from free fall into a
    settled pile."""
    stats = side_by_side([scenes.pile_scene(w, mix=mix, unit=unit) for w in range(3)], substeps=700)
    assert stats["manifolds"].min() > 300


@pytest.mark.parametrize("threads", [128, 256])
@pytest.mark.parametrize("fused", [1, 5, 8])
def test_fused_substeps_and_cta_shapes(threads, fused):
    side_by_side([scenes.pile_scene(w, mix="mixed") for w in range(2)], substeps=400, k=fused, threads=threads, check_every=5)


def test_serial_solver_equals_dataflow_solver():
    side_by_side([scenes.pile_scene(0, mix="mixed")], substeps=250, flags=pxb.FLAG_SERIAL_SOLVER, check_every=10)


@pytest.mark.parametrize("iterations", [0, 1, 8, 20, 255])
def test_iteration_counts(iterations):
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=60, cols=8, iterations=iterations) for w in range(2)]
    side_by_side(sc, substeps=300 if iterations < 100 else 120, check_every=10)


def test_polygon_stack_and_small_drop():
    side_by_side([scenes.stack_scene(w) for w in range(4)], substeps=600, check_every=3)
    side_by_side([scenes.drop_scene(w) for w in range(4)], substeps=600, check_every=3)


def test_sleep_and_wake():
    sc = [scenes.pile_scene(w, mix="circles", n_dynamic=30, cols=6, unit=1000.0) for w in range(4)]
    stats = side_by_side(sc, substeps=2500, k=5, check_every=10)
    assert stats["sleeping"].max() > 0


def test_ragged_batch_and_empty_worlds():
    """Worlds of different sizes in one batch, including one with no dynamic bodies at all."""
    sc = [scenes.pile_scene(0, mix="mixed", n_dynamic=5, cols=3), scenes.pile_scene(1, mix="circles", n_dynamic=0),
          scenes.pile_scene(2, mix="mixed", n_dynamic=200, cols=16), scenes.pile_scene(3, mix="polygons", n_dynamic=1, cols=1)]
    side_by_side(sc, substeps=400, check_every=4)


def test_rotated_static_ground():
    """examples/sprites rotates its static ground with setOrientation between frames."""
    sc = scenes.pile_scene(0, mix="mixed", n_dynamic=60, cols=8)
    sc.angle[0] = np.float32(0.12)
    side_by_side([sc], substeps=500, check_every=5)


def test_capacity_overflow_is_flagged_not_fatal():
    sc = [scenes.pile_scene(0, mix="circles", n_dynamic=60, cols=8)]
    worlds = pxb.HostWorlds(sc)
    for caps, want in ((dict(max_manifolds=32), pxb.OVERFLOW_MANIFOLDS), (dict(max_pairs=64), pxb.OVERFLOW_PAIRS)):
        b = pxb.Batch(worlds, **caps)
        for _ in range(8):
            b.step(50)
        blocks = b.download_blocks()
        assert int(blocks.header[0]["statOverflow"]) & want
        # dropping the excess must leave the step internally consistent (bit 31 = a solver visit
        # count that does not add up, e.g. a manifold slot of an earlier substep linked by mistake)
        assert not int(blocks.header[0]["statOverflow"]) & 0x80000000
        assert b.readback(worlds) == 1  # one world overflowed
        assert worlds.stats(0)["overflow"] & want
        b.close()
    worlds.close()


@pytest.mark.parametrize("mix,unit", [("mixed", 1.0), ("circles", 1000.0)])
def test_against_compiled_reference_through_host_structs(reflib_available, mix, unit):
    """Full drop-in path: host PxWorld structs -> pxBatchNew -> pxBatchStep -> pxBatchReadback
    -> PxBody fields, against the reference stepping its own PxWorld."""
    sc = scenes.pile_scene(4, mix=mix, unit=unit)
    ref = refbind.RefWorld(sc, "wide")
    worlds = pxb.HostWorlds([sc])
    b = pxb.Batch(worlds)
    for _ in range(40):
        b.step(10)
        ref.substep(10)
        assert b.readback(worlds) == 0
        for dyn in (True,):
            rr, rf, ro = ref.bodies(dyn)
            hr, hf, ho = worlds.bodies(0, dyn)
            np.testing.assert_array_equal(ro, ho)
            np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]))  # state, matrix, AABB
            np.testing.assert_array_equal(rf[ro], hf[ro])
    _, ids = ref.manifolds()
    assert worlds.stats(0)["manifolds"] == len(ids)
    b.close()
    worlds.close()


def test_arithmetic_contract_on_device(reflib_available):
    """T0: fast rsqrt over all 2^32 bit patterns (block sums vs the oracle, itself checked
    against the reference on slices), and dense sweeps of rcp/sqrt/div/sin/cos vs the reference."""
    dev = pxb.device_rsqrt_sweep()
    ora = oraclebind.rsqrt_sweep(0, 4096)
    np.testing.assert_array_equal(dev, ora)
    ref = refbind.load("wide", 1.0)
    table = np.ctypeslib.as_array(ref.ref_sin_table(), shape=(256,)).copy()
    rng = np.random.default_rng(5)
    x = np.concatenate([np.linspace(-60, 60, 400001), rng.uniform(-1e4, 1e4, 100000)]).astype(np.float32)
    y = rng.uniform(1e-3, 1e3, x.size).astype(np.float32)
    for op, fn in ((1, ref.ref_math_rcp), (2, ref.ref_math_sqrt), (4, ref.ref_math_sin), (5, ref.ref_math_cos)):
        xin = np.abs(x) if op < 4 else x
        want = np.zeros_like(xin)
        fn(xin, want, xin.size)
        np.testing.assert_array_equal(bits(pxb.device_math(op, xin, None, table)), bits(want))
    want = np.zeros_like(x)
    ref.ref_math_div(np.abs(x), y, want, x.size)
    np.testing.assert_array_equal(bits(pxb.device_math(3, np.abs(x), y, table)), bits(want))


def test_config2_4096_circle_worlds_contact_sets():
    """BASELINE config 2: 4,096 batched circle-pile worlds on one GPU; contact pair sets (and
    everything else) checked bit-exact per world for a sample of worlds against the oracle, and
    size-independent properties for the whole batch: determinism, independence of a world from
    its batch, no overflow."""
    n = 4096
    sc = [scenes.pile_scene(w, mix="circles") for w in range(n)]
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds, flags=pxb.FLAG_TRACE)
    start = b.download_blocks()
    mut0, imm0 = start.mut.copy(), start.imm.copy()
    for _ in range(8):
        b.step(50)
    blocks = b.download_blocks()
    assert int((blocks.header["statOverflow"] != 0).sum()) == 0
    assert blocks.header["statManifolds"].mean() > 250
    final = blocks.mut.copy()
    checksum = sharding.state_checksum(blocks)
    trace = b.trace()
    # (1) sampled worlds against the oracle, pair sets and manifolds of the last substep included
    table = parity.sin_table(worlds)
    sample = [0, 1, 777, 2048, 4095]
    for w in sample:
        L1 = pxb.compute_layout(1, pxb.make_desc(1.0, maxDynamic=b.L.maxDynamic, maxStatic=b.L.maxStatic,
                                                 maxVertices=b.L.maxVertices, maxManifolds=b.L.maxManifolds, maxPairs=b.L.maxPairs))
        m1 = mut0[w * b.L.mutStride:(w + 1) * b.L.mutStride].copy()
        i1 = imm0[w * b.L.immStride:(w + 1) * b.L.immStride].copy()
        t1 = np.zeros(L1.traceStride, np.uint8)
        oraclebind.substeps(L1, m1, i1, table, 400, trace=t1)
        ob = pxb.Blocks(L1, m1, i1)
        dev_w = pxb.Blocks(L1, final[w * b.L.mutStride:(w + 1) * b.L.mutStride], None)
        msg = parity.compare_state(dev_w, ob, f"world {w}")
        assert msg is None, msg
        ot = pxb.Trace(L1, t1)
        npairs, nm = int(ob.header[0]["statPairs"]), int(ob.header[0]["statManifolds"])
        for x, y in zip(trace.world_pairs(w, npairs), ot.world_pairs(0, npairs)):
            np.testing.assert_array_equal(x, y)
        dr, di = trace.world_manifolds(w, nm)
        orr, oi = ot.world_manifolds(0, nm)
        np.testing.assert_array_equal(di, oi)
        np.testing.assert_array_equal(bits(dr), bits(orr))
    # (2) determinism: the same batch from the same start gives the same checksum
    b.host_mutable()[:] = mut0
    b.mutable_h2d()
    for _ in range(8):
        b.step(50)
    assert sharding.state_checksum(b.download_blocks()) == checksum
    b.close()
    worlds.close()


def test_config3_mixed_worlds_sharded_checksum():
    """BASELINE configs 3/4 in miniature: a mixed-shape batch split into two 'rank' shards gives
    the same per-world results as the unsplit batch (checksum of checksums), no overflow."""
    n = 512
    sc = [scenes.pile_scene(w, mix="mixed") for w in range(n)]
    worlds = pxb.HostWorlds(sc)
    whole = pxb.Batch(worlds)
    for _ in range(6):
        whole.step(50)
    wb = whole.download_blocks()
    assert int((wb.header["statOverflow"] != 0).sum()) == 0
    total = sharding.state_checksum(wb)
    whole.close()
    parts = 0
    for r in range(2):
        ids = sharding.shard_range(n, r, 2)
        shard = pxb.Batch(n_worlds=len(ids), unit=1.0, max_dynamic=wb.L.maxDynamic, max_static=wb.L.maxStatic,
                          max_vertices=wb.L.maxVertices)
        shard.upload(worlds, first=0, src_first=ids[0], n=len(ids))
        for _ in range(6):
            shard.step(50)
        parts = (parts + sharding.state_checksum(shard.download_blocks())) & 0xFFFFFFFFFFFFFFFF
        shard.close()
    assert parts == total
    worlds.close()


@pytest.mark.parametrize("unit", [1.0, 1000.0])
def test_edge_scene_big_polygons_and_many_statics(unit):
    """Polygons with 9..32 vertices (the generic SAT loops), 40 statics (long static window, the
    literal non-monotone lower-bound search), all three shape classes -- every 3rd substep."""
    stats = side_by_side([scenes.edge_scene(w, unit=unit) for w in range(3)], substeps=450, check_every=3)
    assert stats["manifolds"].max() > 40


def test_maximum_body_count():
    """255 dynamic bodies: the uint8 slot / sentinel ceiling of the reference."""
    side_by_side([scenes.pile_scene(w, mix="mixed", n_dynamic=255) for w in range(2)], substeps=300, check_every=10)


@pytest.mark.parametrize("solver", ["static", "dynamic"])
@pytest.mark.parametrize("mix,unit,fused,threads", [("mixed", 1.0, 1, 256), ("mixed", 1000.0, 5, 128), ("circles", 1.0, 8, 256),
                                                    ("polygons", 100.0, 5, 256)])
def test_split_pipeline_every_substep(mix, unit, fused, threads, solver):
    """The split pipeline (sweep kernel | solver kernel per substep, hand-over records in global
    memory) forced on small batches, with both of its solvers -- the replayed periodic schedule
    (default) and the dataflow scheduler kernel: pair lists, manifolds and state against the oracle."""
    flags = pxb.FLAG_SPLIT | (pxb.FLAG_DYNAMIC_SOLVER if solver == "dynamic" else 0)
    stats = side_by_side([scenes.pile_scene(w, mix=mix, unit=unit) for w in range(3)], substeps=400, k=fused,
                         threads=threads, flags=flags, check_every=1 if fused == 1 else 2)
    assert stats["manifolds"].min() > 100


@pytest.mark.parametrize("iterations", [0, 1, 2, 8, 20, 255])
def test_static_schedule_iteration_counts(iterations):
    """The replayed schedule at every iteration count the API allows (uint8): 0 (no solve), 1 (only the prologue of
    the period), 255 (254 periods), on full-size 252-body worlds."""
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=252, iterations=iterations) for w in range(2)]
    side_by_side(sc, substeps=120 if iterations < 255 else 40, k=5, flags=pxb.FLAG_SPLIT, check_every=5)


def test_static_schedule_stacks_ragged_worlds_and_overflow():
    """Worlds the schedule builder must cope with: tall polygon stacks (one long dependency chain: deep
    iterations, short periods), a ragged batch with an empty world and a single-body world (no manifolds: nothing
    to schedule), and a world whose manifold capacity overflows (the kept manifolds still form consistent rings:
    the replay's visit count adds up, bit 31 stays clear)."""
    sc = [scenes.stack_scene(w, n_dynamic=40 + 30 * w) for w in range(3)]
    side_by_side(sc, substeps=300, k=1, flags=pxb.FLAG_SPLIT, check_every=3)
    sc = [scenes.pile_scene(0, mix="mixed", n_dynamic=200, cols=16), scenes.pile_scene(1, mix="circles", n_dynamic=0),
          scenes.pile_scene(2, mix="circles", n_dynamic=1, cols=1), scenes.pile_scene(3, mix="polygons", n_dynamic=64, cols=8)]
    side_by_side(sc, substeps=200, k=5, flags=pxb.FLAG_SPLIT, check_every=5)
    worlds = pxb.HostWorlds([scenes.pile_scene(0, mix="circles", n_dynamic=60, cols=8)])
    b = pxb.Batch(worlds, flags=pxb.FLAG_SPLIT, max_manifolds=32)
    for _ in range(8):
        b.step(50)
    flags = int(b.download_blocks().header[0]["statOverflow"])
    assert flags & pxb.OVERFLOW_MANIFOLDS and not flags & 0x80000000
    b.close()
    worlds.close()


def test_split_and_fused_pipelines_agree_on_a_large_batch():
    """600 mixed worlds (more than the split threshold of 4 worlds per SM) stepped by the default
    pipeline and by the forced fused one: identical mutable blocks."""
    n = 600
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=100, cols=10) for w in range(n)]
    wa, wb = pxb.HostWorlds(sc), pxb.HostWorlds(sc)
    a = pxb.Batch(wa)
    b = pxb.Batch(wb, flags=pxb.FLAG_FUSED)
    for k in (50, 5, 1, 8, 50):
        a.step(k)
        b.step(k)
        np.testing.assert_array_equal(np.asarray(a.download_blocks().mut), np.asarray(b.download_blocks().mut))
    for x in (a, b):
        x.close()
    for x in (wa, wb):
        x.close()
