"""GPU tests of the host-driven control path (BASELINE config 5) and of the px->world->step
facade: queued applyImpulse / applyForce / setGravity / setIterations and the clock-driven step
against the reference's own API calls."""
import numpy as np
import pytest

import parity
from oracle import refbind
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes

pytestmark = pytest.mark.gpu
bits = parity.bits


def compare_with_reference(refs, worlds, b):
    assert b.readback(worlds) == 0
    for w, ref in enumerate(refs):
        rr, rf, ro = ref.bodies(True)
        hr, hf, ho = worlds.bodies(w, True)
        np.testing.assert_array_equal(ro, ho, err_msg=f"world {w}: order")
        np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]), err_msg=f"world {w}: state")
        np.testing.assert_array_equal(rf[ro], hf[ro], err_msg=f"world {w}: flags")


@pytest.mark.parametrize("fused", [1, 8])
def test_rl_style_control_matches_reference(reflib_available, fused):
    """Every batch step the host enqueues a per-world applyImpulse on one body, an applyForce on
    another and a setGravity, then steps K substeps (config 5: K=1 vs fused K=8), UNIT_MM."""
    n = 6
    sc = [scenes.stack_scene(w, n_dynamic=30) for w in range(n)]
    refs = [refbind.RefWorld(s, "wide") for s in sc]
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds)
    rng = np.random.default_rng(1)
    for it in range(60):
        for w in range(n):
            body = int(rng.integers(0, 30))
            ix, iy = float(np.float32(rng.uniform(-3, 3))), float(np.float32(rng.uniform(-8, 0)))
            cx, cy = float(np.float32(rng.uniform(-2, 2))), float(np.float32(rng.uniform(-2, 2)))
            b.apply_impulse(w, body, ix, iy, cx, cy)
            refs[w].lib.ref_body_apply_impulse(refs[w].w, body, ix, iy, cx, cy)
            body2 = int(rng.integers(0, 30))
            fx = float(np.float32(rng.uniform(-50, 50)))
            b.apply_force(w, body2, fx, 0.0, 0.0, cy)
            refs[w].lib.ref_body_apply_force(refs[w].w, body2, fx, 0.0, 0.0, cy)
            gx = float(np.float32(rng.uniform(-1, 1)))
            b.set_gravity(w, gx, 9.8)
            refs[w].lib.ref_set_gravity(refs[w].w, gx, 9.8)
        b.step(fused)
        for ref in refs:
            ref.substep(fused)
        if it % 10 == 9:
            compare_with_reference(refs, worlds, b)
    compare_with_reference(refs, worlds, b)
    b.close()
    worlds.close()


def test_batch_wide_setters_and_ordering(reflib_available):
    n = 4
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=40, cols=6) for w in range(n)]
    refs = [refbind.RefWorld(s, "wide") for s in sc]
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds)
    b.step(100)
    for r in refs:
        r.substep(100)
    # ALL-worlds gravity, then a per-world override submitted later must win for that world
    b.set_gravity(pxb.ALL_WORLDS, 1.0, 5.0)
    b.set_gravity(2, -1.0, 3.0)
    b.set_iterations(pxb.ALL_WORLDS, 4)
    b.set_iterations(1, 12)
    for w, r in enumerate(refs):
        r.lib.ref_set_gravity(r.w, *((-1.0, 3.0) if w == 2 else (1.0, 5.0)))
        r.lib.ref_set_iterations(r.w, 12 if w == 1 else 4)
    # a sleeping-or-not body is woken by an impulse (px_body_api.c:37-42)
    b.apply_impulse(0, 3, 0.5, -0.5, 0.0, 0.0)
    refs[0].lib.ref_body_apply_impulse(refs[0].w, 3, 0.5, -0.5, 0.0, 0.0)
    b.step(150)
    for r in refs:
        r.substep(150)
    compare_with_reference(refs, worlds, b)
    with pytest.raises(pxb.PxError):
        b.apply_impulse(99, 0, 1, 1, 0, 0)
    b.close()
    worlds.close()


def test_world_step_facade_keeps_the_clock_semantics(reflib_available):
    """px->world->step on our host API (1-world device batch behind it) against the reference's
    px->world->step under the same stub clock: first call never steps, 0-3 fixed steps per call,
    5 substeps each (px_clock.c:25-52, px_world.c:422-445)."""
    sc = scenes.pile_scene(3, mix="mixed", n_dynamic=40, cols=6, unit=100.0)
    ref = refbind.RefWorld(sc, "wide")
    worlds = pxb.HostWorlds([sc])
    h, w = worlds.lib, worlds.ptrs[0]
    t = 1000
    fired_ref = fired_dev = 0
    schedule = [20] * 30 + [3, 3, 3, 70, 250, 20, 20, 1, 45] + [20] * 20
    for k, dt_ms in enumerate([0] + schedule):
        t += dt_ms
        ref.lib.ref_set_time_ms(t)
        h.pxh_set_time_ms(t)
        a = ref.lib.ref_step(ref.w)
        b = h.pxh_step(w)
        assert a == b, f"call {k}: stepped flag {a} vs {b}"
        fired_ref += a
        fired_dev += b
        rr, rf, ro = ref.bodies(True)
        hr, hf, ho = worlds.bodies(0, True)
        np.testing.assert_array_equal(ro, ho)
        np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]), err_msg=f"call {k}")
        if k == 20:  # host-side edits between frames are carried to the device by the facade
            ref.lib.ref_body_apply_impulse(ref.w, 5, 40.0, -60.0, 0.0, 0.0)
            h.pxh_body_apply_impulse(w, 5, 40.0, -60.0, 0.0, 0.0)
            ref.lib.ref_body_set_position(ref.w, 1, 7, 300.0, 100.0)
            h.pxh_body_set_position(w, 1, 7, 300.0, 100.0)
    assert fired_dev > 40
    h.pxh_draw_debug(w)  # "Contacts: %d  Sleeping: %d" comes from the device statistics
    assert worlds.stats(0)["manifolds"] == len(ref.manifolds()[1])
    worlds.close()


@pytest.mark.gpu
@pytest.mark.parametrize("unit,mix", [(1.0, "mixed"), (1000.0, "circles")])
def test_reference_side_binding_is_a_drop_in(reflib_available, unit, mix):
    """the reference itself, built with integration/px_world_b200.c in place of
    its pxWorldStep (oracle/_ref/libpxref_b200_*.so), steps reference-built PxWorld structs on
    the GPU through the C ABI.  Same scenes, same stub clock, same host-side edits as the
    unmodified reference on the CPU: every body field bit-identical after every call."""
    if not refbind.available("b200", unit):
        pytest.skip("oracle/_ref/libpxref_b200_*.so not built")
    sc = scenes.pile_scene(5, mix=mix, n_dynamic=70, cols=8, unit=unit)
    cpu = refbind.RefWorld(sc, "wide")
    gpu = refbind.RefWorld(sc, "b200")
    t = 5000
    fired = 0
    schedule = [20] * 40 + [3, 70, 250, 20, 1, 45] + [20] * 30
    for k, dt_ms in enumerate([0] + schedule):
        t += dt_ms
        cpu.lib.ref_set_time_ms(t)
        gpu.lib.ref_set_time_ms(t)
        a = cpu.lib.ref_step(cpu.w)
        b = gpu.lib.ref_step(gpu.w)
        assert a == b, f"call {k}: stepped flag {a} vs {b}"
        fired += b
        cr, cf, co = cpu.bodies(True)
        gr, gf, go = gpu.bodies(True)
        np.testing.assert_array_equal(co, go, err_msg=f"call {k}: iteration order")
        np.testing.assert_array_equal(cf, gf, err_msg=f"call {k}: flags")
        np.testing.assert_array_equal(bits(cr[co][:, :18]), bits(gr[co][:, :18]), err_msg=f"call {k}")
        if k == 25:  # host-side API calls and field pokes between frames travel with the world
            for r in (cpu, gpu):
                r.lib.ref_body_apply_impulse(r.w, 9, 3.0, -6.0, 0.0, 0.0)
                r.lib.ref_body_set_position(r.w, 1, 11, 2.0 * unit, 1.0 * unit)
                r.lib.ref_set_gravity(r.w, 0.5, 9.0)
    assert fired > 50
    cpu.close()
    gpu.close()


@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_host_buffer_step_is_the_same_step(chunks):
    """pxBatchStepHost (H2D, K substeps, D2H per world range, pipelined on helper streams) leaves
    exactly the bytes that separate H2D / step / D2H calls leave, in host memory and in HBM."""
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=40, cols=6) for w in range(13)]
    wa, wb = pxb.HostWorlds(sc), pxb.HostWorlds(sc)
    a, b = pxb.Batch(wa), pxb.Batch(wb)
    for _ in range(6):
        a.mutable_h2d()
        a.step(5)
        a.mutable_d2h()
        a.sync()
        b.step_host(5, chunks)
        b.sync()
        ha, hb = a.host_blocks(), b.host_blocks()
        np.testing.assert_array_equal(np.asarray(ha.dyn).view(np.uint32), np.asarray(hb.dyn).view(np.uint32))
        np.testing.assert_array_equal(ha.order, hb.order)
        np.testing.assert_array_equal(ha.flags, hb.flags)
    da, db = a.download_blocks(), b.download_blocks()
    np.testing.assert_array_equal(np.asarray(da.dyn).view(np.uint32), np.asarray(db.dyn).view(np.uint32))
    assert int(da.header["substeps"][0]) == 30
    for x in (a, b):
        x.close()
    for x in (wa, wb):
        x.close()


def test_spawn_and_despawn_between_frames(reflib_available):
    """px->world->newDynamicBody / freeBody on a world that is being stepped on
    the device (examples/draw spawns and frees circles while running).  Slot reuse, the
    swap-remove of activeIndices and the re-sort change the iteration order and with it the
    solver order; the facade re-packs the world, so every frame must still match the reference."""
    sc = scenes.pile_scene(9, mix="mixed", n_dynamic=50, cols=8, unit=1000.0)
    ref = refbind.RefWorld(sc, "wide")
    worlds = pxb.HostWorlds([sc])
    h, w = worlds.lib, worlds.ptrs[0]
    rng = np.random.default_rng(3)
    empty = np.zeros(2, np.float32)
    t = 100
    live = [s for s, d in zip(ref.slots, sc.dynamic) if d]
    for frame in range(90):
        t += 20
        ref.lib.ref_set_time_ms(t)
        h.pxh_set_time_ms(t)
        assert ref.lib.ref_step(ref.w) == h.pxh_step(w)
        rr, rf, ro = ref.bodies(True)
        hr, hf, ho = worlds.bodies(0, True)
        np.testing.assert_array_equal(ro, ho, err_msg=f"frame {frame}: order")
        np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]), err_msg=f"frame {frame}")
        if frame % 7 == 3 and live:  # despawn a random live body
            slot = live.pop(int(rng.integers(0, len(live))))
            ref.lib.ref_free_body(ref.w, 1, slot)
            h.pxh_free_body(w, 1, slot)
        if frame % 5 == 1:  # spawn a circle or a box above the pile; slots are reused
            kind = int(rng.integers(0, 2))
            x, y = float(np.float32(rng.uniform(3000, 9000))), float(np.float32(-2000.0))
            p0, p1 = (250.0, 0.0) if kind == 0 else (400.0, 300.0)
            a = ref.lib.ref_add_body(ref.w, 1, kind, p0, p1, empty, 0, 1.0, x, y, 0.2, 0.5, 0.3)
            b = h.pxh_add_body(w, 1, kind, p0, p1, empty, 0, 1.0, x, y, 0.2, 0.5, 0.3)
            assert a == b
            if a >= 0:
                live.append(a)
    worlds.close()
    ref.close()


def test_vectorised_impulses_equal_single_calls():
    """pxBatchApplyImpulses(n, ...) queues exactly what n pxBatchApplyImpulse calls queue."""
    n = 9
    sc = [scenes.stack_scene(w, n_dynamic=20) for w in range(n)]
    wa, wb = pxb.HostWorlds(sc), pxb.HostWorlds(sc)
    a, b = pxb.Batch(wa), pxb.Batch(wb)
    rng = np.random.default_rng(11)
    for it in range(25):
        ws = rng.integers(0, n, 2 * n).astype(np.uint32)  # several commands per world, any order
        bodies = rng.integers(0, 20, ws.size).astype(np.uint32)
        imp = rng.uniform(-4, 4, (ws.size, 2)).astype(np.float32)
        con = rng.uniform(-2, 2, (ws.size, 2)).astype(np.float32)
        for i in range(ws.size):
            a.apply_impulse(int(ws[i]), int(bodies[i]), float(imp[i, 0]), float(imp[i, 1]), float(con[i, 0]), float(con[i, 1]))
        b.apply_impulses(ws, bodies, imp, con)
        a.set_gravity(pxb.ALL_WORLDS, 0.0, 9.0 + it * 0.01)
        b.set_gravity(pxb.ALL_WORLDS, 0.0, 9.0 + it * 0.01)
        a.step(3)
        b.step(3)
    da, db = a.download_blocks(), b.download_blocks()
    np.testing.assert_array_equal(np.asarray(da.mut), np.asarray(db.mut))
    for x in (a, b):
        x.close()
    for x in (wa, wb):
        x.close()


def test_config1_single_world_1000_step_calls(reflib_available):
    """BASELINE config 1: one world, 252 circles dropped into the pen (+ 4 static boxes = 256
    bodies), 10 iterations, 50 fps, 1000 px->world->step calls with the stub clock advancing
    20 ms per call -- the reference on the CPU, this repo's px->world->step on the device
    (a 1-world batch behind the facade).  996 of the 1000 calls fire; the state is
    compared bit for bit every 50 calls and at the end."""
    sc = scenes.pile_scene(0, mix="circles", n_dynamic=252)
    ref = refbind.RefWorld(sc, "wide")
    worlds = pxb.HostWorlds([sc])
    h, w = worlds.lib, worlds.ptrs[0]
    t, fired = 1000, 0
    for call in range(1000):
        ref.lib.ref_set_time_ms(t)
        h.pxh_set_time_ms(t)
        a, b = ref.lib.ref_step(ref.w), h.pxh_step(w)
        assert a == b, f"call {call}: stepped flag"
        fired += b
        t += 20
        if call % 50 == 49 or call == 999:
            rr, rf, ro = ref.bodies(True)
            hr, hf, ho = worlds.bodies(0, True)
            np.testing.assert_array_equal(ro, ho, err_msg=f"call {call}: order")
            np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]), err_msg=f"call {call}")
            np.testing.assert_array_equal(rf[ro], hf[ro], err_msg=f"call {call}: flags")
    assert fired == 996
    assert worlds.stats(0)["manifolds"] == len(ref.manifolds()[1]) > 300
    worlds.close()
    ref.close()
