"""GPU tests of SURVEY 8(f) ranks 1-2 on a RESIDENT multi-world batch: pose commands
(px->body->setPosition / moveBy / setOrientation / rotate, reference src/px_body.c:85-124) for
dynamic and static bodies, and body lifecycle (px->world->newDynamicBody / newStaticBody /
freeBody, src/px_body_array.c:7-79, src/px_world.c:83-120) without re-creating or re-uploading
the batch.  The checker is the compiled reference making the same API calls."""
import numpy as np
import pytest

import parity
from oracle import refbind
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes

pytestmark = pytest.mark.gpu
bits = parity.bits

POSE_COLS = [0, 1, 5, 10, 11, 12, 13, 14, 15, 16, 17]  # position, angle, orientation, aabb


def compare_with_reference(refs, worlds, b, what=""):
    assert b.readback(worlds) == 0
    for w, ref in enumerate(refs):
        rr, rf, ro = ref.bodies(True)
        hr, hf, ho = worlds.bodies(w, True)
        np.testing.assert_array_equal(ro, ho, err_msg=f"{what} world {w}: order")
        np.testing.assert_array_equal(bits(rr[ro][:, :18]), bits(hr[ro][:, :18]), err_msg=f"{what} world {w}: state")
        np.testing.assert_array_equal(rf[ro], hf[ro], err_msg=f"{what} world {w}: flags")
        sr, _, so = ref.bodies(False)
        hs, _, hso = worlds.bodies(w, False)
        np.testing.assert_array_equal(so, hso, err_msg=f"{what} world {w}: static order")
        np.testing.assert_array_equal(bits(sr[so][:, POSE_COLS]), bits(hs[so][:, POSE_COLS]), err_msg=f"{what} world {w}: static poses")


@pytest.mark.parametrize("fused", [1, 8])
def test_pose_commands_rotate_the_static_ground_every_frame(reflib_available, fused):
    """examples/sprites/main.c:91-95 turns its STATIC ground with the crank every frame; here 64
    resident worlds do it through pxBatchSetOrientation / pxBatchRotate, move statics and teleport,
    nudge and spin dynamic bodies, K = 1 and K = 8 substeps per launch, against the reference."""
    n = 64
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=24, cols=6) for w in range(n)]
    refs = [refbind.RefWorld(s, "wide") for s in sc]
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds)
    statics = [[s for s, d in zip(r.slots, s_.dynamic) if not d] for r, s_ in zip(refs, sc)]
    dynamics = [[s for s, d in zip(r.slots, s_.dynamic) if d] for r, s_ in zip(refs, sc)]
    rng = np.random.default_rng(5)
    for frame in range(40):
        for w in range(n):
            r = refs[w]
            ground = statics[w][0]
            if frame % 2 == 0:
                ang = float(np.float32(0.15 * np.sin(0.3 * frame + w)))
                b.set_orientation(w, pxb.STATIC | ground, ang)
                r.lib.ref_body_set_orientation(r.w, 0, ground, ang)
            else:
                dang = float(np.float32(rng.uniform(-0.02, 0.02)))
                b.rotate(w, pxb.STATIC | ground, dang)
                r.lib.ref_body_rotate(r.w, 0, ground, dang)
            if frame % 9 == 4:  # slide a wall, teleport one body, nudge and spin another
                wall = statics[w][1]
                dx = float(np.float32(rng.uniform(-0.05, 0.05)))
                b.move_by(w, pxb.STATIC | wall, dx, 0.0)
                r.lib.ref_body_move_by(r.w, 0, wall, dx, 0.0)
                body = int(rng.choice(dynamics[w]))
                x, y = float(np.float32(rng.uniform(1.0, 3.0))), float(np.float32(rng.uniform(-3.0, -1.0)))
                b.set_position(w, body, x, y)
                r.lib.ref_body_set_position(r.w, 1, body, x, y)
                body2 = int(rng.choice(dynamics[w]))
                b.move_by(w, body2, 0.01, -0.02)
                r.lib.ref_body_move_by(r.w, 1, body2, 0.01, -0.02)
                b.rotate(w, body2, 0.5)
                r.lib.ref_body_rotate(r.w, 1, body2, 0.5)
                b.set_orientation(w, body, -7.0)  # negative: fmodf then + 2 pi
                r.lib.ref_body_set_orientation(r.w, 1, body, -7.0)
        b.step(fused)
        for r in refs:
            r.substep(fused)
        if frame % 8 == 7:
            compare_with_reference(refs, worlds, b, f"frame {frame}")
    compare_with_reference(refs, worlds, b, "end")
    assert b.alloc_count == 0
    with pytest.raises(pxb.PxError):
        b.apply_impulse(0, pxb.STATIC | 0, 1, 1, 0, 0)  # statics only take pose commands
    b.close()
    worlds.close()


def test_spawn_and_despawn_on_a_resident_batch(reflib_available):
    """64 worlds stay resident while bodies are spawned and freed between frames: dynamic adds
    (slot reuse from the freed stack, appended to the DEVICE's iteration order), frees (swap-remove,
    and the reference's `index >= length` no-op, px_body_array.c:64), a static added (re-sorted
    list) and freed mid-run.  No batch re-creation, no re-upload of dynamic state, no allocation."""
    n = 64
    sc = [scenes.pile_scene(100 + w, mix="mixed", n_dynamic=40, cols=7, unit=100.0) for w in range(n)]
    refs = [refbind.RefWorld(s, "wide") for s in sc]
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds, max_dynamic=64, max_static=8, max_vertices=1024)
    h = worlds.lib
    rng = np.random.default_rng(17)
    empty = np.zeros(2, np.float32)
    tri = np.array([0.0, -30.0, 26.0, 15.0, -26.0, 15.0], np.float32)
    live = [[s for s, d in zip(r.slots, s_.dynamic) if d] for r, s_ in zip(refs, sc)]
    extra_static = [None] * n
    for frame in range(60):
        b.step(5)
        for r in refs:
            r.substep(5)
        if frame % 6 == 5:
            compare_with_reference(refs, worlds, b, f"frame {frame}")
        for w in range(n):
            r, hw = refs[w], worlds.ptrs[w]
            if frame % 7 == 3 and live[w]:  # despawn a random live body
                slot = live[w].pop(int(rng.integers(0, len(live[w]))))
                count_before = len(live[w]) + 1
                r.lib.ref_free_body(r.w, 1, slot)
                h.pxh_free_body(hw, 1, slot)
                b.free_body(w, slot)
                if slot >= count_before:  # the reference's own check made this a no-op: still alive
                    live[w].append(slot)
            if frame % 5 == 1:  # spawn a circle, a box or a triangle above the pile
                kind = int(rng.integers(0, 3))
                x, y = float(np.float32(rng.uniform(200, 500))), float(np.float32(-300.0))
                args = [(12.0, 0.0, empty, 0), (20.0, 15.0, empty, 0), (0.0, 0.0, tri, 3)][kind]
                a = r.lib.ref_add_body(r.w, 1, kind, args[0], args[1], args[2], args[3], 1.0, x, y, 0.2, 0.5, 0.3)
                c = h.pxh_add_body(hw, 1, kind, args[0], args[1], args[2], args[3], 1.0, x, y, 0.2, 0.5, 0.3)
                assert a == c
                if a >= 0:
                    b.add_body(w, worlds, w, True, a)
                    live[w].append(a)
            if frame == 22:  # a static peg: the host re-sorts its static list on every add
                x = float(np.float32(300.0 + w))
                a = r.lib.ref_add_body(r.w, 0, 1, 40.0, 10.0, empty, 0, 0.0, x, 250.0, 0.2, 0.5, 0.3)
                c = h.pxh_add_body(hw, 0, 1, 40.0, 10.0, empty, 0, 0.0, x, 250.0, 0.2, 0.5, 0.3)
                assert a == c and a >= 0
                b.add_body(w, worlds, w, False, a)
                extra_static[w] = a
            if frame == 41 and w % 2 == 0:
                r.lib.ref_free_body(r.w, 0, extra_static[w])
                h.pxh_free_body(hw, 0, extra_static[w])
                b.free_body(w, extra_static[w], dynamic=False, host_worlds=worlds, w=w)
    compare_with_reference(refs, worlds, b, "end")
    assert b.alloc_count == 0, "the resident batch allocated after creation"
    b.close()
    worlds.close()


def test_free_of_a_high_slot_is_the_reference_no_op(reflib_available):
    """px_body_array.c:64 compares the SLOT index with the body COUNT: once a few bodies are gone,
    freeing the highest slot does nothing.  Host, device and reference agree on that."""
    sc = [scenes.pile_scene(7, mix="circles", n_dynamic=12, cols=4)]
    ref = refbind.RefWorld(sc[0], "wide")
    worlds = pxb.HostWorlds(sc)
    b = pxb.Batch(worlds)
    h, hw = worlds.lib, worlds.ptrs[0]
    dyn = [s for s, d in zip(ref.slots, sc[0].dynamic) if d]
    for slot in (dyn[0], dyn[1], dyn[2]):  # 12 -> 9 bodies
        ref.lib.ref_free_body(ref.w, 1, slot)
        h.pxh_free_body(hw, 1, slot)
        b.free_body(0, slot)
    top = max(dyn)  # slot 11 >= 9 live bodies: no-op everywhere
    ref.lib.ref_free_body(ref.w, 1, top)
    h.pxh_free_body(hw, 1, top)
    b.free_body(0, top)
    b.step(50)
    ref.substep(50)
    compare_with_reference([ref], worlds, b)
    assert len(ref.bodies(True)[2]) == 9
    b.close()
    worlds.close()


def test_a_small_batch_does_not_break_a_large_ones_launch():
    """Advisor finding (round 1): the dynamic shared-memory attribute was lowered by the next, smaller
    batch.  A batch above 48 KB of shared memory per world keeps stepping after a tiny one is made."""
    big_sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=120, cols=10) for w in range(3)]
    wa = pxb.HostWorlds(big_sc)
    a = pxb.Batch(wa, max_manifolds=1024, max_pairs=4096)
    assert a.kernel_info()["smem_bytes"] > 48 * 1024
    a.step(20)
    a.sync()
    wb = pxb.HostWorlds([scenes.pile_scene(0, mix="circles", n_dynamic=4, cols=2)])
    small = pxb.Batch(wb)
    small.step(3)
    small.sync()
    a.step(20)
    a.sync()
    ref = pxb.Batch(pxb.HostWorlds(big_sc), max_manifolds=1024, max_pairs=4096)
    ref.step(40)
    np.testing.assert_array_equal(np.asarray(a.download_blocks().mut), np.asarray(ref.download_blocks().mut))
    for x in (a, small, ref):
        x.close()


def _state(worlds, n):
    out = []
    for w in range(n):
        rows, flags, order = worlds.bodies(w, True)
        out.append((bits(rows[order][:, :18]).copy(), flags[order].copy(), order.copy()))
    return out


@pytest.mark.parametrize("devices", [[0, 0, 0], "all"])
def test_batch_group_shards_one_batch_over_devices(devices):
    """PxBatchGroup: contiguous world ranges per device, every call fanned out from one
    host thread, no collective.  A sharded run equals the single-batch run world for world -- with
    three shards on one GPU, and with one shard per GPU when the box has several."""
    import torch
    if devices == "all":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs two GPUs")
        devices = list(range(torch.cuda.device_count()))
    n = 11  # not a multiple of the shard count: the last shard is short
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=40, cols=6) for w in range(n)]
    wa, wb = pxb.HostWorlds(sc), pxb.HostWorlds(sc)
    a = pxb.Batch(wa)
    g = pxb.BatchGroup(wb, devices)
    assert g.n_shards == min(len(devices), n)
    covered = []
    for s in range(g.n_shards):
        _, first, count = g.shard(s)
        covered += list(range(first, first + count))
    assert covered == list(range(n))
    for it in range(6):
        for x in (a, g):
            x.set_gravity(pxb.ALL_WORLDS, 0.1 * it, 9.8)
            x.set_gravity(7, -0.3, 9.0)
            x.apply_impulse(10, 3, 0.5, -1.0, 0.0, 0.0)
            x.set_iterations(2, 6)
        a.step(25)
        ms = g.step_timed(25) if it % 2 else (g.step(25), 0.0)[1]
        assert ms >= 0.0
    assert a.readback(wa) == 0
    assert g.readback(wb) == 0
    for (ra, fa, oa), (rb, fb, ob) in zip(_state(wa, n), _state(wb, n)):
        np.testing.assert_array_equal(oa, ob)
        np.testing.assert_array_equal(ra, rb)
        np.testing.assert_array_equal(fa, fb)
    # host-buffer stepping through the group: every shard's pinned staging in, stepped, out
    g.step_host(5, 2)
    g.sync()
    a.step(5)
    for s in range(g.n_shards):
        h, first, count = g.shard(s)
        L = g.shard_layout(s)
        buf = (pxb.C.c_uint8 * (count * L.mutStride)).from_address(g.lib.pxBatchHostMutable(h))
        got = pxb.Blocks(L, np.frombuffer(buf, np.uint8), None)
        want = a.download_blocks()
        np.testing.assert_array_equal(np.asarray(got.dyn).view(np.uint32), np.asarray(want.dyn[first:first + count]).view(np.uint32))
    a.close()
    g.close()
    wa.close()
    wb.close()
