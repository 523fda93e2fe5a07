"""Scene generators (determinism, reference-compatible pens) and the N>1 sharding path on CPU
(world_size 2, gloo)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import parity
from oracle import oraclebind
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scenes_are_counter_based_and_deterministic():
    a = scenes.pile_scene(7, seed=3, mix="mixed")
    b = scenes.pile_scene(7, seed=3, mix="mixed")
    c = scenes.pile_scene(8, seed=3, mix="mixed")
    for f in ("x", "y", "p0", "p1", "verts", "angle", "kind"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    assert not np.array_equal(a.x, c.x)
    assert a.n_bodies == 256 and a.n_dynamic == 252 and a.n_static == 4
    kinds = a.kind[a.dynamic == 1]
    assert abs(int((kinds == 0).sum()) - 84) <= 1 and abs(int((kinds == 1).sum()) - 84) <= 1
    assert set(a.nverts[a.kind == scenes.KIND_POLYGON]) <= set(range(3, 9))


def test_pen_keeps_the_reference_static_search_monotone():
    """The reference lower-bounds statics on aabb.max.x over a list sorted by aabb.min.x
    (px_body_array.c:104-131); the pen is laid out so that this finds the ground."""
    sc = scenes.pile_scene(0, mix="circles")
    worlds = pxb.HostWorlds([sc])
    rows, flags, order = worlds.bodies(0, dynamic=False)
    maxx = rows[order, 16]
    minx = rows[order, 14]
    assert np.all(np.diff(minx) >= 0)
    dyn_rows, _, dyn_order = worlds.bodies(0, dynamic=True)
    body_minx = dyn_rows[dyn_order, 14]
    for mx in (body_minx.min(), body_minx.max()):
        pred = maxx < mx
        assert not np.any(np.diff(pred.astype(int)) > 0), "predicate must be monotone (true then false)"
    worlds.close()


def test_shard_ranges_partition_the_batch():
    for n, ws in ((65536, 8), (10, 3), (7, 8), (4096, 1)):
        seen = []
        for r in range(ws):
            seen.extend(sharding.shard_range(n, r, ws))
        assert seen == list(range(n))
    assert list(sharding.weak_range(4, 2)) == [8, 9, 10, 11]
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 4)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch.distributed as dist
    import parity
    from oracle import oraclebind
    from pdphyzx_b200 import batch as pxb, scenes, sharding
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    ids = list(sharding.shard_range(6, rank, ws))
    worlds = pxb.HostWorlds([scenes.pile_scene(w, mix="mixed", n_dynamic=24, cols=5) for w in ids])
    L, mut, imm = pxb.pack_worlds(worlds)
    oraclebind.substeps(L, mut, imm, parity.sin_table(worlds), 60)
    blocks = pxb.Blocks(L, mut, imm)
    stats = sharding.all_reduce_stats(dict(worlds=len(ids), manifolds=float(blocks.header["statManifolds"].sum()),
                                           checksum_lo=float(sharding.state_checksum(blocks) & 0xFFFFFF)))
    slowest = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        print(json.dumps(dict(stats=stats, slowest=slowest, ws=ws)))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    """Each rank steps its own contiguous shard with no collective on the stepping path; the
    end-of-run SUM of statistics and the MAX of times match a single-process run."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    got = json.loads(line)
    assert got["ws"] == 2 and got["slowest"] == 2.0 and got["stats"]["worlds"] == 6

    worlds = pxb.HostWorlds([scenes.pile_scene(w, mix="mixed", n_dynamic=24, cols=5) for w in range(6)])
    L, mut, imm = pxb.pack_worlds(worlds)
    oraclebind.substeps(L, mut, imm, parity.sin_table(worlds), 60)
    blocks = pxb.Blocks(L, mut, imm)
    assert got["stats"]["manifolds"] == float(blocks.header["statManifolds"].sum())
    # per-shard checksums were summed; compare with the sum of per-world checksums here
    lo = 0
    for r in range(2):
        ids = list(sharding.shard_range(6, r, 2))
        sub = pxb.Blocks(pxb.compute_layout(len(ids), pxb.make_desc(1.0, maxDynamic=L.maxDynamic, maxStatic=L.maxStatic,
                                                                    maxVertices=L.maxVertices)),
                         mut[ids[0] * L.mutStride:(ids[-1] + 1) * L.mutStride], None)
        lo += sharding.state_checksum(sub) & 0xFFFFFF
    assert got["stats"]["checksum_lo"] == float(lo)
    worlds.close()
