"""The C-ABI library loads on a CPU-only box and exports exactly what include/px_batch.h
declares; host-only entry points (layout, pack/unpack) work without a GPU; the product refuses
to step without one."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity  # noqa: F401  (path setup)
from pdphyzx_b200 import batch as pxb
from pdphyzx_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "px_batch.h")).read()
    return sorted(set(re.findall(r"PX_API\s+[^;(]*?\b(px\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = C.CDLL(os.path.join(ROOT, "pdphyzx_b200", "lib", "libpdphyzx_b200.so"))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in px_batch.h but not exported: {missing}"


def test_binding_covers_header():
    lib = pxb.core()
    assert set(declared_symbols()) <= set(lib._px_symbols) | {"pxBatchDescDefault"}


def test_host_shims_load_and_report_units():
    for unit in (1.0, 100.0, 1000.0):
        assert pxb.host(unit).pxh_unit() == unit


def test_struct_sizes_match_c():
    # PxWorldHeader is 64 bytes by contract (px_batch.h)
    assert C.sizeof(pxb.PxWorldHeader) == 64
    d = pxb.make_desc(1.0, maxDynamic=252, maxStatic=4, maxVertices=800)
    L = pxb.compute_layout(16, d)
    assert L.maxDynamic == 256 and L.maxStatic == 4 and L.nWorlds == 16
    assert L.mutStride % 128 == 0 and L.immStride % 128 == 0
    assert L.offDyn % 16 == 0 and L.offVerts % 16 == 0
    assert L.maxManifolds == 736 and L.maxPairs == 1792  # 23/8 and 7 per body slot
    # unit-derived constants use the reference's association (px_math.h:23, px_polygon.c:329)
    for unit in (1.0, 100.0, 1000.0):
        Lu = pxb.compute_layout(1, pxb.make_desc(unit, maxDynamic=8))
        u = np.float32(unit)
        e = np.float32(0.0001) * u
        assert np.float32(Lu.epsilon) == e and np.float32(Lu.allowedPenetration) == e
        assert np.float32(Lu.epsilonSq) == np.float32(np.float32(e * np.float32(0.0001)) * u)


def test_layout_rejects_bad_descriptors():
    with pytest.raises(pxb.PxError):
        pxb.compute_layout(1, pxb.make_desc(7.0, maxDynamic=8))
    with pytest.raises(pxb.PxError):
        pxb.compute_layout(1, pxb.make_desc(1.0, maxDynamic=300))
    with pytest.raises(pxb.PxError):
        pxb.compute_layout(1, pxb.make_desc(1.0, maxDynamic=8, maxManifolds=5000))


def test_pack_unpack_roundtrip_is_lossless():
    sc = [scenes.pile_scene(w, mix="mixed", n_dynamic=40, cols=6) for w in range(3)]
    worlds = pxb.HostWorlds(sc)
    before = [worlds.bodies(w)[0].copy() for w in range(3)]
    L, mut, imm = pxb.pack_worlds(worlds)
    blocks = pxb.Blocks(L, mut, imm)
    assert list(blocks.header["nDynamic"]) == [40, 40, 40]
    assert list(blocks.header["nStatic"]) == [4, 4, 4]
    assert np.all(blocks.header["iterations"] == 10)
    # perturb the blocks, unpack, and see the perturbation in the host structs
    blocks.dyn[:, pxb.DYN_ROWS.index("vx"), :] += np.float32(1.5)
    rc = pxb.core().pxUnpackWorlds(C.byref(L), worlds.pointer_array(), 3, mut.ctypes.data_as(C.c_void_p), 0)
    assert rc == 0
    for w in range(3):
        rows, flags, order = worlds.bodies(w)
        np.testing.assert_array_equal(rows[order, 2], before[w][order, 2] + np.float32(1.5))
        np.testing.assert_array_equal(rows[order, 0], before[w][order, 0])
    worlds.close()


def test_pack_reports_capacity_errors():
    sc = [scenes.pile_scene(0, mix="mixed", n_dynamic=40, cols=6)]
    worlds = pxb.HostWorlds(sc)
    with pytest.raises(pxb.PxError):
        pxb.pack_worlds(worlds, pxb.make_desc(1.0, maxDynamic=16))
    with pytest.raises(pxb.PxError):
        pxb.pack_worlds(worlds, pxb.make_desc(1.0, maxVertices=8))
    worlds.close()


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never step on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    worlds = pxb.HostWorlds([scenes.drop_scene(0)])
    with pytest.raises(pxb.PxError) as exc:
        pxb.Batch(worlds)
    assert "cuda" in str(exc.value).lower()
    worlds.close()


def test_reference_side_binding_builds_and_loads():
    """integration/px_world_b200.c compiled into the reference's translation unit
    (oracle/_ref/libpxref_b200_*.so) loads and links the C-ABI library;
    stepping it needs a GPU (tests/test_gpu_control.py)."""
    from oracle import refbind
    if not refbind.available("b200", 1.0):
        pytest.skip("oracle/_ref/libpxref_b200_m.so not built (needs /root/reference at build time)")
    lib = refbind.load("b200", 1.0)
    lib.ref_is_b200.restype = C.c_int
    assert lib.ref_is_b200() == 1
    assert refbind.load("wide", 1.0).ref_is_b200() == 0


def test_amalgamated_header_builds_on_its_own(tmp_path):
    """the host side ships as ONE header, the way the reference ships dist/pdphyzx.h
    (reference .nob/nob.c:936-984); tools/amalgamate.py generates it and it compiles alone."""
    import subprocess
    import sys
    out = tmp_path / "pdphyzx_b200.h"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "amalgamate.py"), "--out", str(out), "--check"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    text = out.read_text()
    assert "pxBatchGroupStep" in text and "registerPdPhyzx" in text and "pxPackOneWorld" in text
    assert '#include "px_batch.h"' not in text
