"""Golden vectors produced by the compiled reference (tests/golden/make_golden.py) against the
CPU oracle -- runs anywhere, needs neither /root/reference nor oracle/_ref."""
import os
import sys

import numpy as np
import pytest

import parity
from oracle import oraclebind
from pdphyzx_b200 import batch as pxb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
bits = parity.bits


def check_against_golden(name, cp, blocks, trace):
    key = f"{name}/{cp}"
    h = blocks.header[0]
    order = GOLD[key + "/order"]
    n = len(order)
    np.testing.assert_array_equal(order, blocks.order[0, :n], err_msg=f"{key}: order")
    np.testing.assert_array_equal(bits(GOLD[key + "/state"]), bits(blocks.dyn[0].T[order]), err_msg=f"{key}: state")
    np.testing.assert_array_equal(GOLD[key + "/flags"], blocks.flags[0, order], err_msg=f"{key}: flags")
    pairs = GOLD[key + "/pairs"].astype(np.int64)
    pa, pb, pd = trace.world_pairs(0, int(h["statPairs"]))
    np.testing.assert_array_equal(pairs[0], pa, err_msg=f"{key}: pairs")
    np.testing.assert_array_equal(pairs[1], pb)
    np.testing.assert_array_equal(pairs[2], pd)
    rows, ids = trace.world_manifolds(0, int(h["statManifolds"]))
    np.testing.assert_array_equal(GOLD[key + "/manifold_ids"].astype(np.int64), ids, err_msg=f"{key}: manifold ids")
    np.testing.assert_array_equal(bits(GOLD[key + "/manifold_rows"]), bits(rows), err_msg=f"{key}: manifolds")


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_reference_vectors(name):
    sc = make_golden.scene_case(name)
    worlds = pxb.HostWorlds([sc])
    table = parity.sin_table(worlds)
    L, mut, imm = pxb.pack_worlds(worlds)
    blocks = pxb.Blocks(L, mut, imm)
    trace = np.zeros(L.traceStride, np.uint8)
    done = 0
    for cp in make_golden.CHECKPOINTS:
        oraclebind.substeps(L, mut, imm, table, cp - done, trace=trace)
        done = cp
        check_against_golden(name, cp, blocks, pxb.Trace(L, trace))
    worlds.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
@pytest.mark.parametrize("fused", [1, 5])
def test_device_reproduces_reference_vectors(name, fused):
    """The CUDA path against the reference's own outputs, through the C ABI."""
    sc = make_golden.scene_case(name)
    worlds = pxb.HostWorlds([sc])
    b = pxb.Batch(worlds, flags=pxb.FLAG_TRACE)
    done = 0
    for cp in make_golden.CHECKPOINTS:
        todo = cp - done
        while todo > 0:
            k = min(fused, todo)
            b.step(k)
            todo -= k
        done = cp
        check_against_golden(name, cp, b.download_blocks(), b.trace())
    b.close()
    worlds.close()
