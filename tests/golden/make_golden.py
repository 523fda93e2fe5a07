#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ from the COMPILED REFERENCE (oracle/_ref, built
from the unmodified sources under /root/reference by oracle/Makefile).

Run in the build container (needs oracle/_ref):   python tests/golden/make_golden.py
The reference repository has no tests or fixtures of its own, so these vectors --
outputs of the reference itself on seeded scenes -- are the pins for the oracle restatement and
for the device.  "narrow" cases run the unmodified 64-body reference; "wide" cases the copy
with raised capacities (oracle/widen.sh).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refbind  # noqa: E402
from pdphyzx_b200 import scenes  # noqa: E402

CHECKPOINTS = (1, 25, 100, 300, 600)

CASES = {
    # name: (scene factory kwargs, variant)
    "drop_mixed_m_narrow": (lambda: scenes.drop_scene(11, n_dynamic=18, mix="mixed", unit=1.0), "narrow"),
    "drop_mixed_cm_narrow": (lambda: scenes.drop_scene(12, n_dynamic=18, mix="mixed", unit=100.0), "narrow"),
    "drop_circles_mm_narrow": (lambda: scenes.drop_scene(13, n_dynamic=20, mix="circles", unit=1000.0), "narrow"),
    "pile_mixed_m_wide": (lambda: scenes.pile_scene(14, mix="mixed", n_dynamic=96, cols=10, unit=1.0), "wide"),
    "stack_mm_wide": (lambda: scenes.stack_scene(15, n_dynamic=48), "wide"),
}


def scene_case(name):
    return CASES[name][0]()


def main():
    out = {}
    for name, (factory, variant) in CASES.items():
        sc = factory()
        ref = refbind.RefWorld(sc, variant)
        done = 0
        for cp in CHECKPOINTS:
            ref.substep(cp - done - 1)
            a, b, d = ref.substep_traced()
            done = cp
            rows, flags, order = ref.bodies(True)
            mrows, mids = ref.manifolds()
            mrows = mrows.copy()
            mrows[mids[:, 3] < 2, 8:10] = 0
            key = f"{name}/{cp}"
            out[key + "/state"] = rows[order][:, :14].copy()
            out[key + "/flags"] = flags[order].copy()
            out[key + "/order"] = order.copy()
            out[key + "/pairs"] = np.stack([a, b, d.astype(np.uint16)], axis=0)
            out[key + "/manifold_ids"] = mids
            out[key + "/manifold_rows"] = mrows
        rows, flags, order = ref.bodies(True)
        print(f"{name}: {len(order)} dynamic bodies, {len(mids)} manifolds at substep {done}")
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
