"""Ad-hoc GPU probe (not collected by pytest): quick device-vs-oracle parity sweep + timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity
from pdphyzx_b200 import scenes, batch as pxb

def main():
    t0 = time.time()
    cases = [
        ("circles-M-serial", [scenes.pile_scene(w, mix="circles") for w in range(2)], dict(substeps=300, flags=pxb.FLAG_SERIAL_SOLVER)),
        ("circles-M", [scenes.pile_scene(w, mix="circles") for w in range(4)], dict(substeps=900)),
        ("mixed-M", [scenes.pile_scene(w, mix="mixed") for w in range(4)], dict(substeps=900)),
        ("mixed-MM-k8", [scenes.pile_scene(w, mix="mixed", unit=1000.0) for w in range(4)], dict(substeps=904, k=8)),
        ("mixed-M-128thr", [scenes.pile_scene(w, mix="mixed") for w in range(2)], dict(substeps=600, threads=128)),
        ("stack-MM", [scenes.stack_scene(w) for w in range(4)], dict(substeps=600)),
        ("drop-small", [scenes.drop_scene(w) for w in range(4)], dict(substeps=600)),
    ]
    ok = True
    for name, sc, kw in cases:
        t = time.time()
        try:
            msg, st = parity.run_side_by_side(sc, **kw)
        except Exception as e:
            msg, st = f"EXC {type(e).__name__}: {e}", {}
        print(f"[{name}] {'OK' if msg is None else 'MISMATCH: ' + msg} {st} ({time.time()-t:.1f}s)", flush=True)
        ok = ok and msg is None
    print("probe", "PASS" if ok else "FAIL", f"{time.time()-t0:.1f}s")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
