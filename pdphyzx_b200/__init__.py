"""pdphyzx-b200: the pdphyzx world step for large batches of independent worlds on B200.

The product is C and CUDA (include/pdphyzx.h, include/px_batch.h, pdphyzx_b200/csrc); this
package is the Python-side plumbing used by the tests and the benchmark: scene generators
and a ctypes binding of the C ABI.
"""
from . import scenes  # noqa: F401
from .batch import (ALL_WORLDS, FLAG_PROFILE, FLAG_SERIAL_SOLVER, FLAG_TRACE, OVERFLOW_MANIFOLDS, OVERFLOW_PAIRS, Batch, Blocks,  # noqa: F401
                    HostWorlds, PxError, Trace, compute_layout, core, host, make_desc, pack_worlds)

__all__ = ["scenes", "Batch", "Blocks", "HostWorlds", "PxError", "Trace", "compute_layout", "core", "host",
           "make_desc", "pack_worlds", "ALL_WORLDS", "FLAG_TRACE", "FLAG_SERIAL_SOLVER", "OVERFLOW_PAIRS",
           "OVERFLOW_MANIFOLDS"]
