"""ctypes binding to oracle/libpxoracle.so -- the plain-C restatement of the reference substep.

TEST INFRASTRUCTURE ONLY (see oracle/px_oracle.c).  Consumes the same (mutable, immutable)
blocks as the device; the layout struct comes from the product's host-only
pxBatchComputeLayout.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(os.path.join(_HERE, "libpxoracle.so"))
        vp = C.c_void_p
        lib.pxOracleSubsteps.restype = None
        lib.pxOracleSubsteps.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        lib.pxOracleTimeSubsteps.restype = C.c_double
        lib.pxOracleTimeSubsteps.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.pxOracleMath.restype = None
        lib.pxOracleMath.argtypes = [C.c_int, vp, vp, vp, vp, C.c_uint32]
        lib.pxOracleRsqrtSweep.restype = None
        lib.pxOracleRsqrtSweep.argtypes = [vp, C.c_uint32, C.c_uint32]
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def substeps(layout, mut: np.ndarray, imm: np.ndarray, sin_table: np.ndarray, k: int, first: int = 0, n: int | None = None,
             trace: np.ndarray | None = None):
    """k substeps in place on worlds [first, first+n) of the block pair; optional trace of the last substep."""
    n = layout.nWorlds - first if n is None else n
    load().pxOracleSubsteps(C.byref(layout), _ptr(mut), _ptr(imm), _ptr(sin_table), first, n, k, _ptr(trace))


def time_substeps(layout, mut, imm, sin_table, n_worlds: int, k: int, threads: int) -> float:
    return load().pxOracleTimeSubsteps(C.byref(layout), _ptr(mut), _ptr(imm), _ptr(sin_table), n_worlds, k, threads)


def math(op: int, a: np.ndarray, b: np.ndarray | None, sin_table: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else a, np.float32)
    out = np.zeros_like(a)
    load().pxOracleMath(op, _ptr(a), _ptr(b), _ptr(sin_table), _ptr(out), a.size)
    return out


def rsqrt_sweep(first: int, n: int) -> np.ndarray:
    """Block sums of pxFastRsqrt result bit patterns for 2^20-pattern blocks [first, first+n)."""
    out = np.zeros(n, np.uint64)
    load().pxOracleRsqrtSweep(_ptr(out), first, n)
    return out
