"""ctypes binding of the C-ABI batch layer (include/px_batch.h) and of the host C API shim.

This module is plumbing: it loads ``lib/libpdphyzx_b200.so`` (CUDA kernels + C ABI) and
``lib/libpxhost_{m,cm,mm}.so`` (the pdphyzx host API of include/pdphyzx.h compiled per
PDPHYZX_UNIT) and exposes them to Python.  No physics happens in Python and nothing here
falls back to a CPU implementation: if the CUDA library is missing the import of the
library fails, and stepping without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.environ.get("PX_LIBDIR") or os.path.join(_HERE, "lib")  # PX_LIBDIR: experiment builds
_UNIT_TAG = {1.0: "m", 100.0: "cm", 1000.0: "mm"}

ALL_WORLDS = 0xFFFFFFFF
FLAG_TRACE = 1
FLAG_SERIAL_SOLVER = 2
FLAG_PROFILE = 4
FLAG_SPLIT = 8
FLAG_FUSED = 16
FLAG_DYNAMIC_SOLVER = 32
OVERFLOW_PAIRS = 1
OVERFLOW_MANIFOLDS = 2
STATIC = 0x10000  # PX_BATCH_STATIC: OR into a body slot to address a static body

DYN_ROWS = ("px", "py", "vx", "vy", "av", "angle", "fx", "fy", "torque", "sleep", "m00", "m01", "m10", "m11")
DYNC_ROWS = ("imass", "iinertia", "e", "sf", "df", "radius")
STAT_ROWS = ("e", "sf", "df", "radius")  # immutable block
STATM_ROWS = ("px", "py", "angle", "m00", "m01", "m10", "m11")  # mutable block: static poses
TRACE_M_F32 = 10


class PxBatchDesc(C.Structure):
    _fields_ = [
        ("structSize", C.c_uint32), ("unit", C.c_float), ("maxDynamic", C.c_uint32), ("maxStatic", C.c_uint32),
        ("maxVertices", C.c_uint32), ("maxManifolds", C.c_uint32), ("maxPairs", C.c_uint32), ("device", C.c_int32),
        ("stream", C.c_void_p), ("flags", C.c_uint32), ("threads", C.c_uint32), ("sinTable", C.POINTER(C.c_float)),
    ]


class PxWorldHeader(C.Structure):
    _fields_ = [
        ("gravityX", C.c_float), ("gravityY", C.c_float), ("dt", C.c_float), ("linearSleepThresholdSq", C.c_float),
        ("angularSleepThreshold", C.c_float), ("timeToSleep", C.c_float), ("iterations", C.c_uint32),
        ("nDynamic", C.c_uint32), ("nStatic", C.c_uint32), ("nVertices", C.c_uint32), ("statManifolds", C.c_uint32),
        ("statSleeping", C.c_uint32), ("statOverflow", C.c_uint32), ("statPairs", C.c_uint32), ("substeps", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


HEADER_DTYPE = np.dtype([(n, np.float32 if t is C.c_float else np.uint32) for n, t in PxWorldHeader._fields_])


class PxBatchLayout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "nWorlds", "maxDynamic", "maxStatic", "maxVertices", "maxManifolds", "maxPairs", "mutStride", "immStride",
        "offHeader", "offDyn", "offDynFlags", "offDynOrder", "offStatMut", "offDynConst", "offDynShape", "offStat", "offStatShape",
        "offStatSlot", "offVerts", "traceStride", "offTracePairs", "offTraceManifolds")] + [
        ("unit", C.c_float), ("epsilon", C.c_float), ("epsilonSq", C.c_float), ("allowedPenetration", C.c_float)]


_core = None
_hosts: dict = {}


def core():
    """The C-ABI library.  Raises OSError if it has not been built (python __graft_entry__.py)."""
    global _core
    if _core is not None:
        return _core
    path = os.path.join(_LIBDIR, "libpdphyzx_b200.so")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, u32, i32, f = C.c_void_p, C.c_uint32, C.c_int, C.c_float
    pp = C.POINTER(C.c_void_p)
    sig = {
        "pxBatchNew": (vp, [pp, u32, C.POINTER(PxBatchDesc)]),
        "pxBatchCreate": (vp, [u32, C.POINTER(PxBatchDesc)]),
        "pxBatchUpload": (i32, [vp, pp, u32, u32]),
        "pxBatchFree": (None, [vp]),
        "pxBatchStep": (i32, [vp, u32]),
        "pxBatchStepHost": (i32, [vp, u32, vp, u32]),
        "pxBatchStepTimed": (i32, [vp, u32, C.POINTER(C.c_float)]),
        "pxBatchSync": (i32, [vp]),
        "pxBatchReadback": (i32, [vp, pp, u32, u32, u32]),
        "pxBatchApplyImpulse": (i32, [vp, u32, u32, f, f, f, f]),
        "pxBatchApplyImpulses": (i32, [vp, u32, vp, vp, vp, vp]),
        "pxBatchApplyForce": (i32, [vp, u32, u32, f, f, f, f]),
        "pxBatchSetPosition": (i32, [vp, u32, u32, f, f]),
        "pxBatchMoveBy": (i32, [vp, u32, u32, f, f]),
        "pxBatchSetOrientation": (i32, [vp, u32, u32, f]),
        "pxBatchRotate": (i32, [vp, u32, u32, f]),
        "pxBatchAddBody": (i32, [vp, u32, vp, vp]),
        "pxBatchFreeBody": (i32, [vp, u32, u32, vp]),
        "pxBatchSyncStatics": (i32, [vp, u32, vp]),
        "pxBatchAllocCount": (C.c_uint64, [vp]),
        "pxBatchSetGravity": (i32, [vp, u32, f, f]),
        "pxBatchSetIterations": (i32, [vp, u32, u32]),
        "pxBatchGetLayout": (i32, [vp, C.POINTER(PxBatchLayout)]),
        "pxBatchHostMutable": (vp, [vp]),
        "pxBatchHostImmutable": (vp, [vp]),
        "pxBatchMutableH2D": (i32, [vp, vp, u32, u32]),
        "pxBatchMutableD2H": (i32, [vp, vp, u32, u32]),
        "pxBatchImmutableH2D": (i32, [vp, vp, u32, u32]),
        "pxBatchTraceD2H": (i32, [vp, vp, u32, u32]),
        "pxBatchProfileD2H": (i32, [vp, vp]),
        "pxBatchDeviceMutable": (C.c_uint64, [vp]),
        "pxBatchDeviceImmutable": (C.c_uint64, [vp]),
        "pxBatchLaunchCount": (C.c_uint64, [vp]),
        "pxBatchKernelInfo": (i32, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "pxBatchComputeLayout": (i32, [u32, C.POINTER(PxBatchDesc), C.POINTER(PxBatchLayout)]),
        "pxPackWorlds": (i32, [C.POINTER(PxBatchLayout), pp, u32, vp, vp]),
        "pxUnpackWorlds": (i32, [C.POINTER(PxBatchLayout), pp, u32, vp, u32]),
        "pxBatchFitDesc": (i32, [pp, u32, C.POINTER(PxBatchDesc)]),
        "pxSelfTestMath": (i32, [C.c_int, vp, vp, vp, vp, u32]),
        "pxSelfTestRsqrtSweep": (i32, [vp]),
        "pxBatchStepTimedKernels": (i32, [vp, u32, C.POINTER(C.c_float), C.POINTER(u32)]),
        "pxBatchPipelineInfo": (i32, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "pxBatchTimerStart": (i32, [vp]),
        "pxBatchTimerStop": (i32, [vp, C.POINTER(C.c_float)]),
        "pxBatchGroupNew": (vp, [pp, u32, C.POINTER(PxBatchDesc), C.POINTER(C.c_int32), u32]),
        "pxBatchGroupCreate": (vp, [u32, C.POINTER(PxBatchDesc), C.POINTER(C.c_int32), u32]),
        "pxBatchGroupFree": (None, [vp]),
        "pxBatchGroupShards": (u32, [vp]),
        "pxBatchGroupShard": (vp, [vp, u32, C.POINTER(u32), C.POINTER(u32)]),
        "pxBatchGroupLocate": (vp, [vp, u32, C.POINTER(u32)]),
        "pxBatchGroupUpload": (i32, [vp, pp, u32, u32]),
        "pxBatchGroupStep": (i32, [vp, u32]),
        "pxBatchGroupStepHost": (i32, [vp, u32, u32]),
        "pxBatchGroupStepTimed": (i32, [vp, u32, C.POINTER(C.c_float)]),
        "pxBatchGroupSync": (i32, [vp]),
        "pxBatchGroupReadback": (i32, [vp, pp, u32, u32, u32]),
        "pxBatchGroupSetGravity": (i32, [vp, u32, f, f]),
        "pxBatchGroupSetIterations": (i32, [vp, u32, u32]),
        "pxBatchGroupApplyImpulse": (i32, [vp, u32, u32, f, f, f, f]),
        "pxBatchLastError": (C.c_char_p, []),
        "pxBatchVersion": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._px_symbols = tuple(sig)
    _core = lib
    return lib


def last_error() -> str:
    return core().pxBatchLastError().decode()


class PxError(RuntimeError):
    pass


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise PxError(f"{what} failed ({rc}): {last_error()}")
    return rc


f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def host(unit: float = 1.0):
    """The host C API (include/pdphyzx.h) compiled with PDPHYZX_UNIT = unit, as a flat shim."""
    key = float(unit)
    if key in _hosts:
        return _hosts[key]
    core()
    lib = C.CDLL(os.path.join(_LIBDIR, f"libpxhost_{_UNIT_TAG[key]}.so"))
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "pxh_init": (None, []),
        "pxh_unit": (f, []),
        "pxh_sin_table": (C.POINTER(C.c_float), []),
        "pxh_sizeof_world": (C.c_size_t, []),
        "pxh_set_time_ms": (None, [C.c_uint32]),
        "pxh_get_time_ms": (C.c_uint32, []),
        "pxh_world_new": (vp, [i]),
        "pxh_world_free": (None, [vp]),
        "pxh_set_gravity": (None, [vp, f, f]),
        "pxh_set_iterations": (None, [vp, i]),
        "pxh_set_sleep_params": (None, [vp, f, f, f]),
        "pxh_step": (i, [vp]),
        "pxh_draw_debug": (None, [vp]),
        "pxh_substep_dt": (f, [vp]),
        "pxh_world_stats": (None, [vp, np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")]),
        "pxh_add_body": (i, [vp, i, i, f, f, f32p, i, f, f, f, f, f, f]),
        "pxh_build_worlds": (i, [C.POINTER(vp), i, i32p, i, i, f, f, u8p, u8p, f32p, f32p, u8p, f32p, f32p, f32p, f32p,
                                 f32p, f32p, f32p, C.c_void_p, i]),
        "pxh_body_set_motion": (None, [vp, i, i, f, f, f, f]),
        "pxh_body_apply_impulse": (None, [vp, i, f, f, f, f]),
        "pxh_body_apply_force": (None, [vp, i, f, f, f, f]),
        "pxh_body_set_position": (None, [vp, i, i, f, f]),
        "pxh_body_set_orientation": (None, [vp, i, i, f]),
        "pxh_body_move_by": (None, [vp, i, i, f, f]),
        "pxh_body_rotate": (None, [vp, i, i, f]),
        "pxh_body_is_valid": (i, [vp, i, i]),
        "pxh_free_body": (None, [vp, i, i]),
        "pxh_body_ptr": (vp, [vp, i, i]),
        "pxh_body_row_floats": (i, []),
        "pxh_get_bodies": (i, [vp, i, f32p, u8p, u8p, i]),
        "pxh_get_polygon": (i, [vp, i, i, f32p, f32p, i]),
        "pxh_math": (None, [i, f32p, f32p, f32p, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.pxh_init()
    _hosts[key] = lib
    return lib


class HostWorlds:
    """A list of host PxWorld structs built through the pdphyzx API from Scene objects."""

    def __init__(self, scenes):
        scenes = list(scenes)
        assert scenes, "need at least one scene"
        s0 = scenes[0]
        self.unit = float(s0.unit)
        self.lib = host(self.unit)
        self.n = len(scenes)
        self.ptrs = (C.c_void_p * self.n)()
        for s in scenes:
            assert (s.unit, s.fps, s.iterations, tuple(s.gravity)) == (s0.unit, s0.fps, s0.iterations, tuple(s0.gravity)), \
                "HostWorlds: scenes of one batch share unit, fps, iterations and gravity"
        counts = np.array([s.n_bodies for s in scenes], np.int64)
        start = np.zeros(self.n + 1, np.int32)
        start[1:] = np.cumsum(counts)
        cat = lambda name, dt: np.ascontiguousarray(np.concatenate([getattr(s, name) for s in scenes]).astype(dt))
        motion = None
        if any(s.vx is not None for s in scenes):
            motion = np.ascontiguousarray(np.concatenate([
                np.stack([s.vx, s.vy, s.av, s.angle], axis=1) if s.vx is not None else np.zeros((s.n_bodies, 4), np.float32)
                for s in scenes]).astype(np.float32))
        vmax = max(int(s.verts.shape[1]) for s in scenes)  # vertices per body slot (8 by default, up to 32)
        pad = lambda v: v if v.shape[1] == vmax else np.concatenate([v, np.zeros((v.shape[0], vmax - v.shape[1], 2), v.dtype)], axis=1)
        verts = np.ascontiguousarray(np.concatenate([pad(s.verts) for s in scenes]).astype(np.float32).reshape(-1))
        built = self.lib.pxh_build_worlds(
            self.ptrs, self.n, start, int(s0.fps), int(s0.iterations), float(s0.gravity[0]), float(s0.gravity[1]),
            cat("dynamic", np.uint8), cat("kind", np.uint8), cat("p0", np.float32), cat("p1", np.float32),
            cat("nverts", np.uint8), verts, cat("density", np.float32), cat("x", np.float32), cat("y", np.float32),
            cat("restitution", np.float32), cat("static_friction", np.float32), cat("dynamic_friction", np.float32),
            motion.ctypes.data_as(C.c_void_p) if motion is not None else None, vmax)
        if built != self.n:
            for k in range(built):
                self.lib.pxh_world_free(self.ptrs[k])
            raise PxError(f"host API rejected a body of world {built}")
        self.dt = float(self.lib.pxh_substep_dt(self.ptrs[0]))

    def pointer_array(self):
        return C.cast(self.ptrs, C.POINTER(C.c_void_p))

    def bodies(self, w: int, dynamic: bool = True):
        ncol = self.lib.pxh_body_row_floats()
        rows = np.zeros((255, ncol), np.float32)
        flags = np.zeros(255, np.uint8)
        order = np.zeros(255, np.uint8)
        n = self.lib.pxh_get_bodies(self.ptrs[w], int(dynamic), rows, flags, order, 255)
        return rows, flags, order[:n].copy()

    def polygon(self, w: int, slot: int, dynamic: bool = True):
        v = np.zeros(64, np.float32)
        nrm = np.zeros(64, np.float32)
        n = self.lib.pxh_get_polygon(self.ptrs[w], int(dynamic), int(slot), v, nrm, 32)
        return v[: 2 * n].reshape(n, 2).copy(), nrm[: 2 * n].reshape(n, 2).copy()

    def stats(self, w: int):
        out = np.zeros(3, np.uint32)
        self.lib.pxh_world_stats(self.ptrs[w], out)
        return dict(manifolds=int(out[0]), sleeping=int(out[1]), overflow=int(out[2]))

    def close(self):
        if self.ptrs is not None:
            for k in range(self.n):
                if self.ptrs[k]:
                    self.lib.pxh_world_free(self.ptrs[k])
            self.ptrs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_desc(unit: float, **kw) -> PxBatchDesc:
    d = PxBatchDesc()
    d.structSize = C.sizeof(PxBatchDesc)
    d.unit = float(unit)
    d.device = -1
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def compute_layout(n_worlds: int, desc: PxBatchDesc) -> PxBatchLayout:
    L = PxBatchLayout()
    _check(core().pxBatchComputeLayout(int(n_worlds), C.byref(desc), C.byref(L)), "pxBatchComputeLayout")
    return L


class Blocks:
    """numpy views over a (mutable, immutable) block pair in the PxBatchLayout format."""

    def __init__(self, layout: PxBatchLayout, mut: np.ndarray, imm: np.ndarray | None):
        self.L = layout
        self.mut = mut
        self.imm = imm
        L = layout
        n = mut.shape[0] // L.mutStride
        self.n = n
        self.header = np.ndarray((n,), HEADER_DTYPE, buffer=mut, offset=L.offHeader, strides=(L.mutStride,))
        D, S = L.maxDynamic, L.maxStatic
        self.dyn = np.ndarray((n, len(DYN_ROWS), D), np.float32, buffer=mut, offset=L.offDyn, strides=(L.mutStride, 4 * D, 4))
        self.flags = np.ndarray((n, D), np.uint8, buffer=mut, offset=L.offDynFlags, strides=(L.mutStride, 1))
        self.order = np.ndarray((n, D), np.uint8, buffer=mut, offset=L.offDynOrder, strides=(L.mutStride, 1))
        self.statm = np.ndarray((n, len(STATM_ROWS), S), np.float32, buffer=mut, offset=L.offStatMut, strides=(L.mutStride, 4 * S, 4))
        if imm is not None:
            self.dync = np.ndarray((n, len(DYNC_ROWS), D), np.float32, buffer=imm, offset=L.offDynConst, strides=(L.immStride, 4 * D, 4))
            self.dshape = np.ndarray((n, D), np.uint32, buffer=imm, offset=L.offDynShape, strides=(L.immStride, 4))
            self.stat = np.ndarray((n, len(STAT_ROWS), S), np.float32, buffer=imm, offset=L.offStat, strides=(L.immStride, 4 * S, 4))
            self.sshape = np.ndarray((n, S), np.uint32, buffer=imm, offset=L.offStatShape, strides=(L.immStride, 4))
            self.sslot = np.ndarray((n, S), np.uint8, buffer=imm, offset=L.offStatSlot, strides=(L.immStride, 1))
            self.verts = np.ndarray((n, L.maxVertices, 4), np.float32, buffer=imm, offset=L.offVerts, strides=(L.immStride, 16, 4))

    def row(self, name: str) -> np.ndarray:
        return self.dyn[:, DYN_ROWS.index(name), :]


class Trace:
    """numpy views over trace blocks (last substep's ordered pair and manifold lists)."""

    def __init__(self, layout: PxBatchLayout, buf: np.ndarray):
        L = layout
        n = buf.shape[0] // L.traceStride
        self.pairs = np.ndarray((n, 2, L.maxPairs), np.uint16, buffer=buf, offset=L.offTracePairs, strides=(L.traceStride, 2 * L.maxPairs, 2))
        M = L.maxManifolds
        self.mf = np.ndarray((n, TRACE_M_F32, M), np.float32, buffer=buf, offset=L.offTraceManifolds, strides=(L.traceStride, 4 * M, 4))
        self.ids = np.ndarray((n, M), np.uint32, buffer=buf, offset=L.offTraceManifolds + 4 * TRACE_M_F32 * M, strides=(L.traceStride, 4))

    def world_pairs(self, w: int, n_pairs: int):
        a = self.pairs[w, 0, :n_pairs].astype(np.int64)
        b = self.pairs[w, 1, :n_pairs].astype(np.int64)
        return a, b & 0x7FFF, (b >> 15) & 1

    def world_manifolds(self, w: int, n_m: int):
        ids = self.ids[w, :n_m]
        rows = self.mf[w, :, :n_m].T.copy()
        idm = np.stack([ids & 0xFF, (ids >> 8) & 0xFF, (ids >> 16) & 1, (ids >> 24) & 3], axis=1).astype(np.int64)
        return rows, idm


def pack_worlds(worlds: HostWorlds, desc: PxBatchDesc | None = None, first: int = 0, n: int | None = None):
    """Host-only AoS -> SoA (no GPU): returns (layout, mut bytes, imm bytes) as numpy uint8."""
    lib = core()
    n = worlds.n - first if n is None else n
    d = desc if desc is not None else make_desc(worlds.unit)
    ptrs = C.cast(C.byref(worlds.ptrs, first * C.sizeof(C.c_void_p)), C.POINTER(C.c_void_p))
    _check(lib.pxBatchFitDesc(ptrs, n, C.byref(d)), "pxBatchFitDesc")
    L = compute_layout(n, d)
    mut = np.zeros(n * L.mutStride, np.uint8)
    imm = np.zeros(n * L.immStride, np.uint8)
    _check(lib.pxPackWorlds(C.byref(L), ptrs, n, mut.ctypes.data_as(C.c_void_p), imm.ctypes.data_as(C.c_void_p)), "pxPackWorlds")
    return L, mut, imm


class Batch:
    """A PxBatch: N worlds resident on one GPU (include/px_batch.h)."""

    def __init__(self, worlds: HostWorlds | None = None, n_worlds: int | None = None, unit: float | None = None,
                 flags: int = 0, device: int = -1, stream: int | None = None, threads: int = 0, max_dynamic: int = 0,
                 max_static: int = 0, max_vertices: int = 0, max_manifolds: int = 0, max_pairs: int = 0):
        lib = core()
        self.lib = lib
        self.worlds = worlds
        u = worlds.unit if worlds is not None else float(unit)
        self.desc = make_desc(u, flags=flags, device=device, threads=threads, maxDynamic=max_dynamic,
                              maxStatic=max_static, maxVertices=max_vertices, maxManifolds=max_manifolds,
                              maxPairs=max_pairs)
        if stream:
            self.desc.stream = C.c_void_p(stream)  # NB: 0 (the legacy default stream) means "own stream"
        self._sin_keepalive = None
        if worlds is not None:
            self.desc.sinTable = worlds.lib.pxh_sin_table()
        if worlds is not None and n_worlds is None:
            self.h = lib.pxBatchNew(worlds.pointer_array(), worlds.n, C.byref(self.desc))
        else:
            self.h = lib.pxBatchCreate(int(n_worlds), C.byref(self.desc))
        if not self.h:
            raise PxError(f"pxBatchNew/pxBatchCreate failed: {last_error()}")
        self.L = PxBatchLayout()
        _check(lib.pxBatchGetLayout(self.h, C.byref(self.L)), "pxBatchGetLayout")
        self.n = self.L.nWorlds

    # -- lifecycle -----------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.pxBatchFree(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, worlds: HostWorlds, first: int = 0, src_first: int = 0, n: int | None = None):
        n = worlds.n - src_first if n is None else n
        ptrs = C.cast(C.byref(worlds.ptrs, src_first * C.sizeof(C.c_void_p)), C.POINTER(C.c_void_p))
        _check(self.lib.pxBatchUpload(self.h, ptrs, first, n), "pxBatchUpload")

    def readback(self, worlds: HostWorlds | None = None, first: int = 0, n: int | None = None, flags: int = 0) -> int:
        worlds = worlds or self.worlds
        n = worlds.n if n is None else n
        return _check(self.lib.pxBatchReadback(self.h, worlds.pointer_array(), first, n, flags), "pxBatchReadback")

    # -- stepping ------------------------------------------------------------------------
    def step(self, k: int = 1):
        _check(self.lib.pxBatchStep(self.h, int(k)), "pxBatchStep")

    def step_host(self, k: int = 1, chunks: int = 0, host_ptr: int | None = None):
        """pxBatchStepHost: host buffers in, host buffers out, copies pipelined under compute."""
        _check(self.lib.pxBatchStepHost(self.h, int(k), host_ptr, int(chunks)), "pxBatchStepHost")

    def step_timed(self, k: int = 1) -> float:
        ms = C.c_float(0)
        _check(self.lib.pxBatchStepTimed(self.h, int(k), C.byref(ms)), "pxBatchStepTimed")
        return float(ms.value)

    def step_timed_kernels(self, k: int = 1):
        """One step with every launch timed: dict of summed ms and launch counts per kernel class."""
        ms, n = (C.c_float * 3)(), (C.c_uint32 * 3)()
        _check(self.lib.pxBatchStepTimedKernels(self.h, int(k), ms, n), "pxBatchStepTimedKernels")
        return dict(sweep_ms=float(ms[0]), sweep_launches=int(n[0]), sched_ms=float(ms[1]), sched_launches=int(n[1]),
                    solve_ms=float(ms[2]), solve_launches=int(n[2]))

    def pipeline_info(self):
        a, c, d = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self.lib.pxBatchPipelineInfo(self.h, C.byref(a), C.byref(c), C.byref(d)), "pxBatchPipelineInfo")
        return dict(split=bool(a.value), static_schedule=a.value == 2, solve_ctas_per_sm=int(c.value), sched_ctas_per_sm=int(d.value))

    def sync(self):
        _check(self.lib.pxBatchSync(self.h), "pxBatchSync")

    # -- control -------------------------------------------------------------------------
    def apply_impulse(self, world, body, ix, iy, cx, cy):
        _check(self.lib.pxBatchApplyImpulse(self.h, world, body, ix, iy, cx, cy), "pxBatchApplyImpulse")

    def apply_impulses(self, worlds, bodies, impulse_xy, contact_xy=None):
        """pxBatchApplyImpulses: one applyImpulse per element, queued in order."""
        w = np.ascontiguousarray(worlds, np.uint32)
        bd = np.ascontiguousarray(bodies, np.uint32)
        imp = np.ascontiguousarray(impulse_xy, np.float32).reshape(-1, 2)
        con = None if contact_xy is None else np.ascontiguousarray(contact_xy, np.float32).reshape(-1, 2)
        assert w.size == bd.size == imp.shape[0] and (con is None or con.shape[0] == w.size)
        _check(self.lib.pxBatchApplyImpulses(self.h, w.size, w.ctypes.data, bd.ctypes.data, imp.ctypes.data,
                                             None if con is None else con.ctypes.data), "pxBatchApplyImpulses")

    def apply_force(self, world, body, fx, fy, cx, cy):
        _check(self.lib.pxBatchApplyForce(self.h, world, body, fx, fy, cx, cy), "pxBatchApplyForce")

    def set_position(self, world, body, x, y):
        _check(self.lib.pxBatchSetPosition(self.h, world, body, x, y), "pxBatchSetPosition")

    def move_by(self, world, body, dx, dy):
        _check(self.lib.pxBatchMoveBy(self.h, world, body, dx, dy), "pxBatchMoveBy")

    def set_orientation(self, world, body, radians):
        _check(self.lib.pxBatchSetOrientation(self.h, world, body, radians), "pxBatchSetOrientation")

    def rotate(self, world, body, radians):
        _check(self.lib.pxBatchRotate(self.h, world, body, radians), "pxBatchRotate")

    # -- body lifecycle on the resident batch ----------------------------------------------
    def add_body(self, world: int, host_worlds: "HostWorlds", w: int, dynamic: bool, slot: int):
        """Mirror a body just added to host world `w` (px->world->newDynamicBody / newStaticBody) onto batch world `world`."""
        ptr = host_worlds.lib.pxh_body_ptr(host_worlds.ptrs[w], int(dynamic), int(slot))
        _check(self.lib.pxBatchAddBody(self.h, world, host_worlds.ptrs[w], ptr), "pxBatchAddBody")

    def free_body(self, world: int, slot: int, dynamic: bool = True, host_worlds: "HostWorlds | None" = None, w: int = 0):
        """Mirror px->world->freeBody (already done on the host world for statics) onto batch world `world`."""
        body = int(slot) | (0 if dynamic else STATIC)
        hw = host_worlds.ptrs[w] if host_worlds is not None else None
        _check(self.lib.pxBatchFreeBody(self.h, world, body, hw), "pxBatchFreeBody")

    @property
    def alloc_count(self) -> int:
        return int(self.lib.pxBatchAllocCount(self.h))

    def set_gravity(self, world, x, y):
        _check(self.lib.pxBatchSetGravity(self.h, world, x, y), "pxBatchSetGravity")

    def set_iterations(self, world, it):
        _check(self.lib.pxBatchSetIterations(self.h, world, it), "pxBatchSetIterations")

    # -- raw blocks ----------------------------------------------------------------------
    def _host_view(self, ptr, nbytes):
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        return np.frombuffer(buf, np.uint8)

    def host_mutable(self) -> np.ndarray:
        return self._host_view(self.lib.pxBatchHostMutable(self.h), self.n * self.L.mutStride)

    def host_immutable(self) -> np.ndarray:
        return self._host_view(self.lib.pxBatchHostImmutable(self.h), self.n * self.L.immStride)

    def mutable_h2d(self, first=0, n=None, host_ptr=None):
        _check(self.lib.pxBatchMutableH2D(self.h, host_ptr, first, self.n - first if n is None else n), "pxBatchMutableH2D")

    def mutable_d2h(self, first=0, n=None, host_ptr=None):
        _check(self.lib.pxBatchMutableD2H(self.h, host_ptr, first, self.n - first if n is None else n), "pxBatchMutableD2H")

    def immutable_h2d(self, first=0, n=None, host_ptr=None):
        _check(self.lib.pxBatchImmutableH2D(self.h, host_ptr, first, self.n - first if n is None else n), "pxBatchImmutableH2D")

    def host_blocks(self) -> Blocks:
        """Views over the pinned staging as it is (no copy, no sync)."""
        return Blocks(self.L, self.host_mutable(), self.host_immutable())

    def download_blocks(self) -> Blocks:
        """Sync copy of the device state into the pinned staging; returns views over it."""
        self.mutable_d2h()
        self.sync()
        return Blocks(self.L, self.host_mutable(), self.host_immutable())

    def trace(self, first=0, n=None) -> Trace:
        n = self.n - first if n is None else n
        buf = np.zeros(n * self.L.traceStride, np.uint8)
        _check(self.lib.pxBatchTraceD2H(self.h, buf.ctypes.data_as(C.c_void_p), first, n), "pxBatchTraceD2H")
        return Trace(self.L, buf)

    def profile(self) -> np.ndarray:
        """u64[nWorlds, 16] per-phase SM-clock cycles and counters (px_batch.h, needs FLAG_PROFILE)."""
        out = np.zeros((self.n, 16), np.uint64)
        _check(self.lib.pxBatchProfileD2H(self.h, out.ctypes.data_as(C.c_void_p)), "pxBatchProfileD2H")
        return out

    @property
    def launches(self) -> int:
        return int(self.lib.pxBatchLaunchCount(self.h))

    def kernel_info(self):
        a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self.lib.pxBatchKernelInfo(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "pxBatchKernelInfo")
        return dict(ctas_per_sm=a.value, smem_bytes=b.value, threads=c.value, num_sms=d.value)


class BatchGroup:
    """A PxBatchGroup: one batch of worlds sharded over several GPUs, driven from this thread
    (include/px_batch.h).  `devices` may name a device twice (two shards on one GPU)."""

    def __init__(self, worlds: HostWorlds | None, devices, flags: int = 0, threads: int = 0, n_worlds: int | None = None,
                 unit: float | None = None, **caps):
        lib = core()
        self.lib = lib
        self.worlds = worlds
        u = worlds.unit if worlds is not None else float(unit)
        self.desc = make_desc(u, flags=flags, threads=threads,
                              **{{"max_dynamic": "maxDynamic", "max_static": "maxStatic", "max_vertices": "maxVertices",
                                  "max_manifolds": "maxManifolds", "max_pairs": "maxPairs"}[k]: v for k, v in caps.items()})
        dev = (C.c_int32 * len(devices))(*devices)
        if worlds is not None and n_worlds is None:
            self.desc.sinTable = worlds.lib.pxh_sin_table()
            self.h = lib.pxBatchGroupNew(worlds.pointer_array(), worlds.n, C.byref(self.desc), dev, len(devices))
            self.n = worlds.n
        else:  # empty group with explicit capacities, filled by upload() in chunks
            self.h = lib.pxBatchGroupCreate(int(n_worlds), C.byref(self.desc), dev, len(devices))
            self.n = int(n_worlds)
        if not self.h:
            raise PxError(f"pxBatchGroupNew/Create failed: {last_error()}")
        self.n_shards = int(lib.pxBatchGroupShards(self.h))

    def upload(self, worlds: HostWorlds, first: int = 0):
        _check(self.lib.pxBatchGroupUpload(self.h, worlds.pointer_array(), first, worlds.n), "pxBatchGroupUpload")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pxBatchGroupFree(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, s: int):
        first, count = C.c_uint32(0), C.c_uint32(0)
        h = self.lib.pxBatchGroupShard(self.h, s, C.byref(first), C.byref(count))
        return h, int(first.value), int(count.value)

    def shard_layout(self, s: int) -> PxBatchLayout:
        L = PxBatchLayout()
        _check(self.lib.pxBatchGetLayout(self.shard(s)[0], C.byref(L)), "pxBatchGetLayout")
        return L

    def step(self, k: int = 1):
        _check(self.lib.pxBatchGroupStep(self.h, int(k)), "pxBatchGroupStep")

    def step_host(self, k: int = 1, chunks: int = 0):
        _check(self.lib.pxBatchGroupStepHost(self.h, int(k), int(chunks)), "pxBatchGroupStepHost")

    def step_timed(self, k: int = 1) -> float:
        ms = C.c_float(0)
        _check(self.lib.pxBatchGroupStepTimed(self.h, int(k), C.byref(ms)), "pxBatchGroupStepTimed")
        return float(ms.value)

    def sync(self):
        _check(self.lib.pxBatchGroupSync(self.h), "pxBatchGroupSync")

    def readback(self, worlds: HostWorlds | None = None, flags: int = 0) -> int:
        worlds = worlds or self.worlds
        return _check(self.lib.pxBatchGroupReadback(self.h, worlds.pointer_array(), 0, worlds.n, flags), "pxBatchGroupReadback")

    def set_gravity(self, world, x, y):
        _check(self.lib.pxBatchGroupSetGravity(self.h, world, x, y), "pxBatchGroupSetGravity")

    def set_iterations(self, world, it):
        _check(self.lib.pxBatchGroupSetIterations(self.h, world, it), "pxBatchGroupSetIterations")

    def apply_impulse(self, world, body, ix, iy, cx, cy):
        _check(self.lib.pxBatchGroupApplyImpulse(self.h, world, body, ix, iy, cx, cy), "pxBatchGroupApplyImpulse")

    def launches(self) -> int:
        return sum(int(self.lib.pxBatchLaunchCount(self.shard(s)[0])) for s in range(self.n_shards))


def device_math(op: int, a: np.ndarray, b: np.ndarray | None = None, sin_table: np.ndarray | None = None) -> np.ndarray:
    """Element-wise device evaluation of the fast-math contract (T0 tier); needs a GPU."""
    a = np.ascontiguousarray(a, np.float32)
    bb = np.ascontiguousarray(b if b is not None else a, np.float32)
    out = np.zeros_like(a)
    tp = sin_table.ctypes.data_as(C.c_void_p) if sin_table is not None else None
    _check(core().pxSelfTestMath(op, a.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p), tp,
                                 out.ctypes.data_as(C.c_void_p), a.size), "pxSelfTestMath")
    return out


def device_rsqrt_sweep() -> np.ndarray:
    out = np.zeros(4096, np.uint64)
    _check(core().pxSelfTestRsqrtSweep(out.ctypes.data_as(C.c_void_p)), "pxSelfTestRsqrtSweep")
    return out
