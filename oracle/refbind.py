"""ctypes binding to oracle/_ref/libpxref_{narrow,wide}_{m,cm,mm}.so -- the compiled reference.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.  The libraries are
built by oracle/Makefile from /root/reference/src (where present) and travel prebuilt to the
GPU box.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_UNIT_TAG = {1.0: "m", 100.0: "cm", 1000.0: "mm"}
_LIBS: dict = {}

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")


def available(variant: str = "wide", unit: float = 1.0) -> bool:
    return os.path.exists(_path(variant, unit))


def _path(variant, unit):
    return os.path.join(_HERE, "_ref", f"libpxref_{variant}_{_UNIT_TAG[float(unit)]}.so")


def load(variant: str = "wide", unit: float = 1.0):
    key = (variant, float(unit))
    if key in _LIBS:
        return _LIBS[key]
    lib = C.CDLL(_path(variant, unit))
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "ref_init": (None, []),
        "ref_unit": (f, []),
        "ref_max_bodies": (i, []),
        "ref_max_manifolds": (i, []),
        "ref_sizeof_world": (C.c_size_t, []),
        "ref_sizeof_body": (C.c_size_t, []),
        "ref_sin_table": (C.POINTER(C.c_float), []),
        "ref_set_time_ms": (None, [C.c_uint32]),
        "ref_world_new": (vp, [i]),
        "ref_world_free": (None, [vp]),
        "ref_set_gravity": (None, [vp, f, f]),
        "ref_set_iterations": (None, [vp, i]),
        "ref_set_sleep_params": (None, [vp, f, f, f]),
        "ref_get_world_params": (None, [vp, f32p]),
        "ref_add_body": (i, [vp, i, i, f, f, f32p, i, f, f, f, f, f, f]),
        "ref_body_set_motion": (None, [vp, i, i, f, f, f, f]),
        "ref_body_apply_impulse": (None, [vp, i, f, f, f, f]),
        "ref_body_apply_force": (None, [vp, i, f, f, f, f]),
        "ref_body_set_position": (None, [vp, i, i, f, f]),
        "ref_body_set_orientation": (None, [vp, i, i, f]),
        "ref_body_move_by": (None, [vp, i, i, f, f]),
        "ref_body_rotate": (None, [vp, i, i, f]),
        "ref_free_body": (None, [vp, i, i]),
        "ref_substep": (None, [vp, f, i]),
        "ref_step": (i, [vp]),
        "ref_substep_traced": (i, [vp, f, u16p, u16p, u8p, i]),
        "ref_body_row_floats": (i, []),
        "ref_get_bodies": (i, [vp, i, f32p, u8p, u8p, i]),
        "ref_set_bodies": (None, [vp, i, f32p, u8p, u8p, i]),
        "ref_get_polygon": (i, [vp, i, i, f32p, f32p, i]),
        "ref_get_manifolds": (i, [vp, f32p, u16p, i]),
        "ref_math_rsqrt": (None, [f32p, f32p, i]),
        "ref_math_rcp": (None, [f32p, f32p, i]),
        "ref_math_sqrt": (None, [f32p, f32p, i]),
        "ref_math_div": (None, [f32p, f32p, f32p, i]),
        "ref_math_sin": (None, [f32p, f32p, i]),
        "ref_math_cos": (None, [f32p, f32p, i]),
        "ref_time_substeps": (C.c_double, [C.POINTER(vp), i, i, f, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.ref_init()
    _LIBS[key] = lib
    return lib


BODY_COLS = dict(px=0, py=1, vx=2, vy=3, av=4, angle=5, fx=6, fy=7, torque=8, sleep=9, m00=10, m01=11,
                 m10=12, m11=13, minx=14, miny=15, maxx=16, maxy=17, mass=18, imass=19, inertia=20,
                 iinertia=21, e=22, sf=23, df=24, density=25, radius=26, type=27, nverts=28)
N_MUTABLE_COLS = 18


class RefWorld:
    """One reference PxWorld built from a pdphyzx_b200.scenes.Scene through the reference's own
    px->world->newStaticBody/newDynamicBody."""

    def __init__(self, scene, variant: str = "wide"):
        self.lib = load(variant, scene.unit)
        self.scene = scene
        self.w = self.lib.ref_world_new(int(scene.fps))
        self.lib.ref_set_iterations(self.w, int(scene.iterations))
        self.lib.ref_set_gravity(self.w, float(scene.gravity[0]), float(scene.gravity[1]))
        self.slots = []
        for k in range(scene.n_bodies):
            verts = np.ascontiguousarray(scene.verts[k].reshape(-1), np.float32)
            slot = self.lib.ref_add_body(
                self.w, int(scene.dynamic[k]), int(scene.kind[k]), float(scene.p0[k]), float(scene.p1[k]), verts,
                int(scene.nverts[k]), float(scene.density[k]), float(scene.x[k]), float(scene.y[k]),
                float(scene.restitution[k]), float(scene.static_friction[k]), float(scene.dynamic_friction[k]))
            if slot < 0:
                raise RuntimeError(f"reference rejected body {k}")
            self.slots.append(slot)
            if scene.vx is not None and (scene.vx[k] or scene.vy[k] or scene.av[k] or scene.angle[k]):
                self.lib.ref_body_set_motion(self.w, int(scene.dynamic[k]), slot, float(scene.vx[k]),
                                             float(scene.vy[k]), float(scene.av[k]), float(scene.angle[k]))
        p = np.zeros(8, np.float32)
        self.lib.ref_get_world_params(self.w, p)
        self.params = p
        self.dt = float(p[6])

    def close(self):
        if self.w:
            self.lib.ref_world_free(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def substep(self, n: int = 1, dt: float | None = None):
        self.lib.ref_substep(self.w, self.dt if dt is None else dt, int(n))

    def substep_traced(self, dt: float | None = None, cap: int = 8192):
        a = np.zeros(cap, np.uint16)
        b = np.zeros(cap, np.uint16)
        d = np.zeros(cap, np.uint8)
        n = self.lib.ref_substep_traced(self.w, self.dt if dt is None else dt, a, b, d, cap)
        assert n <= cap
        return a[:n].copy(), b[:n].copy(), d[:n].copy()

    def bodies(self, dynamic: bool = True):
        cap = self.lib.ref_max_bodies()
        ncol = self.lib.ref_body_row_floats()
        rows = np.zeros((cap, ncol), np.float32)
        flags = np.zeros(cap, np.uint8)
        order = np.zeros(cap, np.uint8)
        n = self.lib.ref_get_bodies(self.w, int(dynamic), rows, flags, order, cap)
        return rows, flags, order[:n].copy()

    def set_bodies(self, rows, flags, order, dynamic: bool = True):
        rows = np.ascontiguousarray(rows, np.float32)
        self.lib.ref_set_bodies(self.w, int(dynamic), rows, np.ascontiguousarray(flags, np.uint8),
                                np.ascontiguousarray(order, np.uint8), len(order))

    def polygon(self, slot: int, dynamic: bool = True):
        v = np.zeros(64, np.float32)
        nrm = np.zeros(64, np.float32)
        n = self.lib.ref_get_polygon(self.w, int(dynamic), int(slot), v, nrm, 32)
        return v[: 2 * n].reshape(n, 2).copy(), nrm[: 2 * n].reshape(n, 2).copy()

    def manifolds(self):
        cap = self.lib.ref_max_manifolds()
        rows = np.zeros((cap, 10), np.float32)
        ids = np.zeros((cap, 4), np.uint16)
        n = self.lib.ref_get_manifolds(self.w, rows, ids, cap)
        return rows[:n].copy(), ids[:n].copy()


def time_substeps(worlds, substeps: int, n_threads: int) -> float:
    """Wall seconds for `substeps` x pxWorldStepWithDt over all `worlds` on n_threads pthreads."""
    lib = worlds[0].lib
    arr = (C.c_void_p * len(worlds))(*[w.w for w in worlds])
    return lib.ref_time_substeps(arr, len(worlds), int(substeps), float(worlds[0].dt), int(n_threads))
