"""Synthetic scene generators for batches of independent pdphyzx worlds.

A scene is plain numpy data (no physics): per body the arguments one would hand to the
reference's ``px->world->newDynamicBody / newStaticBody`` (reference src/px_world.c:94-109)
plus the material fields callers poke afterwards (examples/draw/circle.c:43-45).  The same
arrays are fed to the host C API of this repo and, in tests, to the compiled reference, so
both sides construct identical worlds through their own constructors.

Randomness is counter-based -- a 64-bit mix of (seed, world, body, stream) -- so any world
of any batch can be regenerated independently (SURVEY 8d "Seeds").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KIND_CIRCLE, KIND_BOX, KIND_POLYGON = 0, 1, 2
MAX_SCENE_VERTS = 8

UNIT_M, UNIT_CM, UNIT_MM = 1.0, 100.0, 1000.0
UNIT_NAMES = {1.0: "m", 100.0: "cm", 1000.0: "mm"}


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def rand01(seed: int, world, body, stream: int) -> np.ndarray:
    """float32 uniform in [0, 1) keyed by (seed, world, body, stream); vectorised."""
    world = np.asarray(world, dtype=np.uint64)
    body = np.asarray(body, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = _mix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        k = _mix64(k ^ (world * np.uint64(0xD6E8FEB86659FD93)))
        k = _mix64(k ^ (body * np.uint64(0xCA5A826395121157)))
    return ((k >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)


@dataclass
class Scene:
    """One world's construction arguments.  Body arrays are in creation order; static and
    dynamic bodies are created in the order they appear (statics first by convention)."""

    unit: float = UNIT_M
    fps: int = 50
    iterations: int = 10
    gravity: tuple = (0.0, 9.8)
    dynamic: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    p0: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))  # radius | width
    p1: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))  # height
    nverts: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))
    verts: np.ndarray = field(default_factory=lambda: np.zeros((0, MAX_SCENE_VERTS, 2), np.float32))
    density: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    x: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    y: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    restitution: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    static_friction: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    dynamic_friction: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    # optional initial motion (applied after construction; angle via setOrientation)
    vx: np.ndarray | None = None
    vy: np.ndarray | None = None
    av: np.ndarray | None = None
    angle: np.ndarray | None = None

    @property
    def n_bodies(self) -> int:
        return int(self.kind.shape[0])

    @property
    def n_dynamic(self) -> int:
        return int(self.dynamic.sum())

    @property
    def n_static(self) -> int:
        return self.n_bodies - self.n_dynamic


def _stack_verts(rows) -> np.ndarray:
    """[n_bodies, widest, 2]: bodies with few vertices are zero-padded to the widest polygon."""
    if not rows:
        return np.zeros((0, MAX_SCENE_VERTS, 2), np.float32)
    wide = max(v.shape[0] for v in rows)
    out = np.zeros((len(rows), wide, 2), np.float32)
    for i, v in enumerate(rows):
        out[i, :v.shape[0]] = v
    return out


class _Builder:
    def __init__(self):
        self.rows = []

    def add(self, dynamic, kind, p0=0.0, p1=0.0, verts=None, density=0.0, x=0.0, y=0.0, e=0.2, sf=0.5,
            df=0.3, vx=0.0, vy=0.0, av=0.0, angle=0.0):
        n = 0 if verts is None else len(verts)
        v = np.zeros((max(MAX_SCENE_VERTS, n), 2), np.float32)
        if verts is not None:
            v[:n] = np.asarray(verts, np.float32)
        self.rows.append((dynamic, kind, p0, p1, n, v, density, x, y, e, sf, df, vx, vy, av, angle))

    def scene(self, **kw) -> Scene:
        r = self.rows
        col = lambda i, dt: np.array([row[i] for row in r], dtype=dt)
        s = Scene(
            dynamic=col(0, np.uint8), kind=col(1, np.uint8), p0=col(2, np.float32), p1=col(3, np.float32),
            nverts=col(4, np.uint8),
            verts=_stack_verts([row[5] for row in r]),
            density=col(6, np.float32), x=col(7, np.float32), y=col(8, np.float32),
            restitution=col(9, np.float32), static_friction=col(10, np.float32),
            dynamic_friction=col(11, np.float32), vx=col(12, np.float32), vy=col(13, np.float32),
            av=col(14, np.float32), angle=col(15, np.float32), **kw)
        return s


def _pen(b: _Builder, cell: float, cols: int, rows_tall: int, e=0.5, sf=0.5, df=0.3, peg=True):
    """Static pen: ground box, two wall boxes and (peg=True) a ledge box on the right wall
    -> 3 or 4 statics.  y grows downward (Playdate screen convention; gravity is +y).

    The reference finds the first static to test with a binary search on aabb.max.x over a
    list sorted by aabb.min.x (src/px_body_array.c:104-131) and its static AABBs are
    radius-based (src/px_polygon.c:261-266), so a small static in the middle of the ground's
    x-range hides the ground from every body to its right.  The ledge sits against the right
    wall so that max.x stays monotone along the sorted list and the pile actually forms."""
    width = cell * (cols + 1.5)
    height = cell * (rows_tall + 6)
    thick = cell * 1.0
    floor_y = height
    b.add(0, KIND_BOX, p0=width + 2 * thick, p1=thick, x=width * 0.5, y=floor_y + thick * 0.5, e=e, sf=sf, df=df)
    b.add(0, KIND_BOX, p0=thick, p1=height, x=-thick * 0.5, y=height * 0.5, e=e, sf=sf, df=df)
    b.add(0, KIND_BOX, p0=thick, p1=height, x=width + thick * 0.5, y=height * 0.5, e=e, sf=sf, df=df)
    if peg:
        b.add(0, KIND_BOX, p0=cell * 3.0, p1=cell * 0.5, x=width - cell * 1.5, y=floor_y - cell * 5.0, e=e, sf=sf, df=df)
    return width, floor_y


def _polygons(n: np.ndarray, radius: np.ndarray, seed, world, body: np.ndarray) -> np.ndarray:
    """Per body: n[i] points at jittered angles / radii in [0.8, 1.0] * radius[i] -- convex with
    (almost always) all points on the hull for n <= 8; the reference's gift-wrap hull drops
    any that are not.  Returns float32 [len(n), 8, 2], rows beyond n[i] zero."""
    k = np.arange(MAX_SCENE_VERTS)[None, :]
    key = body[:, None] * 16 + k
    ja = rand01(seed, world, key, 11) - np.float32(0.5)
    jr = rand01(seed, world, key, 12)
    nn = n[:, None].astype(np.float32)
    ang = (k.astype(np.float32) + np.float32(0.35) * ja) * (np.float32(2.0 * np.pi) / nn)
    rad = radius[:, None].astype(np.float32) * (np.float32(0.8) + np.float32(0.2) * jr)
    v = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=2).astype(np.float32)
    v[k.repeat(len(n), 0) >= n[:, None]] = 0
    return v


def pile_scene(world: int, seed: int = 1, n_dynamic: int = 252, mix: str = "circles", unit: float = UNIT_M,
               iterations: int = 10, cell: float | None = None, gravity=None, cols: int = 18,
               peg: bool = True, fps: int = 50) -> Scene:
    """Penned pile: `n_dynamic` bodies dropped on a jittered grid `cols` wide into a pen of
    3 static boxes (+1 static ledge box) -- BASELINE configs 1-4.

    mix = "circles": all circles.  mix = "mixed": 1/3 circles, 1/3 boxes, 1/3 convex polygons
    with 3-8 vertices, interleaved by body index.  mix = "polygons": boxes and polygons.
    `cell` is the grid pitch in world length units (default 0.6 for UNIT_M, 12 for the
    pixel-like UNIT_CM / UNIT_MM scenes); body radius is ~0.4 * cell.  Gravity defaults to
    (0, 9.8) m/s^2 for UNIT_M and the examples' (0, 1.025) otherwise (scaled by the unit inside
    setGravity).  Vectorised over bodies: a batch of thousands of worlds builds in seconds.
    """
    if cell is None:
        cell = 0.6 if unit == UNIT_M else 12.0
    if gravity is None:
        gravity = (0.0, 9.8) if unit == UNIT_M else (0.0, 1.025)
    rows_tall = (n_dynamic + cols - 1) // cols
    b = _Builder()
    width, floor_y = _pen(b, cell, cols, rows_tall, peg=peg)
    stat = b.scene()
    ns = stat.n_bodies

    f32 = np.float32
    idx = np.arange(n_dynamic)
    jx = rand01(seed, world, idx, 1) - f32(0.5)
    jy = rand01(seed, world, idx, 2) - f32(0.5)
    jr = rand01(seed, world, idx, 3)
    jk = rand01(seed, world, idx, 4)
    ja = rand01(seed, world, idx, 5)
    cell32 = f32(cell)
    cx = cell32 * (f32(0.75) + (idx % cols).astype(f32) + f32(0.2) * jx)
    cy = f32(floor_y) - cell32 * (f32(2.0) + (idx // cols).astype(f32) * f32(1.05) + f32(0.2) * jy)
    r = cell32 * (f32(0.36) + f32(0.06) * jr)
    if mix == "mixed":
        kind = np.array([KIND_CIRCLE, KIND_BOX, KIND_POLYGON], np.uint8)[idx % 3]
    elif mix == "polygons":
        kind = np.array([KIND_BOX, KIND_POLYGON], np.uint8)[idx % 2]
    else:
        kind = np.full(n_dynamic, KIND_CIRCLE, np.uint8)
    is_c, is_b, is_p = kind == KIND_CIRCLE, kind == KIND_BOX, kind == KIND_POLYGON
    p0 = np.zeros(n_dynamic, f32)
    p1 = np.zeros(n_dynamic, f32)
    p0[is_c] = r[is_c]
    p0[is_b] = (f32(2.0) * r * (f32(0.75) + f32(0.2) * jk))[is_b]
    p1[is_b] = (f32(2.0) * r * (f32(0.95) - f32(0.2) * jk))[is_b]
    nverts = np.zeros(n_dynamic, np.uint8)
    nverts[is_p] = (3 + (jk * f32(6.0)).astype(np.int64) % 6)[is_p]
    verts = np.zeros((n_dynamic, MAX_SCENE_VERTS, 2), f32)
    if is_p.any():
        verts[is_p] = _polygons(nverts[is_p].astype(np.int64), (r * f32(1.1))[is_p], seed, world, idx[is_p])
    angle = np.zeros(n_dynamic, f32)
    angle[is_b] = (f32(0.3) * (ja - f32(0.5)))[is_b]
    angle[is_p] = (f32(6.28) * ja)[is_p]

    cat = lambda a, c: np.concatenate([a, c])
    zeros = np.zeros(ns + n_dynamic, f32)
    return Scene(
        unit=unit, fps=fps, iterations=iterations, gravity=tuple(gravity),
        dynamic=cat(stat.dynamic, np.ones(n_dynamic, np.uint8)), kind=cat(stat.kind, kind),
        p0=cat(stat.p0, p0), p1=cat(stat.p1, p1), nverts=cat(stat.nverts, nverts), verts=cat(stat.verts, verts),
        density=cat(stat.density, np.ones(n_dynamic, f32)), x=cat(stat.x, cx.astype(f32)), y=cat(stat.y, cy.astype(f32)),
        restitution=cat(stat.restitution, np.full(n_dynamic, 0.2, f32)),
        static_friction=cat(stat.static_friction, np.full(n_dynamic, 0.5, f32)),
        dynamic_friction=cat(stat.dynamic_friction, np.full(n_dynamic, 0.3, f32)),
        vx=zeros.copy(), vy=zeros.copy(), av=zeros.copy(), angle=cat(np.zeros(ns, f32), angle))


def stack_scene(world: int, seed: int = 1, n_dynamic: int = 60, unit: float = UNIT_MM, iterations: int = 10,
                cols: int = 6, cell: float = 20.0) -> Scene:
    """Polygon stacking (BASELINE config 5): columns of boxes / polygons resting on a ground
    box, UNIT_MM pixel-like lengths, gravity 9.8 m/s^2."""
    b = _Builder()
    rows_tall = (n_dynamic + cols - 1) // cols
    width = cell * (cols * 1.5 + 1)
    floor_y = cell * (rows_tall + 4)
    b.add(0, KIND_BOX, p0=width * 1.5, p1=cell, x=width * 0.5, y=floor_y + cell * 0.5, e=0.1, sf=0.6, df=0.4)
    idx = np.arange(n_dynamic)
    jk = rand01(seed, world, idx, 4)
    jx = rand01(seed, world, idx, 1) - np.float32(0.5)
    for i in range(n_dynamic):
        col, row = i % cols, i // cols
        cx = cell * (1.0 + 1.5 * col) + float(np.float32(0.04) * np.float32(cell) * jx[i])
        cy = floor_y - cell * (0.5 + row * 1.01)
        if (i + world) % 4 == 3:
            nv = 6 + 2 * (i % 2)
            v = _polygons(np.array([nv]), np.array([cell * 0.5], np.float32), seed, world, np.array([i]))[0, :nv]
            b.add(1, KIND_POLYGON, verts=v, density=1.0, x=cx, y=cy, e=0.1, sf=0.6, df=0.4)
        else:
            w = cell * float(np.float32(0.9) + np.float32(0.1) * jk[i])
            b.add(1, KIND_BOX, p0=w, p1=cell, density=1.0, x=cx, y=cy, e=0.1, sf=0.6, df=0.4)
    return b.scene(unit=unit, fps=50, iterations=iterations, gravity=(0.0, 9.8))


def drop_scene(world: int, seed: int = 1, n_dynamic: int = 20, mix: str = "mixed", unit: float = UNIT_M,
               iterations: int = 8) -> Scene:
    """Small pile that fits the UNMODIFIED reference (<=64 bodies, <=64 manifolds)."""
    return pile_scene(world, seed=seed, n_dynamic=n_dynamic, mix=mix, unit=unit, iterations=iterations,
                      cols=5, peg=False)


def edge_scene(world: int, seed: int = 1, n_dynamic: int = 60, n_static: int = 40, unit: float = UNIT_M,
               iterations: int = 10, max_verts: int = 32) -> Scene:
    """Edge cases of the step in one world: polygons with 9..`max_verts` vertices (the vertex cap
    of the reference is 32, src/includes/px_vec2.h:481), many static boxes of uneven sizes along x
    -- the reference's lower-bound search over statics tests aabb.max.x on a list sorted by
    aabb.min.x (src/px_body_array.c:104-131), which is not monotone here, and more than 32 statics
    exercise the long static window of the sweep -- plus circles and boxes falling through them."""
    b = _Builder()
    cell = 0.6 if unit == UNIT_M else 12.0  # as pile_scene: pixel-like lengths for the CM / MM builds
    f32 = np.float32
    cols = 10
    width = cell * (cols + 1.5)
    rows_tall = (n_dynamic + cols - 1) // cols
    floor_y = cell * (rows_tall + 8)
    b.add(0, KIND_BOX, p0=width * 1.2, p1=cell, x=width * 0.5, y=floor_y + cell * 0.5, e=0.4)
    sidx = np.arange(n_static - 1)
    sx = rand01(seed, world, sidx, 21) * f32(width)
    sy = f32(floor_y) - rand01(seed, world, sidx, 22) * f32(cell * 4.0)
    sw = f32(cell) * (f32(0.2) + f32(2.5) * rand01(seed, world, sidx, 23) ** 2)
    for i in range(n_static - 1):
        b.add(0, KIND_BOX, p0=float(sw[i]), p1=float(cell * 0.3), x=float(sx[i]), y=float(sy[i]), e=0.3, sf=0.6, df=0.4)
    idx = np.arange(n_dynamic)
    jx = rand01(seed, world, idx, 1) - f32(0.5)
    jk = rand01(seed, world, idx, 4)
    for i in range(n_dynamic):
        x = float(cell * (0.75 + (i % cols) + 0.2 * jx[i]))
        y = float(floor_y - cell * (6.0 + (i // cols) * 1.05))
        r = float(cell * 0.38)
        kind = i % 3
        if kind == 0:
            b.add(1, KIND_CIRCLE, p0=r, density=1.0, x=x, y=y)
        elif kind == 1:
            b.add(1, KIND_BOX, p0=1.5 * r, p1=1.8 * r, density=1.0, x=x, y=y, angle=float(0.3 * jx[i]))
        else:
            n = 9 + int(jk[i] * (max_verts - 8)) % (max_verts - 8)
            ang = (np.arange(n, dtype=f32) + f32(0.3) * (rand01(seed, world, i * 64 + np.arange(n), 11) - f32(0.5))) * f32(2 * np.pi / n)
            rad = f32(r * 1.1) * (f32(0.9) + f32(0.1) * rand01(seed, world, i * 64 + np.arange(n), 12))
            b.add(1, KIND_POLYGON, verts=np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1), density=1.0, x=x, y=y,
                  angle=float(6.28 * jk[i]))
    g = (0.0, 9.8) if unit == UNIT_M else (0.0, 1.025)
    return b.scene(unit=unit, iterations=iterations, gravity=g)
