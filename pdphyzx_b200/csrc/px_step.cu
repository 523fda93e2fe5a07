/*
 * px_step.cu -- the pdphyzx world substep as hand-written sm_100a kernels.
 *
 * One CTA per world.  Two pipelines:
 *   fused (small batches)   pxStepKernel<NT, FUSED>: the CTA loads its world's SoA block from HBM with coalesced
 *                           accesses, keeps body state in shared memory (and what only a body's owner thread needs --
 *                           angle, force, torque, sleep time -- in registers) across K fused substeps
 *                           (K x pxWorldStepWithDt, reference src/px_world.c:410-420) and stores the mutable rows back;
 *                           the solver is a dataflow execution on two specialised warps.
 *   split (large batches)   per substep pxStepKernel<256, SWEEP> (phase F of the previous substep + phases A-D, hand-over
 *                           record per world in global memory) and a solver kernel: pxReplayKernel (one warp per world
 *                           derives a periodic schedule of the contact visits and replays it per iteration) or
 *                           pxSolveKernel (the dataflow solver as a kernel of its own).
 * No tensor cores: nothing here is a dense contraction.  Compile with -fmad=false: every multiply-add must round
 * twice, exactly like the scalar CPU build.
 *
 * Phases of a substep and the reference code each one restates
 *   A  AABBs, stable order by aabb.min.x           px_body_array.c:81-102, px_circle.c:43-48
 *   B  sweep: ordered AABB-pass pair list, binned   px_world.c:355-392, px_body.c:126-145,
 *      by shape pair                                px_body_array.c:104-131
 *   C  narrowphase + manifold init + resting        px_collision.c:15-465, px_manifold.c:12-92
 *      contact + sleep flags from manifolds         px_world.c:147-177
 *   D  per-body ordered manifold lists (CSR),       (order = manifold pool order,
 *      integrate forces                             px_world.c:327-345), px_body.c:167-179
 *   E  sequential impulses, `iterations` sweeps     px_world.c:249-260, px_manifold.c:98-215
 *   F  integrate velocities, positional             px_body.c:181-188, px_manifold.c:217-255,
 *      correction, sleep update, clear forces       px_world.c:179-219, 302-311
 *
 * Order preservation.  The solver is Gauss-Seidel over the manifold pool; two manifold
 * visits commute bit-exactly iff they share no dynamic body (statics are never written,
 * px_body.c:148).  Phase D builds, for every dynamic body, the list of manifolds touching it
 * in pool order; phase E is a dataflow execution of the visit DAG those lists define: a
 * visit runs when it is at the head of the lists of both its bodies.  A scheduler warp keeps
 * the ready queue and publishes rounds of up to 32 mutually independent contact visits, a
 * math warp executes them (solverSchedule / solverMath), and iterations pipeline into each
 * other exactly as far as the dependencies allow.  Any such schedule produces the same bits
 * as the sequential loop.  Positional correction (px_manifold.c:217-255) depends only on the
 * manifold's normal/penetration and inverse masses, so each body sums its terms in list
 * order on its own thread.
 */
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "px_step.cuh"
#include <type_traits>

#define PX_FLAG_SLEEPING 8u /* PX_BODY_FLAG_SLEEPING, px_body.h:67 */
#define PX_OVERFLOW_INTERNAL 0x80000000u

/* ------------------------------------------------------------------ arithmetic contract */
/* px_math.h:50-114 -- bit hack + one Newton step; rcp = rsqrt^2; div = a*rcp(b) */
__device__ __forceinline__ float fRsqrt(float x) {
  float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(x) >> 1));
  return y * (1.5f - (x * 0.5f * y * y));
}
__device__ __forceinline__ float fRcp(float x) {
  float r = fRsqrt(x);
  return r * r;
}
__device__ __forceinline__ float fDiv(float a, float b) { return a * fRcp(b); }
__device__ __forceinline__ float fSqrt(float x) { return x * fRsqrt(x); }

#define D_PI 3.141592741f
#define D_2PI (D_PI * 2.0f)

/* px_math.h:226-241; the table is the host's (glibc sinf), never recomputed on the device */
__device__ __forceinline__ float fSin(const float *table, float radians) {
  radians = fmodf(radians, D_2PI);
  if (radians < 0) radians += D_2PI;
  float position = fDiv(radians, D_2PI / 256);
  int ip = (int)position;
  int i1 = ip & 255;
  int i2 = (i1 + 1) & 255;
  float fraction = position - (float)ip;
  return table[i1] * (1.0f - fraction) + table[i2] * fraction;
}

/* ------------------------------------------------- explicit shared-memory accesses */
/* The solver's two loops are single-warp dependency chains: every instruction on them is
 * latency.  They address shared memory by 32-bit window offsets computed once outside the
 * loop (generic pointers make the compiler rebuild the window base every iteration) and use
 * predicated stores instead of branches. */
__device__ __forceinline__ uint32_t sAddr(const void *p) {
  /* the asm pins the offset in a register: rebuilding it from the window base and a kernel
   * parameter every loop iteration is cheaper in registers but sits on the chain */
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));
  return a;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
/* flags another warp writes: acquire keeps later loads behind it, volatile keeps it a real load */
__device__ __forceinline__ uint32_t ldsAcquire(uint32_t a) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ldsVolatile(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
/* pulls a 128-byte line into L2 ahead of its use */
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
/* lines [0, bytes / 128] of a region, spread over the CTA's (or warp's) threads */
__device__ __forceinline__ void prefetchRegion(const uint8_t *base, uint32_t bytes, uint32_t tid, uint32_t nt) {
  for (uint32_t off = tid * 128u; off < bytes; off += nt * 128u) prefetchL2(base + off);
}
/* a global load the compiler may not sink below later volatile asm statements */
__device__ __forceinline__ float4 ldgPinned(const float4 *p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts128If(uint32_t a, float4 v, bool p) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.shared.v4.f32 [%0], {%1, %2, %3, %4}; }" ::"r"(a), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"((uint32_t)p));
}
/* predicated: idle lanes issue nothing (an unconditional add of 0 would pile every idle lane of
 * the warp onto one address, and same-address shared atomics serialise) */
__device__ __forceinline__ uint32_t atomAddSIf(uint32_t a, uint32_t v, bool p) {
  uint32_t old = 0;
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q atom.shared.add.u32 %0, [%1], %2; }"
               : "+r"(old)
               : "r"(a), "r"(v), "r"((uint32_t)p)
               : "memory");
  return old;
}

struct V2 {
  float x, y;
};
__device__ __forceinline__ V2 v2(float x, float y) { return V2{x, y}; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ V2 operator-(V2 a) { return v2(-a.x, -a.y); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float lensq(V2 a) { return a.x * a.x + a.y * a.y; }

struct M2 {
  float m00, m01, m10, m11;
};
__device__ __forceinline__ V2 mul(M2 m, V2 v) { return v2(m.m00 * v.x + m.m01 * v.y, m.m10 * v.x + m.m11 * v.y); }
__device__ __forceinline__ V2 mulT(M2 m, V2 v) { return v2(m.m00 * v.x + m.m10 * v.y, m.m01 * v.x + m.m11 * v.y); }

/* pxVec2Normalize, px_vec2.h:459-472 */
__device__ __forceinline__ V2 normalize(V2 v, float eps) {
  float len = fSqrt(lensq(v));
  if (len <= eps) return v;
  float inv = fRcp(len);
  return v2(v.x * inv, v.y * inv);
}

/* ----------------------------------------------------------------------- shared views */
/* Body-id space: dynamic slot s -> id s, static position q -> id D + q (NB = D + S).
 * Records are float4 so that one LDS.128 fetches what a phase needs of a body or manifold. */
struct Sm {
  float4 *bodyP; /* {pos.x, pos.y, iMass, iInertia}                 statics: iM = iI = 0 */
  float4 *bodyV; /* {vel.x, vel.y, angularVelocity, aabbRadius}     statics: zero velocity */
  float4 *bodyR; /* orientation {m00, m01, m10, m11} */
  uint32_t *shape; /* PX_SHAPE_PACK word by body id */
  const float *dynMat, *statMat; /* global immutable rows: materials are only read at manifold creation */
  uint8_t *flags, *order, *order2, *posOf, *restFlag, *wakeFlag, *sSlot;
  PxWorldHeader *hdr;
  uint32_t *misc;
  unsigned long long *scan;
  /* manifold records, float4[3] per slot, in this CTA's L2-resident scratch slot (global):
   *   [0] {normal.x, normal.y, mixedRestitution, staticFriction}
   *   [1] {dynamicFriction, penetration, c0.x, c0.y}
   *   [2] {c1.x, c1.y, rcpDenom0, rcpDenom1} */
  float4 *man;
  uint32_t *mids;
  uint16_t *validSlot, *base, *listStart, *cnt, *cntB;
  uint32_t *fillB;
  float4 *sBox;    /* {min.x, min.y, max.x, max.y} by sorted position */
  uint32_t *sInfo; /* slot | sleeping << 8 | isPolygon << 9, by sorted position */
  float4 *stBox;
  uint32_t *bins;
  const float4 *verts; /* shared memory or, when the pool does not fit, global through L1 */
  uint16_t *list, *nextw;
  uint32_t *sig;
  uint32_t *queue;
  /* static schedule (region Y'), see PxSmemPlan */
  uint16_t *predA, *predB, *rankOf;
};

/* Manifold records use ordinary (L1-allocating) accesses: a slot is only ever touched by CTAs of
 * one SM, i.e. through one L1, and writers and readers inside the CTA are separated by
 * __syncthreads, which makes global writes visible block-wide.  A record is read once per
 * solver iteration, so L1 hits pay. */
__device__ __forceinline__ float4 manLoad(const Sm &s, uint32_t ms, uint32_t k) { return s.man[3u * ms + k]; }
__device__ __forceinline__ void manStore(const Sm &s, uint32_t ms, float4 a, float4 b, float4 c) {
  s.man[3u * ms + 0u] = a;
  s.man[3u * ms + 1u] = b;
  s.man[3u * ms + 2u] = c;
}
/* {restitution, staticFriction, dynamicFriction} of a body id, from the immutable block */
__device__ __forceinline__ float3 materialOf(const Sm &s, uint32_t id, uint32_t D, uint32_t S) {
  if (id < D) return make_float3(__ldg(s.dynMat + PX_DYNC_E * D + id), __ldg(s.dynMat + PX_DYNC_SF * D + id), __ldg(s.dynMat + PX_DYNC_DF * D + id));
  const uint32_t q = id - D;
  return make_float3(__ldg(s.statMat + PX_STAT_E * S + q), __ldg(s.statMat + PX_STAT_SF * S + q), __ldg(s.statMat + PX_STAT_DF * S + q));
}

/* misc[] slots */
enum { MI_MCOUNT = 0, MI_QTAIL, MI_OVERFLOW, MI_SURVIVORS, MI_SLEEPING, MI_VISITS, MI_MATHROUND, MI_SLOT, MI_WORLD,
       MI_RING = 9 /* static schedule: longest body ring in contact visits */, MI_RDESC = 16 };
#define PX_RING 16u /* round descriptors in flight (misc[MI_RDESC..]) */
/* round descriptor: (round + 1) mod 2^15 << 17 | first queue index of the round << 6 | visits (PX_DESC_END = no more rounds) */
#define PX_DESC(round, head, n) (((((round) + 1u) & 0x7fffu) << 17) | (((head) & 0x7ffu) << 6) | (n))
#define PX_DESC_TAG(d) ((d) >> 17)
#define PX_DESC_HEAD(d) (((d) >> 6) & 0x7ffu)
#define PX_DESC_N(d) ((d) & 63u)
#define PX_DESC_END 63u
#ifndef PX_LAG
#define PX_LAG 8u /* rounds the scheduler warp may run ahead of the math warp */
#endif
/*
 * Rich manifold word (mids[] and ready-queue entries):
 *   slotA | bodyIdB << 8 | bIsDynamic << 17 | contactCount << 18 | manifoldSlot << 20 | contactIndex << 30
 */
#define MID_A(w) ((w) & 0xffu)
#define MID_B(w) (((w) >> 8) & 0x1ffu)
#define MID_BDYN(w) (((w) >> 17) & 1u)
#define MID_COUNT(w) (((w) >> 18) & 3u)
#define MID_SLOT(w) (((w) >> 20) & 0x3ffu)
#define MID_CI(w) (((w) >> 30) & 1u)

/* bins word: pairSeq << 17 | slotA << 9 | bodyIdB */
#define BIN_SEQ(w) ((w) >> 17)
#define BIN_A(w) (((w) >> 9) & 0xffu)
#define BIN_B(w) ((w) & 0x1ffu)

struct DBody {
  V2 pos;
  M2 rot;
  float radius;
  uint32_t nverts;
  const float4 *v; /* {vertex.x, vertex.y, normal.x, normal.y} */
};

__device__ __forceinline__ DBody loadBody(const Sm &s, uint32_t id) {
  DBody b;
  const float4 p = s.bodyP[id], r = s.bodyR[id];
  b.pos = v2(p.x, p.y);
  b.rot = M2{r.x, r.y, r.z, r.w};
  b.radius = s.bodyV[id].w;
  const uint32_t sh = s.shape[id];
  b.nverts = PX_SHAPE_COUNT(sh);
  b.v = s.verts + PX_SHAPE_OFFSET(sh);
  return b;
}

struct Contact {
  V2 normal;
  float pen;
  V2 c0, c1;
  uint32_t count;
};

/* pxCollideCircles, px_collision.c:15-52 */
__device__ __forceinline__ void collideCircles(const DBody &A, const DBody &B, float eps, Contact &m) {
  V2 n = B.pos - A.pos;
  float ra = A.radius, rb = B.radius;
  float combined = ra + rb;
  float d2 = lensq(n);
  m.count = 0;
  if (d2 >= combined * combined) return;
  float dist = fSqrt(d2);
  if (dist < eps) {
    m.pen = ra;
    m.normal = v2(1.0f, 0.0f);
    m.c0 = A.pos;
    m.count = 1;
    return;
  }
  m.pen = combined - dist;
  float inv = fRcp(dist); /* pxVec2Divf: each component times pxFastRcp(dist) */
  m.normal = v2(n.x * inv, n.y * inv);
  m.c0 = A.pos + m.normal * ra;
  m.count = 1;
}

/* pxCollideCirclePolygon, px_collision.c:61-178.  C is the circle, P the polygon; the normal
 * points from the circle to the polygon. */
__device__ void collideCirclePolygon(const DBody &C, const DBody &P, float eps, Contact &m) {
  m.count = 0;
  const float radius = C.radius;
  const M2 rot = P.rot;
  V2 center = mulT(rot, C.pos - P.pos);

  float separation = -FLT_MAX;
  uint32_t face = 0;
  for (uint32_t i = 0; i < P.nverts; i++) {
    float4 vn = P.v[i];
    float sdist = dot(v2(vn.z, vn.w), center - v2(vn.x, vn.y));
    if (sdist > radius) return;
    if (sdist > separation) {
      separation = sdist;
      face = i;
    }
  }
  float4 f1 = P.v[face];
  if (separation < eps) { /* centre inside */
    V2 n = -mul(rot, v2(f1.z, f1.w));
    m.normal = n;
    m.pen = radius;
    m.c0 = C.pos + n * radius;
    m.count = 1;
    return;
  }
  float4 f2 = P.v[face + 1 < P.nverts ? face + 1 : 0];
  V2 v1 = v2(f1.x, f1.y), v2_ = v2(f2.x, f2.y);
  float radiusSq = radius * radius;
  m.pen = radius - separation;

  float dot1 = dot(center - v1, v2_ - v1);
  if (dot1 <= 0) {
    V2 d = center - v1;
    if (dot(d, d) > radiusSq) return;
    m.normal = normalize(mul(rot, v1 - center), eps);
    m.c0 = mul(rot, v1) + P.pos;
    m.count = 1;
    return;
  }
  float dot2 = dot(center - v2_, v1 - v2_);
  if (dot2 <= 0) {
    V2 d = center - v2_;
    if (dot(d, d) > radiusSq) return;
    m.normal = normalize(mul(rot, v2_ - center), eps);
    m.c0 = mul(rot, v2_) + P.pos;
    m.count = 1;
    return;
  }
  V2 n = v2(f1.z, f1.w);
  if (dot(center - v1, n) > radius) return;
  n = -mul(rot, n);
  m.normal = n;
  m.c0 = C.pos + n * radius;
  m.count = 1;
}

/* One face of pxFindAxisLeastPenetration (px_collision.c:180-218, loop body :190-214) with
 * pxPolygonGetSupport (px_polygon.c:326-347) inlined: the separation of B's support point
 * along the normal of face i of A. */
__device__ __forceinline__ float faceSeparation(const DBody &A, const DBody &B, uint32_t i, float epsSq) {
  float4 an = A.v[i];
  V2 nw = mul(A.rot, v2(an.z, an.w));
  V2 n = mulT(B.rot, nw);
  V2 dir = -n;
  V2 sup = v2(0.0f, 0.0f);
  if (!(lensq(dir) < epsSq)) {
    float best = -FLT_MAX;
    float4 b0 = B.v[0];
    sup = v2(b0.x, b0.y);
    for (uint32_t k = 0; k < B.nverts; k++) {
      float4 bv = B.v[k];
      float proj = bv.x * dir.x + bv.y * dir.y;
      if (proj > best) {
        best = proj;
        sup = v2(bv.x, bv.y);
      }
    }
  }
  V2 v = mul(A.rot, v2(an.x, an.y)) + A.pos;
  v = v - B.pos;
  v = mulT(B.rot, v);
  return dot(n, sup - v);
}

/* pxFindAxisLeastPenetration evaluated by the PX_SAT_LANES lanes of a pair group: lane q takes
 * faces q, q + PX_SAT_LANES, ...; the reference keeps the FIRST face among the maxima (strict
 * `>` while scanning upwards), which the (value, lowest index) reduction reproduces. */
#ifndef PX_SAT_LANES
#define PX_SAT_LANES 2u
#endif
__device__ __forceinline__ float leastPenetrationGroup(uint32_t &faceIndex, const DBody &A, const DBody &B, uint32_t q,
                                                       bool valid, float epsSq) {
  float bestD = -FLT_MAX;
  uint32_t bestI = 0xffu;
  if (valid) {
    for (uint32_t i = q; i < A.nverts; i += PX_SAT_LANES) {
      const float d = faceSeparation(A, B, i, epsSq);
      if (d > bestD) {
        bestD = d;
        bestI = i;
      }
    }
  }
#pragma unroll
  for (uint32_t o = 1; o < PX_SAT_LANES; o <<= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, bestD, o);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bestI, o);
    if (od > bestD || (od == bestD && oi < bestI)) {
      bestD = od;
      bestI = oi;
    }
  }
  faceIndex = bestI == 0xffu ? 0u : bestI;
  return bestD;
}

/* The face of A whose normal points most towards `dWorld` (a heuristic pick: any index is valid). */
__device__ __forceinline__ uint32_t facingFace(const DBody &A, V2 dWorld) {
  const V2 dl = mulT(A.rot, dWorld);
  float best = -FLT_MAX;
  uint32_t face = 0;
  for (uint32_t i = 0; i < A.nverts; i++) {
    const float4 v = A.v[i];
    const float p = v.z * dl.x + v.w * dl.y;
    if (p > best) {
      best = p;
      face = i;
    }
  }
  return face;
}

/* Exact early rejection of a polygon pair.  pxCollidePolygons (px_collision.c:299-337) returns
 * without contacts as soon as the LARGEST face separation of either polygon is >= 0; one face
 * whose separation -- evaluated with the reference's own arithmetic (faceSeparation) -- is
 * >= 0 therefore proves the same outcome without looking at the other faces.  The face tried
 * is the one pointing most towards the other body (separatedByFacingFace: the face of X that looks at Y; two lanes
 * try the two polygons of a pair side by side); pairs neither can reject go on to the full two-pass SAT. */
__device__ __forceinline__ bool separatedByFacingFace(const DBody &X, const DBody &Y, float epsSq) {
  return faceSeparation(X, Y, facingFace(X, Y.pos - X.pos), epsSq) >= 0.0f;
}

/* pxClipSegmentToLine, px_collision.c:261-290 (IEEE division).  The reference appends, in this
 * order, the intersection point (signs differ), in[1] (d2 <= 0) and in[0] (d1 <= 0, if room).
 * Written as selects so that nothing is indexed dynamically (no local-memory stack frame). */
__device__ __forceinline__ int clipSegment(V2 &o0, V2 &o1, const V2 in0, const V2 in1, V2 normal, float offset) {
  const float d1 = dot(normal, in0) - offset;
  const float d2 = dot(normal, in1) - offset;
  const bool crossing = d1 * d2 < 0.0f, keep1 = d2 <= 0.0f, keep0 = d1 <= 0.0f;
  if (crossing) {
    /* strictly opposite signs: exactly one endpoint is kept, after the intersection */
    const float alpha = d1 / (d1 - d2);
    o0 = in0 + (in1 - in0) * alpha;
    o1 = keep1 ? in1 : in0;
    return 2;
  }
  o0 = keep1 ? in1 : in0;
  o1 = in0;
  return (keep1 ? 1 : 0) + (keep0 ? 1 : 0);
}

/* pxCollidePolygons after the two SAT passes, px_collision.c:312-435 (pxFindIncidentFace
 * :220-259 inlined): `flip` says B owns the reference face (pxBiasGt, :322-337) */
__device__ void clipPolygons(const DBody &A, const DBody &B, bool flip, uint32_t refIndex, float eps, Contact &m) {
  m.count = 0;
  const DBody &ref = flip ? B : A;
  const DBody &inc = flip ? A : B;

  /* incident face: most anti-parallel normal */
  float4 rf = ref.v[refIndex];
  V2 refNormalW = mul(ref.rot, v2(rf.z, rf.w));
  V2 nInc = mulT(inc.rot, refNormalW);
  uint32_t incFace = 0;
  float minDot = FLT_MAX;
  for (uint32_t i = 0; i < inc.nverts; i++) {
    float4 iv = inc.v[i];
    float d = nInc.x * iv.z + nInc.y * iv.w;
    if (d < minDot) {
      minDot = d;
      incFace = i;
    }
  }
  float4 i0 = inc.v[incFace];
  float4 i1 = inc.v[(incFace + 1) % inc.nverts];
  V2 inc0 = inc.pos + mul(inc.rot, v2(i0.x, i0.y));
  V2 inc1 = inc.pos + mul(inc.rot, v2(i1.x, i1.y));

  V2 ref0 = ref.pos + mul(ref.rot, v2(rf.x, rf.y));
  float4 rf1 = ref.v[(refIndex + 1) % ref.nverts];
  V2 ref1 = ref.pos + mul(ref.rot, v2(rf1.x, rf1.y));

  V2 side = normalize(ref1 - ref0, eps);
  V2 refNormal = v2(-side.y, side.x);
  if (flip) refNormal = -refNormal;

  float refC = dot(refNormal, ref0);
  float negSide = -dot(side, ref0);
  V2 clip1a, clip1b, clip2[2];
  int cp = clipSegment(clip1a, clip1b, inc0, inc1, -side, negSide);
  if (cp < 2) return;
  float posSide = dot(side, ref1);
  cp = clipSegment(clip2[0], clip2[1], clip1a, clip1b, side, posSide);
  if (cp < 2) return;

  m.normal = flip ? -refNormal : refNormal;
  /* px_collision.c:407-432: keep the clipped points behind the reference face, in order */
  const float sep0 = dot(refNormal, clip2[0]) - refC, sep1 = dot(refNormal, clip2[1]) - refC;
  const bool k0 = sep0 <= 0.0f, k1 = sep1 <= 0.0f;
  const V2 zero = v2(0.0f, 0.0f);
  float pen = k0 ? -sep0 : 0.0f;
  if (k1) pen = k0 ? pen + -sep1 : -sep1;
  if (k0 && k1) pen = fDiv(pen, 2.0f);
  m.pen = pen;
  m.c0 = k0 ? clip2[0] : (k1 ? clip2[1] : zero);
  m.c1 = (k0 && k1) ? clip2[1] : zero;
  m.count = (k0 ? 1u : 0u) + (k1 ? 1u : 0u);
}



/* ------------------------------------------------------------------------- block scan */
/* exclusive prefix sum of a packed counter word over the CTA, in thread order */
template <int NT>
__device__ __forceinline__ unsigned long long blockExclusiveScan(unsigned long long v, unsigned long long *scratch,
                                                                 unsigned long long &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  unsigned long long warpBase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; w++) {
    unsigned long long t = scratch[w];
    if (w < warp) warpBase += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return warpBase + inc - v;
}

/* pxBodyAABBsOverlap, px_body.c:126-145: strict comparisons, touching boxes overlap */
__device__ __forceinline__ bool boxesOverlap(const float4 a, const float4 b) {
  if (a.z < b.x || a.x > b.z) return false;
  if (a.w < b.y || a.y > b.w) return false;
  return true;
}

/* pxBodyArrayFindFirstIndexAfterX, px_body_array.c:104-131: predicate on max.x over a list
 * sorted by min.x, replicated literally */
__device__ __forceinline__ uint32_t firstStaticAfterX(const float4 *stBox, uint32_t n, float minX) {
  uint32_t low = 0, high = n;
  while (low < high) {
    uint32_t mid = low + (high - low) / 2;
    if (stBox[mid].z < minX) {
      low = mid + 1;
    } else {
      high = mid;
    }
  }
  return low;
}

/*
 * One contact of pxManifoldApplyImpulse (px_manifold.c:98-215).  The two-contact loop of
 * the reference is executed as two consecutive ready-queue visits of the same manifold
 * (contact 0 then contact 1, nothing in between can touch either body), so every solver
 * round costs one contact.  The caller guarantees nobody else touches the two bodies'
 * velocities while this runs.
 *
 * rcpDenom = pxFastRcp(invMassSum * contactCount) (px_manifold.c:148-156) depends only on
 * positions, the normal and the inverse masses, all frozen during the solver, so it is
 * computed once per contact in the narrowphase epilogue; the friction impulse divides by the
 * same denominator (px_manifold.c:193-194) and reuses it.
 *
 * px_world.c:252-255 skips a manifold when both bodies sleep; the sweep never creates one
 * for two sleeping dynamic bodies (px_world.c:367-369), sleep flags do not change between
 * the two, and statics never sleep, so the test is always false here.
 */
struct ContactIn {
  float4 mA, mB, mC, pA, pB, vA, vB;
};

/* every load of a contact visit depends only on the queue word: one shared-memory round trip */
__device__ __forceinline__ ContactIn loadContact(const Sm &s, uint32_t w) {
  const uint32_t ms = MID_SLOT(w), a = MID_A(w), b = MID_B(w);
  ContactIn in;
  in.mA = manLoad(s, ms, 0);
  in.mB = manLoad(s, ms, 1);
  in.mC = manLoad(s, ms, 2);
  in.pA = s.bodyP[a];
  in.pB = s.bodyP[b];
  in.vA = s.bodyV[a];
  in.vB = s.bodyV[b];
  return in;
}

/* the arithmetic of one contact visit from its frozen operands (normal, lever arms ra = c - posA and rb = c - posB,
 * materials, the precomputed denominator, inverse masses / inertias): vA / vB come back updated */
__device__ __forceinline__ void contactImpulseCore(const V2 n, const V2 ra, const V2 rb, const float e, const float sf, const float df,
                                                   const float rcpDenom, const float iMA, const float iIA, const float iMB,
                                                   const float iIB, const bool bdyn, float4 &vA, float4 &vB, const float eps) {
  V2 va = v2(vA.x, vA.y), vb = v2(vB.x, vB.y);
  float wa = vA.z, wb = vB.z;
  const V2 raPerp = v2(-ra.y, ra.x), rbPerp = v2(-rb.y, rb.x);
  V2 rv = (vb + rbPerp * wb) - (va + raPerp * wa);
  const float vn = dot(rv, n);
  /* The two early exits of the reference (`continue` when the bodies separate, skip friction
   * when the tangent vanishes) are applied as selects at the end: the visit sits on the
   * solver's critical chain, and a branch costs it more than the few discarded operations. */
  const bool approaching = !(vn > 0);

  const float normalMag = -(1.0f + e) * vn;
  const float j = normalMag * rcpDenom;
  const V2 impulse = n * j;
  /* pxBodyApplyImpulse(bodyA, -impulse, ra), px_body.c:147-156 */
  const V2 ni = -impulse;
  const V2 va1 = v2(va.x + ni.x * iMA, va.y + ni.y * iMA);
  const float wa1 = wa + cross(ra, ni) * iIA;
  /* bodyB only if dynamic (px_manifold.c:165-167); a static keeps its exact zeros */
  const V2 vb1 = bdyn ? v2(vb.x + impulse.x * iMB, vb.y + impulse.y * iMB) : vb;
  const float wb1 = bdyn ? wb + cross(rb, impulse) * iIB : wb;

  rv = (vb1 + rbPerp * wb1) - (va1 + raPerp * wa1);
  V2 tangent = rv - n * dot(rv, n);
  const float tlen = fSqrt(lensq(tangent));
  const bool slides = tlen > eps;
  tangent = tangent * fRcp(tlen);
  const float tmag = -dot(rv, tangent);
  const float jt = tmag * rcpDenom;
  const float fmag = (fabsf(jt) < j * sf) ? jt : (-j * df);
  const V2 friction = tangent * fmag;
  const V2 nf = -friction;
  const V2 va2 = v2(va1.x + nf.x * iMA, va1.y + nf.y * iMA);
  const float wa2 = wa1 + cross(ra, nf) * iIA;
  const V2 vb2 = bdyn ? v2(vb1.x + friction.x * iMB, vb1.y + friction.y * iMB) : vb1;
  const float wb2 = bdyn ? wb1 + cross(rb, friction) * iIB : wb1;

  vA.x = approaching ? (slides ? va2.x : va1.x) : va.x;
  vA.y = approaching ? (slides ? va2.y : va1.y) : va.y;
  vA.z = approaching ? (slides ? wa2 : wa1) : wa;
  vB.x = approaching ? (slides ? vb2.x : vb1.x) : vb.x; /* a static B: vb2 == vb1 == vb, exact zeros */
  vB.y = approaching ? (slides ? vb2.y : vb1.y) : vb.y;
  vB.z = approaching ? (slides ? wb2 : wb1) : wb;
}

/* one contact visit of the dynamic solver: operands from the manifold record and the two body rows */
__device__ __forceinline__ void contactImpulse(uint32_t w, ContactIn &in, float eps) {
  const bool second = MID_CI(w) != 0, bdyn = MID_BDYN(w) != 0;
  const float4 mA = in.mA, mB = in.mB, mC = in.mC;
  const float4 pA = in.pA, pB = in.pB;
  const V2 c = second ? v2(mC.x, mC.y) : v2(mB.z, mB.w);
  const V2 pa = v2(pA.x, pA.y), pb = v2(pB.x, pB.y);
  contactImpulseCore(v2(mA.x, mA.y), c - pa, c - pb, mA.z, mA.w, mB.x, second ? mC.w : mC.z, pA.z, pA.w, pB.z, pB.w, bdyn,
                     in.vA, in.vB, eps);
}

__device__ __forceinline__ void solveContact(const Sm &s, uint32_t w, ContactIn &in, float eps) {
  contactImpulse(w, in, eps);
  s.bodyV[MID_A(w)] = in.vA;
  if (MID_BDYN(w)) s.bodyV[MID_B(w)] = in.vB;
}

__device__ __forceinline__ void applyContactImpulse(const Sm &s, uint32_t w, float eps) {
  ContactIn in = loadContact(s, w);
  solveContact(s, w, in, eps);
}

__device__ __forceinline__ void bindSmem(Sm &s, uint8_t *base, const PxSmemPlan &S, const float4 *globalVerts) {
#define BIND(field, type) s.field = (type *)(base + S.field)
  BIND(bodyP, float4); BIND(bodyV, float4); BIND(bodyR, float4); BIND(shape, uint32_t);
  BIND(flags, uint8_t); BIND(order, uint8_t); BIND(order2, uint8_t); BIND(posOf, uint8_t);
  BIND(restFlag, uint8_t); BIND(wakeFlag, uint8_t); BIND(sSlot, uint8_t);
  BIND(hdr, PxWorldHeader); BIND(misc, uint32_t); BIND(scan, unsigned long long);
  BIND(mids, uint32_t);
  BIND(validSlot, uint16_t); BIND(base, uint16_t); BIND(listStart, uint16_t); BIND(cnt, uint16_t); BIND(cntB, uint16_t);
  BIND(fillB, uint32_t);
  BIND(sBox, float4); BIND(sInfo, uint32_t); BIND(stBox, float4);
  BIND(bins, uint32_t);
  BIND(list, uint16_t); BIND(nextw, uint16_t); BIND(sig, uint32_t); BIND(queue, uint32_t);
  BIND(predA, uint16_t); BIND(predB, uint16_t); BIND(rankOf, uint16_t);
#undef BIND
  s.verts = S.vertsInSmem ? (const float4 *)(base + S.verts) : globalVerts;
}

template <int NT>
__device__ __forceinline__ void copyRows(float *dst, const float *src, uint32_t nFloats) {
  /* 16-byte coalesced copy; nFloats is a multiple of 4 and both sides are 16-byte aligned */
  const float4 *s4 = reinterpret_cast<const float4 *>(src);
  float4 *d4 = reinterpret_cast<float4 *>(dst);
  for (uint32_t i = threadIdx.x; i < nFloats / 4; i += NT) d4[i] = s4[i];
}

/* bytes between shared memory and a world's hand-over record (either direction), 16 at a time;
 * both sides are 16-byte aligned and padded, `bytes` is rounded up */
template <int NT>
__device__ __forceinline__ void handCopy(void *dst, const void *src, uint32_t bytes) {
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (uint32_t i = threadIdx.x; i < (bytes + 15u) / 16u; i += NT) d4[i] = s4[i];
}

/* ------------------------------------------------------------------ solver warps */
/* The two solver loops are functions of their own (never inlined): their register allocation
 * is then independent of the rest of the kernel, whose pressure (narrowphase, sweep) would
 * otherwise spill into these latency-critical chains. */
struct SolverCtx {
  uint32_t aQueue, aNext, aMids, aSig, aDesc, aMathRound, aBodyP, aBodyV;
  uint32_t qmask, qcap, iterations, tail0, lagSleepNs;
  const float4 *man;
  float eps;
};

/* returns rounds | (visits is left in *visitsOut) */
__device__ __noinline__ uint32_t solverSchedule(const SolverCtx c, uint32_t *visitsOut) {
  const uint32_t lane = threadIdx.x & 31u, QMASK = c.qmask, iterations = c.iterations;
  const uint32_t aQueue = c.aQueue, aNext = c.aNext, aMids = c.aMids, aSig = c.aSig, aDesc = c.aDesc, aMathRound = c.aMathRound;
  uint32_t head = 0, tail = c.tail0, visits = 0, rounds = 0;
  const uint32_t aDummy = aQueue + 4u * c.qcap; /* queue[qcap] is a write-only sink */
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t mr = 0;
#pragma unroll 2
  while (head != tail) {
    const uint32_t n = min(tail - head, 32u);
    while (rounds - mr >= PX_LAG) { /* far enough ahead: leave the issue slots to the others */
      __nanosleep(c.lagSleepNs);
      mr = ldsVolatile(aMathRound);
    }
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(aDesc + 4u * (rounds & (PX_RING - 1u))),
                 "r"(PX_DESC(rounds, head & QMASK, n))
                 : "memory");
    const bool active = lane < n;
    const uint32_t qw = lds32(aQueue + 4u * ((head + lane) & QMASK));
    const uint32_t w = active ? qw : 0u; /* idle lanes: a harmless word, every index in range */
    const bool again = active && (MID_CI(w) + 1u < MID_COUNT(w)); /* second contact follows */
    const bool sigA = active && !again, sigB = sigA && MID_BDYN(w);
    /* successors of this manifold on body a / body b are fixed for the substep:
     * slot | successor has a dynamic B << 14 | successor's side for this body << 15 */
    const uint32_t nx = lds32(aNext + 4u * MID_SLOT(w)); /* side a | side b << 16 */
    /* (the b half is unset when B is static, the a half when no manifold 0 exists) */
    const uint32_t t0 = sigA ? (nx & 0x3ffu) : 0u, t1 = sigB ? ((nx >> 16) & 0x3ffu) : 0u;
    const uint32_t wt0 = lds32(aMids + 4u * t0), wt1 = lds32(aMids + 4u * t1);
    /* predicated, not branched around */
    const uint32_t old0 = atomAddSIf(aSig + 4u * t0, (nx & 0x8000u) ? 0x10000u : 1u, sigA);
    const uint32_t old1 = atomAddSIf(aSig + 4u * t1, (nx & 0x80000000u) ? 0x10000u : 1u, sigB);
    const uint32_t mine0 = ((nx & 0x8000u) ? (old0 >> 16) : (old0 & 0xffffu)) + 1u;
    const uint32_t other0 = (nx & 0x8000u) ? (old0 & 0xffffu) : (old0 >> 16);
    const uint32_t mine1 = ((nx & 0x80000000u) ? (old1 >> 16) : (old1 & 0xffffu)) + 1u;
    const uint32_t other1 = (nx & 0x80000000u) ? (old1 & 0xffffu) : (old1 >> 16);
    const bool ready0 = sigA && mine0 <= iterations && (!(nx & 0x4000u) || other0 == mine0);
    const bool ready1 = sigB && mine1 <= iterations && (!(nx & 0x40000000u) || other1 == mine1);
    const bool has0 = again || ready0;
    const uint32_t push0 = again ? (w | (1u << 30)) : wt0;
    const uint32_t b0 = __ballot_sync(0xffffffffu, has0);
    const uint32_t b1 = __ballot_sync(0xffffffffu, ready1);
    const uint32_t c0 = __popc(b0);
    sts32(has0 ? aQueue + 4u * ((tail + __popc(b0 & lt)) & QMASK) : aDummy, push0);
    sts32(ready1 ? aQueue + 4u * ((tail + c0 + __popc(b1 & lt)) & QMASK) : aDummy, wt1);
    tail += c0 + __popc(b1);
    head += n;
    visits += sigA ? 1u : 0u;
    rounds++;
    mr = ldsVolatile(aMathRound);
    __syncwarp();
  }
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(aDesc + 4u * (rounds & (PX_RING - 1u))),
               "r"(PX_DESC(rounds, 0u, PX_DESC_END)) /* end of schedule */
               : "memory");
  *visitsOut = visits;
  return rounds;
}

/* Math warp: executes the published rounds in order, one contact visit per lane (the two contacts
 * of a manifold are two consecutive visits: handling both in one visit was measured, twice -- 27 %
 * fewer rounds, each 48 % longer).  While round r runs, the queue word and the three manifold
 * records of round r + 1 are already in flight; only the velocities and the frozen body rows are
 * read at the last moment, from shared memory.  Idle lanes run the arithmetic on manifold 0 and
 * store nothing. */
__device__ __noinline__ void solverMath(const SolverCtx c) {
  const uint32_t lane = threadIdx.x & 31u, QMASK = c.qmask;
  const uint32_t aQueue = c.aQueue, aDesc = c.aDesc, aMathRound = c.aMathRound, aBodyP = c.aBodyP, aBodyV = c.aBodyV;
  const float4 *man = c.man;
  const float eps = c.eps;
  uint32_t r = 0, d = ldsAcquire(aDesc);
  uint32_t w = 0, n = 0;
  float4 mA, mB, mC;
  bool have = false; /* w / n / m* hold round r */
  for (;; r++) {
    if (!have) {
      const uint32_t expect = (r + 1u) & 0x7fffu;
      while (PX_DESC_TAG(d) != expect) d = ldsAcquire(aDesc + 4u * (r & (PX_RING - 1u)));
      n = PX_DESC_N(d);
      if (n == PX_DESC_END) break;
      const uint32_t qw = lds32(aQueue + 4u * ((PX_DESC_HEAD(d) + lane) & QMASK));
      w = lane < n ? qw : 0u;
      mA = ldgPinned(man + 3u * MID_SLOT(w));
      mB = ldgPinned(man + 3u * MID_SLOT(w) + 1u);
      mC = ldgPinned(man + 3u * MID_SLOT(w) + 2u);
    }
    const bool active = lane < n;
    const uint32_t a16 = MID_A(w) * 16u, b16 = MID_B(w) * 16u;
    ContactIn in;
    in.mA = mA;
    in.mB = mB;
    in.mC = mC;
    in.pA = lds128(aBodyP + a16);
    in.pB = lds128(aBodyP + b16);
    in.vA = lds128(aBodyV + a16);
    in.vB = lds128(aBodyV + b16);
    /* next round: descriptor, queue word, records */
    const uint32_t dn = ldsAcquire(aDesc + 4u * ((r + 1u) & (PX_RING - 1u)));
    const uint32_t nNext = PX_DESC_N(dn);
    const bool haveNext = PX_DESC_TAG(dn) == ((r + 2u) & 0x7fffu) && nNext != PX_DESC_END;
    const uint32_t qwN = lds32(aQueue + 4u * ((PX_DESC_HEAD(dn) + lane) & QMASK));
    const uint32_t wNext = (haveNext && lane < nNext) ? qwN : 0u;
    const float4 nA = ldgPinned(man + 3u * MID_SLOT(wNext));
    const float4 nB = ldgPinned(man + 3u * MID_SLOT(wNext) + 1u);
    const float4 nC = ldgPinned(man + 3u * MID_SLOT(wNext) + 2u);
    asm volatile("" : "+f"(in.vA.x), "+f"(in.vB.x), "+f"(in.pA.x), "+f"(in.pB.x));
    contactImpulse(w, in, eps);
    sts128If(aBodyV + a16, in.vA, active);
    sts128If(aBodyV + b16, in.vB, active && MID_BDYN(w));
    __syncwarp();
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(aMathRound), "r"(r + 1u) : "memory");
    have = haveNext;
    d = dn;
    w = wNext;
    n = nNext;
    mA = nA;
    mB = nB;
    mC = nC;
  }
}

#ifndef PX_MINB128
#define PX_MINB128 5
#endif
#define PX_DMAX 256 /* dynamic slots per world (PX_MAX_BODIES rounded up to the copy granule) */

/* ============================================================================ kernel ==== */
template <int NT, int MODE>
#ifndef PX_SWEEP_MINB256
#define PX_SWEEP_MINB256 3
#endif
__global__ void __launch_bounds__(NT, NT == 128 ? PX_MINB128 : (MODE == PX_MODE_SWEEP ? PX_SWEEP_MINB256 : 3)) pxStepKernel(const PxStepParams P) {
  constexpr int R = PX_DMAX / NT; /* body slots (and sorted positions) owned per thread */
  extern __shared__ __align__(16) uint8_t smemRaw[];
  const PxBatchLayout &L = P.L;
  const uint32_t D = L.maxDynamic, S = L.maxStatic, MM = L.maxManifolds, MP = L.maxPairs;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  Sm s;
  bindSmem(s, smemRaw, P.S, nullptr);
  const float eps = L.epsilon, epsSq = L.epsilonSq, allowedPen = L.allowedPenetration;
  const uint32_t QMASK = P.S.qcap - 1;
  /* manifold records of whatever world this CTA is working on: scratch slot blockIdx.x of this
   * launch's region.  Only this CTA ever touches it, through the L1 of the SM it lives on. */
  s.man = P.scratch + (size_t)blockIdx.x * MM * 3u;

  /* Persistent CTA: the grid is one wave of resident CTAs; each takes the next world off a
   * counter until the batch is done (no ragged last wave, and a slow world delays nobody). */
  for (;;) {
  if (tid == 0) s.misc[MI_WORLD] = atomicAdd(P.worldCounter, 1u);
  __syncthreads();
  const uint32_t world = s.misc[MI_WORLD];
  if (world >= L.nWorlds) break;
  uint8_t *gmut = P.mut + (size_t)world * L.mutStride;
  const uint8_t *gimm = P.imm + (size_t)world * L.immStride;
  s.verts = P.S.vertsInSmem ? (const float4 *)(smemRaw + P.S.verts) : (const float4 *)(gimm + L.offVerts);
  s.dynMat = (const float *)(gimm + L.offDynConst);
  s.statMat = (const float *)(gimm + L.offStat);
  const float *gdynC = (const float *)(gmut + L.offDyn);
  float *gdyn = (float *)(gmut + L.offDyn);

  /* -------------------------------------------- load (HBM SoA rows -> smem records) */
  /* global rows are read coalesced (consecutive slots by consecutive threads); the record
   * transpose happens on the shared-memory side */
  {
    const float *gc = (const float *)(gimm + L.offDynConst);
    const uint32_t *gshape = (const uint32_t *)(gimm + L.offDynShape);
    for (uint32_t x = tid; x < D; x += NT) {
      s.bodyP[x] = make_float4(gdynC[PX_DYN_PX * D + x], gdynC[PX_DYN_PY * D + x], gc[PX_DYNC_IMASS * D + x],
                               gc[PX_DYNC_IINERTIA * D + x]);
      s.bodyV[x] = make_float4(gdynC[PX_DYN_VX * D + x], gdynC[PX_DYN_VY * D + x], gdynC[PX_DYN_AV * D + x],
                               gc[PX_DYNC_RADIUS * D + x]);
      s.bodyR[x] = make_float4(gdynC[PX_DYN_M00 * D + x], gdynC[PX_DYN_M01 * D + x], gdynC[PX_DYN_M10 * D + x],
                               gdynC[PX_DYN_M11 * D + x]);
      s.shape[x] = gshape[x];
    }
    const float *gs = (const float *)(gimm + L.offStat);
    const float *gsm = (const float *)(gmut + L.offStatMut); /* static poses: mutable by command, never by the step */
    const uint32_t *gsshape = (const uint32_t *)(gimm + L.offStatShape);
    for (uint32_t q = tid; q < S; q += NT) {
      s.bodyP[D + q] = make_float4(gsm[PX_STATM_PX * S + q], gsm[PX_STATM_PY * S + q], 0.0f, 0.0f);
      s.bodyV[D + q] = make_float4(0.0f, 0.0f, 0.0f, gs[PX_STAT_RADIUS * S + q]);
      s.bodyR[D + q] = make_float4(gsm[PX_STATM_M00 * S + q], gsm[PX_STATM_M01 * S + q], gsm[PX_STATM_M10 * S + q],
                                   gsm[PX_STATM_M11 * S + q]);
      s.shape[D + q] = gsshape[q];
    }
    copyRows<NT>((float *)s.flags, (const float *)(gmut + L.offDynFlags), D / 4);
    copyRows<NT>((float *)s.order, (const float *)(gmut + L.offDynOrder), D / 4);
    copyRows<NT>((float *)s.hdr, (const float *)(gmut + L.offHeader), sizeof(PxWorldHeader) / 4);
    if (tid < 32 && tid != MI_WORLD) s.misc[tid] = 0;
    if (MODE == PX_MODE_SWEEP && P.doFinish && P.solver != PX_SOLVER_STATIC) {
      /* split pipeline, dataflow solver: the lists of the substep being finished come back from the world's hand-over record -- with
       * the state loads above, in ONE round trip to DRAM (whole capacities: the manifold count is not awaited) */
      const uint8_t *h = P.hand + (size_t)world * P.H.stride;
      handCopy<NT>(s.mids, h + P.H.mids, MM * 4u);
      handCopy<NT>(s.list, h + P.H.list, 2u * MM * 2u);
      handCopy<NT>(s.listStart, h + P.H.listStart, (D + 1u) * 2u);
      handCopy<NT>(s.cnt, h + P.H.cnt, D * 2u);
      handCopy<NT>(s.cntB, h + P.H.cntB, D * 2u);
      handCopy<NT>(s.posOf, h + P.H.posOf, D);
      handCopy<NT>(s.restFlag, h + P.H.restFlag, D);
      handCopy<NT>(s.wakeFlag, h + P.H.wakeFlag, D);
    }
  }
  /* thread-private body state: slot x = tid + r * NT lives in this thread's registers */
  float rAngle[R], rFx[R], rFy[R], rTq[R], rSlp[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t x = tid + r * NT;
    const bool in = x < D;
    rAngle[r] = in ? gdynC[PX_DYN_ANGLE * D + x] : 0.0f;
    rFx[r] = in ? gdynC[PX_DYN_FX * D + x] : 0.0f;
    rFy[r] = in ? gdynC[PX_DYN_FY * D + x] : 0.0f;
    rTq[r] = in ? gdynC[PX_DYN_TORQUE * D + x] : 0.0f;
    rSlp[r] = in ? gdynC[PX_DYN_SLEEP * D + x] : 0.0f;
  }
  /* static solver, phase F of the substep just solved: the body's word (its correction moves in H.cvec, the sleep flags
   * of its sorted position), loaded with the state */
  uint32_t rWord[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t x = tid + r * NT;
    rWord[r] = (MODE == PX_MODE_SWEEP && P.solver == PX_SOLVER_STATIC && P.doFinish && x < D)
                   ? ((const uint32_t *)(P.hand + (size_t)world * P.H.stride + P.H.bodyF))[x] : 0u;
  }
  __syncthreads();

  const uint32_t nDyn = min(s.hdr->nDynamic, D), nStat = min(s.hdr->nStatic, S);
  const float dt = s.hdr->dt;
  const V2 g = v2(s.hdr->gravityX, s.hdr->gravityY);
  const uint32_t iterations = s.hdr->iterations & 255u;
  const float linSleepSq = s.hdr->linearSleepThresholdSq, angSleep = s.hdr->angularSleepThreshold;
  const float timeToSleep = s.hdr->timeToSleep;
  uint32_t overflow = 0; /* thread 0 accumulates the world's flags */
  uint32_t statPairs = 0, statManifolds = 0;

  long long profT = 0;
  unsigned long long *prof = P.prof ? P.prof + (size_t)world * PX_PROF_SLOTS : nullptr;
#define PROF_MARK(slot)                              \
  if (prof && tid == 0) {                            \
    long long now = clock64();                       \
    prof[slot] += (unsigned long long)(now - profT); \
    profT = now;                                     \
  }
  if (prof && tid == 0) profT = clock64();
  /* finer marks inside a phase: thread 0's arrival time, accumulated next to (not instead of) the phase total */
  long long profS = 0;
#define PROF_SUB(slot)                               \
  if (prof && tid == 0) {                            \
    long long now = clock64();                       \
    prof[slot] += (unsigned long long)(now - (profS > profT ? profS : profT)); \
    profS = now;                                     \
  }

  /* Split pipeline (PX_MODE_SWEEP): this launch runs phase F of the substep whose solver has
   * just finished (pxSolveKernel) and then phases A-D of the next one; everything the two
   * kernels and phase F exchange beyond the mutable block lives in the world's hand-over
   * record in global memory.  Fused (PX_MODE_FUSED): K whole substeps, nothing leaves the SM. */
  constexpr bool sweepMode = MODE == PX_MODE_SWEEP; /* compile-time: each pipeline carries only its own code */
  /* split pipeline: the solver kernel either runs the dataflow schedule itself (pxSolveKernel) or replays a periodic
   * schedule that phase D derives here, once per substep (pxReplayKernel) */
  const bool staticSched = sweepMode && P.solver == PX_SOLVER_STATIC;
  uint8_t *hand = sweepMode ? P.hand + (size_t)world * P.H.stride : nullptr;
  if (sweepMode) {
    /* dataflow solver: the manifold records cross to the solver kernel and back to phase F in the hand-over record;
     * static solver: nothing outside this launch reads them (the visit operands and correction moves are derived from
     * them in phase D), so they stay in the CTA's L2-resident scratch slot like the fused kernel's */
    s.man = staticSched ? P.scratch + (size_t)blockIdx.x * MM * 3u : (float4 *)(hand + P.H.records);
    statPairs = s.hdr->statPairs;
    statManifolds = s.hdr->statManifolds;
  }
  uint32_t nPairs = 0, M = 0;
  const uint32_t nLoop = sweepMode ? 2u : P.kSubsteps;
  for (uint32_t step = 0; step < nLoop; step++) {
    const bool runSweep = sweepMode ? (step == 1u && P.doSweep) : true;
    const bool runFinish = sweepMode ? (step == 0u && P.doFinish) : true;
    const bool traceNow = P.trace != nullptr && (sweepMode ? P.traceNow != 0u : step + 1 == P.kSubsteps);
    uint8_t *gtrace = traceNow ? P.trace + (size_t)world * L.traceStride : nullptr;
    if (runSweep) {

    /* ====================== phase A: AABB keys, stable order by aabb.min.x ============ */
    /* the order persists across substeps, so it is almost always already sorted */
    if (tid < 32 && tid != MI_OVERFLOW && tid != MI_WORLD && !(sweepMode && tid == MI_SLEEPING)) s.misc[tid] = 0;
    for (uint32_t i = tid; i < D; i += NT) {
      s.restFlag[i] = 0;
      s.wakeFlag[i] = 0;
      s.fillB[i] = 0;
    }
    float *keys = (float *)s.bins; /* region X scratch: sort keys by position */
    for (uint32_t p = tid; p < nDyn; p += NT) {
      const uint32_t slot = s.order[p];
      keys[p] = s.bodyP[slot].x - s.bodyV[slot].w;
    }
    __syncthreads();
    {
      /* Stable sort by key (px_body_array.c:81-102 moves an element left only past strictly
       * greater keys, so equal keys keep their order and the result is the unique stable
       * permutation).  The list is nearly sorted: odd-even transposition passes that swap
       * only strictly out-of-order neighbours reach the same permutation in a few passes;
       * rank-by-counting is the fallback. */
      bool unsorted = false;
      for (uint32_t p = tid; p + 1 < nDyn; p += NT) unsorted |= keys[p] > keys[p + 1];
      int dirty = __syncthreads_or(unsorted);
      for (int pass = 0; dirty && pass < 12; pass++) {
#pragma unroll
        for (uint32_t parity = 0; parity < 2; parity++) {
          for (uint32_t i = tid; 2 * i + parity + 1 < nDyn; i += NT) {
            const uint32_t p0 = 2 * i + parity;
            const float k0 = keys[p0], k1 = keys[p0 + 1];
            if (k0 > k1) {
              keys[p0] = k1;
              keys[p0 + 1] = k0;
              const uint8_t o0 = s.order[p0];
              s.order[p0] = s.order[p0 + 1];
              s.order[p0 + 1] = o0;
            }
          }
          __syncthreads();
        }
        unsorted = false;
        for (uint32_t p = tid; p + 1 < nDyn; p += NT) unsorted |= keys[p] > keys[p + 1];
        dirty = __syncthreads_or(unsorted);
      }
      if (dirty) {
        /* keys compared as order-preserving integers: a total order even if a diverged world
         * holds NaN positions, so the ranks are always a permutation (no out-of-range slot) */
        auto ordered = [](float f) {
          const uint32_t u = __float_as_uint(f);
          return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        };
        for (uint32_t p = tid; p < nDyn; p += NT) {
          const float key = keys[p];
          const uint32_t ukey = ordered(key);
          uint32_t rank = 0;
          for (uint32_t q = 0; q < nDyn; q++) {
            const float kq = keys[q];
            const uint32_t uq = ordered(kq);
            /* -0.0f == 0.0f in the reference's float comparison: equal keys keep their order */
            const bool less = (kq == kq && key == key) ? kq < key : uq < ukey;
            const bool same = (kq == kq && key == key) ? kq == key : uq == ukey;
            rank += (less || (same && q < p)) ? 1u : 0u;
          }
          s.order2[rank] = s.order[p];
        }
        __syncthreads();
        for (uint32_t p = tid; p < nDyn; p += NT) s.order[p] = s.order2[p];
        __syncthreads();
      }
    }
    /* sorted-position records for the sweep: radius-based AABBs (px_circle.c:43-48,
     * px_polygon.c:261-266) */
    for (uint32_t p = tid; p < nDyn; p += NT) {
      const uint32_t slot = s.order[p];
      const float4 bp = s.bodyP[slot];
      const float r = s.bodyV[slot].w;
      s.sBox[p] = make_float4(bp.x - r, bp.y - r, bp.x + r, bp.y + r);
      const uint32_t poly = PX_SHAPE_COUNT(s.shape[slot]) != 0;
      s.sInfo[p] = slot | ((s.flags[slot] & PX_FLAG_SLEEPING) ? 0x100u : 0u) | (poly << 9);
      s.sSlot[p] = (uint8_t)slot;
      s.posOf[slot] = (uint8_t)p;
    }
    for (uint32_t q = tid; q < nStat; q += NT) {
      const float4 bp = s.bodyP[D + q];
      const float r = s.bodyV[D + q].w;
      s.stBox[q] = make_float4(bp.x - r, bp.y - r, bp.x + r, bp.y + r);
    }
    __syncthreads();
    PROF_MARK(0)

    /* ====================== phase B: sweep -> ordered pair list, binned by shape pair == */
    /* pass 1 counts, a packed 4x16-bit block scan gives every body its first sequence
     * number and its first index in each of the three type bins, pass 2 writes.  The
     * reference's early-exit `break` acts as `continue`; over a sorted list
     * both give the same pairs, so the dynamic scan stops at the first min.x > A.max.x.
     * Pass 1 remembers which candidates of the scan window passed the AABB test in a bit
     * mask, so pass 2 only revisits those. */
    unsigned long long exr[R];
    uint32_t startStatR[R], dynMaskR[R], statMaskR[R];
    unsigned long long carry = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t p = r * NT + tid;
      unsigned long long counts = 0; /* all | cc << 16 | cp << 32 | pp << 48 */
      uint32_t startStat = 0, dynMask = 0, statMask = 0;
      if (p < nDyn) {
        const float4 a = s.sBox[p];
        const uint32_t ai = s.sInfo[p];
        const uint32_t aSleep = ai & 0x100u, aPoly = (ai >> 9) & 1u;
        for (uint32_t q = p + 1; q < nDyn; q++) {
          const float4 b = s.sBox[q];
          if (b.x > a.z) break;
          const uint32_t bi = s.sInfo[q];
          if (aSleep & bi) continue; /* both sleeping */
          if (!boxesOverlap(a, b)) continue;
          counts += 1ull + (1ull << (16 * (1 + aPoly + ((bi >> 9) & 1u))));
          const uint32_t k = q - p - 1;
          dynMask |= k < 32 ? (1u << k) : 0u;
          if (k >= 32) dynMask = 0xffffffffu, counts |= 1ull << 63; /* window too long: rescan */
        }
        startStat = firstStaticAfterX(s.stBox, nStat, a.x);
        for (uint32_t q = startStat; q < nStat; q++) {
          const float4 b = s.stBox[q];
          if (b.x > a.z) continue;
          if (!boxesOverlap(a, b)) continue;
          const uint32_t bPoly = PX_SHAPE_COUNT(s.shape[D + q]) != 0;
          counts += 1ull + (1ull << (16 * (1 + aPoly + bPoly)));
          const uint32_t k = q - startStat;
          statMask |= k < 32 ? (1u << k) : 0u;
        }
      }
      const bool rescan = (counts >> 63) != 0;
      counts &= ~(1ull << 63);
      unsigned long long total;
      exr[r] = blockExclusiveScan<NT>(counts, s.scan, total) + carry;
      carry += total;
      startStatR[r] = startStat;
      dynMaskR[r] = rescan ? 0xffffffffu : dynMask;
      statMaskR[r] = statMask;
      if (rescan) exr[r] |= 1ull << 63;
      if (p < nDyn) s.base[p] = (uint16_t)min((uint32_t)(exr[r] & 0xffffu), MP);
    }
    const uint32_t totalPairsRaw = (uint32_t)(carry & 0xffffu);
    nPairs = min(totalPairsRaw, MP);
    const uint32_t nCC = (uint32_t)((carry >> 16) & 0xffffu), nCP = (uint32_t)((carry >> 32) & 0xffffu);
    const uint32_t nPP = (uint32_t)((carry >> 48) & 0x7fffu);
    if (tid == 0) {
      s.base[nDyn] = (uint16_t)nPairs;
      if (totalPairsRaw > MP) overflow |= PX_BATCH_OVERFLOW_PAIRS;
    }
    /* bin regions inside bins[]: [0,nCC) circle-circle, [nCC, nCC+nCP) circle-polygon (either
     * order), then polygon-polygon; entries whose sequence number is >= MP are dropped.
     * bins[] doubles as the sort-key scratch of phase A: all key reads are behind us. */
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t p = r * NT + tid;
      if (p >= nDyn) continue;
      const uint32_t ai = s.sInfo[p];
      const uint32_t aSlot = ai & 0xffu, aPoly = (ai >> 9) & 1u;
      const bool rescan = (exr[r] >> 63) != 0;
      const unsigned long long ex = exr[r] & ~(1ull << 63);
      uint32_t seq = (uint32_t)(ex & 0xffffu);
      uint32_t binAt[3] = {(uint32_t)((ex >> 16) & 0xffffu), nCC + (uint32_t)((ex >> 32) & 0xffffu),
                           nCC + nCP + (uint32_t)((ex >> 48) & 0x7fffu)};
      uint16_t *tpA = gtrace ? (uint16_t *)(gtrace + L.offTracePairs) : nullptr;
      auto emit = [&](uint32_t bId, uint32_t bPoly, uint32_t traceB) {
        const uint32_t t = aPoly + bPoly;
        const uint32_t at = binAt[t]++;
        /* pair overflow (flagged): pairs beyond either capacity are dropped and the world's
         * results are unspecified from here on -- but every sequence number below MP must still
         * be reset, or phase D would link last substep's manifold slot to it */
        if (seq < MP) {
          s.validSlot[seq] = 0;
          if (at < MP) s.bins[at] = (seq << 17) | (aSlot << 9) | bId;
          if (tpA) {
            tpA[seq] = (uint16_t)aSlot;
            tpA[MP + seq] = (uint16_t)traceB;
          }
        } else if (at < MP) {
          s.bins[at] = 0xffffffffu;
        }
        seq++;
      };
      if (!rescan) {
        uint32_t m = dynMaskR[r];
        while (m) {
          const uint32_t k = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t bi = s.sInfo[p + 1 + k];
          emit(bi & 0xffu, (bi >> 9) & 1u, 0x8000u | (bi & 0xffu));
        }
      } else {
        const float4 a = s.sBox[p];
        const uint32_t aSleep = ai & 0x100u;
        for (uint32_t q = p + 1; q < nDyn; q++) {
          const float4 b = s.sBox[q];
          if (b.x > a.z) break;
          const uint32_t bi = s.sInfo[q];
          if (aSleep & bi) continue;
          if (!boxesOverlap(a, b)) continue;
          emit(bi & 0xffu, (bi >> 9) & 1u, 0x8000u | (bi & 0xffu));
        }
      }
      if (nStat <= 32 + startStatR[r]) {
        uint32_t m = statMaskR[r];
        while (m) {
          const uint32_t k = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t q = startStatR[r] + k;
          const uint32_t bPoly = PX_SHAPE_COUNT(s.shape[D + q]) != 0;
          emit(D + q, bPoly, tpA ? (uint32_t)(gimm + L.offStatSlot)[q] : 0u);
        }
      } else {
        const float4 a = s.sBox[p];
        for (uint32_t q = startStatR[r]; q < nStat; q++) {
          const float4 b = s.stBox[q];
          if (b.x > a.z) continue;
          if (!boxesOverlap(a, b)) continue;
          const uint32_t bPoly = PX_SHAPE_COUNT(s.shape[D + q]) != 0;
          emit(D + q, bPoly, tpA ? (uint32_t)(gimm + L.offStatSlot)[q] : 0u);
        }
      }
    }
    if (P.S.vertsInSmem) {
      /* region X holds the vertex pool during the narrowphase; the solver's region Y overlays
       * it afterwards, so it is refreshed from L2 every substep */
      copyRows<NT>((float *)(smemRaw + P.S.verts), (const float *)(gimm + L.offVerts), L.maxVertices * 4);
    }
    __syncthreads();
    PROF_MARK(1)

    /* ====================== phase C: narrowphase ===================================== */
    /* C1 circle pairs -> C2a exact quick rejection of polygon pairs, survivors compacted ->
     * C2b two-pass SAT on the survivors -> C3 clipping -> C4 manifold init, one thread per
     * manifold.  C1 and C3 only stage the contact geometry of a hit. */
    uint32_t statSurvivors = 0, statPoly = 0;
    {
      const uint32_t nAll = min(nCC + nCP + nPP, MP), nCircle = min(nCC + nCP, nAll), nPoly = nAll - nCircle;
      /* a pair with contacts takes the next manifold slot and parks {normal, penetration | c0, c1} there */
      auto stage = [&](uint32_t word, uint32_t a, uint32_t b, const Contact &m) {
        const uint32_t ms = atomicAdd(&s.misc[MI_MCOUNT], 1u);
        if (ms >= MM) { /* manifold overflow, flagged below */
          s.validSlot[BIN_SEQ(word)] = 0;
          return;
        }
        s.man[3u * ms + 0u] = make_float4(m.normal.x, m.normal.y, m.pen, 0.0f);
        s.man[3u * ms + 1u] = make_float4(m.c0.x, m.c0.y, m.c1.x, m.c1.y);
        s.mids[ms] = a | (b << 8) | ((b < D ? 1u : 0u) << 17) | (m.count << 18) | (ms << 20);
        s.validSlot[BIN_SEQ(word)] = (uint16_t)(ms + 1);
      };

      /* C1: circle-circle and circle-polygon pairs, one pair per thread */
      for (uint32_t idx = tid; idx < nCircle; idx += NT) {
        const uint32_t word = s.bins[idx];
        if (word == 0xffffffffu) continue;
        const uint32_t a = BIN_A(word), b = BIN_B(word);
        Contact m;
        m.count = 0;
        m.c1 = v2(0.0f, 0.0f);
        {
          DBody A = loadBody(s, a), B = loadBody(s, b);
          if (idx < nCC) {
            collideCircles(A, B, eps, m);
          } else if (A.nverts == 0) {
            /* pxCollide, px_collision.c:445-459: the polygon-circle order swaps the arguments
             * and flips the normal */
            collideCirclePolygon(A, B, eps, m);
          } else {
            collideCirclePolygon(B, A, eps, m);
            m.normal = -m.normal;
          }
        }
        if (m.count) stage(word, a, b, m);
      }

      PROF_SUB(8)
      /* C2a: polygon pairs, one pair per thread: most AABB-pass pairs do not touch, and one
       * well-chosen face proves it.  Survivors (bin indices) are compacted with a warp ballot
       * into the sweep's dead AABB array; their order does not matter -- a manifold's place in
       * the pool comes from its pair's sequence number. */
      uint16_t *surv = (uint16_t *)s.sBox;
      /* pass 1: reject; a survivor leaves its size class (1: both polygons <= 4 vertices, 2: <= 6,
       * 3: larger) in validSlot[seq], which the sweep zeroed */
      for (uint32_t base = 0; base < nPoly; base += NT / 2u) {
        /* two lanes per pair: the even one tries A's face that looks at B, the odd one B's face that looks at A */
        const uint32_t k = base + (tid >> 1), q = tid & 1u;
        const uint32_t word = k < nPoly ? s.bins[nCircle + k] : 0xffffffffu;
        const bool valid = word != 0xffffffffu;
        const uint32_t a = valid ? BIN_A(word) : 0u, b = valid ? BIN_B(word) : 0u;
        const DBody X = loadBody(s, q ? b : a), Y = loadBody(s, q ? a : b);
        bool separated = valid && separatedByFacingFace(X, Y, epsSq);
        separated |= __shfl_xor_sync(0xffffffffu, separated, 1) != 0;
        if (valid && !separated && q == 0) {
          const uint32_t big = max(X.nverts, Y.nverts);
          s.validSlot[BIN_SEQ(word)] = (uint16_t)(1u + (big > 4u ? 1u : 0u) + (big > 6u ? 1u : 0u));
        }
      }
      __syncthreads();
      /* pass 2: survivors compacted class by class, so that the lanes of a warp of the SAT pass
       * below run loops of similar length */
      for (uint32_t cls = 1; cls <= 3; cls++) {
        for (uint32_t base = 0; base < nPoly; base += NT) {
          const uint32_t k = base + tid;
          bool mine = false;
          if (k < nPoly) {
            const uint32_t word = s.bins[nCircle + k];
            mine = word != 0xffffffffu && s.validSlot[BIN_SEQ(word)] == cls;
          }
          const uint32_t mask = __ballot_sync(0xffffffffu, mine);
          uint32_t at = 0;
          if (lane == 0 && mask) at = atomicAdd(&s.misc[MI_SURVIVORS], __popc(mask));
          at = __shfl_sync(0xffffffffu, at, 0);
          if (mine) surv[at + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)(nCircle + k);
        }
      }
      __syncthreads();
      PROF_SUB(9)
      const uint32_t nSurv = s.misc[MI_SURVIVORS];
      statSurvivors = nSurv;
      statPoly = nPoly;

      /* C2b: polygon-polygon SAT, PX_SAT_LANES lanes per surviving pair (the faces of a polygon
       * are spread over the lanes of its group).  Colliding pairs leave
       * {1 << 15 | flip << 8 | reference face} in validSlot[seq] for the clip pass; separated pairs
       * get 0 back. */
      {
        const uint32_t q = tid & (PX_SAT_LANES - 1u), grp = tid / PX_SAT_LANES;
        for (uint32_t base = 0; base < nSurv; base += NT / PX_SAT_LANES) {
          const uint32_t k = base + grp;
          const bool valid = k < nSurv;
          const uint32_t word = valid ? s.bins[surv[k]] : 0u;
          const uint32_t a = BIN_A(word), b = BIN_B(word);
          const DBody A = loadBody(s, a), B = loadBody(s, b);
          uint32_t faceA, faceB = 0;
          const float penA = leastPenetrationGroup(faceA, A, B, q, valid, epsSq);
          const bool alive = valid && !(penA >= 0.0f);
          /* the second pass is skipped only when no pair of the warp needs it */
          float penB = 0.0f;
          if (__any_sync(0xffffffffu, alive)) penB = leastPenetrationGroup(faceB, B, A, q, alive, epsSq);
          if (valid && q == 0) { /* the size class written by C2a is replaced by the verdict */
            const bool hit = alive && !(penB >= 0.0f);
            const bool flip = !(penA >= penB * 0.95f + penA * 0.01f); /* pxBiasGt */
            s.validSlot[BIN_SEQ(word)] = hit ? (uint16_t)(0x8000u | (flip ? 0x100u : 0u) | (flip ? faceB : faceA)) : (uint16_t)0;
          }
        }
      }
      __syncthreads();
      PROF_SUB(10)

      /* C3: reference-face clipping for the SAT survivors, one pair per thread */
      for (uint32_t k = tid; k < nSurv; k += NT) {
        const uint32_t word = s.bins[surv[k]];
        const uint32_t sat = s.validSlot[BIN_SEQ(word)];
        if (!sat) continue;
        const uint32_t a = BIN_A(word), b = BIN_B(word);
        Contact m;
        m.c1 = v2(0.0f, 0.0f);
        {
          DBody A = loadBody(s, a), B = loadBody(s, b);
          clipPolygons(A, B, (sat & 0x100u) != 0, sat & 0xffu, eps, m);
        }
        if (m.count) {
          stage(word, a, b, m);
        } else {
          s.validSlot[BIN_SEQ(word)] = 0;
        }
      }
      __syncthreads();
      PROF_SUB(11)

      /* C4: pxManifoldInit + resting contact + solver denominators + sleep flags, one manifold
       * per thread (every lane busy) */
      const uint32_t nStaged = min(s.misc[MI_MCOUNT], MM);
      const V2 gstep = g * dt;
      const float gsq = lensq(gstep);
      for (uint32_t ms = tid; ms < nStaged; ms += NT) {
        const uint32_t w = s.mids[ms];
        const uint32_t a = MID_A(w), b = MID_B(w), bdyn = MID_BDYN(w), count = MID_COUNT(w);
        const float4 g0 = s.man[3u * ms + 0u], g1 = s.man[3u * ms + 1u];
        const V2 normal = v2(g0.x, g0.y);
        const float3 mtA = materialOf(s, a, D, S), mtB = materialOf(s, b, D, S);
        const float4 pA = s.bodyP[a], pB = s.bodyP[b], vA = s.bodyV[a], vB = s.bodyV[b];
        /* pxManifoldInit, px_manifold.c:12-31: geometric-mean friction via pxFastSqrt */
        const float sfm = fSqrt(mtA.y * mtB.y), dfm = fSqrt(mtA.z * mtB.z);
        /* pxManifoldDetectRestingContact, px_manifold.c:46-92, on the velocities as they are
         * before force integration (a static's are exact zeros) */
        float minE = mtA.x < mtB.x ? mtA.x : mtB.x;
        const V2 pa = v2(pA.x, pA.y), pb = v2(pB.x, pB.y);
        const V2 va = v2(vA.x, vA.y), vb = v2(vB.x, vB.y);
        const float wa = vA.z, wb = vB.z;
        const float fcount = (float)count;
        float rcp[2] = {0.0f, 0.0f};
        bool restingHit = false;
#pragma unroll
        for (uint32_t i = 0; i < 2; i++) {
          if (i < count) {
            const V2 c = i == 0 ? v2(g1.x, g1.y) : v2(g1.z, g1.w);
            const V2 ra = c - pa, rb = c - pb;
            const V2 rv = (vb + v2(-rb.y, rb.x) * wb) - (va + v2(-ra.y, ra.x) * wa);
            /* the reference stops at the first hit; later contacts cannot un-hit it */
            restingHit = restingHit || (lensq(rv) < gsq + eps);
            /* per-contact solver denominator, px_manifold.c:142-156 */
            const float raCrossN = cross(ra, normal), rbCrossN = cross(rb, normal);
            const float invMassSum = pA.z + pB.z + raCrossN * raCrossN * pA.w + rbCrossN * rbCrossN * pB.w;
            rcp[i] = fRcp(invMassSum * fcount);
          }
        }
        if (restingHit) minE = 0.0f;
        manStore(s, ms, make_float4(normal.x, normal.y, minE, sfm), make_float4(dfm, g0.z, g1.x, g1.y),
                 make_float4(g1.z, g1.w, rcp[0], rcp[1]));
        /* manifold pass of pxUpdateSleepStates, px_world.c:147-177: flags set by SLOT */
        const bool resting = (minE == 0.0f);
        if (!bdyn && resting) {
          s.restFlag[a] = 1;
        } else if (bdyn && !resting) {
          if (s.flags[a] & PX_FLAG_SLEEPING) s.wakeFlag[b] = 1;
          if (s.flags[b] & PX_FLAG_SLEEPING) s.wakeFlag[a] = 1;
        }
        if (bdyn) atomicAdd(&s.fillB[b], 1u); /* bodyB-role manifold count of b */
        if (staticSched && g0.z > allowedPen) {
          /* pxManifoldPositionalCorrection, px_manifold.c:217-255: the two moves depend on nothing the solver changes;
           * phase F (in pxReplayKernel) adds them to the integrated positions body by body, in pool order */
          const float excess = g0.z - allowedPen;
          const float mag = fDiv(excess * 0.2f, pA.z + pB.z);
          const V2 corr = normal * mag;
          const V2 dA = -(corr * pA.z), dB = corr * pB.z;
          ((float4 *)(smemRaw + P.S.cm))[ms] = make_float4(dA.x, dA.y, dB.x, dB.y);
          s.mids[ms] = w | 0x80000000u; /* bit 31: the manifold corrects positions */
        }
      }
    }
    __syncthreads();
    PROF_MARK(2)
    if (prof && tid == 0) prof[7] += ((unsigned long long)statPoly << 32) + statSurvivors;
    const uint32_t mRaw = s.misc[MI_MCOUNT];
    M = min(mRaw, MM);
    if (tid == 0 && mRaw > MM) overflow |= PX_BATCH_OVERFLOW_MANIFOLDS;

    /* ====================== phase D: per-body manifold lists, integrate forces ======== */
    /* Body x (sorted position p) meets its manifolds in pool order: first the ones where it
     * is bodyB (their A-bodies come earlier in the sweep, ordered by A's sorted position),
     * then its own block where it is bodyA, in sequence order.  Region X is dead from here;
     * region Y (list, nextw, sig, queue) overlays it. */
    uint32_t rankStartR[R];
    carry = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t p = r * NT + tid;
      uint32_t nA = 0, nB = 0, x = 0;
      if (p < nDyn) {
        x = s.sSlot[p];
        for (uint32_t seq = s.base[p]; seq < s.base[p + 1]; seq++) nA += s.validSlot[seq] != 0;
        nB = s.fillB[x];
      }
      unsigned long long total;
      unsigned long long ex = blockExclusiveScan<NT>((unsigned long long)(nA + nB) | ((unsigned long long)nA << 32), s.scan, total) + carry;
      carry += total;
      rankStartR[r] = (uint32_t)(ex >> 32);
      if (p < nDyn) {
        s.listStart[x] = (uint16_t)(ex & 0xffffffffu);
        s.cnt[x] = (uint16_t)(nA + nB);
        s.cntB[x] = (uint16_t)nB;
        s.fillB[x] = 0;
      }
    }
    if (staticSched) {
      for (uint32_t i = tid; i < M; i += NT) s.predB[i] = 0xffffu; /* by pool rank; stays so for a static B */
    } else {
      for (uint32_t i = tid; i < M; i += NT) s.sig[i] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t p = r * NT + tid;
      if (p >= nDyn) continue;
      const uint32_t x = s.sSlot[p];
      const uint32_t startX = s.listStart[x], nBx = s.cntB[x];
      uint32_t k = 0;
      for (uint32_t seq = s.base[p]; seq < s.base[p + 1]; seq++) {
        uint32_t vs = s.validSlot[seq];
        if (!vs) continue;
        const uint32_t ms = vs - 1;
        const uint32_t ids = s.mids[ms];
        s.list[startX + nBx + k] = (uint16_t)ms;
        if (staticSched) s.rankOf[ms] = (uint16_t)(rankStartR[r] + k); /* place in the manifold pool (px_world.c:327-345) */
        if (MID_BDYN(ids)) { /* unordered placement in B's bodyB segment, sorted below */
          const uint32_t y = MID_B(ids);
          s.list[s.listStart[y] + atomicAdd(&s.fillB[y], 1u)] = (uint16_t)ms;
        }
        if (gtrace) { /* ordered manifold dump */
          const uint32_t rank = rankStartR[r] + k;
          const float4 mA = manLoad(s, ms, 0), mB = manLoad(s, ms, 1), mC = manLoad(s, ms, 2);
          float *mf = (float *)(gtrace + L.offTraceManifolds);
          mf[0 * MM + rank] = mA.x;
          mf[1 * MM + rank] = mA.y;
          mf[2 * MM + rank] = mB.y;
          mf[3 * MM + rank] = mA.w;
          mf[4 * MM + rank] = mB.x;
          mf[5 * MM + rank] = mA.z;
          mf[6 * MM + rank] = mB.z;
          mf[7 * MM + rank] = mB.w;
          mf[8 * MM + rank] = mC.x;
          mf[9 * MM + rank] = mC.y;
          uint32_t bId = MID_B(ids);
          uint32_t slotB = MID_BDYN(ids) ? bId : (uint32_t)(gimm + L.offStatSlot)[bId - D];
          ((uint32_t *)(mf + (size_t)PX_TRACE_M_F32 * MM))[rank] = MID_A(ids) | (slotB << 8) | (MID_BDYN(ids) << 16) | (MID_COUNT(ids) << 24);
        }
        k++;
      }
    }
    __syncthreads();
    /* by slot owner: order the bodyB segment by the A-body's sorted position, derive each
     * manifold's successor on this body, signal the first manifold, integrate forces */
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t x = tid + r * NT;
      if (x >= D || !(s.flags[x] & 4u)) continue;
      const uint32_t start = s.listStart[x], n = s.cnt[x], nBx = s.cntB[x];
      for (uint32_t i = 1; i < nBx; i++) { /* insertion sort, a handful of entries */
        const uint32_t cur = s.list[start + i];
        const uint32_t key = s.posOf[MID_A(s.mids[cur])];
        int jj = (int)i - 1;
        while (jj >= 0) {
          const uint32_t other = s.list[start + jj];
          if (s.posOf[MID_A(s.mids[other])] <= key) break;
          s.list[start + jj + 1] = (uint16_t)other;
          jj--;
        }
        s.list[start + jj + 1] = (uint16_t)cur;
      }
      if (staticSched) {
        /* the body's ring of visits: every manifold learns its predecessor on this body (the last one closes the
         * ring: its edge carries the iteration boundary) and how many rounds that predecessor takes */
        uint32_t ring = 0;
        for (uint32_t i = 0; i < n; i++) {
          const uint32_t ms = s.list[start + i];
          const uint32_t pred = s.list[start + (i ? i - 1 : n - 1)];
          const uint32_t own = MID_COUNT(s.mids[ms]);
          /* pool rank of the predecessor | closes the ring << 10 | its contacts << 11 | own contacts << 13 */
          const uint32_t word = s.rankOf[pred] | (i ? 0u : 0x400u) | (MID_COUNT(s.mids[pred]) << 11) | (own << 13);
          if (i < nBx) s.predB[s.rankOf[ms]] = (uint16_t)word;
          else s.predA[s.rankOf[ms]] = (uint16_t)word;
          ring += own;
        }
        if (ring) atomicMax(&s.misc[MI_RING], ring);
        /* phase F's inputs for this body (pxReplayKernel): its positional-correction moves in pool order, and the
         * sleep flags of its SORTED POSITION -- set by slot, tested by position: the reference's behaviour */
        {
          const float4 *cmS = (const float4 *)(smemRaw + P.S.cm);
          float2 *cvec = (float2 *)(hand + P.H.cvec) + start;
          uint32_t k = 0;
          for (uint32_t i = 0; i < n; i++) {
            const uint32_t ms = s.list[start + i];
            if (s.mids[ms] >> 31) {
              const float4 c = cmS[ms];
              cvec[k++] = i >= nBx ? make_float2(c.x, c.y) : make_float2(c.z, c.w);
            }
          }
          const uint32_t p = s.posOf[x];
          ((uint32_t *)(hand + P.H.bodyF))[x] = start | (k << 16) | (s.wakeFlag[p] ? 0x40000000u : 0u) | (s.restFlag[p] ? 0x80000000u : 0u);
        }
      } else {
      for (uint32_t i = 0; i < n; i++) {
        const uint32_t ms = s.list[start + i];
        const uint32_t side = i < nBx ? 1u : 0u;
        const uint32_t in = i + 1 == n ? 0u : i + 1;
        const uint32_t succ = s.list[start + in];
        s.nextw[2 * ms + side] = (uint16_t)(succ | (MID_BDYN(s.mids[succ]) << 14) | ((in < nBx ? 1u : 0u) << 15));
      }
      if (n) { /* the body's first manifold receives its first signal from the body itself */
        const uint32_t first = s.list[start];
        atomicAdd(&s.sig[first], nBx ? 0x10000u : 1u);
      }
      }
      /* pxIntegrateForces, px_world.c:231-239 / px_body.c:167-179 (resting-contact detection
       * above already consumed the pre-integration velocities) */
      if (!(s.flags[x] & PX_FLAG_SLEEPING)) {
        const float halfDt = fDiv(dt, 2.0f);
        const float4 bp = s.bodyP[x];
        float4 bv = s.bodyV[x];
        const V2 dv = (v2(rFx[r], rFy[r]) * bp.z + g) * halfDt;
        bv.x = bv.x + dv.x;
        bv.y = bv.y + dv.y;
        bv.z = bv.z + rTq[r] * bp.w * halfDt;
        s.bodyV[x] = bv;
      }
    }
    __syncthreads();
    /* manifolds at the head of the lists of all their dynamic bodies start the solver */
    if (iterations && !(P.flags & PX_BATCH_FLAG_SERIAL_SOLVER) && !staticSched) {
      for (uint32_t ms = tid; ms < M; ms += NT) {
        const uint32_t w = s.mids[ms];
        const uint32_t sg = s.sig[ms];
        if ((sg & 0xffffu) == 1u && (!MID_BDYN(w) || (sg >> 16) == 1u)) {
          s.queue[atomicAdd(&s.misc[MI_QTAIL], 1u) & QMASK] = w;
        }
      }
    }
    __syncthreads();
    PROF_MARK(3)
    if constexpr (sweepMode) {
      if (staticSched && iterations) {
        /* static schedule: the rings in pool-rank space and, by rank, the frozen operands of every contact visit
         * {n, ra | rb, rcpDenom, e | sf, df, ids, inverse inertia of A} (entry 2 * rank + contact); pxReplayKernel
         * orders them */
        float4 *pl0 = (float4 *)(hand + P.H.vrec), *pl1 = pl0 + PX_PLANE, *pl2 = pl1 + PX_PLANE;
        for (uint32_t v = tid; v < 2u * M; v += NT) {
          const uint32_t m = v >> 1, ci = v & 1u, w = s.mids[m];
          if (ci < MID_COUNT(w)) {
            const uint32_t a = MID_A(w), b = MID_B(w), at = 2u * s.rankOf[m] + ci;
            const float4 mA = manLoad(s, m, 0), mB = manLoad(s, m, 1), mC = manLoad(s, m, 2);
            const float4 pA = s.bodyP[a], pB = s.bodyP[b];
            const V2 cpt = ci ? v2(mC.x, mC.y) : v2(mB.z, mB.w);
            const V2 ra = cpt - v2(pA.x, pA.y), rb = cpt - v2(pB.x, pB.y); /* px_manifold.c:112-113 */
            pl0[at] = make_float4(mA.x, mA.y, ra.x, ra.y);
            pl1[at] = make_float4(rb.x, rb.y, ci ? mC.w : mC.z, mA.z);
            pl2[at] = make_float4(mA.w, mB.x, __uint_as_float(a | (b << 8) | (MID_BDYN(w) << 17)), pA.w);
          }
        }
        handCopy<NT>(hand + P.H.predA, s.predA, M * 2u);
        handCopy<NT>(hand + P.H.predB, s.predB, M * 2u);
      }
      PROF_MARK(12)
      /* hand the substep over: what the solver kernel needs (dynamic: manifold ids, successor words, initial
       * signals, the first ready visits; static: the visit stream written above and its segment table) and what
       * phase F needs after it (the per-body lists) */
      if (staticSched) {
        if (tid == 0) ((uint32_t *)(hand + P.H.scalars))[PX_HS_RING] = s.misc[MI_RING];
      } else {
      handCopy<NT>(hand + P.H.mids, s.mids, M * 4u);
      handCopy<NT>(hand + P.H.nextw, s.nextw, M * 4u);
      handCopy<NT>(hand + P.H.sig, s.sig, M * 4u);
      handCopy<NT>(hand + P.H.queue, s.queue, min(s.misc[MI_QTAIL], P.S.qcap) * 4u);
      handCopy<NT>(hand + P.H.list, s.list, 2u * M * 2u);
      handCopy<NT>(hand + P.H.listStart, s.listStart, (D + 1u) * 2u);
      handCopy<NT>(hand + P.H.cnt, s.cnt, D * 2u);
      handCopy<NT>(hand + P.H.cntB, s.cntB, D * 2u);
      handCopy<NT>(hand + P.H.posOf, s.posOf, D);
      handCopy<NT>(hand + P.H.restFlag, s.restFlag, D);
      handCopy<NT>(hand + P.H.wakeFlag, s.wakeFlag, D);
      }
      if (tid == 0) {
        uint32_t *sc = (uint32_t *)(hand + P.H.scalars);
        sc[0] = M;
        sc[1] = s.misc[MI_QTAIL];
      }
      statPairs = nPairs;
      statManifolds = M;
    }
    } /* runSweep */

    /* ====================== phase E: sequential impulses, dataflow order ============== */
    if (sweepMode) {
      /* the solver is a kernel of its own (pxSolveKernel) */
    } else if (P.flags & PX_BATCH_FLAG_SERIAL_SOLVER) {
      /* debug: literal pool-order walk by one thread */
      if (tid == 0) {
        for (uint32_t it = 0; it < iterations; it++) {
          for (uint32_t p = 0; p < nDyn; p++) {
            for (uint32_t seq = s.base[p]; seq < s.base[p + 1]; seq++) {
              uint32_t vs = s.validSlot[seq];
              if (!vs) continue;
              const uint32_t w = s.mids[vs - 1];
              applyContactImpulse(s, w, eps);
              if (MID_COUNT(w) > 1) applyContactImpulse(s, w | (1u << 30), eps);
            }
          }
        }
      }
    } else if (warp < 2 && iterations) {
      /*
       * Two specialised warps.  The ORDER in which manifold visits may run depends only on the
       * manifold lists, never on a velocity, so scheduling is split from the arithmetic:
       *   warp 0  scheduler: runs the dataflow bookkeeping and publishes rounds of up to 32
       *           mutually independent visits (solverSchedule);
       *   warp 1  math: executes the rounds in order, one contact visit per lane (solverMath).
       * The solver is bound by its dependency chain -- more than 200 rounds per substep in a
       * settled 252-body pile at 10 iterations whatever the round width (profiles/sched_sim.py) --
       * so the two chains (ready queue / signals here, load / impulse / store there) run side by
       * side instead of adding up.
       *
       * sig[m] counts the signals manifold m has received from its A body (low half) and its
       * B body (high half).  Visit v of a manifold may run once each of its dynamic bodies has
       * delivered v + 1 signals: one from the body at list-build time if the manifold heads
       * that body's list, one each time the body's previous manifold (cyclically) finishes a
       * visit.  A signal is one shared-memory atomicAdd; its return value shows both counts
       * as they were, so exactly one of the (at most two) signalers sees the other side
       * already complete and enqueues the visit.
       *
       * Round r is the queue segment [head_r, head_r + n_r); rdesc[r % PX_RING] = PX_DESC(r,
       * head_r, n_r) publishes it (release store; the segment's words were stored in earlier
       * rounds).  Everything signalled in round r is pushed behind the segments of rounds <= r,
       * so a visit is always published in a later round than the visits it depends on.  The
       * scheduler stays at most PX_LAG rounds ahead of the math warp, which bounds the
       * descriptor ring and keeps unexecuted segments from being overwritten (qcap >= MM + 32 *
       * (PX_LAG + 1)).
       */
      SolverCtx c;
      c.aQueue = sAddr(s.queue); c.aNext = sAddr(s.nextw); c.aMids = sAddr(s.mids); c.aSig = sAddr(s.sig);
      c.aDesc = sAddr(s.misc + MI_RDESC); c.aMathRound = sAddr(s.misc + MI_MATHROUND);
      c.aBodyP = sAddr(s.bodyP); c.aBodyV = sAddr(s.bodyV);
      c.qmask = QMASK; c.qcap = P.S.qcap; c.iterations = iterations; c.tail0 = s.misc[MI_QTAIL]; c.lagSleepNs = P.lagSleepNs;
      c.man = s.man; c.eps = eps;
      if (warp == 0) {
        uint32_t visits = 0;
        const uint32_t rounds = solverSchedule(c, &visits);
        visits = __reduce_add_sync(0xffffffffu, visits);
        if (lane == 0) s.misc[MI_VISITS] = visits;
        if (prof && lane == 0) prof[6] += rounds;
      } else {
        solverMath(c);
      }
    }
    __syncthreads();
    if (!sweepMode && tid == 0 && iterations && !(P.flags & PX_BATCH_FLAG_SERIAL_SOLVER) && s.misc[MI_VISITS] != M * iterations)
      overflow |= PX_OVERFLOW_INTERNAL;
    PROF_MARK(4)

    if (runFinish) {
    /* ====================== phase F: integrate velocities, correct positions, sleep ==== */
    uint32_t sleepingLocal = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t x = tid + r * NT;
      if (x >= D || !(s.flags[x] & 4u)) continue;
      const uint32_t p = staticSched ? 0u : s.posOf[x];
      uint32_t fl = s.flags[x];
      float4 bp = s.bodyP[x];
      float4 bv = s.bodyV[x];
      float posx = bp.x, posy = bp.y;
      /* pxBodyIntegrateVelocity, px_body.c:181-188 */
      if (!(fl & PX_FLAG_SLEEPING)) {
        posx = posx + bv.x * dt;
        posy = posy + bv.y * dt;
        float radians = rAngle[r] + bv.z * dt;
        radians = fmodf(radians, D_2PI);
        if (radians < 0) radians += D_2PI;
        rAngle[r] = radians;
        const float c = fSin(P.sinTable, radians + D_PI * 0.5f), sn = fSin(P.sinTable, radians);
        s.bodyR[x] = make_float4(c, -sn, sn, c);
      }
      /* pxCorrectPositions, px_world.c:288-292: this body's terms, in pool order */
      if (staticSched) { /* precomputed by the previous launch's phase D: they depend on nothing the solver changes */
        const float2 *cvec = (const float2 *)(hand + P.H.cvec);
        const uint32_t start = rWord[r] & 0xffffu, n = (rWord[r] >> 16) & 0x3fffu;
        for (uint32_t i = 0; i < n; i++) {
          const float2 d = cvec[start + i];
          posx = posx + d.x;
          posy = posy + d.y;
        }
      } else {
        const uint32_t start = s.listStart[x], n = s.cnt[x], nBx = s.cntB[x];
        for (uint32_t i = 0; i < n; i++) {
          const uint32_t ms = s.list[start + i];
          const float pen = manLoad(s, ms, 1).y;
          if (pen <= allowedPen) continue;
          const uint32_t ids = s.mids[ms];
          const float4 mA = manLoad(s, ms, 0);
          const float iMA = s.bodyP[MID_A(ids)].z, iMB = s.bodyP[MID_B(ids)].z;
          const float excess = pen - allowedPen;
          const float mag = fDiv(excess * 0.2f, iMA + iMB);
          const V2 corr = v2(mA.x, mA.y) * mag;
          if (i >= nBx) { /* bodyA: moveBy(-(corr * iMA)) */
            const V2 d = -(corr * iMA);
            posx = posx + d.x;
            posy = posy + d.y;
          } else {
            const V2 d = corr * iMB;
            posx = posx + d.x;
            posy = posy + d.y;
          }
        }
      }
      bp.x = posx;
      bp.y = posy;

      /* body pass of pxUpdateSleepStates, px_world.c:179-219: the flags are TESTED by sorted
       * position p although they were set by slot -- the reference's behaviour */
      float slp = rSlp[r];
      const bool wake = staticSched ? (rWord[r] & 0x40000000u) != 0 : s.wakeFlag[p] != 0;
      const bool rest = staticSched ? (rWord[r] & 0x80000000u) != 0 : s.restFlag[p] != 0;
      if (wake) {
        slp = 0.0f;
        fl &= ~PX_FLAG_SLEEPING;
      } else {
        const bool below = (bv.x * bv.x + bv.y * bv.y <= linSleepSq) && (fabsf(bv.z) <= angSleep);
        if (!below) {
          slp = 0.0f;
          fl &= ~PX_FLAG_SLEEPING;
        } else if (!(fl & PX_FLAG_SLEEPING)) {
          slp = slp + dt;
          if (rest) slp = slp + dt;
          if (slp >= timeToSleep) {
            bv.x = 0.0f;
            bv.y = 0.0f;
            bv.z = 0.0f;
            rFx[r] = 0.0f;
            rFy[r] = 0.0f;
            rTq[r] = 0.0f;
            fl |= PX_FLAG_SLEEPING;
          }
        }
      }
      rSlp[r] = slp;
      /* pxClearForces, px_world.c:302-311 */
      if (!(slp >= timeToSleep)) {
        rFx[r] = 0.0f;
        rFy[r] = 0.0f;
        rTq[r] = 0.0f;
      }
      sleepingLocal += (fl & PX_FLAG_SLEEPING) ? 1u : 0u;
      /* positions are read by other bodies' correction terms?  No: corrections use only the
       * manifold and inverse masses, so the record can be written back right away */
      s.bodyP[x] = bp;
      s.bodyV[x] = bv;
      s.flags[x] = (uint8_t)fl;
    }
    if (sweepMode || step + 1 == P.kSubsteps) {
      if (sleepingLocal) atomicAdd(&s.misc[MI_SLEEPING], sleepingLocal);
      if (!sweepMode) {
        statPairs = nPairs;
        statManifolds = M;
      }
    }
    __syncthreads();
    PROF_MARK(5)
    } /* runFinish */
  }

  /* ------------------------------------------- store (smem records -> HBM SoA rows) */
  {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t x = tid + r * NT;
      if (x < D) {
        const float4 bp = s.bodyP[x], bv = s.bodyV[x], br = s.bodyR[x];
        gdyn[PX_DYN_PX * D + x] = bp.x;
        gdyn[PX_DYN_PY * D + x] = bp.y;
        gdyn[PX_DYN_VX * D + x] = bv.x;
        gdyn[PX_DYN_VY * D + x] = bv.y;
        gdyn[PX_DYN_AV * D + x] = bv.z;
        gdyn[PX_DYN_ANGLE * D + x] = rAngle[r];
        gdyn[PX_DYN_FX * D + x] = rFx[r];
        gdyn[PX_DYN_FY * D + x] = rFy[r];
        gdyn[PX_DYN_TORQUE * D + x] = rTq[r];
        gdyn[PX_DYN_SLEEP * D + x] = rSlp[r];
        gdyn[PX_DYN_M00 * D + x] = br.x;
        gdyn[PX_DYN_M01 * D + x] = br.y;
        gdyn[PX_DYN_M10 * D + x] = br.z;
        gdyn[PX_DYN_M11 * D + x] = br.w;
      }
    }
    copyRows<NT>((float *)(gmut + L.offDynFlags), (const float *)s.flags, D / 4);
    copyRows<NT>((float *)(gmut + L.offDynOrder), (const float *)s.order, D / 4);
    if (tid == 0) {
      PxWorldHeader *gh = (PxWorldHeader *)(gmut + L.offHeader);
      gh->statManifolds = statManifolds;
      gh->statPairs = statPairs;
      gh->statSleeping = (sweepMode && !P.doFinish) ? s.hdr->statSleeping : s.misc[MI_SLEEPING];
      gh->statOverflow = s.hdr->statOverflow | overflow | s.misc[MI_OVERFLOW];
      gh->substeps = s.hdr->substeps + (sweepMode ? (P.doFinish ? 1u : 0u) : P.kSubsteps);
    }
  }
  __syncthreads(); /* the next world's load overwrites what the store above still reads */
  } /* world loop */
}

/* ======================================================================= solver kernel ==== */
/*
 * Split pipeline, phase E: one CTA of two specialised warps (scheduler, math -- see pxStepKernel)
 * per world.  It needs a fraction of the sweep kernel's shared memory and registers
 * (bodies' position/mass and velocity rows, manifold ids, successor words, signals, ready queue),
 * so more than twice as many worlds are resident per SM while their dependency chains run, and
 * no throughput-oriented warp competes with them.  Reads the velocity rows the sweep kernel left in
 * the mutable block (after force integration), writes them back solved.
 */
#ifndef PX_SOLVE_MINB
#define PX_SOLVE_MINB 10
#endif
__global__ void __launch_bounds__(PX_SOLVE_THREADS, PX_SOLVE_MINB) pxSolveKernel(const PxStepParams P) {
  constexpr uint32_t NT = PX_SOLVE_THREADS;
  extern __shared__ __align__(16) uint8_t smemRaw[];
  const PxBatchLayout &L = P.L;
  const uint32_t D = L.maxDynamic, S = L.maxStatic, MM = L.maxManifolds;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  Sm s;
  bindSmem(s, smemRaw, P.solveS, nullptr);
  const uint32_t QMASK = P.solveS.qcap - 1;
  for (;;) {
    if (tid == 0) s.misc[MI_WORLD] = atomicAdd(P.worldCounter, 1u);
    __syncthreads();
    const uint32_t world = s.misc[MI_WORLD];
    if (world >= L.nWorlds) break;
    uint8_t *gmut = P.mut + (size_t)world * L.mutStride;
    const uint8_t *gimm = P.imm + (size_t)world * L.immStride;
    const uint8_t *hand = P.hand + (size_t)world * P.H.stride;
    float *gdyn = (float *)(gmut + L.offDyn);
    const PxWorldHeader *gh = (const PxWorldHeader *)(gmut + L.offHeader);
    const uint32_t iterations = gh->iterations & 255u;
    const uint32_t M = min(((const uint32_t *)(hand + P.H.scalars))[0], MM);
    const uint32_t qtail = min(((const uint32_t *)(hand + P.H.scalars))[1], P.solveS.qcap);
    if (M && iterations) { /* uniform: nothing to solve otherwise */
      const float *gc = (const float *)(gimm + L.offDynConst);
      for (uint32_t x = tid; x < D; x += NT) {
        s.bodyP[x] = make_float4(gdyn[PX_DYN_PX * D + x], gdyn[PX_DYN_PY * D + x], gc[PX_DYNC_IMASS * D + x],
                                 gc[PX_DYNC_IINERTIA * D + x]);
        s.bodyV[x] = make_float4(gdyn[PX_DYN_VX * D + x], gdyn[PX_DYN_VY * D + x], gdyn[PX_DYN_AV * D + x], 0.0f);
      }
      const float *gsm = (const float *)(gmut + L.offStatMut);
      for (uint32_t q = tid; q < S; q += NT) {
        s.bodyP[D + q] = make_float4(gsm[PX_STATM_PX * S + q], gsm[PX_STATM_PY * S + q], 0.0f, 0.0f);
        s.bodyV[D + q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
      handCopy<NT>(s.mids, hand + P.H.mids, M * 4u);
      handCopy<NT>(s.nextw, hand + P.H.nextw, M * 4u);
      handCopy<NT>(s.sig, hand + P.H.sig, M * 4u);
      handCopy<NT>(s.queue, hand + P.H.queue, qtail * 4u);
      if (tid < 32 && tid != MI_WORLD) s.misc[tid] = 0;
      __syncthreads();
      s.man = (float4 *)(const_cast<uint8_t *>(hand) + P.H.records);
      SolverCtx c;
      c.aQueue = sAddr(s.queue); c.aNext = sAddr(s.nextw); c.aMids = sAddr(s.mids); c.aSig = sAddr(s.sig);
      c.aDesc = sAddr(s.misc + MI_RDESC); c.aMathRound = sAddr(s.misc + MI_MATHROUND);
      c.aBodyP = sAddr(s.bodyP); c.aBodyV = sAddr(s.bodyV);
      c.qmask = QMASK; c.qcap = P.solveS.qcap; c.iterations = iterations; c.tail0 = qtail; c.lagSleepNs = P.lagSleepNs;
      c.man = s.man; c.eps = L.epsilon;
      long long t0 = 0;
      unsigned long long *prof = P.prof ? P.prof + (size_t)world * PX_PROF_SLOTS : nullptr;
      if (prof && tid == 0) t0 = clock64();
      if (warp == 0) {
        uint32_t visits = 0;
        const uint32_t rounds = solverSchedule(c, &visits);
        visits = __reduce_add_sync(0xffffffffu, visits);
        if (lane == 0) s.misc[MI_VISITS] = visits;
        if (prof && lane == 0) prof[6] += rounds;
      } else {
        solverMath(c);
      }
      __syncthreads();
      if (prof && tid == 0) prof[4] += (unsigned long long)(clock64() - t0);
      for (uint32_t x = tid; x < D; x += NT) {
        const float4 bv = s.bodyV[x];
        gdyn[PX_DYN_VX * D + x] = bv.x;
        gdyn[PX_DYN_VY * D + x] = bv.y;
        gdyn[PX_DYN_AV * D + x] = bv.z;
      }
      if (tid == 0 && s.misc[MI_VISITS] != M * iterations) {
        PxWorldHeader *wh = (PxWorldHeader *)(gmut + L.offHeader);
        wh->statOverflow |= PX_OVERFLOW_INTERNAL;
      }
    }
    __syncthreads(); /* the next world's load overwrites what the store above still reads */
  }
}

/* ===================================================================== static schedule ==== */
/*
 * Split pipeline with the static solver: a PERIODIC schedule of the substep's contact visits, derived once per
 * substep (here) and replayed for every iteration (pxReplayKernel) -- the dataflow solver re-derives the same order
 * `iterations` times.
 *
 * The visits of one body form a ring in pool order (its manifold list; a 2-contact manifold takes two rounds), and
 * the ring's closing edge carries the iteration boundary.  Give every manifold a start round t(m) inside "its"
 * iteration and let visit k of manifold m run in round t(m) + k * II.  That is a valid execution of the sequential
 * Gauss-Seidel loop (px_world.c:249-260) iff
 *     t(m) >= t(p) + contacts(p)                  for the predecessor p of m in each of its bodies' rings, and
 *     t(first) + II >= t(last) + contacts(last)   for every body (the closing edge, weight -II),
 * because two visits that share a dynamic body then never overlap and keep their pool order -- and any such
 * execution produces the sequential loop's bits.  The least such t is the longest-path solution of those difference
 * constraints.  Pool rank is a topological order of all but the closing edges, so ONE warp relaxes the manifolds in
 * rank order, 32 at a time (each block repeated until it is stable: its inner chains are a body's few consecutive
 * manifolds), and one pass settles everything the closing edges' current values imply; 3-5 passes reach the fixed
 * point.  It exists iff II is at least the critical cycle ratio of the rings' graph -- on settled piles the longest ring
 * plus 0..5 rounds (profiles/sched_sim.py).  A period that is too short shows after PX_SCHED_PASSES passes, which is
 * cheap, so the periods are tried from tight to loose: ring + 2, + 4, + 8, 2 * ring + 8, and finally the closing
 * edges are dropped and II = the depth of one iteration (iterations back to back: always valid).
 *
 * Visits are then sorted by (round mod II, round div II) with a counting sort: round R of the replay is slot
 * R mod II, and the visits of the iterations alive in it are one contiguous segment of that order (segment starts ->
 * H.seg).  The visits' frozen operands, which the sweep kernel left by pool rank, are permuted into that order
 * (H.vrec -> H.stream), so the replay reads them coalesced.  Work per world: about ten thousand warp instructions,
 * on one warp; twenty-odd worlds per SM hide its latency.
 */
#ifndef PX_REPLAY_MINB
#define PX_REPLAY_MINB 20
#endif
#ifndef PX_SCHED_PASSES
#define PX_SCHED_PASSES 7u
#endif
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
/* a global load of data that is read exactly once: evicted first from L2 */
__device__ __forceinline__ float4 ldgStream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

/* Schedule of one world, by one warp (see above).  Shared memory: tlev u16[MM], khist u16[PX_KSM + 2]; the ring words
 * are read from the hand-over record (coalesced, by rank).  Leaves the visit stream and the segment table in the
 * world's hand-over record. */
__device__ __forceinline__ void buildSchedule(const PxStepParams &P, uint8_t *smem, uint8_t *hand, const uint32_t M, const uint32_t ring,
                                              unsigned long long *prof, uint32_t &IIout, uint32_t &QNout, uint32_t &depthOut,
                                              uint32_t &visitsOut) {
  const uint32_t MM = P.L.maxManifolds, lane = threadIdx.x;
  uint16_t *tlev = (uint16_t *)(smem + P.solveS.tlev);
  uint32_t *h32 = (uint32_t *)(smem + P.solveS.khist);
  uint16_t *h16 = (uint16_t *)h32;
  const uint16_t *gA = (const uint16_t *)(hand + P.H.predA), *gB = (const uint16_t *)(hand + P.H.predB);
  const uint32_t aT = sAddr(tlev);
  long long t0 = (prof && lane == 0) ? clock64() : 0;
  uint32_t II = 1, passes = 0, depth = 0;
  bool converged = false, sequential = false;
  /* The margin that worked for this world last substep is tried first (a pile's critical cycle changes slowly):
   * a failed attempt costs a relaxation run to its pass cap, which the slowest world of a launch pays in full.
   * hint word (persists in the hand-over record): margin | first-attempt successes in a row << 8; after 64 of those
   * the margin is tightened by one round. */
  uint32_t *hintWord = (uint32_t *)(hand + P.H.scalars) + PX_HS_HINT;
  const uint32_t hint = *hintWord, margin0 = max(hint & 0xffu, PX_II_MARGIN);
  uint32_t marginUsed = margin0, attemptUsed = 0;
  for (uint32_t attempt = 0; attempt < 5 && !converged; attempt++) {
    marginUsed = margin0 + (attempt == 1 ? 2u : attempt == 2 ? 5u : 0u);
    attemptUsed = attempt;
    II = min(attempt < 3 ? ring + marginUsed : 2u * ring + 8u, 2u * MM + 8u);
    sequential = attempt == 4;
    const int closing = sequential ? 0x8000 : (int)II; /* dropped edges: a cost no start round can reach */
    for (uint32_t r = lane; r < M; r += 32u) tlev[r] = 0;
    __syncwarp();
    /* a block of 32 ranks is relaxed again only if a block holding one of its predecessors has moved since: lane b
     * keeps the set of blocks that block b reads (maxManifolds <= 1024: at most 32 blocks) */
    uint32_t myDeps = 0, movedPrev = 0xffffffffu;
    for (uint32_t pass = 0; pass < (sequential ? 3u : PX_SCHED_PASSES) && !converged; pass++) {
      uint32_t movedNow = 0;
      /* the ring words of the next block are on their way while this block is relaxed */
      uint32_t paNext = lane < M ? gA[lane] : 0u, pbNext = lane < M ? gB[lane] : 0xffffu;
      for (uint32_t base = 0; base < M; base += 32u) {
        const uint32_t r = base + lane, blk = base >> 5;
        const bool valid = r < M;
        /* rank | closes the ring << 10 | predecessor's contacts << 11 | own contacts << 13; 0xffff: static B */
        const uint32_t pa = paNext, pb = pbNext;
        paNext = r + 32u < M ? gA[r + 32u] : 0u;
        pbNext = r + 32u < M ? gB[r + 32u] : 0xffffu;
        if (pass && !(__shfl_sync(0xffffffffu, myDeps, blk) & (movedPrev | movedNow))) continue;
        if (!pass) {
          const uint32_t deps = __reduce_or_sync(0xffffffffu, valid ? (1u << ((pa & 0x3ffu) >> 5)) | (pb == 0xffffu ? 0u : 1u << ((pb & 0x3ffu) >> 5)) : 0u);
          if (lane == blk) myDeps = deps;
        }
        const uint32_t adA = aT + 2u * (pa & 0x3ffu), adB = aT + 2u * ((pb == 0xffffu ? pa : pb) & 0x3ffu);
        const int addA = (int)((pa >> 11) & 3u) - ((pa & 0x400u) ? closing : 0);
        const int addB = pb == 0xffffu ? -0x10000 : (int)((pb >> 11) & 3u) - ((pb & 0x400u) ? closing : 0);
        int cur = valid ? (int)lds16(aT + 2u * r) : 0x7fffffff;
        for (uint32_t it = 0; it < 34u; it++) { /* a block's inner chains: at most 31 hops unless the period is infeasible */
          const int v = max((int)lds16(adA) + addA, (int)lds16(adB) + addB);
          const bool up = v > cur;
          if (up) {
            cur = v;
            sts16(aT + 2u * r, (uint32_t)v);
          }
          __syncwarp();
          if (!__any_sync(0xffffffffu, up)) break;
          movedNow |= 1u << blk;
        }
      }
      passes++;
      converged = movedNow == 0;
      movedPrev = movedNow;
    }
    if (converged) {
      /* depth of one iteration; own contact count: bits 13-14 of the A word */
      depth = 0;
      for (uint32_t r = lane; r < M; r += 32u) {
        /* the manifold's contact count joins its start round (bits 14-15): the passes below read shared memory only */
        const uint32_t cnt = ((uint32_t)gA[r] >> 13) & 3u, t = tlev[r];
        depth = max(depth, t + cnt);
        tlev[r] = (uint16_t)(t | (cnt << 14));
      }
      depth = max(__reduce_max_sync(0xffffffffu, depth), 1u);
      if (sequential) II = depth;
      /* keys = II * phases <= depth - 1 + II must fit the table; the sequential schedule's (= depth <= 2 * MM) always do */
      if (II * ((depth - 1u) / II + 1u) > PX_KSM(MM)) converged = false;
    }
  }
  __syncwarp();
  if (lane == 0) {
    uint32_t m = margin0, age = hint >> 8;
    if (attemptUsed == 0) {
      if (++age >= 64u) {
        age = 0;
        if (m > PX_II_MARGIN) m--;
      }
    } else {
      m = min(attemptUsed < 3u ? marginUsed : margin0 + 6u, 24u);
      age = 0;
    }
    *hintWord = m | (age << 8);
  }
  const uint32_t QN = (depth - 1u) / II + 1u, nKeys = II * QN;
  if (prof && lane == 0) {
    const long long now = clock64();
    prof[15] += (unsigned long long)(now - t0);
    prof[13] += ((unsigned long long)passes << 32) + II;
  }
  /* visits per key, two 16-bit counters per word; every visit's key is kept for the scatter (0xffff: a manifold's
   * absent second contact) */
  uint16_t *keys = (uint16_t *)(smem + P.solveS.vkeys);
  const uint32_t magic = 0xffffffffu / II + 1u; /* tau / II == umulhi(tau, magic) for tau * II < 2^32 */
  for (uint32_t i = lane; i < (nKeys + 3u) / 2u; i += 32u) h32[i] = 0;
  __syncwarp();
#pragma unroll 2
  for (uint32_t base = 0; base < 2u * M; base += 32u) {
    const uint32_t v = base + lane, r = v >> 1, ci = v & 1u;
    uint32_t key = 0xffffu;
    const uint32_t tl = r < M ? tlev[r] : 0u;
    if (ci < (tl >> 14)) {
      const uint32_t tau = (tl & 0x3fffu) + ci, q = __umulhi(tau, magic);
      key = (tau - q * II) * QN + q;
      atomicAdd(&h32[key >> 1], (key & 1u) ? 0x10000u : 1u);
    }
    if (r < M) keys[v] = (uint16_t)key;
  }
  __syncwarp();
  /* exclusive prefix over the keys, in place: segment starts, copied out (entry nKeys = the number of visits),
   * then used as the cursors of the scatter below */
  uint16_t *seg = (uint16_t *)(hand + P.H.seg);
  uint32_t run = 0;
  for (uint32_t base = 0; base <= nKeys; base += 32u) {
    const uint32_t k = base + lane;
    const uint32_t n = (k < nKeys) ? h16[k] : 0u;
    uint32_t inc = n;
#pragma unroll
    for (uint32_t o = 1; o < 32u; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    const uint32_t ex = run + inc - n;
    __syncwarp();
    if (k <= nKeys) {
      h16[k] = (uint16_t)ex;
      seg[k] = (uint16_t)ex;
    }
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  /* permute the visit operands from pool-rank order into schedule order, two blocks of 32 in flight */
  const float4 *in0 = (const float4 *)(hand + P.H.vrec), *in1 = in0 + PX_PLANE, *in2 = in1 + PX_PLANE;
  float4 *out0 = (float4 *)(hand + P.H.stream), *out1 = out0 + PX_PLANE, *out2 = out1 + PX_PLANE;
  /* The operands come out of DRAM (the sweep kernel wrote them a launch ago): the loads of the next 64 visits are
   * issued before the current 64 are scattered, two register sets taking turns. */
  struct Vis {
    float4 a[2], b[2], c[2];
    uint32_t key[2];
  };
  auto fetch = [&](Vis &x, uint32_t base) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t v = base + 32u * u + lane;
      x.key[u] = v < 2u * M ? keys[v] : 0xffffu;
      if (x.key[u] != 0xffffu) {
        x.a[u] = ldgStream(in0 + v); /* read once: first out of L2, which the visit stream should keep */
        x.b[u] = ldgStream(in1 + v);
        x.c[u] = ldgStream(in2 + v);
      }
    }
  };
  auto scatter = [&](const Vis &x) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (x.key[u] != 0xffffu) {
        const uint32_t old = atomicAdd(&h32[x.key[u] >> 1], (x.key[u] & 1u) ? 0x10000u : 1u);
        const uint32_t pos = (x.key[u] & 1u) ? old >> 16 : old & 0xffffu;
        out0[pos] = x.a[u];
        out1[pos] = x.b[u];
        out2[pos] = x.c[u];
      }
    }
  };
  {
    Vis x, y;
    fetch(x, 0);
#pragma unroll 1
    for (uint32_t base = 0; base < 2u * M; base += 128u) {
      fetch(y, base + 64u);
      scatter(x);
      if (base + 64u >= 2u * M) break;
      fetch(x, base + 128u);
      scatter(y);
    }
  }
  if (prof && lane == 0) prof[14] += (unsigned long long)(clock64() - t0);
  IIout = II;
  QNout = QN;
  depthOut = depth;
  visitsOut = run;
  if (!converged && lane == 0) atomicOr((uint32_t *)(hand + P.H.scalars) + PX_HS_READY, 0x80000000u); /* unreachable */
}

/* ======================================================================= replay kernel ==== */
/*
 * Split pipeline, phase E with the static schedule: ONE warp per world first derives the schedule (buildSchedule)
 * and then replays it.  Round R is slot R mod II of the period; the visits alive in it -- phases q with
 * 0 <= R div II - q < iterations -- are one contiguous segment of the visit stream, processed 32 at a time, one
 * contact per lane.  The stream (three coalesced float4 planes in the hand-over record) holds everything frozen
 * during the solve, so a visit costs three coalesced global loads, two velocity gathers and two mass gathers from
 * shared memory and two predicated velocity stores; there is no scheduler, no atomics and no second warp.  Shared
 * memory per world is the schedule tables, then -- over them -- the velocity and mass rows (10 KB at 736 manifolds),
 * so residency is bound by registers and the dependent chains of twenty-odd worlds per SM interleave.  The next
 * chunk's operands are in flight while the current one computes.
 */
__global__ void __launch_bounds__(PX_REPLAY_THREADS, PX_REPLAY_MINB) pxReplayKernel(const PxStepParams P) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  const PxBatchLayout &L = P.L;
  const uint32_t D = L.maxDynamic, S = L.maxStatic, MM = L.maxManifolds;
  const uint32_t lane = threadIdx.x;
  float4 *bodyV = (float4 *)(smemRaw + P.solveS.bodyV);
  float *bodyI = (float *)(smemRaw + P.solveS.bodyI); /* inverse inertia; the inverse mass rides in bodyV.w */
  uint16_t *segS = (uint16_t *)(smemRaw + P.solveS.segS); /* PX_SEG_SMEM entries behind the body rows */
  const uint32_t aBodyV = sAddr(bodyV), aBodyI = sAddr(bodyI);
  const float eps = L.epsilon;
  for (;;) {
    uint32_t world = 0;
    if (lane == 0) world = atomicAdd(P.worldCounter, 1u);
    world = __shfl_sync(0xffffffffu, world, 0);
    if (world >= L.nWorlds) break;
    uint8_t *gmut = P.mut + (size_t)world * L.mutStride;
    const uint8_t *gimm = P.imm + (size_t)world * L.immStride;
    uint8_t *hand = P.hand + (size_t)world * P.H.stride;
    float *gdyn = (float *)(gmut + L.offDyn);
    PxWorldHeader *gh = (PxWorldHeader *)(gmut + L.offHeader);
    const uint32_t iterations = gh->iterations & 255u;
    uint32_t *sc = (uint32_t *)(hand + P.H.scalars);
    const uint32_t M = min(sc[PX_HS_M], MM), ring = sc[PX_HS_RING];
    const bool solving = M && iterations; /* uniform */
    unsigned long long *prof = P.prof ? P.prof + (size_t)world * PX_PROF_SLOTS : nullptr;
    /* what is read late is started on its way out of DRAM now: the visit operands by rank (after the relaxation) */
    if (solving)
      for (uint32_t k = 0; k < 3u; k++) prefetchRegion(hand + P.H.vrec + k * PX_PLANE * 16u, 2u * M * 16u, lane, PX_REPLAY_THREADS);
    uint32_t II = 0, QN = 0, depth = 0, nVisits = 0;
    if (solving) {
      buildSchedule(P, smemRaw, hand, M, ring, prof, II, QN, depth, nVisits);
      __syncwarp(); /* the schedule tables are dead: the body rows take their place; the stream is in global memory */
      /* the segment table is read at the head of every round, on the warp's critical path: keep it in shared memory */
      if (II * QN + 1u <= PX_SEG_SMEM) {
        const uint16_t *gseg = (const uint16_t *)(hand + P.H.seg);
        for (uint32_t k = lane; k <= II * QN; k += PX_REPLAY_THREADS) segS[k] = gseg[k];
      }
    }
    if (solving) {
      const float *gc = (const float *)(gimm + L.offDynConst);
      for (uint32_t x = lane; x < D; x += PX_REPLAY_THREADS) {
        bodyV[x] = make_float4(gdyn[PX_DYN_VX * D + x], gdyn[PX_DYN_VY * D + x], gdyn[PX_DYN_AV * D + x], gc[PX_DYNC_IMASS * D + x]);
        bodyI[x] = gc[PX_DYNC_IINERTIA * D + x];
      }
      for (uint32_t q = lane; q < S; q += PX_REPLAY_THREADS) {
        bodyV[D + q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        bodyI[D + q] = 0.0f;
      }
      __syncwarp();
    }
    long long t0 = 0;
    if (prof && lane == 0) t0 = clock64();
    /* phase E; `segInSmem` (a std::bool_constant) says where the segment table is: the copy in shared memory, or -- a
     * table too long for it -- the one in the hand-over record.  Two instantiations, so that the common one reads the
     * table with 32-bit shared-memory addresses instead of generic loads. */
    auto replay = [&](auto segInSmem) {
      constexpr bool SEG_SMEM = decltype(segInSmem)::value;

      /* Everything below is one warp's instruction stream and the kernel is issue-bound: the chunk iterator keeps
       * its state in registers (row = slot * QN, the alive phase window [qlo, qhi1) changes only when the period
       * wraps), the three operand planes are PX_PLANE entries apart so that one address serves all three loads. */
      const uint16_t *gseg = (const uint16_t *)(hand + P.H.seg);
      const float4 *pl = (const float4 *)(hand + P.H.stream);
      asm volatile("" : "+l"(gseg), "+l"(pl)); /* opaque: kept in registers, not rebuilt from the parameters every chunk */
      uint32_t aSeg = sAddr(segS);
      asm volatile("" : "+r"(aSeg));
      auto segAt = [&](uint32_t i) -> uint32_t {
        if constexpr (SEG_SMEM) {
          uint32_t v;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(aSeg + 2u * i));
          return v;
        } else {
          return gseg[i];
        }
      };
      const uint32_t nR = depth + (iterations - 1u) * II;
      uint32_t R = 0, slot = 0, row = 0, Q = 0, qlo = 0, qhi1 = 1, beg = 0, end = 0;
      beg = segAt(0);
      end = segAt(1);
      /* a chunk is [base, min(base + 32, lim)) of the stream; lim == 0: the schedule is exhausted.  The alive window
       * slides by increments when the period wraps (Q -> Q + 1: qlo = max(0, Q + 1 - iterations) and
       * qhi1 = min(Q, QN - 1) + 1): no min / max on the iterator's path, which runs on the uniform datapath. */
      auto nextChunk = [&](uint32_t &base, uint32_t &lim) {
        while (beg >= end) {
          if (++R >= nR) {
            base = lim = 0;
            return;
          }
          row += QN;
          if (++slot == II) {
            slot = row = 0;
            if (Q + 1u >= iterations) qlo++;
            if (Q + 1u < QN) qhi1++;
            Q++;
          }
          if (qlo < qhi1) {
            beg = segAt(row + qlo);
            end = segAt(row + qhi1);
          }
        }
        base = beg;
        lim = end;
        beg += 32u;
      };
      uint32_t chunks = 0, visits = 0; /* visits: per lane, summed at the end */
      /* one chunk: start the NEXT chunk's three operand loads, then run this one.  Two register sets take turns
       * (no copies: a copy at the end of the step would wait for the loads it had just issued). */
      auto step = [&](const float4 &r0, const float4 &r1, const float4 &r2, const uint32_t base, const uint32_t lim, float4 &n0,
                      float4 &n1, float4 &n2, uint32_t &baseNext, uint32_t &limNext) {
        nextChunk(baseNext, limNext);
        /* idle lanes repeat the chunk's last visit and store nothing; past the end (limNext == 0): entry `lane`, discarded */
        const float4 *at = pl + min(baseNext + lane, limNext - 1u);
        n0 = ldgPinned(at);
        n1 = ldgPinned(at + PX_PLANE);
        n2 = ldgPinned(at + 2u * PX_PLANE);
        const uint32_t ids = __float_as_uint(r2.z);
        const uint32_t a16 = (ids & 0xffu) * 16u, b16 = ((ids >> 8) & 0x1ffu) * 16u;
        const bool bdyn = ((ids >> 17) & 1u) != 0, active = base + lane < lim;
        float4 vA = lds128(aBodyV + a16), vB = lds128(aBodyV + b16);
        /* A's inverse inertia fills the record's twelfth word: one gather (three or four shared-memory wavefronts) less */
        const float iIB = __uint_as_float(lds32(aBodyI + (b16 >> 2)));
        contactImpulseCore(v2(r0.x, r0.y), v2(r0.z, r0.w), v2(r1.x, r1.y), r1.w, r2.x, r2.y, r1.z, vA.w, r2.w, vB.w, iIB, bdyn, vA, vB,
                           eps);
        sts128If(aBodyV + a16, vA, active);
        sts128If(aBodyV + b16, vB, active && bdyn);
        __syncwarp();
        visits += active ? 1u : 0u;
        chunks++;
      };
      uint32_t bA, lA, bB = 0, lB = 0;
      nextChunk(bA, lA);
      float4 a0, a1, a2, b0, b1, b2;
      {
        const float4 *at = pl + min(bA + lane, lA - 1u);
        a0 = ldgPinned(at);
        a1 = ldgPinned(at + PX_PLANE);
        a2 = ldgPinned(at + 2u * PX_PLANE);
      }
#pragma unroll 1
      while (lA) {
        step(a0, a1, a2, bA, lA, b0, b1, b2, bB, lB);
        if (!lB) break;
        step(b0, b1, b2, bB, lB, a0, a1, a2, bA, lA);
      }
      visits = __reduce_add_sync(0xffffffffu, visits);
      if (prof && lane == 0) {
        const long long now = clock64();
        prof[4] += (unsigned long long)(now - t0);
        prof[6] += chunks;
        t0 = now;
      }
      if (lane == 0 && (visits != nVisits * iterations || (sc[PX_HS_READY] & 0x80000000u))) gh->statOverflow |= PX_OVERFLOW_INTERNAL;
    };
    if (solving) {
      if (II * QN + 1u <= PX_SEG_SMEM) replay(std::true_type{});
      else replay(std::false_type{});
    }

    if (solving) {
      for (uint32_t x = lane; x < D; x += PX_REPLAY_THREADS) {
        const float4 bv = bodyV[x];
        gdyn[PX_DYN_VX * D + x] = bv.x;
        gdyn[PX_DYN_VY * D + x] = bv.y;
        gdyn[PX_DYN_AV * D + x] = bv.z;
      }
    }
    __syncwarp(); /* the next world's tables overwrite what the store above still reads */
  }
}

/* ======================================================================= finish kernel ==== */
/*
 * Split pipeline with the static solver, phase F of the LAST substep of a step (the others are finished at the head of
 * the next sweep launch): one thread per body, one CTA per world.  Integrate velocities (px_body.c:181-188), positional
 * correction (px_world.c:288-292: the body's moves, precomputed by the sweep kernel, added in pool order), sleep update
 * (px_world.c:179-219, with the flags of the body's sorted position), clear forces (px_world.c:302-311) -- the same
 * expressions as phase F of pxStepKernel.  Pure streaming: every row is read and written once, coalesced.
 */
__global__ void __launch_bounds__(PX_DMAX) pxFinishKernel(const PxStepParams P) {
  const PxBatchLayout &L = P.L;
  const uint32_t D = L.maxDynamic, x = threadIdx.x, world = blockIdx.x;
  uint8_t *gmut = P.mut + (size_t)world * L.mutStride;
  const uint8_t *hand = P.hand + (size_t)world * P.H.stride;
  PxWorldHeader *gh = (PxWorldHeader *)(gmut + L.offHeader);
  float *gdyn = (float *)(gmut + L.offDyn);
  uint8_t *gflags = gmut + L.offDynFlags;
  const float dt = gh->dt, linSleepSq = gh->linearSleepThresholdSq, angSleep = gh->angularSleepThreshold;
  const float timeToSleep = gh->timeToSleep;
  uint32_t asleep = 0;
  uint32_t fl = x < D ? gflags[x] : 0u;
  if (fl & 4u) {
    const uint32_t word = ((const uint32_t *)(hand + P.H.bodyF))[x];
    const float2 *cvec = (const float2 *)(hand + P.H.cvec);
    float posx = gdyn[PX_DYN_PX * D + x], posy = gdyn[PX_DYN_PY * D + x], angle = gdyn[PX_DYN_ANGLE * D + x];
    float fx = gdyn[PX_DYN_FX * D + x], fy = gdyn[PX_DYN_FY * D + x], tq = gdyn[PX_DYN_TORQUE * D + x];
    float slp = gdyn[PX_DYN_SLEEP * D + x];
    float bvx = gdyn[PX_DYN_VX * D + x], bvy = gdyn[PX_DYN_VY * D + x], bvz = gdyn[PX_DYN_AV * D + x];
    if (!(fl & PX_FLAG_SLEEPING)) {
      posx = posx + bvx * dt;
      posy = posy + bvy * dt;
      float radians = angle + bvz * dt;
      radians = fmodf(radians, D_2PI);
      if (radians < 0) radians += D_2PI;
      angle = radians;
      const float c = fSin(P.sinTable, radians + D_PI * 0.5f), sn = fSin(P.sinTable, radians);
      gdyn[PX_DYN_M00 * D + x] = c;
      gdyn[PX_DYN_M01 * D + x] = -sn;
      gdyn[PX_DYN_M10 * D + x] = sn;
      gdyn[PX_DYN_M11 * D + x] = c;
    }
    {
      const uint32_t start = word & 0xffffu, n = (word >> 16) & 0x3fffu;
      for (uint32_t i = 0; i < n; i++) {
        const float2 d = cvec[start + i];
        posx = posx + d.x;
        posy = posy + d.y;
      }
    }
    if (word & 0x40000000u) { /* wake flag of the body's sorted position */
      slp = 0.0f;
      fl &= ~PX_FLAG_SLEEPING;
    } else {
      const bool below = (bvx * bvx + bvy * bvy <= linSleepSq) && (fabsf(bvz) <= angSleep);
      if (!below) {
        slp = 0.0f;
        fl &= ~PX_FLAG_SLEEPING;
      } else if (!(fl & PX_FLAG_SLEEPING)) {
        slp = slp + dt;
        if (word & 0x80000000u) slp = slp + dt; /* rest flag of the body's sorted position */
        if (slp >= timeToSleep) {
          bvx = 0.0f;
          bvy = 0.0f;
          bvz = 0.0f;
          fx = 0.0f;
          fy = 0.0f;
          tq = 0.0f;
          fl |= PX_FLAG_SLEEPING;
        }
      }
    }
    if (!(slp >= timeToSleep)) {
      fx = 0.0f;
      fy = 0.0f;
      tq = 0.0f;
    }
    asleep = (fl & PX_FLAG_SLEEPING) ? 1u : 0u;
    gdyn[PX_DYN_PX * D + x] = posx;
    gdyn[PX_DYN_PY * D + x] = posy;
    gdyn[PX_DYN_VX * D + x] = bvx;
    gdyn[PX_DYN_VY * D + x] = bvy;
    gdyn[PX_DYN_AV * D + x] = bvz;
    gdyn[PX_DYN_ANGLE * D + x] = angle;
    gdyn[PX_DYN_FX * D + x] = fx;
    gdyn[PX_DYN_FY * D + x] = fy;
    gdyn[PX_DYN_TORQUE * D + x] = tq;
    gdyn[PX_DYN_SLEEP * D + x] = slp;
    gflags[x] = (uint8_t)fl;
  }
  const uint32_t total = __syncthreads_count(asleep != 0u);
  if (x == 0) {
    gh->statSleeping = total;
    gh->substeps = gh->substeps + 1u;
  }
}

extern "C" int pxLaunchFinish(const PxStepParams *P, void *stream) {
  if (P->L.nWorlds == 0) return (int)cudaErrorInvalidValue;
  pxFinishKernel<<<P->L.nWorlds, PX_DMAX, 0, (cudaStream_t)stream>>>(*P);
  return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------- control commands */
/* One thread per world applies that world's queued API calls in submission order, merged
 * with the batch-wide (PX_ALL_WORLDS) calls by sequence number. */
__global__ void pxCommandKernel(const PxBatchLayout L, uint8_t *mut, const uint8_t *imm, const PxCommand *cmds,
                                const uint32_t *worldStart, const PxCommand *global, uint32_t nGlobal,
                                const PxBodyPayload *payloads, const float *sinTable) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= L.nWorlds) return;
  uint8_t *gmut = mut + (size_t)w * L.mutStride;
  const uint8_t *gimm = imm + (size_t)w * L.immStride;
  PxWorldHeader *h = (PxWorldHeader *)(gmut + L.offHeader);
  float *dyn = (float *)(gmut + L.offDyn);
  float *statm = (float *)(gmut + L.offStatMut);
  uint8_t *flags = gmut + L.offDynFlags;
  uint8_t *order = gmut + L.offDynOrder;
  const float *dc = (const float *)(gimm + L.offDynConst);
  const uint8_t *statSlot = gimm + L.offStatSlot;
  const uint32_t D = L.maxDynamic, S = L.maxStatic;
  uint32_t i = worldStart ? worldStart[w] : 0, iEnd = worldStart ? worldStart[w + 1] : 0, gi = 0;
  while (i < iEnd || gi < nGlobal) {
    const PxCommand *c;
    if (gi < nGlobal && (i >= iEnd || global[gi].seq < cmds[i].seq)) {
      c = &global[gi++];
    } else {
      c = &cmds[i++];
    }
    const uint32_t op = c->op & 0xffu, body = (c->op >> 8) & 0xffu;
    const bool isStatic = (c->op & PX_CMD_STATIC_BIT) != 0;
    if (op == PX_CMD_GRAVITY) { /* pxWorldSetGravity, px_world.c:57-63 */
      h->gravityX = c->a * L.unit;
      h->gravityY = c->b * L.unit;
    } else if (op == PX_CMD_ITERATIONS) {
      h->iterations = (uint32_t)c->a & 255u;
    } else if (op >= PX_CMD_SET_POSITION && op <= PX_CMD_ROTATE) {
      /* pxBodySetPosition / MoveBy / SetOrientation / Rotate, px_body.c:85-124: plain pose edits
       * of a dynamic OR static body; none of them wakes the body, the AABB follows the position
       * (it is derived on the device) */
      float *px, *py, *angle, *m00, *m01, *m10, *m11;
      if (!isStatic) {
        if (body >= D) continue;
        px = dyn + PX_DYN_PX * D + body; py = dyn + PX_DYN_PY * D + body; angle = dyn + PX_DYN_ANGLE * D + body;
        m00 = dyn + PX_DYN_M00 * D + body; m01 = dyn + PX_DYN_M01 * D + body;
        m10 = dyn + PX_DYN_M10 * D + body; m11 = dyn + PX_DYN_M11 * D + body;
      } else {
        /* statics are stored by position in the static iteration order */
        uint32_t q = 0;
        const uint32_t nS = min(h->nStatic, S);
        while (q < nS && statSlot[q] != body) q++;
        if (q >= nS) continue;
        px = statm + PX_STATM_PX * S + q; py = statm + PX_STATM_PY * S + q; angle = statm + PX_STATM_ANGLE * S + q;
        m00 = statm + PX_STATM_M00 * S + q; m01 = statm + PX_STATM_M01 * S + q;
        m10 = statm + PX_STATM_M10 * S + q; m11 = statm + PX_STATM_M11 * S + q;
      }
      if (op == PX_CMD_SET_POSITION) {
        *px = c->a;
        *py = c->b;
      } else if (op == PX_CMD_MOVE_BY) { /* pxVec2AddAssign */
        *px = *px + c->a;
        *py = *py + c->b;
      } else {
        float radians = op == PX_CMD_ROTATE ? *angle + c->a : c->a;
        radians = fmodf(radians, D_2PI); /* px_body.c:110-112 */
        if (radians < 0) radians += D_2PI;
        *angle = radians;
        const float cs = fSin(sinTable, radians + D_PI * 0.5f), sn = fSin(sinTable, radians); /* pxMat2Orientation */
        *m00 = cs;
        *m01 = -sn;
        *m10 = sn;
        *m11 = cs;
      }
    } else if (op == PX_CMD_ADD_DYNAMIC) {
      /* device half of pxBodyArrayAdd, px_body_array.c:7-59: the host chose the slot by the
       * reference's rule (freed stack first, else `length`); here the body's mutable rows are
       * written and the slot is appended to the iteration order */
      const uint32_t n = h->nDynamic;
      if (body >= D || n >= D) continue;
      const PxBodyPayload *pl = payloads + (uint32_t)c->a;
      for (uint32_t r = 0; r < PX_DYN_ROWS; r++) dyn[r * D + body] = pl->dyn[r];
      flags[body] = (uint8_t)pl->flags;
      order[n] = (uint8_t)body;
      h->nDynamic = n + 1;
    } else if (op == PX_CMD_FREE_DYNAMIC) {
      /* pxBodyArrayRemove, px_body_array.c:61-79: a no-op when the SLOT index is >= the body
       * COUNT (the reference's own check), else swap-remove from the iteration order */
      const uint32_t n = h->nDynamic;
      if (body >= n) continue;
      for (uint32_t k = 0; k < n; k++) {
        if (order[k] == body) {
          order[k] = order[n - 1];
          h->nDynamic = n - 1;
          flags[body] = 0; /* device-side only: the slot no longer holds a valid body for the per-slot phases */
          break;
        }
      }
    } else if (body < D && !isStatic) {
      /* pxBodyApiApplyImpulse / pxBodyApiApplyForce, px_body_api.c:20-42: scale by the unit,
       * accumulate, wake */
      const float sx = c->a * L.unit, sy = c->b * L.unit;
      if (op == PX_CMD_IMPULSE) {
        dyn[PX_DYN_VX * D + body] += sx * dc[PX_DYNC_IMASS * D + body];
        dyn[PX_DYN_VY * D + body] += sy * dc[PX_DYNC_IMASS * D + body];
        dyn[PX_DYN_AV * D + body] += (c->c * sy - c->d * sx) * dc[PX_DYNC_IINERTIA * D + body];
      } else if (op == PX_CMD_FORCE) {
        dyn[PX_DYN_FX * D + body] += sx;
        dyn[PX_DYN_FY * D + body] += sy;
        dyn[PX_DYN_TORQUE * D + body] += c->c * sy - c->d * sx;
      }
      dyn[PX_DYN_SLEEP * D + body] = 0.0f; /* pxBodyWakeUp, px_body.c:195-202 */
      flags[body] &= (uint8_t)~PX_FLAG_SLEEPING;
    }
  }
}

/* ------------------------------------------------------------------- arithmetic self-test */
/* T0 parity tier: the device's restatement of the reference's inline math (px_math.h:50-114,
 * 226-252) evaluated element-wise, and an exhaustive sweep of pxFastRsqrt over all 2^32 bit
 * patterns reduced to one 64-bit sum per 2^20-pattern block. */
__global__ void pxMathKernel(int op, const float *a, const float *b, const float *sinTable, float *out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = a[i];
  float y = 0.0f;
  switch (op) {
  case 0: y = fRsqrt(x); break;
  case 1: y = fRcp(x); break;
  case 2: y = fSqrt(x); break;
  case 3: y = fDiv(x, b[i]); break;
  case 4: y = fSin(sinTable, x); break;
  case 5: y = fSin(sinTable, x + D_PI * 0.5f); break;
  }
  out[i] = y;
}

__global__ void pxRsqrtSweepKernel(unsigned long long *sums /*[4096]*/) {
  /* block b covers bit patterns [b << 20, (b + 1) << 20) */
  const uint32_t base = blockIdx.x << 20;
  unsigned long long acc = 0;
  for (uint32_t k = threadIdx.x; k < (1u << 20); k += blockDim.x) {
    /* NaN payloads are outside the contract (x86 keeps the operand's payload, the GPU returns
     * the canonical NaN); NaNs only arise from negative or NaN inputs, which the step never
     * produces.  NaN-ness itself must agree. */
    const float y = fRsqrt(__uint_as_float(base + k));
    acc += (y != y) ? 0x7fc00000u : __float_as_uint(y);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; w++) t += part[w];
    sums[blockIdx.x] = t;
  }
}

extern "C" int pxLaunchMath(int op, const float *a, const float *b, const float *sinTable, float *out, uint32_t n, void *stream) {
  pxMathKernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, a, b, sinTable, out, n);
  return (int)cudaGetLastError();
}

extern "C" int pxLaunchRsqrtSweep(unsigned long long *sums, void *stream) {
  pxRsqrtSweepKernel<<<4096, 256, 0, (cudaStream_t)stream>>>(sums);
  return (int)cudaGetLastError();
}

/* ---------------------------------------------------------------------- host launchers */
static uint32_t alignUp(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

extern "C" void pxPlanSmem(const PxBatchLayout *L, int vertsInSmem, int staticSolver, PxSmemPlan *S) {
  const uint32_t D = L->maxDynamic, St = L->maxStatic, NB = D + St, V = L->maxVertices;
  const uint32_t MM = L->maxManifolds, MP = L->maxPairs;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t at = o;
    o += alignUp(bytes, 16);
    return at;
  };
  S->bodyP = take(NB * 16); S->bodyV = take(NB * 16); S->bodyR = take(NB * 16); S->shape = take(NB * 4);
  S->flags = take(D); S->order = take(D); S->order2 = take(D); S->posOf = take(D);
  S->restFlag = take(D); S->wakeFlag = take(D); S->sSlot = take(D);
  S->hdr = take(sizeof(PxWorldHeader));
  S->misc = take(32 * 4);
  S->scan = take(8 * 8);
  S->mids = take(MM * 4);
  S->listStart = take((D + 1) * 2);
  S->cnt = take(D * 2); S->cntB = take(D * 2);
  S->validSlot = take(MP * 2);
  S->base = take((D + 1) * 2);
  S->fillB = take(D * 4);
  uint32_t qcap = 32;
  while (qcap < MM + 32 * (PX_LAG + 1)) qcap <<= 1; /* ready visits + published, unexecuted rounds */
  S->qcap = qcap;
  /* region X (sweep + narrowphase) overlaid by region Y (solver) */
  const uint32_t rx = o;
  S->sBox = take(D * 16 > MP * 2 ? D * 16 : MP * 2); /* AABBs by sorted position; after the sweep, u16[MP] SAT survivors */
  S->sInfo = take(D * 4); S->stBox = take(St * 16);
  S->bins = take((MP > D ? MP : D) * 4);
  S->vertsInSmem = vertsInSmem ? 1u : 0u;
  S->verts = vertsInSmem ? take(V * 16) : 0;
  const uint32_t rxEnd = o;
  o = rx;
  S->list = take(2 * MM * 2);
  const uint32_t ry = o;
  S->nextw = take(2 * MM * 2); S->sig = take(MM * 4); S->queue = take((qcap + 4) * 4); /* + the dummy sink at [qcap] */
  const uint32_t ryEnd = o;
  S->cm = 0;
  S->predA = S->predB = S->rankOf = S->tlev = S->khist = S->vkeys = S->bodyI = S->segS = 0;
  if (staticSolver) { /* region Y': the rings in pool-rank space take the place of the dynamic solver's tables */
    o = ry;
    S->predA = take(MM * 2); S->predB = take(MM * 2); S->rankOf = take(MM * 2);
    S->cm = take(MM * 16); /* written in phase C4, when bins / survivors / vertex pool under it are dead */
  }
  o = o > ryEnd ? o : ryEnd;
  o = o > rxEnd ? o : rxEnd;
  S->total = o;
}

/* The kernel may use any dynamic shared-memory size up to the device's opt-in maximum: the
 * attribute is raised to that maximum once per device and never lowered, so batches of different
 * sizes (and devices) cannot invalidate each other's launches. */
template <int NT, int MODE>
static cudaError_t configureT() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  int optin = 0;
  err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(pxStepKernel<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}

template <int NT, int MODE>
static cudaError_t launchTM(const PxStepParams *P, uint32_t grid, cudaStream_t stream) {
  cudaError_t err = configureT<NT, MODE>();
  if (err != cudaSuccess) return err;
  pxStepKernel<NT, MODE><<<grid, NT, P->S.total, stream>>>(*P);
  return cudaGetLastError();
}
template <int NT>
static cudaError_t launchT(const PxStepParams *P, uint32_t grid, cudaStream_t stream) {
  return P->mode == PX_MODE_SWEEP ? launchTM<NT, PX_MODE_SWEEP>(P, grid, stream) : launchTM<NT, PX_MODE_FUSED>(P, grid, stream);
}

extern "C" int pxLaunchStep(const PxStepParams *P, uint32_t threads, uint32_t grid, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (grid == 0) return (int)cudaErrorInvalidValue;
  switch (threads) {
  case 128: return (int)launchT<128>(P, grid, st);
  case 256: return (int)launchT<256>(P, grid, st);
  default: return (int)cudaErrorInvalidValue;
  }
}

template <int NT>
static cudaError_t occupancyT(const PxStepParams *P, int *ctasPerSm) {
  if (P->mode == PX_MODE_SWEEP) {
    cudaError_t err = configureT<NT, PX_MODE_SWEEP>();
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, pxStepKernel<NT, PX_MODE_SWEEP>, NT, P->S.total);
  }
  cudaError_t err = configureT<NT, PX_MODE_FUSED>();
  if (err != cudaSuccess) return err;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, pxStepKernel<NT, PX_MODE_FUSED>, NT, P->S.total);
}

extern "C" int pxStepOccupancy(const PxStepParams *P, uint32_t threads, int *ctasPerSm) {
  switch (threads) {
  case 128: return (int)occupancyT<128>(P, ctasPerSm);
  case 256: return (int)occupancyT<256>(P, ctasPerSm);
  default: return (int)cudaErrorInvalidValue;
  }
}

extern "C" int pxLaunchCommands(const PxBatchLayout *L, uint8_t *mut, const uint8_t *imm, const PxCommand *cmds,
                                const uint32_t *worldStart, const PxCommand *global, uint32_t nGlobal,
                                const PxBodyPayload *payloads, const float *sinTable, void *stream) {
  const uint32_t threads = 128, blocks = (L->nWorlds + threads - 1) / threads;
  pxCommandKernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*L, mut, imm, cmds, worldStart, global, nGlobal, payloads, sinTable);
  return (int)cudaGetLastError();
}

/* ---------------------------------------------------------------- split pipeline, host side */
extern "C" void pxPlanHandover(const PxBatchLayout *L, const PxSmemPlan *S, PxHandLayout *H) {
  const uint32_t D = L->maxDynamic, MM = L->maxManifolds;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t at = o;
    o += alignUp(bytes, 16) + 16; /* + one granule: copies are rounded up to 16 bytes */
    return at;
  };
  H->mids = take(MM * 4); H->nextw = take(2 * MM * 2); H->sig = take(MM * 4); H->queue = take((S->qcap + 4) * 4);
  H->list = take(2 * MM * 2); H->listStart = take((D + 1) * 2); H->cnt = take(D * 2); H->cntB = take(D * 2);
  H->posOf = take(D); H->restFlag = take(D); H->wakeFlag = take(D);
  H->scalars = take(32);
  H->records = S->rankOf ? 0 : take(MM * 3 * 16); /* static solver: the records never leave the sweep launch (scratch slot) */
  H->stream = H->seg = 0;
  H->predA = H->predB = H->vrec = H->bodyF = H->cvec = 0;
  if (S->rankOf) { /* static schedule */
    H->stream = take(3 * PX_PLANE * 16);
    H->seg = take((PX_KCAP(MM) + 2) * 2);
    H->predA = take(MM * 2); H->predB = take(MM * 2);
    H->vrec = take(3 * PX_PLANE * 16);
    H->bodyF = take(D * 4);
    H->cvec = take(2 * MM * 8);
  }
  H->stride = alignUp(o, 128);
}

extern "C" void pxPlanSolveSmem(const PxBatchLayout *L, const PxSmemPlan *S, int staticSolver, PxSmemPlan *out) {
  const uint32_t NB = L->maxDynamic + L->maxStatic, MM = L->maxManifolds;
  memset(out, 0, sizeof(*out));
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t at = o;
    o += alignUp(bytes, 16);
    return at;
  };
  if (staticSolver) { /* pxReplayKernel: the schedule tables, then -- in their place -- the velocity and mass rows */
    out->bodyV = take(NB * 16); out->bodyI = take(NB * 4); out->segS = take(PX_SEG_SMEM * 2);
    const uint32_t rows = o;
    o = 0;
    out->tlev = take(MM * 2); out->khist = take((PX_KSM(MM) + 2) * 2); out->vkeys = take(PX_VCAP(MM) * 2);
    out->qcap = S->qcap;
    out->total = o > rows ? o : rows;
    return;
  }
  out->bodyP = take(NB * 16); out->bodyV = take(NB * 16);
  out->hdr = take(sizeof(PxWorldHeader)); out->misc = take(32 * 4);
  out->mids = take(MM * 4 + 16); out->nextw = take(2 * MM * 2 + 16); out->sig = take(MM * 4 + 16);
  out->qcap = S->qcap;
  out->queue = take((S->qcap + 4) * 4 + 16);
  out->total = o;
}

static cudaError_t configureSolve() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  int optin = 0;
  err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(pxSolveKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}

static cudaError_t configureReplay() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  int optin = 0;
  err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(pxReplayKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  if (err != cudaSuccess) return err;
  /* many small CTAs: ask for the largest shared-memory carve-out, or the default split caps residency */
  static const char *cv = getenv("PX_REPLAY_CARVEOUT"); /* tuning switch: percent of the SM's L1/shared memory asked for as shared */
  err = cudaFuncSetAttribute(pxReplayKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) return err;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}

extern "C" int pxLaunchSolve(const PxStepParams *P, uint32_t grid, void *stream) {
  if (grid == 0) return (int)cudaErrorInvalidValue;
  if (P->solver == PX_SOLVER_STATIC) {
    cudaError_t err = configureReplay();
    if (err != cudaSuccess) return (int)err;
    pxReplayKernel<<<grid, PX_REPLAY_THREADS, P->solveS.total, (cudaStream_t)stream>>>(*P);
    return (int)cudaGetLastError();
  }
  cudaError_t err = configureSolve();
  if (err != cudaSuccess) return (int)err;
  pxSolveKernel<<<grid, PX_SOLVE_THREADS, P->solveS.total, (cudaStream_t)stream>>>(*P);
  return (int)cudaGetLastError();
}

extern "C" int pxSolveOccupancy(const PxStepParams *P, int *ctasPerSm) {
  if (P->solver == PX_SOLVER_STATIC) {
    cudaError_t err = configureReplay();
    if (err != cudaSuccess) return (int)err;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, pxReplayKernel, PX_REPLAY_THREADS, P->solveS.total);
  }
  cudaError_t err = configureSolve();
  if (err != cudaSuccess) return (int)err;
  return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, pxSolveKernel, PX_SOLVE_THREADS, P->solveS.total);
}
