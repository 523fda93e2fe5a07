/*
 * px_oracle.c -- CPU RESTATEMENT of the reference world substep.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain, sequential C restatement of pxWorldStepWithDt (reference src/px_world.c:410-420)
 * and everything below it, operating directly on the per-world SoA blocks of the batch
 * layer (include/px_batch.h) so the CUDA kernel and this checker consume identical bytes.
 * It follows the reference statement by statement; every function cites the lines it
 * restates.  It deliberately shares no code with the CUDA kernel.
 *
 * Parity pin: the reference has no tests or golden vectors, so this file is
 * pinned against the compiled reference itself (oracle/_ref, built from /root/reference by
 * oracle/Makefile): tests/test_oracle_vs_reference.py checks bit-exact equality of pair
 * lists, manifolds and full body state over hundreds of substeps, and tests/golden/ holds
 * vectors generated from the compiled reference by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product never does; it has no CPU path.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA: SURVEY 0.11).
 */
#include "px_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- arithmetic contract (reference src/includes/px_math.h:50-114) ------------------- */

static inline float fastRsqrt(float x) {
  union {
    float f;
    uint32_t u;
  } c = {.f = x};
  c.u = 0x5f3759dfu - (c.u >> 1);
  c.f *= 1.5f - (x * 0.5f * c.f * c.f);
  return c.f;
}
static inline float fastRcp(float x) {
  x = fastRsqrt(x);
  return x * x;
}
static inline float fastDiv(float a, float b) { return a * fastRcp(b); }
static inline float fastSqrt(float x) { return x * fastRsqrt(x); }

#define O_PI 3.141592741f
#define O_2PI (O_PI * 2.0f)

/* px_math.h:226-252 */
static float fastSin(const float *table, float radians) {
  if ((radians = fmodf(radians, O_2PI)) < 0) {
    radians += O_2PI;
  }
  float position = fastDiv(radians, O_2PI / 256);
  int i1 = (int)position & 255;
  int i2 = (i1 + 1) & 255;
  float fraction = position - (int)position;
  return table[i1] * (1.0f - fraction) + table[i2] * fraction;
}
static float fastCos(const float *table, float radians) { return fastSin(table, radians + O_PI * 0.5f); }

typedef struct {
  float x, y;
} V2;
static inline V2 v2(float x, float y) { return (V2){x, y}; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vmulf(V2 a, float s) { return v2(a.x * s, a.y * s); }
static inline V2 vneg(V2 a) { return v2(-a.x, -a.y); }
static inline float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline float vlensq(V2 a) { return a.x * a.x + a.y * a.y; }

typedef struct {
  float m00, m01, m10, m11;
} M2;
static inline M2 mtranspose(M2 m) { return (M2){m.m00, m.m10, m.m01, m.m11}; }
static inline V2 mmul(M2 m, V2 v) { return v2(m.m00 * v.x + m.m01 * v.y, m.m10 * v.x + m.m11 * v.y); }

/* ---- one world, as views into its blocks ---------------------------------------------- */

typedef struct {
  const PxBatchLayout *L;
  PxWorldHeader *h;
  float *dyn;           /* [PX_DYN_ROWS][D] */
  uint8_t *flags;       /* [D] */
  uint8_t *order;       /* [D] */
  const float *dync;    /* [PX_DYNC_ROWS][D] */
  const uint32_t *dshape;
  const float *stat;    /* [PX_STAT_ROWS][S] materials + radius (immutable block) */
  const float *statm;   /* [PX_STATM_ROWS][S] static poses (mutable block) */
  const uint32_t *sshape;
  const uint8_t *sslot;
  const float *verts;   /* float4 per vertex */
  const float *sinTable;
  uint32_t D, S;
} World;

#define DYN(w, row, slot) ((w)->dyn[(row) * (w)->D + (slot)])
#define DYNC(w, row, slot) ((w)->dync[(row) * (w)->D + (slot)])
#define STAT(w, row, pos) ((w)->stat[(row) * (w)->S + (pos)])
#define STATM(w, row, pos) ((w)->statm[(row) * (w)->S + (pos)])

#define FLAG_SLEEPING 8u /* PX_BODY_FLAG_SLEEPING, px_body.h:67 */

/* A body as the narrowphase and solver see it.  `dynamic` bodies are addressed by slot,
 * statics by their position in the static iteration order. */
typedef struct {
  int dynamic;
  uint32_t id;
  V2 pos;
  M2 rot;
  float radius; /* circle radius, or polygon maxRadius (AABB half extent) */
  uint32_t nverts, voff;
} Body;

static Body loadBody(const World *w, int dynamic, uint32_t id) {
  Body b;
  b.dynamic = dynamic;
  b.id = id;
  if (dynamic) {
    b.pos = v2(DYN(w, PX_DYN_PX, id), DYN(w, PX_DYN_PY, id));
    b.rot = (M2){DYN(w, PX_DYN_M00, id), DYN(w, PX_DYN_M01, id), DYN(w, PX_DYN_M10, id), DYN(w, PX_DYN_M11, id)};
    b.radius = DYNC(w, PX_DYNC_RADIUS, id);
    b.nverts = PX_SHAPE_COUNT(w->dshape[id]);
    b.voff = PX_SHAPE_OFFSET(w->dshape[id]);
  } else {
    b.pos = v2(STATM(w, PX_STATM_PX, id), STATM(w, PX_STATM_PY, id));
    b.rot = (M2){STATM(w, PX_STATM_M00, id), STATM(w, PX_STATM_M01, id), STATM(w, PX_STATM_M10, id), STATM(w, PX_STATM_M11, id)};
    b.radius = STAT(w, PX_STAT_RADIUS, id);
    b.nverts = PX_SHAPE_COUNT(w->sshape[id]);
    b.voff = PX_SHAPE_OFFSET(w->sshape[id]);
  }
  return b;
}
static inline V2 bodyVertex(const World *w, const Body *b, uint32_t i) {
  /* pxVec2ArrayGet returns (0,0) out of range (px_vec2.h:495-501) */
  if (i >= b->nverts) return v2(0.0f, 0.0f);
  const float *p = w->verts + (size_t)(b->voff + i) * 4;
  return v2(p[0], p[1]);
}
static inline V2 bodyNormal(const World *w, const Body *b, uint32_t i) {
  if (i >= b->nverts) return v2(0.0f, 0.0f);
  const float *p = w->verts + (size_t)(b->voff + i) * 4;
  return v2(p[2], p[3]);
}

typedef struct {
  uint16_t a;     /* dynamic slot of bodyA */
  uint16_t b;     /* dynamic slot, or static position, of bodyB */
  uint8_t bdyn;
  uint8_t count;
  V2 normal;
  float penetration, sf, df, e;
  V2 c[2];
} Manifold;

/* ---- narrowphase (reference src/px_collision.c) --------------------------------------- */

/* px_collision.c:15-52 */
static void collideCircles(const World *w, const Body *A, const Body *B, Manifold *m) {
  V2 normal = vsub(B->pos, A->pos);
  float ra = A->radius, rb = B->radius;
  float combined = ra + rb;
  float d2 = vlensq(normal);
  if (d2 >= combined * combined) {
    m->count = 0;
    return;
  }
  float dist = fastSqrt(d2);
  if (dist < w->L->epsilon) {
    m->penetration = ra;
    m->normal = v2(1, 0);
    m->count = 1;
    m->c[0] = A->pos;
    return;
  }
  m->penetration = combined - dist;
  m->normal = v2(fastDiv(normal.x, dist), fastDiv(normal.y, dist)); /* pxVec2Divf */
  m->count = 1;
  m->c[0] = vadd(A->pos, vmulf(m->normal, ra));
}

/* pxVec2Normalize, px_vec2.h:459-472 */
static V2 normalize(const World *w, V2 v) {
  float len = fastSqrt(vlensq(v));
  if (len <= w->L->epsilon) return v;
  float inv = fastRcp(len);
  return v2(v.x * inv, v.y * inv);
}

/* px_collision.c:61-178 */
static void collideCirclePolygon(const World *w, const Body *C, const Body *P, Manifold *m) {
  m->count = 0;
  float radius = C->radius;
  M2 rot = P->rot;
  V2 center = mmul(mtranspose(rot), vsub(C->pos, P->pos));

  float separation = -FLT_MAX;
  uint32_t face = 0;
  for (uint32_t i = 0; i < P->nverts; i++) {
    float s = vdot(bodyNormal(w, P, i), vsub(center, bodyVertex(w, P, i)));
    if (s > radius) return;
    if (s > separation) {
      separation = s;
      face = i;
    }
  }

  if (separation < w->L->epsilon) { /* centre inside the polygon */
    V2 normal = vneg(mmul(rot, bodyNormal(w, P, face)));
    m->normal = normal;
    m->penetration = radius;
    m->count = 1;
    m->c[0] = vadd(C->pos, vmulf(normal, radius));
    return;
  }

  V2 v1 = bodyVertex(w, P, face);
  V2 v2_ = bodyVertex(w, P, face + 1 < P->nverts ? face + 1 : 0);
  float radiusSq = radius * radius;
  m->penetration = radius - separation;

  float dot1 = vdot(vsub(center, v1), vsub(v2_, v1));
  if (dot1 <= 0) { /* vertex region of v1 */
    V2 d = vsub(center, v1);
    if (vdot(d, d) > radiusSq) return;
    m->normal = normalize(w, mmul(rot, vsub(v1, center)));
    m->count = 1;
    m->c[0] = vadd(mmul(rot, v1), P->pos);
    return;
  }
  float dot2 = vdot(vsub(center, v2_), vsub(v1, v2_));
  if (dot2 <= 0) { /* vertex region of v2 */
    V2 d = vsub(center, v2_);
    if (vdot(d, d) > radiusSq) return;
    m->normal = normalize(w, mmul(rot, vsub(v2_, center)));
    m->count = 1;
    m->c[0] = vadd(mmul(rot, v2_), P->pos);
    return;
  }
  V2 n = bodyNormal(w, P, face); /* face region */
  if (vdot(vsub(center, v1), n) > radius) return;
  n = vneg(mmul(rot, n));
  m->normal = n;
  m->count = 1;
  m->c[0] = vadd(C->pos, vmulf(n, radius));
}

/* pxPolygonGetSupport, src/px_polygon.c:326-347 */
static V2 polygonSupport(const World *w, const Body *P, V2 dir) {
  if (vlensq(dir) < w->L->epsilonSq) return v2(0, 0);
  float best = -FLT_MAX;
  V2 bestV = bodyVertex(w, P, 0);
  for (uint32_t i = 0; i < P->nverts; i++) {
    V2 v = bodyVertex(w, P, i);
    float proj = vdot(v, dir);
    if (proj > best) {
      bestV = v;
      best = proj;
    }
  }
  return bestV;
}

/* pxFindAxisLeastPenetration, px_collision.c:180-218 */
static float leastPenetration(const World *w, uint32_t *faceIndex, const Body *A, const Body *B) {
  float bestD = -FLT_MAX;
  uint32_t bestI = 0;
  for (uint32_t i = 0; i < A->nverts; i++) {
    V2 n = bodyNormal(w, A, i);
    V2 nw = mmul(A->rot, n);
    M2 buT = mtranspose(B->rot);
    n = mmul(buT, nw);
    V2 s = polygonSupport(w, B, vneg(n));
    V2 v = bodyVertex(w, A, i);
    v = vadd(mmul(A->rot, v), A->pos);
    v = vsub(v, B->pos);
    v = mmul(buT, v);
    float d = vdot(n, vsub(s, v));
    if (d > bestD) {
      bestD = d;
      bestI = i;
    }
  }
  *faceIndex = bestI;
  return bestD;
}

/* pxFindIncidentFace, px_collision.c:220-259 */
static void incidentFace(const World *w, V2 out[2], const Body *ref, const Body *inc, uint32_t refIndex) {
  V2 refNormal = mmul(ref->rot, bodyNormal(w, ref, refIndex));
  V2 n = mmul(mtranspose(inc->rot), refNormal);
  uint32_t face = 0;
  float minDot = FLT_MAX;
  for (uint32_t i = 0; i < inc->nverts; i++) {
    float d = vdot(n, bodyNormal(w, inc, i));
    if (d < minDot) {
      minDot = d;
      face = i;
    }
  }
  out[0] = vadd(inc->pos, mmul(inc->rot, bodyVertex(w, inc, face)));
  uint32_t next = (face + 1) % inc->nverts;
  out[1] = vadd(inc->pos, mmul(inc->rot, bodyVertex(w, inc, next)));
}

/* pxClipSegmentToLine, px_collision.c:261-290 (the one true division of the hot path) */
static int clipSegment(V2 out[2], const V2 in[2], V2 normal, float offset) {
  int sp = 0;
  float d1 = vdot(normal, in[0]) - offset;
  float d2 = vdot(normal, in[1]) - offset;
  if (d1 * d2 < 0.0f) {
    float alpha = d1 / (d1 - d2);
    out[sp] = vadd(in[0], vmulf(vsub(in[1], in[0]), alpha));
    sp++;
  }
  if (d2 <= 0.0f) out[sp++] = in[1];
  if (d1 <= 0.0f) out[sp++] = in[0];
  return sp;
}

/* pxCollidePolygons, px_collision.c:299-435 */
static void collidePolygons(const World *w, const Body *A, const Body *B, Manifold *m) {
  m->count = 0;
  uint32_t faceA, faceB;
  float penA = leastPenetration(w, &faceA, A, B);
  if (penA >= 0.0f) return;
  float penB = leastPenetration(w, &faceB, B, A);
  if (penB >= 0.0f) return;

  const Body *ref, *inc;
  uint32_t refIndex;
  int flip;
  if (penA >= penB * 0.95f + penA * 0.01f) { /* pxBiasGt, px_math.h:166-168 */
    ref = A;
    inc = B;
    refIndex = faceA;
    flip = 0;
  } else {
    ref = B;
    inc = A;
    refIndex = faceB;
    flip = 1;
  }

  V2 incFace[2];
  incidentFace(w, incFace, ref, inc, refIndex);

  V2 refFace[2];
  refFace[0] = vadd(ref->pos, mmul(ref->rot, bodyVertex(w, ref, refIndex)));
  refIndex = (refIndex + 1) % ref->nverts;
  refFace[1] = vadd(ref->pos, mmul(ref->rot, bodyVertex(w, ref, refIndex)));

  V2 side = normalize(w, vsub(refFace[1], refFace[0]));
  V2 refNormal = v2(-side.y, side.x);
  if (flip) refNormal = vneg(refNormal);

  float refC = vdot(refNormal, refFace[0]);
  float negSide = -vdot(side, refFace[0]);
  V2 clip1[2];
  int cp = clipSegment(clip1, incFace, vneg(side), negSide);
  if (cp < 2) return;
  float posSide = vdot(side, refFace[1]);
  V2 clip2[2];
  cp = clipSegment(clip2, clip1, side, posSide);
  if (cp < 2) return;

  m->normal = flip ? vneg(refNormal) : refNormal;
  int n = 0;
  for (int i = 0; i < cp; i++) {
    float separation = vdot(refNormal, clip2[i]) - refC;
    if (separation <= 0.0f) {
      m->c[n] = clip2[i];
      if (n == 0) {
        m->penetration = -separation;
      } else {
        m->penetration += -separation;
      }
      n++;
      if (n >= 2) break;
    }
  }
  if (n > 1) m->penetration = fastDiv(m->penetration, (float)n);
  m->count = (uint8_t)n;
}

/* pxCollide, px_collision.c:437-465: circle = 0 vertices */
static void collide(const World *w, const Body *A, const Body *B, Manifold *m) {
  int polyA = A->nverts != 0, polyB = B->nverts != 0;
  if (!polyA && polyB) {
    collideCirclePolygon(w, A, B, m);
  } else if (!polyA && !polyB) {
    collideCircles(w, A, B, m);
  } else if (polyA && !polyB) {
    collideCirclePolygon(w, B, A, m);
    m->normal = vneg(m->normal);
  } else {
    collidePolygons(w, A, B, m);
  }
}

/* ---- the substep ---------------------------------------------------------------------- */

typedef struct {
  float minx, miny, maxx, maxy;
} Box;

typedef struct {
  Manifold *items;
  uint32_t count, cap;
  uint16_t *pairA, *pairB; /* ordered AABB-pass pair log (slot ids) */
  uint32_t pairs, pairCap;
  Box *dynBox, *statBox;
} Scratch;

static inline Box bodyBox(V2 pos, float r) {
  /* radius-based AABB: pxCircleUpdateAABB px_circle.c:43-48, pxPolygonUpdateAABB px_polygon.c:261-266 */
  return (Box){pos.x - r, pos.y - r, pos.x + r, pos.y + r};
}

/* pxBodyAABBsOverlap, px_body.c:126-145: touching counts as overlap */
static inline int boxesOverlap(Box a, Box b) {
  if (a.maxx < b.minx || a.minx > b.maxx) return 0;
  if (a.maxy < b.miny || a.miny > b.maxy) return 0;
  return 1;
}

static inline float bodyMaterial(const World *w, int dynamic, uint32_t id, int which) {
  /* which: 0 e, 1 sf, 2 df */
  return dynamic ? DYNC(w, PX_DYNC_E + which, id) : STAT(w, PX_STAT_E + which, id);
}

/* pxDetectCollision, px_world.c:327-345 (+ pxManifoldInit px_manifold.c:12-31) */
static void detectCollision(World *w, Scratch *s, uint32_t slotA, int bdyn, uint32_t idB) {
  Box a = s->dynBox[slotA];
  Box b = bdyn ? s->dynBox[idB] : s->statBox[idB];
  if (!boxesOverlap(a, b)) return;

  if (s->pairs >= s->pairCap) {
    w->h->statOverflow |= PX_BATCH_OVERFLOW_PAIRS;
    return;
  }
  s->pairA[s->pairs] = (uint16_t)slotA;
  s->pairB[s->pairs] = (uint16_t)((bdyn ? 0x8000u : 0u) | (bdyn ? idB : w->sslot[idB]));
  s->pairs++;

  Manifold m;
  memset(&m, 0, sizeof(m));
  m.a = (uint16_t)slotA;
  m.b = (uint16_t)idB;
  m.bdyn = (uint8_t)bdyn;
  m.sf = fastSqrt(bodyMaterial(w, 1, slotA, 1) * bodyMaterial(w, bdyn, idB, 1));
  m.df = fastSqrt(bodyMaterial(w, 1, slotA, 2) * bodyMaterial(w, bdyn, idB, 2));
  Body A = loadBody(w, 1, slotA);
  Body B = loadBody(w, bdyn, idB);
  collide(w, &A, &B, &m);
  if (m.count == 0) return;
  if (s->count >= s->cap) {
    w->h->statOverflow |= PX_BATCH_OVERFLOW_MANIFOLDS;
    return;
  }
  s->items[s->count++] = m;
}

/* pxBodyArraySortByAxis, px_body_array.c:81-102 */
static void sortByMinX(World *w, const Scratch *s) {
  uint32_t n = w->h->nDynamic;
  for (uint32_t i = 1; i < n; i++) {
    uint8_t cur = w->order[i];
    float key = s->dynBox[cur].minx;
    int j = (int)i - 1;
    while (j >= 0) {
      if (s->dynBox[w->order[j]].minx <= key) break;
      w->order[j + 1] = w->order[j];
      j--;
    }
    w->order[j + 1] = cur;
  }
}

/* pxBodyArrayFindFirstIndexAfterX, px_body_array.c:104-131: lower bound with the predicate
 * aabb.max.x < minX over a list sorted by aabb.min.x -- replicated literally */
static uint32_t firstStaticAfterX(const World *w, const Scratch *s, float minX) {
  uint32_t n = w->h->nStatic;
  if (n == 0) return 0;
  uint32_t low = 0, high = n;
  while (low < high) {
    uint32_t mid = low + (high - low) / 2;
    if (s->statBox[mid].maxx < minX) {
      low = mid + 1;
    } else {
      high = mid;
    }
  }
  return low;
}

/* pxGenerateCollisionInfo, px_world.c:355-392.  The `break`s of the reference sit inside the
 * one-shot inner `for` of its iteration macro and therefore act as `continue`. */
static void generateCollisionInfo(World *w, Scratch *s) {
  uint32_t nd = w->h->nDynamic, ns = w->h->nStatic;
  for (uint32_t i = 0; i < nd; i++) {
    uint32_t slot = w->order[i];
    s->dynBox[slot] = bodyBox(v2(DYN(w, PX_DYN_PX, slot), DYN(w, PX_DYN_PY, slot)), DYNC(w, PX_DYNC_RADIUS, slot));
  }
  for (uint32_t i = 0; i < ns; i++) {
    s->statBox[i] = bodyBox(v2(STATM(w, PX_STATM_PX, i), STATM(w, PX_STATM_PY, i)), STAT(w, PX_STAT_RADIUS, i));
  }
  sortByMinX(w, s);
  for (uint32_t i = 0; i < nd; i++) {
    uint32_t a = w->order[i];
    float minX = s->dynBox[a].minx, maxX = s->dynBox[a].maxx;
    for (uint32_t j = i + 1; j < nd; j++) {
      uint32_t b = w->order[j];
      if ((w->flags[a] & FLAG_SLEEPING) && (w->flags[b] & FLAG_SLEEPING)) continue;
      if (s->dynBox[b].minx > maxX) continue;
      detectCollision(w, s, a, 1, b);
    }
    uint32_t start = firstStaticAfterX(w, s, minX);
    for (uint32_t j = start; j < ns; j++) {
      if (s->statBox[j].minx > maxX) continue;
      detectCollision(w, s, a, 0, j);
    }
  }
}

typedef struct {
  V2 vel;
  float av;
  V2 pos;
  float iM, iI;
} Mover;

static Mover loadMover(const World *w, int dynamic, uint32_t id) {
  Mover m;
  if (dynamic) {
    m.vel = v2(DYN(w, PX_DYN_VX, id), DYN(w, PX_DYN_VY, id));
    m.av = DYN(w, PX_DYN_AV, id);
    m.pos = v2(DYN(w, PX_DYN_PX, id), DYN(w, PX_DYN_PY, id));
    m.iM = DYNC(w, PX_DYNC_IMASS, id);
    m.iI = DYNC(w, PX_DYNC_IINERTIA, id);
  } else { /* statics: zero velocity and zero inverse mass (px_body.c:30-44) */
    m.vel = v2(0, 0);
    m.av = 0;
    m.pos = v2(STATM(w, PX_STATM_PX, id), STATM(w, PX_STATM_PY, id));
    m.iM = 0;
    m.iI = 0;
  }
  return m;
}

/* pxManifoldDetectRestingContact, px_manifold.c:46-92 */
static void detectRestingContact(const World *w, Manifold *m) {
  if (m->count == 0) return;
  float e = bodyMaterial(w, 1, m->a, 0);
  float eb = bodyMaterial(w, m->bdyn, m->b, 0);
  float minE = e < eb ? e : eb; /* pxMin */
  V2 gstep = vmulf(v2(w->h->gravityX, w->h->gravityY), w->h->dt);
  float gsq = vlensq(gstep);
  Mover A = loadMover(w, 1, m->a), B = loadMover(w, m->bdyn, m->b);
  for (int i = 0; i < m->count; i++) {
    V2 ra = vsub(m->c[i], A.pos), rb = vsub(m->c[i], B.pos);
    V2 va = vadd(A.vel, vmulf(v2(-ra.y, ra.x), A.av));
    V2 vb = vadd(B.vel, vmulf(v2(-rb.y, rb.x), B.av));
    V2 rv = vsub(vb, va);
    if (vlensq(rv) < gsq + w->L->epsilon) {
      minE = 0.0f;
      break;
    }
  }
  m->e = minE;
}

/* pxBodyApplyImpulse, px_body.c:147-156 */
static void applyImpulse(World *w, uint32_t slot, V2 impulse, V2 contact) {
  float iM = DYNC(w, PX_DYNC_IMASS, slot), iI = DYNC(w, PX_DYNC_IINERTIA, slot);
  DYN(w, PX_DYN_VX, slot) += impulse.x * iM;
  DYN(w, PX_DYN_VY, slot) += impulse.y * iM;
  DYN(w, PX_DYN_AV, slot) += vcross(contact, impulse) * iI;
}

/* pxManifoldApplyImpulse, px_manifold.c:98-215 */
static void manifoldApplyImpulse(World *w, const Manifold *m) {
  if (!m->count) return;
  for (int i = 0; i < m->count; i++) {
    Mover A = loadMover(w, 1, m->a), B = loadMover(w, m->bdyn, m->b);
    V2 ra = vsub(m->c[i], A.pos), rb = vsub(m->c[i], B.pos);
    V2 raPerp = v2(-ra.y, ra.x), rbPerp = v2(-rb.y, rb.x);
    V2 va = vadd(A.vel, vmulf(raPerp, A.av));
    V2 vb = vadd(B.vel, vmulf(rbPerp, B.av));
    V2 rv = vsub(vb, va);
    float vn = vdot(rv, m->normal);
    if (vn > 0) continue;

    float raCrossN = vcross(ra, m->normal), rbCrossN = vcross(rb, m->normal);
    float invMassSum = A.iM + B.iM + raCrossN * raCrossN * A.iI + rbCrossN * rbCrossN * B.iI;
    float normalMag = -(1.0f + m->e) * vn;
    float denom = invMassSum * m->count;
    float j = fastDiv(normalMag, denom);
    V2 impulse = vmulf(m->normal, j);
    applyImpulse(w, m->a, vneg(impulse), ra);
    if (m->bdyn) applyImpulse(w, m->b, impulse, rb);

    A = loadMover(w, 1, m->a);
    B = loadMover(w, m->bdyn, m->b);
    va = vadd(A.vel, vmulf(raPerp, A.av));
    vb = vadd(B.vel, vmulf(rbPerp, B.av));
    rv = vsub(vb, va);
    V2 tangent = vsub(rv, vmulf(m->normal, vdot(rv, m->normal))); /* pxVec2Tangent */
    float tlen = fastSqrt(vlensq(tangent));
    if (tlen > w->L->epsilon) {
      tangent = vmulf(tangent, fastRcp(tlen));
    } else {
      continue;
    }
    float tmag = -vdot(rv, tangent);
    float tdenom = invMassSum * m->count;
    float jt = fastDiv(tmag, tdenom);
    V2 friction;
    if (fabsf(jt) < j * m->sf) {
      friction = vmulf(tangent, jt);
    } else {
      friction = vmulf(tangent, -j * m->df);
    }
    applyImpulse(w, m->a, vneg(friction), ra);
    if (m->bdyn) applyImpulse(w, m->b, friction, rb);
  }
}

/* pxBodyMoveBy, px_body.c:95-102 (the AABB is re-derived from the position on demand) */
static void moveBy(World *w, uint32_t slot, V2 d) {
  DYN(w, PX_DYN_PX, slot) += d.x;
  DYN(w, PX_DYN_PY, slot) += d.y;
}

/* pxManifoldCorrectPosition, px_manifold.c:217-255 */
static void correctPosition(World *w, const Manifold *m) {
  if (m->penetration <= w->L->allowedPenetration || m->count == 0) return;
  float iMA = DYNC(w, PX_DYNC_IMASS, m->a);
  float iMB = m->bdyn ? DYNC(w, PX_DYNC_IMASS, m->b) : 0.0f;
  float excess = m->penetration - w->L->allowedPenetration;
  float mag = fastDiv(excess * 0.2f, iMA + iMB);
  V2 corr = vmulf(m->normal, mag);
  moveBy(w, m->a, vneg(vmulf(corr, iMA)));
  if (m->bdyn) moveBy(w, m->b, vmulf(corr, iMB));
}

/* pxUpdateSleepStates, px_world.c:141-219.  Flags are SET by slot (worldIndex) and TESTED by
 * the position in the sorted iteration order -- the reference's index mix-up, kept. */
static void updateSleepStates(World *w, const Scratch *s) {
  uint8_t resting[256], wake[256];
  memset(resting, 0, sizeof(resting));
  memset(wake, 0, sizeof(wake));
  for (uint32_t k = 0; k < s->count; k++) {
    const Manifold *m = &s->items[k];
    int isResting = (m->e == 0.0f);
    if (!m->bdyn && isResting) {
      resting[m->a] = 1;
    } else if (m->bdyn && !isResting) {
      if (w->flags[m->a] & FLAG_SLEEPING) wake[m->b] = 1;
      if (w->flags[m->b] & FLAG_SLEEPING) wake[m->a] = 1;
    }
  }
  float dt = w->h->dt;
  for (uint32_t idx = 0; idx < w->h->nDynamic; idx++) {
    uint32_t slot = w->order[idx];
    if (wake[idx]) { /* pxBodyWakeUp, px_body.c:195-202 */
      DYN(w, PX_DYN_SLEEP, slot) = 0.0f;
      w->flags[slot] &= (uint8_t)~FLAG_SLEEPING;
      continue;
    }
    V2 vel = v2(DYN(w, PX_DYN_VX, slot), DYN(w, PX_DYN_VY, slot));
    int below = (vlensq(vel) <= w->h->linearSleepThresholdSq) &&
                (fabsf(DYN(w, PX_DYN_AV, slot)) <= w->h->angularSleepThreshold);
    if (!below) {
      DYN(w, PX_DYN_SLEEP, slot) = 0.0f;
      w->flags[slot] &= (uint8_t)~FLAG_SLEEPING;
      continue;
    }
    if (w->flags[slot] & FLAG_SLEEPING) continue;
    DYN(w, PX_DYN_SLEEP, slot) += dt;
    if (resting[idx]) DYN(w, PX_DYN_SLEEP, slot) += dt;
    if (DYN(w, PX_DYN_SLEEP, slot) >= w->h->timeToSleep) {
      DYN(w, PX_DYN_VX, slot) = 0;
      DYN(w, PX_DYN_VY, slot) = 0;
      DYN(w, PX_DYN_AV, slot) = 0.0f;
      DYN(w, PX_DYN_FX, slot) = 0;
      DYN(w, PX_DYN_FY, slot) = 0;
      DYN(w, PX_DYN_TORQUE, slot) = 0;
      w->flags[slot] |= FLAG_SLEEPING;
    }
  }
}

static void writeTrace(const World *w, const Scratch *s, void *trace) {
  const PxBatchLayout *L = w->L;
  uint16_t *pa = (uint16_t *)((uint8_t *)trace + L->offTracePairs);
  uint16_t *pb = pa + L->maxPairs;
  memcpy(pa, s->pairA, sizeof(uint16_t) * s->pairs);
  memcpy(pb, s->pairB, sizeof(uint16_t) * s->pairs);
  float *mf = (float *)((uint8_t *)trace + L->offTraceManifolds);
  uint32_t *ids = (uint32_t *)(mf + (size_t)PX_TRACE_M_F32 * L->maxManifolds);
  uint32_t M = L->maxManifolds;
  for (uint32_t k = 0; k < s->count; k++) {
    const Manifold *m = &s->items[k];
    mf[0 * M + k] = m->normal.x;
    mf[1 * M + k] = m->normal.y;
    mf[2 * M + k] = m->penetration;
    mf[3 * M + k] = m->sf;
    mf[4 * M + k] = m->df;
    mf[5 * M + k] = m->e;
    mf[6 * M + k] = m->c[0].x;
    mf[7 * M + k] = m->c[0].y;
    mf[8 * M + k] = m->count > 1 ? m->c[1].x : 0.0f;
    mf[9 * M + k] = m->count > 1 ? m->c[1].y : 0.0f;
    uint32_t slotB = m->bdyn ? m->b : w->sslot[m->b];
    ids[k] = (uint32_t)m->a | (slotB << 8) | ((uint32_t)m->bdyn << 16) | ((uint32_t)m->count << 24);
  }
}

/* pxWorldStepWithDt, px_world.c:410-420 */
static void substep(World *w, Scratch *s, void *trace) {
  PxWorldHeader *h = w->h;
  const float dt = h->dt;
  const V2 g = v2(h->gravityX, h->gravityY);
  s->count = 0; /* pxManifoldPoolClear */
  s->pairs = 0;
  generateCollisionInfo(w, s);

  for (uint32_t k = 0; k < s->count; k++) detectRestingContact(w, &s->items[k]); /* px_world.c:126-130 */

  /* pxIntegrateForces px_world.c:231-239, pxBodyIntegrateForces px_body.c:167-179 */
  for (uint32_t i = 0; i < h->nDynamic; i++) {
    uint32_t slot = w->order[i];
    if (w->flags[slot] & FLAG_SLEEPING) continue;
    float halfDt = fastDiv(dt, 2);
    float iM = DYNC(w, PX_DYNC_IMASS, slot), iI = DYNC(w, PX_DYNC_IINERTIA, slot);
    V2 f = v2(DYN(w, PX_DYN_FX, slot), DYN(w, PX_DYN_FY, slot));
    V2 dv = vmulf(vadd(vmulf(f, iM), g), halfDt);
    DYN(w, PX_DYN_VX, slot) += dv.x;
    DYN(w, PX_DYN_VY, slot) += dv.y;
    DYN(w, PX_DYN_AV, slot) += DYN(w, PX_DYN_TORQUE, slot) * iI * halfDt;
  }

  /* pxApplyImpulses px_world.c:249-260 */
  for (uint32_t it = 0; it < (h->iterations & 255u); it++) {
    for (uint32_t k = 0; k < s->count; k++) {
      const Manifold *m = &s->items[k];
      int sleepA = w->flags[m->a] & FLAG_SLEEPING;
      int sleepB = m->bdyn ? (w->flags[m->b] & FLAG_SLEEPING) : 0;
      if (sleepA && sleepB) continue;
      manifoldApplyImpulse(w, m);
    }
  }

  /* pxIntegrateVelocities px_world.c:271-279, pxBodyIntegrateVelocity px_body.c:181-188,
   * pxBodySetOrientation px_body.c:104-116, pxMat2Orientation px_mat2.h:54-59 */
  for (uint32_t i = 0; i < h->nDynamic; i++) {
    uint32_t slot = w->order[i];
    if (w->flags[slot] & FLAG_SLEEPING) continue;
    moveBy(w, slot, vmulf(v2(DYN(w, PX_DYN_VX, slot), DYN(w, PX_DYN_VY, slot)), dt));
    float radians = DYN(w, PX_DYN_ANGLE, slot) + DYN(w, PX_DYN_AV, slot) * dt;
    if ((radians = fmodf(radians, O_2PI)) < 0) radians += O_2PI;
    DYN(w, PX_DYN_ANGLE, slot) = radians;
    float c = fastCos(w->sinTable, radians), sn = fastSin(w->sinTable, radians);
    DYN(w, PX_DYN_M00, slot) = c;
    DYN(w, PX_DYN_M01, slot) = -sn;
    DYN(w, PX_DYN_M10, slot) = sn;
    DYN(w, PX_DYN_M11, slot) = c;
  }

  for (uint32_t k = 0; k < s->count; k++) correctPosition(w, &s->items[k]); /* px_world.c:288-292 */

  updateSleepStates(w, s);

  /* pxClearForces px_world.c:302-311 */
  for (uint32_t i = 0; i < h->nDynamic; i++) {
    uint32_t slot = w->order[i];
    if (DYN(w, PX_DYN_SLEEP, slot) >= h->timeToSleep) continue;
    DYN(w, PX_DYN_FX, slot) = 0;
    DYN(w, PX_DYN_FY, slot) = 0;
    DYN(w, PX_DYN_TORQUE, slot) = 0;
  }

  uint32_t sleeping = 0;
  for (uint32_t i = 0; i < h->nDynamic; i++) sleeping += (w->flags[w->order[i]] & FLAG_SLEEPING) ? 1 : 0;
  h->statManifolds = s->count;
  h->statPairs = s->pairs;
  h->statSleeping = sleeping;
  h->substeps++;
  if (trace) writeTrace(w, s, trace);
}

static void bindWorld(World *w, const PxBatchLayout *L, void *mut, const void *imm, const float *sinTable, uint32_t index) {
  uint8_t *m = (uint8_t *)mut + (size_t)index * L->mutStride;
  const uint8_t *c = (const uint8_t *)imm + (size_t)index * L->immStride;
  w->L = L;
  w->h = (PxWorldHeader *)(m + L->offHeader);
  w->dyn = (float *)(m + L->offDyn);
  w->flags = m + L->offDynFlags;
  w->order = m + L->offDynOrder;
  w->dync = (const float *)(c + L->offDynConst);
  w->dshape = (const uint32_t *)(c + L->offDynShape);
  w->stat = (const float *)(c + L->offStat);
  w->statm = (const float *)(m + L->offStatMut);
  w->sshape = (const uint32_t *)(c + L->offStatShape);
  w->sslot = c + L->offStatSlot;
  w->verts = (const float *)(c + L->offVerts);
  w->sinTable = sinTable;
  w->D = L->maxDynamic;
  w->S = L->maxStatic;
}

static void scratchInit(Scratch *s, const PxBatchLayout *L) {
  s->cap = L->maxManifolds;
  s->pairCap = L->maxPairs;
  s->items = (Manifold *)malloc(sizeof(Manifold) * s->cap);
  s->pairA = (uint16_t *)malloc(sizeof(uint16_t) * s->pairCap);
  s->pairB = (uint16_t *)malloc(sizeof(uint16_t) * s->pairCap);
  s->dynBox = (Box *)malloc(sizeof(Box) * 256);
  s->statBox = (Box *)malloc(sizeof(Box) * 256);
  s->count = s->pairs = 0;
}
static void scratchFree(Scratch *s) {
  free(s->items);
  free(s->pairA);
  free(s->pairB);
  free(s->dynBox);
  free(s->statBox);
}

void pxOracleSubsteps(const PxBatchLayout *L, void *mut, const void *imm, const float *sinTable, uint32_t first,
                      uint32_t n, uint32_t k, void *trace) {
  Scratch s;
  scratchInit(&s, L);
  for (uint32_t i = 0; i < n; i++) {
    World w;
    bindWorld(&w, L, mut, imm, sinTable, first + i);
    void *t = trace ? (uint8_t *)trace + (size_t)i * L->traceStride : NULL;
    for (uint32_t step = 0; step < k; step++) substep(&w, &s, t);
  }
  scratchFree(&s);
}

typedef struct {
  const PxBatchLayout *L;
  void *mut;
  const void *imm;
  const float *sinTable;
  uint32_t first, n, k;
} Job;

static void *worker(void *arg) {
  Job *j = (Job *)arg;
  pxOracleSubsteps(j->L, j->mut, j->imm, j->sinTable, j->first, j->n, j->k, NULL);
  return NULL;
}

double pxOracleTimeSubsteps(const PxBatchLayout *L, void *mut, const void *imm, const float *sinTable, uint32_t nWorlds,
                            uint32_t k, uint32_t nThreads) {
  if (nThreads < 1) nThreads = 1;
  if (nThreads > nWorlds) nThreads = nWorlds;
  pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * nThreads);
  Job *jobs = (Job *)malloc(sizeof(Job) * nThreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (uint32_t t = 0; t < nThreads; t++) {
    uint32_t a = (uint32_t)((uint64_t)nWorlds * t / nThreads), b = (uint32_t)((uint64_t)nWorlds * (t + 1) / nThreads);
    jobs[t] = (Job){L, mut, imm, sinTable, a, b - a, k};
    pthread_create(&tid[t], NULL, worker, &jobs[t]);
  }
  for (uint32_t t = 0; t < nThreads; t++) pthread_join(tid[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(tid);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* T0 arithmetic probes */
void pxOracleMath(int op, const float *a, const float *b, const float *sinTable, float *out, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    switch (op) {
    case 0: out[i] = fastRsqrt(a[i]); break;
    case 1: out[i] = fastRcp(a[i]); break;
    case 2: out[i] = fastSqrt(a[i]); break;
    case 3: out[i] = fastDiv(a[i], b[i]); break;
    case 4: out[i] = fastSin(sinTable, a[i]); break;
    case 5: out[i] = fastCos(sinTable, a[i]); break;
    default: out[i] = 0.0f;
    }
  }
}

/* pxFastRsqrt over bit patterns [first << 20, (first + n) << 20): one 64-bit sum of the result
 * bit patterns per 2^20 block (same reduction as the device self-test) */
void pxOracleRsqrtSweep(uint64_t *sums, uint32_t first, uint32_t n) {
  for (uint32_t b = 0; b < n; b++) {
    uint64_t acc = 0;
    uint32_t base = (first + b) << 20;
    for (uint32_t k = 0; k < (1u << 20); k++) {
      union {
        float f;
        uint32_t u;
      } c;
      c.u = base + k;
      c.f = fastRsqrt(c.f);
      acc += (c.f != c.f) ? 0x7fc00000u : c.u; /* NaN payloads are outside the contract */
    }
    sums[b] = acc;
  }
}
