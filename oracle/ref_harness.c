/*
 * ref_harness.c -- unity translation unit around the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing the product ships may call into this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Why one TU: the reference keeps `static float pxSinTable[]` in a header
 * (src/includes/px_math.h:206) that only registerPdPhyzx() fills (src/px_api.c:10), so a
 * multi-TU build gives px_body.c an all-zero table.  The project itself only ships the
 * amalgamated single header, i.e. one TU.
 *
 * REF_SRC is given on the command line (-I$(REF_SRC) -I$(REF_SRC)/includes): either
 * /root/reference/src (narrow, unmodified) or build/ref_wide (oracle/widen.sh copy).
 * The flat ref_* functions below exist so that Python (ctypes) never needs the layout
 * of PxWorld / PxBody.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pd_api.h"
#include "pd_stub.h"

/* ---- the reference, verbatim, in dependency order --------------------------------- */
#include "px_api.c"
#include "px_body.c"
#include "px_body_api.c"
#include "px_body_array.c"
#include "px_circle.c"
#include "px_clock.c"
#include "px_collider.c"
#include "px_collision.c"
#include "px_manifold.c"
#include "px_manifold_pool.c"
#include "px_polygon.c"
#include "px_world_api.c"

/*
 * Pair trace: px_world.c calls pxManifoldPoolAcquire at exactly one site
 * (src/px_world.c:333) where bodyA/bodyB are in scope and the AABB test has just passed.
 * Redirecting that one call logs the ordered AABB-pass pair list without editing the file.
 */
typedef struct {
  uint16_t *a;     /* slot of bodyA (always a dynamic body) */
  uint16_t *b;     /* slot of bodyB */
  uint8_t *bdyn;   /* 1 if bodyB is dynamic */
  uint32_t cap, len;
} RefPairLog;

static __thread RefPairLog *tl_pairlog = NULL;

static inline PxManifold *refLoggedAcquire(PxManifoldPool *pool, PxBody *bodyA, PxBody *bodyB) {
  RefPairLog *log = tl_pairlog;
  if (log) {
    if (log->len < log->cap) {
      log->a[log->len] = bodyA->worldIndex;
      log->b[log->len] = bodyB->worldIndex;
      log->bdyn[log->len] = pxBodyIsDynamic(bodyB) ? 1 : 0;
    }
    log->len++;
  }
  return pxManifoldPoolAcquire(pool);
}
#define pxManifoldPoolAcquire(pool) refLoggedAcquire((pool), bodyA, bodyB)
#ifdef PXREF_B200
/* "b200" variant: the reference with the one-function binding of integration/px_world_b200.c
 * (INTEGRATION.md): its own pxWorldStep is renamed, the binding's definition is what
 * px_world_api.c binds into PxWorldAPI.step.  Everything else is the reference, verbatim. */
#define pxWorldStep pxWorldStepCpu
#endif
#include "px_world.c"
#undef pxManifoldPoolAcquire
#ifdef PXREF_B200
#undef pxWorldStep
#include "../integration/px_world_b200.c"
#endif

/* ---- flat C ABI for the test harness ------------------------------------------------ */

static PdPhyzxAPI *g_px = NULL;

#define REF_EXPORT __attribute__((visibility("default")))

/* constructor: the reference's own destructors (src/px_world_api.c:41-42) call
 * pd->system->realloc at unload, so `pd` must be valid even if nobody calls ref_init */
REF_EXPORT __attribute__((constructor)) void ref_init(void) {
  if (!g_px) {
    g_px = registerPdPhyzx(pdStubAPI());
  }
}

REF_EXPORT float ref_unit(void) { return PDPHYZX_UNIT; }
REF_EXPORT int ref_max_bodies(void) { return PX_MAX_BODIES; }
REF_EXPORT int ref_max_manifolds(void) { return PX_MAX_MANIFOLDS; }
REF_EXPORT size_t ref_sizeof_world(void) { return sizeof(PxWorld); }
REF_EXPORT size_t ref_sizeof_body(void) { return sizeof(PxBody); }
REF_EXPORT const float *ref_sin_table(void) { return pxSinTable; }
REF_EXPORT void ref_set_time_ms(uint32_t ms) { pdStubSetTimeMs(ms); }

REF_EXPORT PxWorld *ref_world_new(int fps) {
  ref_init();
  return g_px->world->new((uint8_t)fps);
}
REF_EXPORT void ref_world_free(PxWorld *w) {
#ifdef PXREF_B200
  pxWorldB200Release(w);
#endif
  g_px->world->free(w);
}
REF_EXPORT int ref_is_b200(void) {
#ifdef PXREF_B200
  return 1;
#else
  return 0;
#endif
}
REF_EXPORT void ref_set_gravity(PxWorld *w, float x, float y) { g_px->world->setGravity(w, x, y); }
REF_EXPORT void ref_set_iterations(PxWorld *w, int it) { g_px->world->setIterations(w, (uint8_t)it); }
REF_EXPORT void ref_set_sleep_params(PxWorld *w, float linSq, float ang, float tts) {
  w->linearSleepThresholdSq = linSq;
  w->angularSleepThreshold = ang;
  w->timeToSleep = tts;
}
REF_EXPORT void ref_get_world_params(PxWorld *w, float *out /*[8]*/) {
  out[0] = w->gravity.x;
  out[1] = w->gravity.y;
  out[2] = w->linearSleepThresholdSq;
  out[3] = w->angularSleepThreshold;
  out[4] = w->timeToSleep;
  out[5] = w->clock.fixedTimestep;
  out[6] = pxFastDiv(pxClockGetFixedDeltaTime(&w->clock), 5); /* px_world.c:432 */
  out[7] = (float)w->iterations;
}

/*
 * kind: 0 circle (p0 = radius), 1 box (p0 = width, p1 = height), 2 polygon (verts, nverts).
 * Goes through px->world->newDynamicBody / newStaticBody so that hull, normals, mass and
 * the initial AABB are produced by reference code.  Returns the slot (worldIndex) or -1.
 */
REF_EXPORT int ref_add_body(PxWorld *w, int dynamic, int kind, float p0, float p1, const float *verts,
                            int nverts, float density, float x, float y, float e, float sf, float df) {
  PxShape shape;
  memset(&shape, 0, sizeof(shape));
  if (kind == 0) {
    shape.circle.radius = p0;
  } else if (kind == 1) {
    shape.box.width = p0;
    shape.box.height = p1;
  } else {
    for (int i = 0; i < nverts && i < PX_MAX_POLY_VERTEX_COUNT; i++) {
      shape.polygon.vertices[i] = pxVec2(verts[2 * i], verts[2 * i + 1]);
    }
    shape.polygon.vertexCount = (size_t)nverts;
  }
  PxBody *body = dynamic ? g_px->world->newDynamicBody(w, shape, density, pxVec2(x, y))
                         : g_px->world->newStaticBody(w, shape, pxVec2(x, y));
  if (!body) {
    return -1;
  }
  /* callers poke material fields directly (examples/draw/circle.c:43-45) */
  body->restitution = e;
  body->staticFriction = sf;
  body->dynamicFriction = df;
  return body->worldIndex;
}

static PxBodyArray *refArray(PxWorld *w, int dynamic) {
  return dynamic ? &w->dynamicBodies : &w->staticBodies;
}

REF_EXPORT void ref_body_set_motion(PxWorld *w, int dynamic, int slot, float vx, float vy, float av,
                                    float angle) {
  PxBody *b = &refArray(w, dynamic)->items[slot];
  b->velocity = pxVec2(vx, vy);
  b->angularVelocity = av;
  g_px->body->setOrientation(b, angle);
}
REF_EXPORT void ref_body_apply_impulse(PxWorld *w, int slot, float ix, float iy, float cx, float cy) {
  g_px->body->applyImpulse(&w->dynamicBodies.items[slot], ix, iy, cx, cy);
}
REF_EXPORT void ref_body_apply_force(PxWorld *w, int slot, float fx, float fy, float cx, float cy) {
  g_px->body->applyForce(&w->dynamicBodies.items[slot], fx, fy, cx, cy);
}
REF_EXPORT void ref_body_set_position(PxWorld *w, int dynamic, int slot, float x, float y) {
  g_px->body->setPosition(&refArray(w, dynamic)->items[slot], x, y);
}
REF_EXPORT void ref_body_set_orientation(PxWorld *w, int dynamic, int slot, float rad) {
  g_px->body->setOrientation(&refArray(w, dynamic)->items[slot], rad);
}
REF_EXPORT void ref_body_move_by(PxWorld *w, int dynamic, int slot, float dx, float dy) {
  g_px->body->moveBy(&refArray(w, dynamic)->items[slot], pxVec2(dx, dy));
}
REF_EXPORT void ref_body_rotate(PxWorld *w, int dynamic, int slot, float rad) {
  g_px->body->rotate(&refArray(w, dynamic)->items[slot], rad);
}
REF_EXPORT void ref_free_body(PxWorld *w, int dynamic, int slot) {
  g_px->world->freeBody(w, &refArray(w, dynamic)->items[slot]);
}

/* n substeps of pxWorldStepWithDt (src/px_world.c:410-420) */
REF_EXPORT void ref_substep(PxWorld *w, float dt, int n) {
  for (int i = 0; i < n; i++) {
    pxWorldStepWithDt(w, dt);
  }
}
/* the clock-driven px->world->step (src/px_world.c:422-445); returns `stepped` */
REF_EXPORT int ref_step(PxWorld *w) { return g_px->world->step(w) ? 1 : 0; }

/* one substep with the ordered AABB-pass pair list captured; returns the pair count */
REF_EXPORT int ref_substep_traced(PxWorld *w, float dt, uint16_t *a, uint16_t *b, uint8_t *bdyn, int cap) {
  RefPairLog log = {a, b, bdyn, (uint32_t)cap, 0};
  tl_pairlog = &log;
  pxWorldStepWithDt(w, dt);
  tl_pairlog = NULL;
  return (int)log.len;
}

/*
 * Body dump, indexed by SLOT (row = worldIndex), REF_BODY_F32 floats per row:
 *  0 pos.x 1 pos.y 2 vel.x 3 vel.y 4 angularVelocity 5 orientationAngle 6 force.x 7 force.y
 *  8 torque 9 sleepTime 10..13 orientation m00 m01 m10 m11 14..17 aabb min.x min.y max.x max.y
 *  18 mass 19 iMass 20 I 21 iI 22 restitution 23 staticFriction 24 dynamicFriction 25 density
 *  26 radius|maxRadius 27 type (0 circle, 1 polygon) 28 vertexCount
 * flags[row] = PxBody.flags; order[i] = activeIndices[i].  Returns the active length.
 */
#define REF_BODY_F32 29
REF_EXPORT int ref_body_row_floats(void) { return REF_BODY_F32; }

REF_EXPORT int ref_get_bodies(PxWorld *w, int dynamic, float *rows, uint8_t *flags, uint8_t *order,
                              int rowCap) {
  PxBodyArray *arr = refArray(w, dynamic);
  for (int i = 0; i < arr->length; i++) {
    int slot = arr->activeIndices[i];
    order[i] = (uint8_t)slot;
    if (slot >= rowCap) {
      continue;
    }
    PxBody *b = &arr->items[slot];
    float *r = rows + (size_t)slot * REF_BODY_F32;
    r[0] = b->position.x;
    r[1] = b->position.y;
    r[2] = b->velocity.x;
    r[3] = b->velocity.y;
    r[4] = b->angularVelocity;
    r[5] = b->orientationAngle;
    r[6] = b->force.x;
    r[7] = b->force.y;
    r[8] = b->torque;
    r[9] = b->sleepTime;
    r[10] = b->orientation.m00;
    r[11] = b->orientation.m01;
    r[12] = b->orientation.m10;
    r[13] = b->orientation.m11;
    r[14] = b->aabb.min.x;
    r[15] = b->aabb.min.y;
    r[16] = b->aabb.max.x;
    r[17] = b->aabb.max.y;
    r[18] = b->mass;
    r[19] = b->iMass;
    r[20] = b->momentOfInertia;
    r[21] = b->iMomentOfInertia;
    r[22] = b->restitution;
    r[23] = b->staticFriction;
    r[24] = b->dynamicFriction;
    r[25] = b->density;
    if (b->collider.type == PX_CIRCLE) {
      r[26] = b->collider.shape.circle.radius;
      r[27] = 0.0f;
      r[28] = 0.0f;
    } else {
      r[26] = b->collider.shape.polygon.maxRadius;
      r[27] = 1.0f;
      r[28] = (float)b->collider.shape.polygon.vertices.length;
    }
    flags[slot] = b->flags;
  }
  return arr->length;
}

/*
 * Inverse of ref_get_bodies for the MUTABLE columns 0..17 plus flags and the iteration
 * order: lets a test or the CPU baseline start the reference from a state produced
 * elsewhere (e.g. a settled pile read back from the device).  Immutable columns ignored.
 */
REF_EXPORT void ref_set_bodies(PxWorld *w, int dynamic, const float *rows, const uint8_t *flags,
                               const uint8_t *order, int n) {
  PxBodyArray *arr = refArray(w, dynamic);
  for (int i = 0; i < n && i < arr->length; i++) {
    int slot = order[i];
    arr->activeIndices[i] = (uint8_t)slot;
    PxBody *b = &arr->items[slot];
    const float *r = rows + (size_t)slot * REF_BODY_F32;
    b->position = pxVec2(r[0], r[1]);
    b->velocity = pxVec2(r[2], r[3]);
    b->angularVelocity = r[4];
    b->orientationAngle = r[5];
    b->force = pxVec2(r[6], r[7]);
    b->torque = r[8];
    b->sleepTime = r[9];
    b->orientation = (PxMat2){.m00 = r[10], .m01 = r[11], .m10 = r[12], .m11 = r[13]};
    b->aabb.min = pxVec2(r[14], r[15]);
    b->aabb.max = pxVec2(r[16], r[17]);
    b->flags = flags[slot];
  }
}

/* polygon hull of one body: verts[2*i..], normals[2*i..]; returns the vertex count (0: circle) */
REF_EXPORT int ref_get_polygon(PxWorld *w, int dynamic, int slot, float *verts, float *normals, int cap) {
  PxBody *b = &refArray(w, dynamic)->items[slot];
  if (b->collider.type != PX_POLYGON) {
    return 0;
  }
  PxPolygonData *p = &b->collider.shape.polygon;
  for (int i = 0; i < p->vertices.length && i < cap; i++) {
    verts[2 * i] = p->vertices.items[i].x;
    verts[2 * i + 1] = p->vertices.items[i].y;
    normals[2 * i] = p->normals.items[i].x;
    normals[2 * i + 1] = p->normals.items[i].y;
  }
  return p->vertices.length;
}

/*
 * Manifold dump of the pool as the last substep left it (solver order), REF_MANIFOLD_F32
 * floats per row: 0 normal.x 1 normal.y 2 penetration 3 staticFriction 4 dynamicFriction
 * 5 mixedRestitution 6 c0.x 7 c0.y 8 c1.x 9 c1.y ; ids row: slotA, slotB, bIsDynamic, contactCount.
 */
#define REF_MANIFOLD_F32 10
REF_EXPORT int ref_get_manifolds(PxWorld *w, float *rows, uint16_t *ids, int cap) {
  int n = w->contacts.length;
  for (int i = 0; i < n && i < cap; i++) {
    PxManifold *m = &w->contacts.items[i];
    float *r = rows + (size_t)i * REF_MANIFOLD_F32;
    r[0] = m->normal.x;
    r[1] = m->normal.y;
    r[2] = m->penetration;
    r[3] = m->staticFriction;
    r[4] = m->dynamicFriction;
    r[5] = m->mixedRestitution;
    r[6] = m->contacts.items[0].x;
    r[7] = m->contacts.items[0].y;
    r[8] = m->contacts.items[1].x;
    r[9] = m->contacts.items[1].y;
    ids[4 * i + 0] = m->bodyA->worldIndex;
    ids[4 * i + 1] = m->bodyB->worldIndex;
    ids[4 * i + 2] = pxBodyIsDynamic(m->bodyB) ? 1 : 0;
    ids[4 * i + 3] = m->contacts.length;
  }
  return n;
}

/* arithmetic known-answer probes (T0 tier): evaluate the reference's inline math on arrays */
REF_EXPORT void ref_math_rsqrt(const float *x, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastRsqrt(x[i]);
}
REF_EXPORT void ref_math_rcp(const float *x, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastRcp(x[i]);
}
REF_EXPORT void ref_math_sqrt(const float *x, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastSqrt(x[i]);
}
REF_EXPORT void ref_math_div(const float *a, const float *b, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastDiv(a[i], b[i]);
}
REF_EXPORT void ref_math_sin(const float *x, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastSin(x[i]);
}
REF_EXPORT void ref_math_cos(const float *x, float *y, int n) {
  for (int i = 0; i < n; i++) y[i] = pxFastCos(x[i]);
}

/*
 * CPU baseline timing: `nThreads` pthreads, each stepping a disjoint contiguous range of
 * the given worlds by `substeps` x pxWorldStepWithDt.  Worlds are independent, so this is
 * the reference's best case on the host.  Returns wall seconds (CLOCK_MONOTONIC) for the
 * slowest thread's span, measured around the stepping only.
 */
typedef struct {
  PxWorld **worlds;
  int first, last, substeps;
  float dt;
} RefJob;

static void *refWorker(void *arg) {
  RefJob *job = (RefJob *)arg;
  for (int i = job->first; i < job->last; i++) {
    for (int s = 0; s < job->substeps; s++) {
      pxWorldStepWithDt(job->worlds[i], job->dt);
    }
  }
  return NULL;
}

REF_EXPORT double ref_time_substeps(PxWorld **worlds, int nWorlds, int substeps, float dt, int nThreads) {
  if (nThreads < 1) nThreads = 1;
  if (nThreads > nWorlds) nThreads = nWorlds;
  pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
  RefJob *jobs = (RefJob *)malloc(sizeof(RefJob) * (size_t)nThreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nThreads; t++) {
    jobs[t].worlds = worlds;
    jobs[t].first = (int)((long long)nWorlds * t / nThreads);
    jobs[t].last = (int)((long long)nWorlds * (t + 1) / nThreads);
    jobs[t].substeps = substeps;
    jobs[t].dt = dt;
    pthread_create(&tid[t], NULL, refWorker, &jobs[t]);
  }
  for (int t = 0; t < nThreads; t++) {
    pthread_join(tid[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(tid);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
