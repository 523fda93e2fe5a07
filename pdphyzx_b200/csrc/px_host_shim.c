/*
 * px_host_shim.c -- flat C entry points over the host API (include/pdphyzx.h) for callers
 * that cannot see C structs (Python/ctypes).  Built once per PDPHYZX_UNIT into
 * pdphyzx_b200/lib/libpxhost_{m,cm,mm}.so, each linked against libpdphyzx_b200.so.
 *
 * This is the one translation unit that defines PDPHYZX_IMPLEMENTATION; the PlaydateAPI it
 * registers is the host stub (include/pd_host/), whose millisecond clock the caller drives.
 * Every pxh_* function is a thin pass-through to the API call it names.
 */
#include <string.h>

#include "pd_api.h"
#include "pd_stub.h"

#define PDPHYZX_IMPLEMENTATION
#include "pdphyzx.h"

#define PXH __attribute__((visibility("default")))

static PdPhyzxAPI *g_px = NULL;

PXH void pxh_init(void) {
  if (!g_px) {
    g_px = registerPdPhyzx(pdStubAPI());
  }
}

PXH float pxh_unit(void) { return PDPHYZX_UNIT; }
PXH const float *pxh_sin_table(void) { return pxSinTable; }
PXH size_t pxh_sizeof_world(void) { return sizeof(PxWorld); }
PXH void pxh_set_time_ms(uint32_t ms) { pdStubSetTimeMs(ms); }
PXH uint32_t pxh_get_time_ms(void) { return pdStubGetTimeMs(); }

PXH PxWorld *pxh_world_new(int fps) {
  pxh_init();
  return g_px->world->new((uint8_t)fps);
}
PXH void pxh_world_free(PxWorld *w) { g_px->world->free(w); }
PXH void pxh_set_gravity(PxWorld *w, float x, float y) { g_px->world->setGravity(w, x, y); }
PXH void pxh_set_iterations(PxWorld *w, int it) { g_px->world->setIterations(w, (uint8_t)it); }
PXH void pxh_set_sleep_params(PxWorld *w, float linSq, float ang, float tts) {
  w->linearSleepThresholdSq = linSq;
  w->angularSleepThreshold = ang;
  w->timeToSleep = tts;
}
PXH int pxh_step(PxWorld *w) { return g_px->world->step(w) ? 1 : 0; }
PXH void pxh_draw_debug(PxWorld *w) { g_px->world->drawDebug(w); }
PXH float pxh_substep_dt(PxWorld *w) { return pxWorldSubstepDt(w); }
PXH void pxh_world_stats(PxWorld *w, uint32_t *out /*[3]*/) {
  out[0] = w->contacts.length;
  out[1] = w->contacts.sleeping;
  out[2] = w->contacts.overflow;
}

/* kind: 0 circle (p0 radius), 1 box (p0 width, p1 height), 2 polygon (verts[2*nverts]).
 * Returns the slot (PxBody.worldIndex) or -1 when the API returned NULL. */
PXH int pxh_add_body(PxWorld *w, int dynamic, int kind, float p0, float p1, const float *verts, int nverts,
                     float density, float x, float y, float e, float sf, float df) {
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
  body->restitution = e;
  body->staticFriction = sf;
  body->dynamicFriction = df;
  return body->worldIndex;
}

/*
 * Bulk scene loader: builds `nWorlds` worlds from concatenated per-body arrays (world w owns
 * bodies [bodyStart[w], bodyStart[w+1])) through px->world->new / setIterations / setGravity
 * / newStaticBody / newDynamicBody / body->setOrientation.  motion may be NULL, else
 * [nBodies][4] = vx, vy, angularVelocity, angle.  Returns the number of worlds built.
 */
PXH int pxh_build_worlds(PxWorld **out, int nWorlds, const int *bodyStart, int fps, int iterations, float gx, float gy,
                         const uint8_t *dynamic, const uint8_t *kind, const float *p0, const float *p1,
                         const uint8_t *nverts, const float *verts /*[n][vertsPerBody][2]*/, const float *density, const float *x,
                         const float *y, const float *e, const float *sf, const float *df, const float *motion, int vertsPerBody) {
  pxh_init();
  for (int w = 0; w < nWorlds; w++) {
    PxWorld *world = g_px->world->new((uint8_t)fps);
    g_px->world->setIterations(world, (uint8_t)iterations);
    g_px->world->setGravity(world, gx, gy);
    for (int k = bodyStart[w]; k < bodyStart[w + 1]; k++) {
      int slot = pxh_add_body(world, dynamic[k], kind[k], p0[k], p1[k], verts + (size_t)k * 2 * (size_t)vertsPerBody, nverts[k], density[k],
                              x[k], y[k], e[k], sf[k], df[k]);
      if (slot < 0) {
        g_px->world->free(world);
        return w;
      }
      if (motion) {
        const float *m = motion + (size_t)k * 4;
        if (m[0] != 0 || m[1] != 0 || m[2] != 0 || m[3] != 0) {
          PxBody *b = dynamic[k] ? &world->dynamicBodies.items[slot] : &world->staticBodies.items[slot];
          b->velocity = pxVec2(m[0], m[1]);
          b->angularVelocity = m[2];
          g_px->body->setOrientation(b, m[3]);
        }
      }
    }
    out[w] = world;
  }
  return nWorlds;
}

static PxBodyArray *pxhArray(PxWorld *w, int dynamic) { return dynamic ? &w->dynamicBodies : &w->staticBodies; }

PXH void pxh_body_set_motion(PxWorld *w, int dynamic, int slot, float vx, float vy, float av, float angle) {
  PxBody *b = &pxhArray(w, dynamic)->items[slot];
  b->velocity = pxVec2(vx, vy);
  b->angularVelocity = av;
  g_px->body->setOrientation(b, angle);
}
PXH void pxh_body_apply_impulse(PxWorld *w, int slot, float ix, float iy, float cx, float cy) {
  g_px->body->applyImpulse(&w->dynamicBodies.items[slot], ix, iy, cx, cy);
}
PXH void pxh_body_apply_force(PxWorld *w, int slot, float fx, float fy, float cx, float cy) {
  g_px->body->applyForce(&w->dynamicBodies.items[slot], fx, fy, cx, cy);
}
PXH void pxh_body_set_position(PxWorld *w, int dynamic, int slot, float x, float y) {
  g_px->body->setPosition(&pxhArray(w, dynamic)->items[slot], x, y);
}
PXH void pxh_body_set_orientation(PxWorld *w, int dynamic, int slot, float rad) {
  g_px->body->setOrientation(&pxhArray(w, dynamic)->items[slot], rad);
}
PXH void pxh_body_move_by(PxWorld *w, int dynamic, int slot, float dx, float dy) {
  g_px->body->moveBy(&pxhArray(w, dynamic)->items[slot], pxVec2(dx, dy));
}
PXH void pxh_body_rotate(PxWorld *w, int dynamic, int slot, float rad) {
  g_px->body->rotate(&pxhArray(w, dynamic)->items[slot], rad);
}
PXH int pxh_body_is_valid(PxWorld *w, int dynamic, int slot) {
  return g_px->body->isValid(&pxhArray(w, dynamic)->items[slot]) ? 1 : 0;
}
PXH void *pxh_body_ptr(PxWorld *w, int dynamic, int slot) { return &pxhArray(w, dynamic)->items[slot]; }
PXH void pxh_free_body(PxWorld *w, int dynamic, int slot) {
  g_px->world->freeBody(w, &pxhArray(w, dynamic)->items[slot]);
}

/* Body dump in the same row format as oracle/ref_harness.c (29 floats per slot). */
#define PXH_BODY_F32 29
PXH int pxh_body_row_floats(void) { return PXH_BODY_F32; }
PXH int pxh_get_bodies(PxWorld *w, int dynamic, float *rows, uint8_t *flags, uint8_t *order, int rowCap) {
  PxBodyArray *arr = pxhArray(w, dynamic);
  for (int i = 0; i < arr->length; i++) {
    int slot = arr->activeIndices[i];
    order[i] = (uint8_t)slot;
    if (slot >= rowCap) {
      continue;
    }
    PxBody *b = &arr->items[slot];
    float *r = rows + (size_t)slot * PXH_BODY_F32;
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
    memcpy(r + 10, b->orientation.v, 4 * sizeof(float));
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
    int poly = b->collider.type == PX_POLYGON;
    r[26] = poly ? b->collider.shape.polygon.maxRadius : b->collider.shape.circle.radius;
    r[27] = poly ? 1.0f : 0.0f;
    r[28] = poly ? (float)b->collider.shape.polygon.vertices.length : 0.0f;
    flags[slot] = b->flags;
  }
  return arr->length;
}

PXH int pxh_get_polygon(PxWorld *w, int dynamic, int slot, float *verts, float *normals, int cap) {
  PxBody *b = &pxhArray(w, dynamic)->items[slot];
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

/* inline math of the header, for the arithmetic known-answer tests */
PXH void pxh_math(int op, const float *a, const float *b, float *out, int n) {
  pxh_init();
  for (int i = 0; i < n; i++) {
    switch (op) {
    case 0: out[i] = pxFastRsqrt(a[i]); break;
    case 1: out[i] = pxFastRcp(a[i]); break;
    case 2: out[i] = pxFastSqrt(a[i]); break;
    case 3: out[i] = pxFastDiv(a[i], b[i]); break;
    case 4: out[i] = pxFastSin(a[i]); break;
    case 5: out[i] = pxFastCos(a[i]); break;
    default: out[i] = 0.0f;
    }
  }
}
