/*
 * px_pack_inl.h -- AoS <-> SoA conversion between ONE host PxWorld and its two device blocks
 * (PxBatchLayout, include/px_batch.h), written against field NAMES only.
 *
 * Include it after a header that defines PxWorld / PxBodyArray / PxBody / PxPolygonData with
 * the reference's field names (reference src/includes/px_world.h:31-44, px_body_array.h:19-33,
 * px_body.h:82-120, px_collider.h:46-50).  Two such headers exist:
 *   - include/pdphyzx.h of this repo (pdphyzx_b200/csrc/px_pack.c is that instantiation);
 *   - the reference's own headers (integration/px_world_b200.c: the binding a reference
 *     maintainer adds).
 * The struct layouts differ (capacities, extra fields); the names do not, so the same
 * source packs both.
 *
 * Optional macros:
 *   PX_PACK_FAIL(rc, fmt, ...)   statement that reports an error and returns rc
 *                                (default: fprintf(stderr) + return)
 *   PX_PACK_WORLD_HAS_STATS      PxWorld.contacts has .sleeping and .overflow (this repo's
 *                                header does, the reference's pool does not)
 */
#ifndef PX_PACK_INL_H
#define PX_PACK_INL_H

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "px_batch.h"

#ifndef PX_PACK_FAIL
#define PX_PACK_FAIL(rc, ...)     \
  do {                            \
    fprintf(stderr, __VA_ARGS__); \
    fputc('\n', stderr);          \
    return (rc);                  \
  } while (0)
#endif

/* px->world->step runs 5 substeps per fired fixed step (reference src/px_world.c:427) */
#define PX_PACK_SUBSTEPS_PER_STEP 5

static inline uint32_t pxPackVertexCount(const PxBody *b) {
  return b->collider.type == PX_POLYGON ? (uint32_t)b->collider.shape.polygon.vertices.length : 0u;
}

/* the AABB half-extent: circle radius or the polygon's pre-recentring maxRadius
 * (reference src/px_circle.c:43-48, src/px_polygon.c:261-266) */
static inline float pxPackRadius(const PxBody *b) {
  return b->collider.type == PX_CIRCLE ? b->collider.shape.circle.radius : b->collider.shape.polygon.maxRadius;
}

/* capacities one world needs: highest dynamic slot + 1, static count, polygon vertices */
static inline void pxPackWorldNeeds(const PxWorld *world, uint32_t *maxDyn, uint32_t *maxStat, uint32_t *maxVerts) {
  uint32_t verts = 0;
  for (uint32_t i = 0; i < world->dynamicBodies.length; i++) {
    uint32_t slot = world->dynamicBodies.activeIndices[i];
    if (slot + 1 > *maxDyn) *maxDyn = slot + 1;
    verts += pxPackVertexCount(&world->dynamicBodies.items[slot]);
  }
  for (uint32_t i = 0; i < world->staticBodies.length; i++) {
    verts += pxPackVertexCount(&world->staticBodies.items[world->staticBodies.activeIndices[i]]);
  }
  if (world->staticBodies.length > *maxStat) *maxStat = world->staticBodies.length;
  if (verts > *maxVerts) *maxVerts = verts;
}

static inline uint32_t pxPackShape(const PxBody *b, float *verts4, uint32_t *cursor, uint32_t cap, int *err) {
  uint32_t count = pxPackVertexCount(b);
  if (count == 0) {
    return PX_SHAPE_PACK(0, 0);
  }
  if (*cursor + count > cap) {
    *err = 1;
    return PX_SHAPE_PACK(0, 0);
  }
  const PxPolygonData *p = &b->collider.shape.polygon;
  uint32_t off = *cursor;
  for (uint32_t i = 0; i < count; i++) {
    float *v = verts4 + (size_t)(off + i) * 4;
    v[0] = p->vertices.items[i].x;
    v[1] = p->vertices.items[i].y;
    v[2] = p->normals.items[i].x;
    v[3] = p->normals.items[i].y;
  }
  *cursor += count;
  return PX_SHAPE_PACK(off, count);
}

/* host -> blocks: every field pxWorldStepWithDt reads (reference src/px_world.c:410-420) */
static inline int pxPackOneWorld(const PxBatchLayout *L, const PxWorld *world, uint32_t w, uint8_t *mut, uint8_t *imm) {
  const uint32_t D = L->maxDynamic, S = L->maxStatic;
  memset(mut, 0, L->mutStride);
  memset(imm, 0, L->immStride);

  const PxBodyArray *dyn = &world->dynamicBodies;
  const PxBodyArray *stat = &world->staticBodies;
  if (stat->length > S) {
    PX_PACK_FAIL(-1, "world %u: %u static bodies exceed maxStatic %u", w, (unsigned)stat->length, S);
  }

  PxWorldHeader *h = (PxWorldHeader *)(mut + L->offHeader);
  h->gravityX = world->gravity.x;
  h->gravityY = world->gravity.y;
  h->dt = pxFastDiv(world->clock.fixedTimestep, PX_PACK_SUBSTEPS_PER_STEP); /* px_world.c:432 */
  h->linearSleepThresholdSq = world->linearSleepThresholdSq;
  h->angularSleepThreshold = world->angularSleepThreshold;
  h->timeToSleep = world->timeToSleep;
  h->iterations = world->iterations;
  h->nDynamic = dyn->length;
  h->nStatic = stat->length;

  float *dynF = (float *)(mut + L->offDyn);
  uint8_t *dynFlags = mut + L->offDynFlags;
  uint8_t *dynOrder = mut + L->offDynOrder;
  float *dynC = (float *)(imm + L->offDynConst);
  uint32_t *dynShape = (uint32_t *)(imm + L->offDynShape);
  float *statF = (float *)(imm + L->offStat);
  uint32_t *statShape = (uint32_t *)(imm + L->offStatShape);
  uint8_t *statSlot = imm + L->offStatSlot;
  float *verts4 = (float *)(imm + L->offVerts);
  uint32_t vcursor = 0;
  int verr = 0;

  for (uint32_t i = 0; i < dyn->length; i++) {
    uint32_t slot = dyn->activeIndices[i];
    if (slot >= D) {
      PX_PACK_FAIL(-1, "world %u: dynamic slot %u exceeds maxDynamic %u", w, slot, D);
    }
    const PxBody *b = &dyn->items[slot];
    dynOrder[i] = (uint8_t)slot;
    dynF[PX_DYN_PX * D + slot] = b->position.x;
    dynF[PX_DYN_PY * D + slot] = b->position.y;
    dynF[PX_DYN_VX * D + slot] = b->velocity.x;
    dynF[PX_DYN_VY * D + slot] = b->velocity.y;
    dynF[PX_DYN_AV * D + slot] = b->angularVelocity;
    dynF[PX_DYN_ANGLE * D + slot] = b->orientationAngle;
    dynF[PX_DYN_FX * D + slot] = b->force.x;
    dynF[PX_DYN_FY * D + slot] = b->force.y;
    dynF[PX_DYN_TORQUE * D + slot] = b->torque;
    dynF[PX_DYN_SLEEP * D + slot] = b->sleepTime;
    dynF[PX_DYN_M00 * D + slot] = b->orientation.m00;
    dynF[PX_DYN_M01 * D + slot] = b->orientation.m01;
    dynF[PX_DYN_M10 * D + slot] = b->orientation.m10;
    dynF[PX_DYN_M11 * D + slot] = b->orientation.m11;
    dynFlags[slot] = b->flags;
    dynC[PX_DYNC_IMASS * D + slot] = b->iMass;
    dynC[PX_DYNC_IINERTIA * D + slot] = b->iMomentOfInertia;
    dynC[PX_DYNC_E * D + slot] = b->restitution;
    dynC[PX_DYNC_SF * D + slot] = b->staticFriction;
    dynC[PX_DYNC_DF * D + slot] = b->dynamicFriction;
    dynC[PX_DYNC_RADIUS * D + slot] = pxPackRadius(b);
    dynShape[slot] = pxPackShape(b, verts4, &vcursor, L->maxVertices, &verr);
  }
  for (uint32_t i = 0; i < stat->length; i++) {
    uint32_t slot = stat->activeIndices[i];
    const PxBody *b = &stat->items[slot];
    statSlot[i] = (uint8_t)slot;
    statF[PX_STAT_PX * S + i] = b->position.x;
    statF[PX_STAT_PY * S + i] = b->position.y;
    statF[PX_STAT_M00 * S + i] = b->orientation.m00;
    statF[PX_STAT_M01 * S + i] = b->orientation.m01;
    statF[PX_STAT_M10 * S + i] = b->orientation.m10;
    statF[PX_STAT_M11 * S + i] = b->orientation.m11;
    statF[PX_STAT_E * S + i] = b->restitution;
    statF[PX_STAT_SF * S + i] = b->staticFriction;
    statF[PX_STAT_DF * S + i] = b->dynamicFriction;
    statF[PX_STAT_RADIUS * S + i] = pxPackRadius(b);
    statShape[i] = pxPackShape(b, verts4, &vcursor, L->maxVertices, &verr);
  }
  if (verr) {
    PX_PACK_FAIL(-1, "world %u: polygon vertices exceed maxVertices %u", w, L->maxVertices);
  }
  h->nVertices = vcursor;
  return 0;
}

/* blocks -> host: everything pxWorldStepWithDt mutates.  Returns 1 if the world overflowed a
 * device capacity during its last step, 0 if not, < 0 on error. */
static inline int pxUnpackOneWorld(const PxBatchLayout *L, PxWorld *world, uint32_t w, const uint8_t *mut, uint32_t flags) {
  const uint32_t D = L->maxDynamic;
  const PxWorldHeader *h = (const PxWorldHeader *)(mut + L->offHeader);
  world->contacts.length = (__typeof__(world->contacts.length))h->statManifolds;
#ifdef PX_PACK_WORLD_HAS_STATS
  world->contacts.sleeping = (uint16_t)h->statSleeping;
  world->contacts.overflow = h->statOverflow;
#endif
  const int overflowed = h->statOverflow ? 1 : 0;
  if (flags & PX_READBACK_STATS_ONLY) {
    return overflowed;
  }
  PxBodyArray *dyn = &world->dynamicBodies;
  if (h->nDynamic != dyn->length) {
    PX_PACK_FAIL(-1, "world %u: body count changed on the host since upload (%u vs %u)", w, (unsigned)dyn->length, h->nDynamic);
  }
  const float *dynF = (const float *)(mut + L->offDyn);
  const uint8_t *dynFlags = mut + L->offDynFlags;
  const uint8_t *dynOrder = mut + L->offDynOrder;
  for (uint32_t i = 0; i < dyn->length; i++) {
    uint32_t slot = dynOrder[i];
    dyn->activeIndices[i] = (uint8_t)slot;
    PxBody *b = &dyn->items[slot];
    b->position.x = dynF[PX_DYN_PX * D + slot];
    b->position.y = dynF[PX_DYN_PY * D + slot];
    b->velocity.x = dynF[PX_DYN_VX * D + slot];
    b->velocity.y = dynF[PX_DYN_VY * D + slot];
    b->angularVelocity = dynF[PX_DYN_AV * D + slot];
    b->orientationAngle = dynF[PX_DYN_ANGLE * D + slot];
    b->force.x = dynF[PX_DYN_FX * D + slot];
    b->force.y = dynF[PX_DYN_FY * D + slot];
    b->torque = dynF[PX_DYN_TORQUE * D + slot];
    b->sleepTime = dynF[PX_DYN_SLEEP * D + slot];
    b->orientation.m00 = dynF[PX_DYN_M00 * D + slot];
    b->orientation.m01 = dynF[PX_DYN_M01 * D + slot];
    b->orientation.m10 = dynF[PX_DYN_M10 * D + slot];
    b->orientation.m11 = dynF[PX_DYN_M11 * D + slot];
    b->flags = dynFlags[slot];
    /* the AABB always follows the position (pxBodyMoveBy, reference src/px_body.c:95-102) */
    float r = pxPackRadius(b);
    b->aabb.min = pxVec2Subf(b->position, r);
    b->aabb.max = pxVec2Addf(b->position, r);
  }
  return overflowed;
}

#endif /* PX_PACK_INL_H */
