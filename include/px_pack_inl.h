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
  float *statM = (float *)(mut + L->offStatMut);
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
    statM[PX_STATM_PX * S + i] = b->position.x;
    statM[PX_STATM_PY * S + i] = b->position.y;
    statM[PX_STATM_ANGLE * S + i] = b->orientationAngle;
    statM[PX_STATM_M00 * S + i] = b->orientation.m00;
    statM[PX_STATM_M01 * S + i] = b->orientation.m01;
    statM[PX_STATM_M10 * S + i] = b->orientation.m10;
    statM[PX_STATM_M11 * S + i] = b->orientation.m11;
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
  /* static poses change only through pose commands (pxBatchSetPosition ... pxBatchRotate); the
   * step never writes them */
  PxBodyArray *stat = &world->staticBodies;
  const uint32_t S = L->maxStatic;
  const float *statM = (const float *)(mut + L->offStatMut);
  for (uint32_t i = 0; i < stat->length && i < h->nStatic && i < S; i++) {
    PxBody *b = &stat->items[stat->activeIndices[i]];
    b->position.x = statM[PX_STATM_PX * S + i];
    b->position.y = statM[PX_STATM_PY * S + i];
    b->orientationAngle = statM[PX_STATM_ANGLE * S + i];
    b->orientation.m00 = statM[PX_STATM_M00 * S + i];
    b->orientation.m01 = statM[PX_STATM_M01 * S + i];
    b->orientation.m10 = statM[PX_STATM_M10 * S + i];
    b->orientation.m11 = statM[PX_STATM_M11 * S + i];
    float r = pxPackRadius(b);
    b->aabb.min = pxVec2Subf(b->position, r);
    b->aabb.max = pxVec2Addf(b->position, r);
  }
  return overflowed;
}

/* ---- body lifecycle on a resident world (pxBatchAddBody / pxBatchFreeBody) -------------------
 * The immutable block has a host mirror; a body added to a live world gets its constants, shape
 * word and vertices written there.  Vertex ranges are not reference-counted: at every add the
 * occupied ranges are those of the bodies that are live in the host world, so the ranges of freed
 * bodies are reclaimed implicitly. */
static inline void pxPackMarkRange(uint8_t *used, uint32_t cap, uint32_t shapeWord) {
  const uint32_t n = PX_SHAPE_COUNT(shapeWord), off = PX_SHAPE_OFFSET(shapeWord);
  for (uint32_t k = 0; k < n && off + k < cap; k++) used[off + k] = 1;
}

/* first fit of `count` vertices in the pool; returns the offset or UINT32_MAX */
static inline uint32_t pxPackFirstFit(const uint8_t *used, uint32_t cap, uint32_t count) {
  uint32_t run = 0;
  for (uint32_t i = 0; i < cap; i++) {
    run = used[i] ? 0 : run + 1;
    if (run == count) return i + 1 - count;
  }
  return UINT32_MAX;
}

static inline void pxPackWriteVerts(const PxBody *b, float *verts4, uint32_t off) {
  const PxPolygonData *p = &b->collider.shape.polygon;
  const uint32_t count = pxPackVertexCount(b);
  for (uint32_t i = 0; i < count; i++) {
    float *v = verts4 + (size_t)(off + i) * 4;
    v[0] = p->vertices.items[i].x;
    v[1] = p->vertices.items[i].y;
    v[2] = p->normals.items[i].x;
    v[3] = p->normals.items[i].y;
  }
}

/* vertex-pool occupancy of the live bodies of `world` as the mirror `imm` describes them;
 * `skip` (may be NULL) is left out, statics are included on request.  `used` has maxVertices bytes. */
static inline void pxPackOccupancy(const PxBatchLayout *L, const PxWorld *world, const uint8_t *imm, const PxBody *skip,
                                   int withStatics, uint8_t *used) {
  memset(used, 0, L->maxVertices);
  const uint32_t *dynShape = (const uint32_t *)(imm + L->offDynShape);
  const uint32_t *statShape = (const uint32_t *)(imm + L->offStatShape);
  const uint8_t *statSlot = imm + L->offStatSlot;
  for (uint32_t i = 0; i < world->dynamicBodies.length; i++) {
    const uint32_t slot = world->dynamicBodies.activeIndices[i];
    if (&world->dynamicBodies.items[slot] == skip || slot >= L->maxDynamic) continue;
    pxPackMarkRange(used, L->maxVertices, dynShape[slot]);
  }
  if (withStatics) {
    /* the mirror's static rows are by position; mark those whose slot is still live on the host */
    for (uint32_t q = 0; q < L->maxStatic; q++) {
      if (!PX_SHAPE_COUNT(statShape[q])) continue;
      for (uint32_t i = 0; i < world->staticBodies.length; i++) {
        if (world->staticBodies.activeIndices[i] == statSlot[q] && &world->staticBodies.items[statSlot[q]] != skip) {
          pxPackMarkRange(used, L->maxVertices, statShape[q]);
          break;
        }
      }
    }
  }
}

/* A dynamic body the host has just added to `world` (px->world->newDynamicBody: the reference's
 * slot-reuse rule already chose body->worldIndex, src/px_body_array.c:32-42): write its
 * constants, shape and vertices into the mirror `imm` and its mutable rows into `dynRows`
 * (PX_DYN_ROWS floats) / `flagsOut`.  `used` is scratch of maxVertices bytes. */
static inline int pxPackAddDynamicBody(const PxBatchLayout *L, const PxWorld *world, const PxBody *b, uint8_t *imm, float *dynRows,
                                       uint32_t *flagsOut, uint8_t *used) {
  const uint32_t D = L->maxDynamic, slot = b->worldIndex;
  if (slot >= D) {
    PX_PACK_FAIL(-1, "added body: slot %u exceeds maxDynamic %u", slot, D);
  }
  float *dynC = (float *)(imm + L->offDynConst);
  uint32_t *dynShape = (uint32_t *)(imm + L->offDynShape);
  const uint32_t count = pxPackVertexCount(b);
  uint32_t shape = PX_SHAPE_PACK(0, 0);
  if (count) {
    pxPackOccupancy(L, world, imm, b, 1, used);
    const uint32_t off = pxPackFirstFit(used, L->maxVertices, count);
    if (off == UINT32_MAX) {
      PX_PACK_FAIL(-1, "added body: no room for %u vertices in the world's pool of %u (raise PxBatchDesc.maxVertices)", count,
                   L->maxVertices);
    }
    pxPackWriteVerts(b, (float *)(imm + L->offVerts), off);
    shape = PX_SHAPE_PACK(off, count);
  }
  dynShape[slot] = shape;
  dynC[PX_DYNC_IMASS * D + slot] = b->iMass;
  dynC[PX_DYNC_IINERTIA * D + slot] = b->iMomentOfInertia;
  dynC[PX_DYNC_E * D + slot] = b->restitution;
  dynC[PX_DYNC_SF * D + slot] = b->staticFriction;
  dynC[PX_DYNC_DF * D + slot] = b->dynamicFriction;
  dynC[PX_DYNC_RADIUS * D + slot] = pxPackRadius(b);
  dynRows[PX_DYN_PX] = b->position.x;
  dynRows[PX_DYN_PY] = b->position.y;
  dynRows[PX_DYN_VX] = b->velocity.x;
  dynRows[PX_DYN_VY] = b->velocity.y;
  dynRows[PX_DYN_AV] = b->angularVelocity;
  dynRows[PX_DYN_ANGLE] = b->orientationAngle;
  dynRows[PX_DYN_FX] = b->force.x;
  dynRows[PX_DYN_FY] = b->force.y;
  dynRows[PX_DYN_TORQUE] = b->torque;
  dynRows[PX_DYN_SLEEP] = b->sleepTime;
  dynRows[PX_DYN_M00] = b->orientation.m00;
  dynRows[PX_DYN_M01] = b->orientation.m01;
  dynRows[PX_DYN_M10] = b->orientation.m10;
  dynRows[PX_DYN_M11] = b->orientation.m11;
  *flagsOut = b->flags;
  return 0;
}

/* The static bodies of `world` as they are on the host now (after px->world->newStaticBody,
 * which re-sorts them, src/px_world.c:94-101, or freeBody): rewrites the static rows of the
 * mirror `imm` and of the mutable image `mut` (header.nStatic and the pose rows only). */
static inline int pxPackStaticBodies(const PxBatchLayout *L, const PxWorld *world, uint8_t *mut, uint8_t *imm, uint8_t *used) {
  const uint32_t S = L->maxStatic;
  const PxBodyArray *stat = &world->staticBodies;
  if (stat->length > S) {
    PX_PACK_FAIL(-1, "%u static bodies exceed maxStatic %u", (unsigned)stat->length, S);
  }
  pxPackOccupancy(L, world, imm, NULL, 0, used); /* live dynamic polygons keep their ranges */
  float *statF = (float *)(imm + L->offStat);
  float *statM = (float *)(mut + L->offStatMut);
  uint32_t *statShape = (uint32_t *)(imm + L->offStatShape);
  uint8_t *statSlot = imm + L->offStatSlot;
  memset(statF, 0, PX_STAT_ROWS * S * 4);
  memset(statM, 0, PX_STATM_ROWS * S * 4);
  memset(statShape, 0, S * 4);
  memset(statSlot, 0, S);
  for (uint32_t i = 0; i < stat->length; i++) {
    const uint32_t slot = stat->activeIndices[i];
    const PxBody *b = &stat->items[slot];
    statSlot[i] = (uint8_t)slot;
    statM[PX_STATM_PX * S + i] = b->position.x;
    statM[PX_STATM_PY * S + i] = b->position.y;
    statM[PX_STATM_ANGLE * S + i] = b->orientationAngle;
    statM[PX_STATM_M00 * S + i] = b->orientation.m00;
    statM[PX_STATM_M01 * S + i] = b->orientation.m01;
    statM[PX_STATM_M10 * S + i] = b->orientation.m10;
    statM[PX_STATM_M11 * S + i] = b->orientation.m11;
    statF[PX_STAT_E * S + i] = b->restitution;
    statF[PX_STAT_SF * S + i] = b->staticFriction;
    statF[PX_STAT_DF * S + i] = b->dynamicFriction;
    statF[PX_STAT_RADIUS * S + i] = pxPackRadius(b);
    const uint32_t count = pxPackVertexCount(b);
    if (count) {
      const uint32_t off = pxPackFirstFit(used, L->maxVertices, count);
      if (off == UINT32_MAX) {
        PX_PACK_FAIL(-1, "static body: no room for %u vertices in the world's pool of %u", count, L->maxVertices);
      }
      pxPackWriteVerts(b, (float *)(imm + L->offVerts), off);
      for (uint32_t k = 0; k < count; k++) used[off + k] = 1;
      statShape[i] = PX_SHAPE_PACK(off, count);
    }
  }
  ((PxWorldHeader *)(mut + L->offHeader))->nStatic = stat->length;
  return 0;
}

#endif /* PX_PACK_INL_H */
