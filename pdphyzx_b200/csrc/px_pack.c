/*
 * px_pack.c -- host-only half of the batch layer: layout arithmetic and the AoS <-> SoA
 * conversion between host PxWorld structs (include/pdphyzx.h) and the per-world blocks the
 * device consumes (include/px_batch.h, PxBatchLayout).  Plain C: pdphyzx.h uses `new` as a
 * member name and cannot be read by a C++ compiler.
 *
 * Field provenance (reference): PxBody src/includes/px_body.h:82-120, PxBodyArray
 * src/includes/px_body_array.h:19-33, PxWorld src/includes/px_world.h:31-44,
 * PxPolygonData src/includes/px_collider.h:46-50.
 */
#include <stdio.h>
#include <string.h>

#include "pdphyzx.h"
#include "px_batch.h"
#include "px_internal.h"

static uint32_t alignUp(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

int pxBatchComputeLayout(uint32_t nWorlds, const PxBatchDesc *desc, PxBatchLayout *L) {
  if (!desc || !L) {
    return pxSetError(-1, "pxBatchComputeLayout: NULL argument");
  }
  if (!(desc->unit == 1.0f || desc->unit == 100.0f || desc->unit == 1000.0f)) {
    return pxSetError(-1, "PxBatchDesc.unit must be 1, 100 or 1000 (got %g)", (double)desc->unit);
  }
  /* 256 is accepted so that a padded capacity read back from a layout can be fed in again */
  if (desc->maxDynamic == 0 || desc->maxDynamic > 256 || desc->maxStatic > 256) {
    return pxSetError(-1, "PxBatchDesc: maxDynamic must be 1..256 and maxStatic 0..256 (got %u, %u)",
                      desc->maxDynamic, desc->maxStatic);
  }
  memset(L, 0, sizeof(*L));
  L->nWorlds = nWorlds;
  L->maxDynamic = alignUp(desc->maxDynamic, 16); /* byte rows are copied 16 bytes at a time */
  L->maxStatic = alignUp(desc->maxStatic ? desc->maxStatic : 1, 4);
  L->maxVertices = alignUp(desc->maxVertices ? desc->maxVertices : 4, 4);
  /* defaults sized from dense piles (about 2.5 manifolds and 5.5 AABB-pass pairs per body);
   * a world that needs more sets its PX_BATCH_OVERFLOW_* flag and the caller raises these */
  L->maxManifolds = desc->maxManifolds ? desc->maxManifolds : (11 * L->maxDynamic) / 4;
  L->maxPairs = desc->maxPairs ? desc->maxPairs : 7 * L->maxDynamic;
  L->maxManifolds = alignUp(L->maxManifolds, 32);
  L->maxPairs = alignUp(L->maxPairs, 32);
  if (L->maxVertices >= (1u << 24) || L->maxPairs > 32767 || L->maxManifolds > 1024) {
    return pxSetError(-1, "PxBatchDesc: capacity too large");
  }
  /* the sweep counts pairs in packed 16-bit fields: all-pairs worst case must fit */
  if (L->maxDynamic * (L->maxDynamic - 1) / 2 + L->maxDynamic * L->maxStatic > 65535) {
    return pxSetError(-1, "PxBatchDesc: %u dynamic x %u static bodies exceed the 65535 candidate-pair bound",
                      L->maxDynamic, L->maxStatic);
  }
  const uint32_t D = L->maxDynamic, S = L->maxStatic, V = L->maxVertices;

  uint32_t o = 0;
  L->offHeader = o;
  o += alignUp((uint32_t)sizeof(PxWorldHeader), 16);
  L->offDyn = o;
  o += PX_DYN_ROWS * D * 4;
  L->offDynFlags = o;
  o += alignUp(D, 16);
  L->offDynOrder = o;
  o += alignUp(D, 16);
  L->mutStride = alignUp(o, 128);

  o = 0;
  L->offDynConst = o;
  o += PX_DYNC_ROWS * D * 4;
  L->offDynShape = o;
  o += D * 4;
  L->offStat = o;
  o += PX_STAT_ROWS * S * 4;
  L->offStatShape = o;
  o += S * 4;
  L->offStatSlot = o;
  o += alignUp(S, 16);
  L->offVerts = o;
  o += V * 16;
  L->immStride = alignUp(o, 128);

  o = 0;
  L->offTracePairs = o;
  o += alignUp(2 * L->maxPairs * 2, 16);
  L->offTraceManifolds = o;
  o += (PX_TRACE_M_F32 + 1) * L->maxManifolds * 4;
  L->traceStride = alignUp(o, 128);

  /* unit-derived constants with the reference's own expressions and association:
   * PX_EPSILON = 0.0001f * U (px_math.h:23); PX_EPSILON * PX_EPSILON expands to
   * 0.0001f * U * 0.0001f * U (px_polygon.c:329); PX_ALLOWED_PENETRATION (px_manifold.h:15) */
  volatile float U = desc->unit;
  L->unit = U;
  L->epsilon = 0.0001f * U;
  L->epsilonSq = 0.0001f * U * 0.0001f * U;
  L->allowedPenetration = 0.0001f * U;
  return 0;
}

static uint32_t bodyVertexCount(const PxBody *b) {
  return b->collider.type == PX_POLYGON ? b->collider.shape.polygon.vertices.length : 0;
}

int pxBatchFitDesc(struct PxWorld *const *worlds, uint32_t n, PxBatchDesc *desc) {
  if (!worlds || !desc) {
    return pxSetError(-1, "pxBatchFitDesc: NULL argument");
  }
  uint32_t maxDyn = 1, maxStat = 1, maxVerts = 4;
  for (uint32_t w = 0; w < n; w++) {
    const PxWorld *world = worlds[w];
    if (!world) {
      return pxSetError(-1, "pxBatchFitDesc: world %u is NULL", w);
    }
    uint32_t verts = 0;
    for (uint32_t i = 0; i < world->dynamicBodies.length; i++) {
      uint32_t slot = world->dynamicBodies.activeIndices[i];
      if (slot + 1 > maxDyn) maxDyn = slot + 1;
      verts += bodyVertexCount(&world->dynamicBodies.items[slot]);
    }
    for (uint32_t i = 0; i < world->staticBodies.length; i++) {
      verts += bodyVertexCount(&world->staticBodies.items[world->staticBodies.activeIndices[i]]);
    }
    if (world->staticBodies.length > maxStat) maxStat = world->staticBodies.length;
    if (verts > maxVerts) maxVerts = verts;
  }
  if (!desc->maxDynamic) desc->maxDynamic = maxDyn;
  if (!desc->maxStatic) desc->maxStatic = maxStat;
  if (!desc->maxVertices) desc->maxVertices = maxVerts;
  if (desc->maxDynamic < maxDyn || desc->maxStatic < maxStat || desc->maxVertices < maxVerts) {
    return pxSetError(-1, "PxBatchDesc capacities (%u dyn, %u static, %u verts) too small for the worlds (%u, %u, %u)",
                      desc->maxDynamic, desc->maxStatic, desc->maxVertices, maxDyn, maxStat, maxVerts);
  }
  return 0;
}

static uint32_t packShape(const PxBody *b, float *verts4, uint32_t *cursor, uint32_t cap, int *err) {
  uint32_t count = bodyVertexCount(b);
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

static float bodyRadius(const PxBody *b) {
  /* the AABB half-extent: circle radius or the polygon's pre-recentring maxRadius
   * (px_circle.c:43-48, px_polygon.c:261-266) */
  return b->collider.type == PX_CIRCLE ? b->collider.shape.circle.radius : b->collider.shape.polygon.maxRadius;
}

int pxPackWorlds(const PxBatchLayout *L, struct PxWorld *const *worlds, uint32_t n, void *mutBase, void *immBase) {
  if (!L || !worlds || !mutBase || !immBase) {
    return pxSetError(-1, "pxPackWorlds: NULL argument");
  }
  const uint32_t D = L->maxDynamic, S = L->maxStatic;
  for (uint32_t w = 0; w < n; w++) {
    const PxWorld *world = worlds[w];
    if (!world) {
      return pxSetError(-1, "pxPackWorlds: world %u is NULL", w);
    }
    uint8_t *mut = (uint8_t *)mutBase + (size_t)w * L->mutStride;
    uint8_t *imm = (uint8_t *)immBase + (size_t)w * L->immStride;
    memset(mut, 0, L->mutStride);
    memset(imm, 0, L->immStride);

    const PxBodyArray *dyn = &world->dynamicBodies;
    const PxBodyArray *stat = &world->staticBodies;
    if (stat->length > S) {
      return pxSetError(-1, "world %u: %u static bodies exceed maxStatic %u", w, stat->length, S);
    }

    PxWorldHeader *h = (PxWorldHeader *)(mut + L->offHeader);
    h->gravityX = world->gravity.x;
    h->gravityY = world->gravity.y;
    h->dt = pxFastDiv(world->clock.fixedTimestep, PX_SUBSTEPS_PER_STEP); /* px_world.c:432 */
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
        return pxSetError(-1, "world %u: dynamic slot %u exceeds maxDynamic %u", w, slot, D);
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
      dynC[PX_DYNC_RADIUS * D + slot] = bodyRadius(b);
      dynShape[slot] = packShape(b, verts4, &vcursor, L->maxVertices, &verr);
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
      statF[PX_STAT_RADIUS * S + i] = bodyRadius(b);
      statShape[i] = packShape(b, verts4, &vcursor, L->maxVertices, &verr);
    }
    if (verr) {
      return pxSetError(-1, "world %u: polygon vertices exceed maxVertices %u", w, L->maxVertices);
    }
    h->nVertices = vcursor;
  }
  return 0;
}

int pxUnpackWorlds(const PxBatchLayout *L, struct PxWorld *const *worlds, uint32_t n, const void *mutBase, uint32_t flags) {
  if (!L || !worlds || !mutBase) {
    return pxSetError(-1, "pxUnpackWorlds: NULL argument");
  }
  const uint32_t D = L->maxDynamic;
  int overflowed = 0;
  for (uint32_t w = 0; w < n; w++) {
    PxWorld *world = worlds[w];
    if (!world) {
      return pxSetError(-1, "pxUnpackWorlds: world %u is NULL", w);
    }
    const uint8_t *mut = (const uint8_t *)mutBase + (size_t)w * L->mutStride;
    const PxWorldHeader *h = (const PxWorldHeader *)(mut + L->offHeader);
    world->contacts.length = (uint16_t)h->statManifolds;
    world->contacts.sleeping = (uint16_t)h->statSleeping;
    world->contacts.overflow = h->statOverflow;
    if (h->statOverflow) {
      overflowed++;
    }
    if (flags & PX_READBACK_STATS_ONLY) {
      continue;
    }
    PxBodyArray *dyn = &world->dynamicBodies;
    if (h->nDynamic != dyn->length) {
      return pxSetError(-1, "world %u: body count changed on the host since upload (%u vs %u)", w, dyn->length, h->nDynamic);
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
      /* the AABB always follows the position (pxBodyMoveBy, px_body.c:95-102) */
      float r = bodyRadius(b);
      b->aabb.min = pxVec2Subf(b->position, r);
      b->aabb.max = pxVec2Addf(b->position, r);
    }
  }
  return overflowed;
}
