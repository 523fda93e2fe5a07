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

/* the per-world AoS <-> SoA conversion is shared, by field name, with the reference-side
 * binding (integration/px_world_b200.c): include/px_pack_inl.h */
#define PX_PACK_FAIL(rc, ...) return pxSetError((rc), __VA_ARGS__)
#define PX_PACK_WORLD_HAS_STATS 1
#include "px_pack_inl.h"

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
  /* defaults sized from dense piles: a settled 252-circle pile peaks at 705 manifolds (SURVEY 0.2,
   * 6), a mixed one at about 660, so 23/8 = 2.875 per body slot (736 at 256 slots; the ready queue
   * of the solver stays at 1024 entries up to exactly that) and 7 AABB-pass pairs per slot;
   * a world that needs more sets its PX_BATCH_OVERFLOW_* flag and the caller raises these */
  L->maxManifolds = desc->maxManifolds ? desc->maxManifolds : (23 * L->maxDynamic) / 8;
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
  L->offStatMut = o;
  o += PX_STATM_ROWS * S * 4;
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

int pxBatchFitDesc(struct PxWorld *const *worlds, uint32_t n, PxBatchDesc *desc) {
  if (!worlds || !desc) {
    return pxSetError(-1, "pxBatchFitDesc: NULL argument");
  }
  uint32_t maxDyn = 1, maxStat = 1, maxVerts = 4;
  for (uint32_t w = 0; w < n; w++) {
    if (!worlds[w]) {
      return pxSetError(-1, "pxBatchFitDesc: world %u is NULL", w);
    }
    pxPackWorldNeeds(worlds[w], &maxDyn, &maxStat, &maxVerts);
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

/* ---- pxPackWorlds / pxUnpackWorlds: worlds are independent, so large batches are converted by
 * several host threads (PX_PACK_THREADS, default = online cores, at most one per 64 worlds) */
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

typedef struct {
  const PxBatchLayout *L;
  struct PxWorld *const *worlds;
  uint32_t first, last;
  uint8_t *mut, *imm;
  const uint8_t *mutIn;
  uint32_t flags;
  int rc, overflowed;
} PxPackJob;

static void *packJob(void *arg) {
  PxPackJob *j = (PxPackJob *)arg;
  for (uint32_t w = j->first; w < j->last && j->rc == 0; w++) {
    j->rc = pxPackOneWorld(j->L, j->worlds[w], w, j->mut + (size_t)w * j->L->mutStride, j->imm + (size_t)w * j->L->immStride);
  }
  return NULL;
}

static void *unpackJob(void *arg) {
  PxPackJob *j = (PxPackJob *)arg;
  for (uint32_t w = j->first; w < j->last; w++) {
    int rc = pxUnpackOneWorld(j->L, j->worlds[w], w, j->mutIn + (size_t)w * j->L->mutStride, j->flags);
    if (rc < 0) {
      j->rc = rc;
      return NULL;
    }
    j->overflowed += rc;
  }
  return NULL;
}

static uint32_t packThreads(uint32_t n) {
  const char *env = getenv("PX_PACK_THREADS");
  long t = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
  if (t < 1) t = 1;
  if (t > 64) t = 64;
  if ((uint32_t)t > (n + 63) / 64) t = (long)((n + 63) / 64);
  return (uint32_t)(t < 1 ? 1 : t);
}

static int runJobs(PxPackJob *proto, uint32_t n, void *(*fn)(void *), int *overflowed) {
  const uint32_t T = packThreads(n);
  PxPackJob jobs[64];
  pthread_t tid[64];
  for (uint32_t t = 0; t < T; t++) {
    jobs[t] = *proto;
    jobs[t].first = (uint32_t)((uint64_t)n * t / T);
    jobs[t].last = (uint32_t)((uint64_t)n * (t + 1) / T);
    jobs[t].rc = 0;
    jobs[t].overflowed = 0;
  }
  uint32_t started = 0;
  for (uint32_t t = 1; t < T; t++) {
    if (pthread_create(&tid[t], NULL, fn, &jobs[t]) != 0) break;
    started = t;
  }
  fn(&jobs[0]);
  for (uint32_t t = started + 1; t < T; t++) fn(&jobs[t]); /* threads that could not be started run here */
  for (uint32_t t = 1; t <= started; t++) pthread_join(tid[t], NULL);
  int total = 0;
  for (uint32_t t = 0; t < T; t++) {
    if (jobs[t].rc != 0) return jobs[t].rc;
    total += jobs[t].overflowed;
  }
  if (overflowed) *overflowed = total;
  return 0;
}

int pxPackWorlds(const PxBatchLayout *L, struct PxWorld *const *worlds, uint32_t n, void *mutBase, void *immBase) {
  if (!L || !worlds || !mutBase || !immBase) {
    return pxSetError(-1, "pxPackWorlds: NULL argument");
  }
  for (uint32_t w = 0; w < n; w++) {
    if (!worlds[w]) {
      return pxSetError(-1, "pxPackWorlds: world %u is NULL", w);
    }
  }
  PxPackJob proto;
  memset(&proto, 0, sizeof(proto));
  proto.L = L;
  proto.worlds = worlds;
  proto.mut = (uint8_t *)mutBase;
  proto.imm = (uint8_t *)immBase;
  return runJobs(&proto, n, packJob, NULL);
}

int pxUnpackWorlds(const PxBatchLayout *L, struct PxWorld *const *worlds, uint32_t n, const void *mutBase, uint32_t flags) {
  if (!L || !worlds || !mutBase) {
    return pxSetError(-1, "pxUnpackWorlds: NULL argument");
  }
  for (uint32_t w = 0; w < n; w++) {
    if (!worlds[w]) {
      return pxSetError(-1, "pxUnpackWorlds: world %u is NULL", w);
    }
  }
  PxPackJob proto;
  memset(&proto, 0, sizeof(proto));
  proto.L = L;
  proto.worlds = worlds;
  proto.mutIn = (const uint8_t *)mutBase;
  proto.flags = flags;
  int overflowed = 0;
  int rc = runJobs(&proto, n, unpackJob, &overflowed);
  return rc != 0 ? rc : overflowed;
}

/* ---- body lifecycle on a resident world: host half (see px_pack_inl.h) */
int pxPackBodyIsDynamic(const void *body) { return body && ((const PxBody *)body)->density > 0; /* the reference's own test, px_world.c:116-117 */ }

int pxPackAddDynamic(const PxBatchLayout *L, const struct PxWorld *world, const void *bodyPtr, void *immWorld, float *dynRows,
                     uint32_t *flags, uint32_t *slot) {
  const PxBody *body = (const PxBody *)bodyPtr;
  if (!L || !world || !body || !immWorld || !dynRows || !flags || !slot) {
    return pxSetError(-1, "pxPackAddDynamic: NULL argument");
  }
  uint8_t *used = (uint8_t *)malloc(L->maxVertices ? L->maxVertices : 1);
  if (!used) return pxSetError(-1, "pxPackAddDynamic: out of memory");
  *slot = body->worldIndex;
  int rc = pxPackAddDynamicBody(L, world, body, (uint8_t *)immWorld, dynRows, flags, used);
  free(used);
  return rc;
}

int pxPackStatics(const PxBatchLayout *L, const struct PxWorld *world, void *mutWorld, void *immWorld) {
  if (!L || !world || !mutWorld || !immWorld) {
    return pxSetError(-1, "pxPackStatics: NULL argument");
  }
  uint8_t *used = (uint8_t *)malloc(L->maxVertices ? L->maxVertices : 1);
  if (!used) return pxSetError(-1, "pxPackStatics: out of memory");
  int rc = pxPackStaticBodies(L, world, (uint8_t *)mutWorld, (uint8_t *)immWorld, used);
  free(used);
  return rc;
}
