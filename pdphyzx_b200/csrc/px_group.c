/*
 * px_group.c -- PxBatchGroup (include/px_batch.h): one batch of worlds sharded over the GPUs of
 * a box, driven from one host thread.  Worlds are independent, so this layer is pure
 * fan-out over per-device PxBatch objects: no collective, no peer traffic, nothing here touches
 * CUDA directly.  Every per-device call is asynchronous, so the devices run side by side.
 */
#include <stdlib.h>
#include <string.h>

#include "px_batch.h"
#include "px_internal.h"

struct PxBatchGroup {
  uint32_t nShards, nWorlds, perShard;
  PxBatch **shard;
  uint32_t *first, *count;
};

void pxBatchGroupFree(PxBatchGroup *g) {
  if (!g) return;
  for (uint32_t s = 0; s < g->nShards; s++) pxBatchFree(g->shard[s]);
  free(g->shard);
  free(g->first);
  free(g->count);
  free(g);
}

static PxBatchGroup *groupMake(struct PxWorld *const *worlds, uint32_t nWorlds, const PxBatchDesc *desc, const int32_t *devices,
                               uint32_t nDevices) {
  if (!desc || nWorlds == 0 || nDevices == 0) {
    pxSetError(-1, "pxBatchGroup: need a descriptor, worlds and at least one device");
    return NULL;
  }
  if (desc->stream) {
    pxSetError(-1, "pxBatchGroup: PxBatchDesc.stream must be NULL (every device gets its own stream)");
    return NULL;
  }
  if (nDevices > nWorlds) nDevices = nWorlds;
  PxBatchGroup *g = (PxBatchGroup *)calloc(1, sizeof(*g));
  if (!g) return NULL;
  g->nWorlds = nWorlds;
  g->perShard = (nWorlds + nDevices - 1) / nDevices; /* contiguous ranges of ceil(N / G) worlds */
  g->nShards = (nWorlds + g->perShard - 1) / g->perShard;
  g->shard = (PxBatch **)calloc(g->nShards, sizeof(PxBatch *));
  g->first = (uint32_t *)calloc(g->nShards, sizeof(uint32_t));
  g->count = (uint32_t *)calloc(g->nShards, sizeof(uint32_t));
  if (!g->shard || !g->first || !g->count) {
    pxBatchGroupFree(g);
    pxSetError(-1, "pxBatchGroup: out of memory");
    return NULL;
  }
  PxBatchDesc d = *desc;
  if (worlds && pxBatchFitDesc(worlds, nWorlds, &d) != 0) { /* one set of capacities for all shards */
    pxBatchGroupFree(g);
    return NULL;
  }
  for (uint32_t s = 0; s < g->nShards; s++) {
    g->first[s] = s * g->perShard;
    g->count[s] = nWorlds - g->first[s] < g->perShard ? nWorlds - g->first[s] : g->perShard;
    d.device = devices ? devices[s] : (int32_t)s;
    g->shard[s] = worlds ? pxBatchNew(worlds + g->first[s], g->count[s], &d) : pxBatchCreate(g->count[s], &d);
    if (!g->shard[s]) {
      pxBatchGroupFree(g);
      return NULL;
    }
  }
  return g;
}

PxBatchGroup *pxBatchGroupNew(struct PxWorld *const *worlds, uint32_t nWorlds, const PxBatchDesc *desc, const int32_t *devices,
                              uint32_t nDevices) {
  if (!worlds) {
    pxSetError(-1, "pxBatchGroupNew: NULL worlds");
    return NULL;
  }
  return groupMake(worlds, nWorlds, desc, devices, nDevices);
}

PxBatchGroup *pxBatchGroupCreate(uint32_t nWorlds, const PxBatchDesc *desc, const int32_t *devices, uint32_t nDevices) {
  return groupMake(NULL, nWorlds, desc, devices, nDevices);
}

uint32_t pxBatchGroupShards(const PxBatchGroup *g) { return g ? g->nShards : 0; }

PxBatch *pxBatchGroupShard(const PxBatchGroup *g, uint32_t shard, uint32_t *firstWorld, uint32_t *nWorlds) {
  if (!g || shard >= g->nShards) return NULL;
  if (firstWorld) *firstWorld = g->first[shard];
  if (nWorlds) *nWorlds = g->count[shard];
  return g->shard[shard];
}

PxBatch *pxBatchGroupLocate(const PxBatchGroup *g, uint32_t world, uint32_t *localWorld) {
  if (!g || world >= g->nWorlds) {
    pxSetError(-1, "pxBatchGroupLocate: world %u out of range", world);
    return NULL;
  }
  const uint32_t s = world / g->perShard;
  if (localWorld) *localWorld = world - g->first[s];
  return g->shard[s];
}

/* the part of [first, first + n) that shard s owns: local first world and count */
static uint32_t overlap(const PxBatchGroup *g, uint32_t s, uint32_t first, uint32_t n, uint32_t *localFirst, uint32_t *srcOffset) {
  const uint32_t lo = first > g->first[s] ? first : g->first[s];
  const uint32_t endA = first + n, endB = g->first[s] + g->count[s];
  const uint32_t hi = endA < endB ? endA : endB;
  if (hi <= lo) return 0;
  *localFirst = lo - g->first[s];
  *srcOffset = lo - first;
  return hi - lo;
}

static int checkRange(const PxBatchGroup *g, uint32_t first, uint32_t n, const char *who) {
  if (!g) return pxSetError(-1, "%s: NULL group", who);
  if (first > g->nWorlds || n > g->nWorlds - first) return pxSetError(-1, "%s: worlds [%u, %u) out of range (%u)", who, first, first + n, g->nWorlds);
  return 0;
}

int pxBatchGroupUpload(PxBatchGroup *g, struct PxWorld *const *worlds, uint32_t first, uint32_t n) {
  if (checkRange(g, first, n, "pxBatchGroupUpload")) return -1;
  for (uint32_t s = 0; s < g->nShards; s++) {
    uint32_t lf = 0, off = 0, c = overlap(g, s, first, n, &lf, &off);
    if (c && pxBatchUpload(g->shard[s], worlds + off, lf, c) != 0) return -1;
  }
  return 0;
}

int pxBatchGroupStep(PxBatchGroup *g, uint32_t kSubsteps) {
  if (!g) return pxSetError(-1, "pxBatchGroupStep: NULL group");
  for (uint32_t s = 0; s < g->nShards; s++) {
    int rc = pxBatchStep(g->shard[s], kSubsteps); /* asynchronous: the next device starts right away */
    if (rc != 0) return rc;
  }
  return 0;
}

int pxBatchGroupStepHost(PxBatchGroup *g, uint32_t kSubsteps, uint32_t chunks) {
  if (!g) return pxSetError(-1, "pxBatchGroupStepHost: NULL group");
  for (uint32_t s = 0; s < g->nShards; s++) {
    int rc = pxBatchStepHost(g->shard[s], kSubsteps, NULL, chunks);
    if (rc != 0) return rc;
  }
  return 0;
}

int pxBatchGroupStepTimed(PxBatchGroup *g, uint32_t kSubsteps, float *ms) {
  if (!g) return pxSetError(-1, "pxBatchGroupStepTimed: NULL group");
  for (uint32_t s = 0; s < g->nShards; s++) {
    if (pxBatchSync(g->shard[s]) != 0) return -1;
  }
  for (uint32_t s = 0; s < g->nShards; s++) {
    if (pxBatchTimerStart(g->shard[s]) != 0) return -1;
    int rc = pxBatchStep(g->shard[s], kSubsteps);
    if (rc != 0) return rc;
  }
  float worst = 0.0f;
  for (uint32_t s = 0; s < g->nShards; s++) {
    float t = 0.0f;
    if (pxBatchTimerStop(g->shard[s], &t) != 0) return -1;
    if (t > worst) worst = t;
  }
  if (ms) *ms = worst;
  return 0;
}

int pxBatchGroupSync(PxBatchGroup *g) {
  if (!g) return pxSetError(-1, "pxBatchGroupSync: NULL group");
  for (uint32_t s = 0; s < g->nShards; s++) {
    if (pxBatchSync(g->shard[s]) != 0) return -1;
  }
  return 0;
}

int pxBatchGroupReadback(PxBatchGroup *g, struct PxWorld *const *worlds, uint32_t first, uint32_t n, uint32_t flags) {
  if (checkRange(g, first, n, "pxBatchGroupReadback")) return -1;
  /* all device -> host copies are queued before the first one is waited for */
  for (uint32_t s = 0; s < g->nShards; s++) {
    uint32_t lf = 0, off = 0, c = overlap(g, s, first, n, &lf, &off);
    if (c && pxBatchMutableD2H(g->shard[s], NULL, lf, c) != 0) return -1;
  }
  int overflowed = 0;
  for (uint32_t s = 0; s < g->nShards; s++) {
    uint32_t lf = 0, off = 0, c = overlap(g, s, first, n, &lf, &off);
    if (!c) continue;
    if (pxBatchSync(g->shard[s]) != 0) return -1;
    PxBatchLayout L;
    if (pxBatchGetLayout(g->shard[s], &L) != 0) return -1;
    int rc = pxUnpackWorlds(&L, worlds + off, c, (const uint8_t *)pxBatchHostMutable(g->shard[s]) + (size_t)lf * L.mutStride, flags);
    if (rc < 0) return rc;
    overflowed += rc;
  }
  return overflowed;
}

int pxBatchGroupSetGravity(PxBatchGroup *g, uint32_t world, float x, float y) {
  if (!g) return pxSetError(-1, "pxBatchGroupSetGravity: NULL group");
  if (world == PX_ALL_WORLDS) {
    for (uint32_t s = 0; s < g->nShards; s++) {
      if (pxBatchSetGravity(g->shard[s], PX_ALL_WORLDS, x, y) != 0) return -1;
    }
    return 0;
  }
  uint32_t local = 0;
  PxBatch *b = pxBatchGroupLocate(g, world, &local);
  return b ? pxBatchSetGravity(b, local, x, y) : -1;
}

int pxBatchGroupSetIterations(PxBatchGroup *g, uint32_t world, uint32_t iterations) {
  if (!g) return pxSetError(-1, "pxBatchGroupSetIterations: NULL group");
  if (world == PX_ALL_WORLDS) {
    for (uint32_t s = 0; s < g->nShards; s++) {
      if (pxBatchSetIterations(g->shard[s], PX_ALL_WORLDS, iterations) != 0) return -1;
    }
    return 0;
  }
  uint32_t local = 0;
  PxBatch *b = pxBatchGroupLocate(g, world, &local);
  return b ? pxBatchSetIterations(b, local, iterations) : -1;
}

int pxBatchGroupApplyImpulse(PxBatchGroup *g, uint32_t world, uint32_t body, float ix, float iy, float cx, float cy) {
  uint32_t local = 0;
  PxBatch *b = pxBatchGroupLocate(g, world, &local);
  return b ? pxBatchApplyImpulse(b, local, body, ix, iy, cx, cy) : -1;
}
