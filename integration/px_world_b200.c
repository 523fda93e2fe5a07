/*
 * px_world_b200.c -- THE REFERENCE-SIDE BINDING: what a pdphyzx maintainer adds to the
 * reference tree (as src/px_world_b200.c) to run px->world->step on a B200 through
 * libpdphyzx_b200.so.  See INTEGRATION.md.
 *
 * It is compiled inside the reference's single translation unit, AFTER px_world.c, against
 * the reference's OWN headers (PxWorld / PxBody as the reference lays them out); the only
 * files of this repo it sees are include/px_batch.h (the C ABI) and include/px_pack_inl.h
 * (field-name based AoS <-> SoA packing).  It replaces exactly one reference function:
 *
 *     bool pxWorldStep(PxWorld *world)                    reference src/px_world.c:422-445
 *
 * keeping the clock loop (pxClockBeginFrame / ShouldStep / Advance, src/px_clock.c:26-52,
 * <= maxStepsPerFrame fixed steps per call, first call never steps) on the host and moving
 * the 5 x pxWorldStepWithDt per fired step (src/px_world.c:427-437) to the device.
 *
 * The one-line change in the reference: build px_world.c with
 *     -DpxWorldStep=pxWorldStepCpu
 * (or wrap its definition in #ifndef PDPHYZX_B200) so that px_world_api.c:29 binds this
 * definition into PxWorldAPI.step.  oracle/Makefile builds exactly that variant
 * (oracle/_ref/libpxref_b200_*.so) and tests/test_gpu_control.py steps reference-built worlds
 * through it, comparing with the reference's CPU step bit for bit.
 */
#include <stdlib.h>

#include "px_batch.h"
#include "px_pack_inl.h"

/* The reference's PxWorld has no spare field, so device batches are kept in a side table
 * (a maintainer would rather add `void *batch;` to PxWorld). */
#define PX_B200_MAX_WORLDS 64
static struct {
  PxWorld *world;
  PxBatch *batch;
  PxBatchLayout layout;
} pxB200Table[PX_B200_MAX_WORLDS];

static int pxB200Find(PxWorld *world, int create) {
  int freeAt = -1;
  for (int i = 0; i < PX_B200_MAX_WORLDS; i++) {
    if (pxB200Table[i].world == world) return i;
    if (!pxB200Table[i].world && freeAt < 0) freeAt = i;
  }
  if (create && freeAt >= 0) {
    pxB200Table[freeAt].world = world;
    pxB200Table[freeAt].batch = NULL;
  }
  return create ? freeAt : -1;
}

/* call from pxWorldFree (reference src/px_world.c:49-55) */
void pxWorldB200Release(PxWorld *world) {
  int i = pxB200Find(world, 0);
  if (i >= 0) {
    pxBatchFree(pxB200Table[i].batch);
    pxB200Table[i].world = NULL;
    pxB200Table[i].batch = NULL;
  }
}

static PxBatch *pxB200Batch(PxWorld *world, PxBatchLayout **layout) {
  int i = pxB200Find(world, 1);
  if (i < 0) {
    return NULL;
  }
  uint32_t needDyn = 1, needStat = 1, needVerts = 4;
  pxPackWorldNeeds(world, &needDyn, &needStat, &needVerts);
  PxBatchLayout *L = &pxB200Table[i].layout;
  if (pxB200Table[i].batch && (needDyn > L->maxDynamic || needStat > L->maxStatic || needVerts > L->maxVertices)) {
    pxBatchFree(pxB200Table[i].batch); /* bodies were added since: grow */
    pxB200Table[i].batch = NULL;
  }
  if (!pxB200Table[i].batch) {
    PxBatchDesc desc = pxBatchDescDefault(PDPHYZX_UNIT);
    /* headroom so that newDynamicBody between steps rarely reallocates */
    desc.maxDynamic = needDyn + 16 > 255 ? 255 : needDyn + 16;
    desc.maxStatic = needStat + 4;
    desc.maxVertices = needVerts + 64;
    desc.sinTable = pxSinTable; /* the host's table (src/includes/px_math.h:206-218) */
    pxB200Table[i].batch = pxBatchCreate(1, &desc);
    if (!pxB200Table[i].batch || pxBatchGetLayout(pxB200Table[i].batch, L) != 0) {
      return NULL;
    }
  }
  *layout = L;
  return pxB200Table[i].batch;
}

bool pxWorldStep(PxWorld *world) {
  PxClock *clock = &world->clock;
  uint32_t fired = 0;

  pxClockBeginFrame(clock);
  while (pxClockShouldStep(clock) && fired < clock->maxStepsPerFrame) {
    pxClockAdvance(clock);
    fired++;
  }
  if (!fired) {
    return false;
  }

  PxBatchLayout *L = NULL;
  PxBatch *b = pxB200Batch(world, &L);
  if (!b) {
    fprintf(stderr, "pdphyzx: no CUDA batch back end: %s\n", pxBatchLastError());
    abort(); /* the reference asserts on allocation failure (src/px_world.c:23); no CPU fallback */
  }
  /* host-side edits since the last call (positions, forces, materials, gravity, new bodies)
   * travel with the world: pack -> H2D -> K substeps -> D2H -> unpack */
  int rc = pxPackOneWorld(L, world, 0, (uint8_t *)pxBatchHostMutable(b), (uint8_t *)pxBatchHostImmutable(b));
  if (rc == 0) rc = pxBatchImmutableH2D(b, NULL, 0, 1);
  if (rc == 0) rc = pxBatchMutableH2D(b, NULL, 0, 1);
  if (rc == 0) rc = pxBatchStep(b, fired * PX_PACK_SUBSTEPS_PER_STEP);
  if (rc == 0) rc = pxBatchMutableD2H(b, NULL, 0, 1);
  if (rc == 0) rc = pxBatchSync(b);
  if (rc == 0) rc = pxUnpackOneWorld(L, world, 0, (const uint8_t *)pxBatchHostMutable(b), 0);
  if (rc < 0) {
    fprintf(stderr, "pdphyzx: device step failed: %s\n", pxBatchLastError());
    abort();
  }
  return true;
}
