/*
 * px_batch.h -- the C-ABI batch layer: N independent pdphyzx worlds stepped on one B200.
 *
 * This is the drop-in boundary.  The reference has no FFI; its only entry into the hot path
 * is `bool (*step)(PxWorld *)` (reference src/includes/px_world_api.h:34 -> pxWorldStep,
 * src/px_world.c:422-445), whose body is 5 x pxWorldStepWithDt per fired fixed step
 * (src/px_world.c:410-420, 435-437).  pxBatchStep(b, K) is K x pxWorldStepWithDt for every
 * world of the batch, executed by the sm_100a kernel in pdphyzx_b200/csrc/px_step.cu.
 * Everything in this header is plain C: pointers, sizes, ints.  No CUDA or torch types.
 *
 * Each function below names the reference interface it replaces or batches.
 *
 * Error convention: int rc; 0 = ok, < 0 = CUDA/driver/argument error (text in
 * pxBatchLastError()), > 0 = number of worlds whose last step overflowed a capacity (their
 * PxContactStats.overflow / PxWorldHeader.statOverflow says which).  Nothing aborts.
 * Threading: one host thread per PxBatch at a time.
 */
#ifndef PX_BATCH_H
#define PX_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PX_API __attribute__((visibility("default")))
#else
#define PX_API
#endif

typedef struct PxBatch PxBatch;
struct PxWorld; /* include/pdphyzx.h */

/* PxBatchDesc.flags */
#define PX_BATCH_FLAG_TRACE 1u        /* keep the last substep's ordered pair + manifold lists */
#define PX_BATCH_FLAG_SERIAL_SOLVER 2u /* debug: one thread walks the manifolds in pool order */
#define PX_BATCH_FLAG_PROFILE 4u       /* accumulate SM-clock cycles per kernel phase per world */
#define PX_BATCH_FLAG_SPLIT 8u         /* force the split pipeline (sweep kernel + solver kernel per substep) */
#define PX_BATCH_FLAG_FUSED 16u        /* force the fused pipeline (one kernel, K substeps); default: by batch size */
#define PX_BATCH_FLAG_DYNAMIC_SOLVER 32u /* split pipeline: solve with the dataflow scheduler kernel instead of replaying the
                                            periodic schedule the sweep kernel derives (the default) */

/* PxWorldHeader.statOverflow bits */
#define PX_BATCH_OVERFLOW_PAIRS 1u     /* more AABB-pass pairs than maxPairs: extra pairs dropped */
#define PX_BATCH_OVERFLOW_MANIFOLDS 2u /* more manifolds than maxManifolds: extra manifolds dropped */
/* (the reference dereferences NULL when its 64-manifold pool runs out, src/px_world.c:333-336;
 *  there is no behaviour to match, so the device drops the excess and raises the flag) */

/* pxBatchReadback flags */
#define PX_READBACK_STATS_ONLY 1u

typedef struct PxBatchDesc {
  uint32_t structSize;   /* sizeof(PxBatchDesc), for forward compatibility */
  float unit;            /* PDPHYZX_UNIT of the host code: 1, 100 or 1000 (px_unit.h:18-20) */
  uint32_t maxDynamic;   /* per-world capacities; 0 = take from the worlds given to pxBatchNew */
  uint32_t maxStatic;
  uint32_t maxVertices;  /* polygon vertices per world (sum over bodies) */
  uint32_t maxManifolds; /* 0 = 2.875 * maxDynamic (<= 1024) */
  uint32_t maxPairs;     /* 0 = 7 * maxDynamic */
  int32_t device;        /* CUDA device ordinal, -1 = current device */
  void *stream;          /* cudaStream_t to launch on; NULL = a stream owned by the batch */
  uint32_t flags;        /* PX_BATCH_FLAG_* */
  uint32_t threads;      /* threads per world-CTA: 0 = default */
  const float *sinTable; /* the host's pxSinTable (256 floats, px_math.h:206-218); NULL = the
                            library fills one with sinf exactly as pxInitSinTable does */
} PxBatchDesc;

static inline PxBatchDesc pxBatchDescDefault(float unit) {
  PxBatchDesc d = {0};
  d.structSize = (uint32_t)sizeof(PxBatchDesc);
  d.unit = unit;
  d.device = -1;
  return d;
}

/*
 * Per-world header at the start of every world's mutable block (64 bytes).  The first ten
 * fields are the world-level inputs pxWorldStepWithDt reads every substep (PxWorld fields,
 * src/includes/px_world.h:36-41, and the clock-derived dt, src/px_world.c:432); the stat*
 * fields are written by the device after every pxBatchStep.
 */
typedef struct PxWorldHeader {
  float gravityX, gravityY;     /* already unit-scaled, as px->world->setGravity stores it */
  float dt;                     /* substep dt */
  float linearSleepThresholdSq;
  float angularSleepThreshold;
  float timeToSleep;
  uint32_t iterations;
  uint32_t nDynamic, nStatic;   /* active lengths (PxBodyArray.length) */
  uint32_t nVertices;
  uint32_t statManifolds;       /* manifolds kept by the last substep (= contacts.length) */
  uint32_t statSleeping;        /* sleeping dynamic bodies after the last substep */
  uint32_t statOverflow;        /* PX_BATCH_OVERFLOW_* accumulated since upload */
  uint32_t statPairs;           /* AABB-pass pairs of the last substep */
  uint32_t substeps;            /* substeps executed since upload */
  uint32_t reserved;
} PxWorldHeader;

/* rows of the per-world dynamic-body float matrix f32[PX_DYN_ROWS][maxDynamic], indexed by
 * SLOT (PxBody.worldIndex); mutable state in the order of px_body.h:82-120 */
enum {
  PX_DYN_PX = 0, PX_DYN_PY, PX_DYN_VX, PX_DYN_VY, PX_DYN_AV, PX_DYN_ANGLE, PX_DYN_FX, PX_DYN_FY,
  PX_DYN_TORQUE, PX_DYN_SLEEP, PX_DYN_M00, PX_DYN_M01, PX_DYN_M10, PX_DYN_M11, PX_DYN_ROWS
};
/* rows of the immutable dynamic matrix f32[PX_DYNC_ROWS][maxDynamic] */
enum { PX_DYNC_IMASS = 0, PX_DYNC_IINERTIA, PX_DYNC_E, PX_DYNC_SF, PX_DYNC_DF, PX_DYNC_RADIUS, PX_DYNC_ROWS };
/* Static bodies are indexed by POSITION in the static iteration order
 * (staticBodies.activeIndices), not by slot.  Their pose is mutable -- px->body->setPosition /
 * setOrientation / moveBy / rotate work on static bodies too (reference examples/sprites/main.c:91-95
 * turns the ground every frame) -- and lives in the mutable block: f32[PX_STATM_ROWS][maxStatic];
 * the step itself never writes it.  Materials and the AABB radius are immutable:
 * f32[PX_STAT_ROWS][maxStatic]. */
enum { PX_STATM_PX = 0, PX_STATM_PY, PX_STATM_ANGLE, PX_STATM_M00, PX_STATM_M01, PX_STATM_M10, PX_STATM_M11, PX_STATM_ROWS };
enum { PX_STAT_E = 0, PX_STAT_SF, PX_STAT_DF, PX_STAT_RADIUS, PX_STAT_ROWS };

/*
 * Where things are inside the two per-world blocks (all offsets in bytes from the start of
 * the world's block, all 16-byte aligned).  World w lives at mut + w * mutStride and
 * imm + w * immStride.  The same layout is used in pinned host staging and in HBM.
 */
typedef struct PxBatchLayout {
  uint32_t nWorlds;
  uint32_t maxDynamic, maxStatic, maxVertices, maxManifolds, maxPairs; /* maxDynamic/maxStatic padded to x4 */
  uint32_t mutStride, immStride;
  /* mutable block */
  uint32_t offHeader;    /* PxWorldHeader */
  uint32_t offDyn;       /* f32[PX_DYN_ROWS][maxDynamic] */
  uint32_t offDynFlags;  /* u8[maxDynamic]   PxBody.flags by slot */
  uint32_t offDynOrder;  /* u8[maxDynamic]   dynamicBodies.activeIndices */
  uint32_t offStatMut;   /* f32[PX_STATM_ROWS][maxStatic] static poses, by position */
  /* immutable block */
  uint32_t offDynConst;  /* f32[PX_DYNC_ROWS][maxDynamic] */
  uint32_t offDynShape;  /* u32[maxDynamic]  (vertexOffset << 8) | vertexCount, count 0 = circle; see PX_SHAPE_* */
  uint32_t offStat;      /* f32[PX_STAT_ROWS][maxStatic] */
  uint32_t offStatShape; /* u32[maxStatic] */
  uint32_t offStatSlot;  /* u8[maxStatic]    slot (worldIndex) of the static at each position */
  uint32_t offVerts;     /* float4[maxVertices] = {vertex.x, vertex.y, normal.x, normal.y} */
  /* trace block (PX_BATCH_FLAG_TRACE), per world at trace + w * traceStride */
  uint32_t traceStride;
  uint32_t offTracePairs;     /* u16[2][maxPairs]: slotA, then (bIsDynamic << 15) | slotB */
  uint32_t offTraceManifolds; /* f32[PX_TRACE_M_F32][maxManifolds] then u32 ids[maxManifolds] */
  /* unit-derived constants, computed on the host with the reference's expressions */
  float unit;
  float epsilon;          /* 0.0001f * U                   px_math.h:23 */
  float epsilonSq;        /* ((0.0001f * U) * 0.0001f) * U  as PX_EPSILON * PX_EPSILON expands */
  float allowedPenetration; /* 0.0001f * U                 px_manifold.h:15 */
} PxBatchLayout;

#define PX_SHAPE_COUNT(s) ((uint32_t)(s) & 255u)
#define PX_SHAPE_OFFSET(s) ((uint32_t)(s) >> 8)
#define PX_SHAPE_PACK(offset, count) ((uint32_t)(((uint32_t)(offset) << 8) | (uint32_t)(count)))
#define PX_MAX_DEVICE_VERTS 32 /* per polygon, as PX_MAX_POLY_VERTEX_COUNT */

/* trace manifold rows: normal.x normal.y penetration staticFriction dynamicFriction
 * mixedRestitution c0.x c0.y c1.x c1.y; ids = slotA | slotB << 8 | bIsDynamic << 16 | count << 24 */
#define PX_TRACE_M_F32 10

/* ------------------------------------------------------------------------------ lifecycle */

/* Pack `nWorlds` host worlds (AoS PxWorld, include/pdphyzx.h) into device SoA and return the
 * batch.  Capacities left 0 in `desc` are derived from the worlds.  Copies, among the rest,
 * the iteration orders, sleep state, clock-derived dt, iterations, gravity, thresholds and
 * the host's pxSinTable.  Does not take ownership of the worlds.  NULL on error.
 * Batches: PxWorld construction state consumed by pxWorldStepWithDt (src/px_world.c:410). */
PX_API PxBatch *pxBatchNew(struct PxWorld *const *worlds, uint32_t nWorlds, const PxBatchDesc *desc);

/* Empty batch of `nWorlds` worlds with explicit capacities (all desc.max* must be non-zero);
 * fill it with pxBatchUpload in chunks -- for batches whose AoS form would not fit host RAM. */
PX_API PxBatch *pxBatchCreate(uint32_t nWorlds, const PxBatchDesc *desc);

/* (Re)pack host worlds into batch worlds [first, first + n): the host -> device direction of
 * every field the step reads.  Used by px->world->step to carry host-side edits over. */
PX_API int pxBatchUpload(PxBatch *b, struct PxWorld *const *worlds, uint32_t first, uint32_t n);

PX_API void pxBatchFree(PxBatch *b);

/* -------------------------------------------------------------------------------- stepping */

/* K x pxWorldStepWithDt (src/px_world.c:410-420) for every world.  Small batches: one kernel launch, state
 * resident in shared memory across the K substeps.  Large batches: a sweep kernel and a solver kernel per
 * substep, the two halves of the batch on two internal streams that fork from and join the batch's stream.
 * Asynchronous on the batch's stream either way. */
PX_API int pxBatchStep(PxBatch *b, uint32_t kSubsteps);

/* One step with HOST buffers: the mutable blocks of every world are copied host -> HBM from
 * `hostMut` (layout and size of pxBatchHostMutable(); NULL = that staging buffer; pinned memory
 * for the copies to be asynchronous), stepped K substeps and copied back, in `chunks`
 * contiguous world ranges (0 = default: five, the first and the last half as long as the inner ones)
 * pipelined over three internal streams so that the PCIe copies of one range run under the substeps
 * of another.  Ordered after everything queued on
 * the batch's stream; pxBatchSync (or later work on the stream) waits for all of it.  This is
 * what a px->world->step over N host-resident worlds costs end to end. */
PX_API int pxBatchStepHost(PxBatch *b, uint32_t kSubsteps, void *hostMut, uint32_t chunks);

/* CUDA events on the batch's stream around whatever is queued between the two calls; Stop blocks
 * until the work is done; *ms = device time.  (Start/Stop of several batches on several devices
 * can be interleaved: pxBatchGroupStepTimed.) */
PX_API int pxBatchTimerStart(PxBatch *b);
PX_API int pxBatchTimerStop(PxBatch *b, float *ms);
/* One step with EVERY kernel launch bracketed by CUDA events; blocks.  ms[c] / n[c]: summed device time and number of
 * the launches of class c -- 0 the sweep kernel (or the one fused launch), 1 the finish kernel (phase F of the
 * step's last substep; static solver only), 2 the solver kernel (dataflow, or schedule + replay).  The roofline numbers of bench.py come from here. */
PX_API int pxBatchStepTimedKernels(PxBatch *b, uint32_t kSubsteps, float ms[3], uint32_t n[3]);
/* which pipeline the batch runs: *split = 0 for the fused kernel, 1 for sweep + dataflow-solver kernels per substep,
 * 2 for sweep + schedule-and-replay kernels per substep (large batches); resident CTAs per SM of the solver
 * kernel; *schedCtasPerSm is reserved (0) */
PX_API int pxBatchPipelineInfo(const PxBatch *b, int *split, int *solveCtasPerSm, int *schedCtasPerSm);
/* Same, bracketed by CUDA events on the launching stream; blocks; *ms = device time. */
PX_API int pxBatchStepTimed(PxBatch *b, uint32_t kSubsteps, float *ms);

/* Wait for the stream; 0 or < 0.  Capacity overflows are reported per world by
 * pxBatchReadback (return value and PxContactStats.overflow). */
PX_API int pxBatchSync(PxBatch *b);

/* Sync, then write position, velocity, angularVelocity, orientationAngle, orientation, aabb,
 * force, torque, sleepTime, flags and the dynamic activeIndices back into host worlds
 * [first, first + n), plus PxWorld.contacts statistics.  Device -> host direction of
 * everything pxWorldStepWithDt mutates (PxBody fields, src/includes/px_body.h:82-120). */
PX_API int pxBatchReadback(PxBatch *b, struct PxWorld *const *worlds, uint32_t first, uint32_t n, uint32_t flags);

/* ------------------------------------------------------------- control path (queued, K=1..) */
/* Applied in call order at the start of the next pxBatchStep, before its first substep (and
 * before any device -> host copy: pxBatchReadback / pxBatchMutableD2H see their effect).
 * body = slot (PxBody.worldIndex) of a dynamic body; world = index in the batch or
 * PX_ALL_WORLDS for the setters. */
#define PX_ALL_WORLDS 0xffffffffu
/* px->body->applyImpulse: scales by the unit, v += J*iM, w += cross(c, J)*iI, wakes the body
 * (src/px_body_api.c:37-42, src/px_body.c:147-156, 195-202) */
PX_API int pxBatchApplyImpulse(PxBatch *b, uint32_t world, uint32_t body, float ix, float iy, float cx, float cy);
/* n applyImpulse calls in one (per-step control of many worlds): element i acts on body bodies[i]
 * of world worlds[i] with impulse impulseXY[2i..], contact contactXY[2i..] (NULL = body centre);
 * queued and ordered exactly like n single calls */
PX_API int pxBatchApplyImpulses(PxBatch *b, uint32_t n, const uint32_t *worlds, const uint32_t *bodies, const float *impulseXY,
                                const float *contactXY);
/* px->body->applyForce (src/px_body_api.c:20-26, src/px_body.c:158-165) */
PX_API int pxBatchApplyForce(PxBatch *b, uint32_t world, uint32_t body, float fx, float fy, float cx, float cy);
/* px->body->setPosition / moveBy / setOrientation / rotate (src/px_body_api.c:57-60,
 * src/px_body.c:85-124): plain pose edits, no wake-up; the orientation matrix comes from the same
 * sine table as the step's.  `body` is the slot of a dynamic body, or PX_BATCH_STATIC | slot for
 * a static one (examples/sprites/main.c:91-95 turns its static ground every frame). */
#define PX_BATCH_STATIC 0x10000u
PX_API int pxBatchSetPosition(PxBatch *b, uint32_t world, uint32_t body, float x, float y);
PX_API int pxBatchMoveBy(PxBatch *b, uint32_t world, uint32_t body, float dx, float dy);
PX_API int pxBatchSetOrientation(PxBatch *b, uint32_t world, uint32_t body, float radians);
PX_API int pxBatchRotate(PxBatch *b, uint32_t world, uint32_t body, float radians);
/* px->world->setGravity: stores (x, y) * unit (src/px_world.c:57-63) */
PX_API int pxBatchSetGravity(PxBatch *b, uint32_t world, float x, float y);
/* px->world->setIterations (src/px_world.c:65-71) */
PX_API int pxBatchSetIterations(PxBatch *b, uint32_t world, uint32_t iterations);

/* ------------------------------------------------------ body lifecycle on a resident batch */
/* The host keeps the reference's PxBodyArray bookkeeping (slot reuse from the freed stack, swap-remove
 * from activeIndices, src/px_body_array.c:7-79); these calls mirror one host-side
 * px->world->newDynamicBody / newStaticBody / freeBody onto the resident world without re-creating
 * or re-uploading it.  Queued like the control commands; no device allocation.
 *
 * pxBatchAddBody: `body` is what the host call returned (it lives in hostWorld, its worldIndex is
 * the slot the reference's rule chose).  Dynamic: constants, shape and vertices go to the world's
 * immutable block, the mutable rows are written on the device and the slot is appended to the
 * device's iteration order (which, not the host's stale copy, is the order the reference would
 * have).  Static: the world's static list is re-packed from hostWorld (the host re-sorts it on
 * every add, src/px_world.c:94-101; host static poses are authoritative at that moment).
 * pxBatchFreeBody: dynamic = slot, swap-removed from the device's order with the reference's own
 * `index >= length` no-op rule (src/px_body_array.c:64); static = PX_BATCH_STATIC | slot, re-packs
 * the static list from hostWorld (after the host-side freeBody). */
/* (PxBody is an untagged typedef in the reference, src/includes/px_body.h:82-120, hence the void pointer) */
PX_API int pxBatchAddBody(PxBatch *b, uint32_t world, const struct PxWorld *hostWorld, const void *body /* const PxBody * */);
PX_API int pxBatchFreeBody(PxBatch *b, uint32_t world, uint32_t body, const struct PxWorld *hostWorld);
PX_API int pxBatchSyncStatics(PxBatch *b, uint32_t world, const struct PxWorld *hostWorld);
/* device + pinned allocations the batch has made since creation (0 unless a command burst outgrew
 * the buffers sized at creation) */
PX_API uint64_t pxBatchAllocCount(const PxBatch *b);

/* ------------------------------------------------------------------ raw blocks (SoA level) */

PX_API int pxBatchGetLayout(const PxBatch *b, PxBatchLayout *out);
/* Pinned host staging of the two blocks (layout above); valid until pxBatchFree. */
PX_API void *pxBatchHostMutable(PxBatch *b);
PX_API void *pxBatchHostImmutable(PxBatch *b);
/* Async copies of worlds [first, first + n) between host staging (or `host` if non-NULL,
 * same layout, ideally pinned) and HBM on the batch's stream. */
PX_API int pxBatchMutableH2D(PxBatch *b, const void *host, uint32_t first, uint32_t n);
PX_API int pxBatchMutableD2H(PxBatch *b, void *host, uint32_t first, uint32_t n);
PX_API int pxBatchImmutableH2D(PxBatch *b, const void *host, uint32_t first, uint32_t n);
/* Copy the trace block of worlds [first, first + n) into `host` (n * traceStride bytes). */
PX_API int pxBatchTraceD2H(PxBatch *b, void *host, uint32_t first, uint32_t n);
/* PX_BATCH_FLAG_PROFILE: copy out u64[nWorlds][16] accumulated since creation: SM-clock cycles per
 * phase (0 A sort, 1 B sweep, 2 C narrowphase, 3 D lists, 4 E solver, 5 F integrate/correct/sleep),
 * 6 solver rounds, 7 polygon pairs << 32 | pairs left after the quick rejection, 8-11 cycles of the
 * narrowphase stages (circle pairs, quick rejection, SAT, clipping; the rest of 2 is manifold init) */
PX_API int pxBatchProfileD2H(PxBatch *b, uint64_t *host);
/* Device pointers (for zero-copy consumers on the same GPU), as integers */
PX_API uint64_t pxBatchDeviceMutable(PxBatch *b);
PX_API uint64_t pxBatchDeviceImmutable(PxBatch *b);
/* kernels launched by this batch since creation (bench.py's gpu_launches) */
PX_API uint64_t pxBatchLaunchCount(const PxBatch *b);
/* occupancy facts of the step kernel for this layout: CTAs per SM, dynamic smem bytes, threads */
PX_API int pxBatchKernelInfo(const PxBatch *b, int *ctasPerSm, int *smemBytes, int *threads, int *numSms);

/* Compute a layout from capacities without touching a GPU (host-only; used by the packer,
 * the oracle and tests). */
PX_API int pxBatchComputeLayout(uint32_t nWorlds, const PxBatchDesc *desc, PxBatchLayout *out);

/* Host-only AoS <-> SoA conversion against an arbitrary buffer pair (no GPU needed). */
PX_API int pxPackWorlds(const PxBatchLayout *layout, struct PxWorld *const *worlds, uint32_t n, void *mut, void *imm);
PX_API int pxUnpackWorlds(const PxBatchLayout *layout, struct PxWorld *const *worlds, uint32_t n, const void *mut, uint32_t flags);
/* Capacities the given worlds need (fills desc.max* that are 0). */
PX_API int pxBatchFitDesc(struct PxWorld *const *worlds, uint32_t n, PxBatchDesc *desc);

/* Arithmetic self-tests on the current device (parity tier T0): op 0 rsqrt, 1 rcp, 2 sqrt,
 * 3 div(a, b), 4 sin, 5 cos -- the device restatement of reference px_math.h:50-114, 226-252
 * applied to host arrays; and pxFastRsqrt over all 2^32 bit patterns as 4096 block sums of the
 * result bit patterns (block i covers patterns [i << 20, (i + 1) << 20)). */
PX_API int pxSelfTestMath(int op, const float *a, const float *b, const float *sinTable256, float *out, uint32_t n);
PX_API int pxSelfTestRsqrtSweep(uint64_t *sums4096);

/* --------------------------------------------------------------- several GPUs, one caller */
/*
 * Worlds are independent, so a batch shards over the GPUs of a box with no collective and no peer
 * traffic: a PxBatchGroup owns one PxBatch -- slab, stream, staging -- per device and
 * fans every call out from the calling host thread; the per-device work is asynchronous, so the
 * devices run side by side.  World w of the group lives in shard w / ceil(nWorlds / nDevices)
 * (contiguous ranges, the same rule as pdphyzx_b200/sharding.py).  devices == NULL: devices
 * 0 .. nDevices-1.  The control and lifecycle calls of a single world go to its shard:
 * pxBatchGroupLocate gives the shard's PxBatch and the world's index in it.
 */
typedef struct PxBatchGroup PxBatchGroup;
PX_API PxBatchGroup *pxBatchGroupNew(struct PxWorld *const *worlds, uint32_t nWorlds, const PxBatchDesc *desc,
                                     const int32_t *devices, uint32_t nDevices);
PX_API PxBatchGroup *pxBatchGroupCreate(uint32_t nWorlds, const PxBatchDesc *desc, const int32_t *devices, uint32_t nDevices);
PX_API void pxBatchGroupFree(PxBatchGroup *g);
PX_API uint32_t pxBatchGroupShards(const PxBatchGroup *g);
PX_API PxBatch *pxBatchGroupShard(const PxBatchGroup *g, uint32_t shard, uint32_t *firstWorld, uint32_t *nWorlds);
PX_API PxBatch *pxBatchGroupLocate(const PxBatchGroup *g, uint32_t world, uint32_t *localWorld);
PX_API int pxBatchGroupUpload(PxBatchGroup *g, struct PxWorld *const *worlds, uint32_t first, uint32_t n);
PX_API int pxBatchGroupStep(PxBatchGroup *g, uint32_t kSubsteps);
PX_API int pxBatchGroupStepHost(PxBatchGroup *g, uint32_t kSubsteps, uint32_t chunks);
/* one step on every device, device-timed: *ms = the slowest device's time */
PX_API int pxBatchGroupStepTimed(PxBatchGroup *g, uint32_t kSubsteps, float *ms);
PX_API int pxBatchGroupSync(PxBatchGroup *g);
PX_API int pxBatchGroupReadback(PxBatchGroup *g, struct PxWorld *const *worlds, uint32_t first, uint32_t n, uint32_t flags);
/* setters for one world or PX_ALL_WORLDS, routed to the owning shard(s) */
PX_API int pxBatchGroupSetGravity(PxBatchGroup *g, uint32_t world, float x, float y);
PX_API int pxBatchGroupSetIterations(PxBatchGroup *g, uint32_t world, uint32_t iterations);
PX_API int pxBatchGroupApplyImpulse(PxBatchGroup *g, uint32_t world, uint32_t body, float ix, float iy, float cx, float cy);

PX_API const char *pxBatchLastError(void);
PX_API const char *pxBatchVersion(void);

#ifdef __cplusplus
}
#endif
#endif /* PX_BATCH_H */
