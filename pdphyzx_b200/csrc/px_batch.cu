/*
 * px_batch.cu -- the C-ABI batch layer (include/px_batch.h): device memory, pinned staging,
 * stream, command queue and launches for one batch of worlds on one B200.
 *
 * There is no CPU path here.  Every stepping entry point launches the sm_100a kernel in
 * px_step.cu or fails with an error; the AoS<->SoA conversion (px_pack.c) is the only host
 * arithmetic and it only moves values.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "px_batch.h"
#include "px_internal.h"
#include "px_step.cuh"

#define PX_AUX 3           /* helper streams: world ranges of a step run side by side on them */
#define PX_REGIONS 4u       /* scratch regions (fused pipeline): the batch's stream and the helper streams of pxBatchStepHost */
#define PX_COUNTER_RING 1024u /* >= launches that can be in flight at once (a host-buffer step queues (2K + 1) x chunks) */

struct PxBatch {
  PxBatchLayout L;
  PxSmemPlan plan;
  uint32_t flags;
  int device;
  cudaStream_t stream;
  bool ownStream;
  uint32_t threads;
  uint8_t *dMut, *dImm, *dTrace;
  float *dSin;
  unsigned long long *dProf;
  float4 *dScratch;   /* manifold records of the resident CTAs (px_step.cuh): PX_REGIONS regions of gridMax slots */
  uint32_t *dCounters; /* world counters of the persistent grids, one per launch in flight (ring) */
  uint32_t gridMax;    /* CTAs of a full launch: resident capacity of the device, at most nWorlds */
  /* split pipeline (sweep kernel + solver kernel per substep, px_step.cu) */
  bool split;
  bool staticSched;    /* split pipeline: the sweep kernel derives a periodic visit schedule, pxReplayKernel replays it */
  PxHandLayout hand;
  PxSmemPlan solvePlan;
  uint8_t *dHand;      /* hand-over records, one per world */
  uint32_t gridSolve;
  int solveCtasPerSm;

  uint8_t *hMut, *hImm; /* pinned staging, same layout */
  float sinTable[256];
  /* control path */
  std::vector<PxCommand> cmds, globalCmds;
  std::vector<PxBodyPayload> payloads; /* mutable rows of bodies added since the last flush */
  PxBodyPayload *dPayloads;
  size_t dPayloadCap;
  uint64_t allocs; /* device / pinned allocations made so far (pxBatchAllocCount) */
  uint32_t cmdSeq;
  PxCommand *dCmds, *dGlobal;
  uint32_t *dWorldStart;
  size_t dCmdCap, dGlobalCap;
  PxCommand *hCmds;
  uint32_t *hWorldStart;
  size_t hCmdCap;
  uint64_t launches;
  int ctasPerSm, numSms;
  /* pxBatchStepHost: two helper streams pipeline the chunks of a host-buffer step */
  cudaStream_t aux[PX_AUX];
  cudaEvent_t evFork, evJoin[PX_AUX];
  bool auxReady;
  cudaEvent_t evTimer[2]; /* pxBatchTimerStart / pxBatchTimerStop */
  bool timerReady;
};

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t _e = (call);                                                                            \
    if (_e != cudaSuccess) return pxSetError(-(int)_e - 1000, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

static int useDevice(const PxBatch *b) {
  CU(cudaSetDevice(b->device));
  return 0;
}

extern "C" PxBatch *pxBatchCreate(uint32_t nWorlds, const PxBatchDesc *desc) {
  if (!desc || nWorlds == 0) {
    pxSetError(-1, "pxBatchCreate: need a descriptor and at least one world");
    return nullptr;
  }
  PxBatch *b = new PxBatch();
  memset(&b->L, 0, sizeof(b->L));
  b->dMut = b->dImm = b->dTrace = nullptr;
  b->dSin = nullptr;
  b->dProf = nullptr;
  b->dScratch = nullptr;
  b->auxReady = false;
  b->timerReady = false;
  b->dCounters = nullptr;
  b->gridMax = 0;
  b->split = false;
  b->staticSched = false;
  b->dHand = nullptr;
  b->gridSolve = 0;
  b->solveCtasPerSm = 0;

  b->hMut = b->hImm = nullptr;
  b->dCmds = b->dGlobal = nullptr;
  b->dWorldStart = nullptr;
  b->hCmds = nullptr;
  b->hWorldStart = nullptr;
  b->dCmdCap = b->dGlobalCap = b->hCmdCap = 0;
  b->dPayloads = nullptr;
  b->dPayloadCap = 0;
  b->allocs = 0;
  b->cmdSeq = 0;
  b->launches = 0;
  b->stream = nullptr;
  b->ownStream = false;
  b->flags = desc->flags;
  if (pxBatchComputeLayout(nWorlds, desc, &b->L) != 0) {
    delete b;
    return nullptr;
  }

  auto fail = [&](const char *what, cudaError_t e) -> PxBatch * {
    pxSetError(-(int)e - 1000, "pxBatchCreate: %s: %s", what, cudaGetErrorString(e));
    pxBatchFree(b);
    return nullptr;
  };
  cudaError_t e;
  int dev = desc->device;
  if (dev < 0) {
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return fail("cudaGetDevice (is a CUDA device present?)", e);
  }
  b->device = dev;
  if ((e = cudaSetDevice(dev)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  b->numSms = prop.multiProcessorCount;
  if (desc->stream) {
    b->stream = (cudaStream_t)desc->stream;
  } else {
    if ((e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    b->ownStream = true;
  }
  /* Pipeline: fused (one kernel, K substeps, state never leaves the SM) for small batches and
   * single worlds; split (a sweep kernel and a solver kernel per substep) once the batch is a few
   * waves deep -- the solver's latency-bound chains then run at more than twice the residency and
   * without throughput warps stealing their issue slots.  PX_BATCH_FLAG_SPLIT / _FUSED force one. */
  {
    const char *env = getenv("PX_PIPELINE"); /* tuning switch: "split" or "fused" */
    if (desc->flags & PX_BATCH_FLAG_SERIAL_SOLVER) b->split = false; /* the debug walk lives in the fused kernel */
    else if (desc->flags & PX_BATCH_FLAG_SPLIT) b->split = true;
    else if (desc->flags & PX_BATCH_FLAG_FUSED) b->split = false;
    else if (env) b->split = env[0] == 's';
    else b->split = nWorlds >= 4u * (uint32_t)b->numSms;
    /* the split pipeline's solver: the static (replayed) schedule unless the dataflow solver is asked for */
    const char *sv = getenv("PX_SOLVER"); /* tuning switch: "dynamic" or "static" */
    b->staticSched = b->split && !(desc->flags & PX_BATCH_FLAG_DYNAMIC_SOLVER) && !(sv && sv[0] == 'd');
  }
  pxPlanSmem(&b->L, 1, b->staticSched ? 1 : 0, &b->plan);
  /* fused: 128 threads, 5 CTAs/SM keep the SM's issue slots busier than 3 CTAs of 256 threads;
   * split: the sweep kernel alone is throughput work, 256 threads x 3 CTAs leave it a larger L1 */
  b->threads = desc->threads ? desc->threads : (b->split ? 256 : 128);
  if (b->threads != 128 && b->threads != 256) {
    pxSetError(-1, "PxBatchDesc.threads must be 128 or 256");
    pxBatchFree(b);
    return nullptr;
  }
  {
    /* keep the polygon vertex pool in shared memory unless reading it through L1 instead
     * lets one more world be resident per SM (the solver phase is latency-bound: resident
     * worlds are what hides it) */
    PxSmemPlan slim;
    pxPlanSmem(&b->L, 0, b->staticSched ? 1 : 0, &slim);
    const size_t perSm = (size_t)prop.sharedMemPerMultiprocessor, reserve = (size_t)prop.reservedSharedMemPerBlock;
    size_t fat = perSm / (b->plan.total + reserve), thin = perSm / (slim.total + reserve);
    if (b->plan.total <= (uint32_t)prop.sharedMemPerBlockOptin) {
      /* what counts is the occupancy the kernel really gets (registers may bind before shared memory) */
      PxStepParams Q;
      memset(&Q, 0, sizeof(Q));
      Q.L = b->L;
      Q.mode = b->split ? PX_MODE_SWEEP : PX_MODE_FUSED;
      int of = 0, ot = 0;
      Q.S = b->plan;
      const int r1 = pxStepOccupancy(&Q, b->threads, &of);
      Q.S = slim;
      const int r2 = pxStepOccupancy(&Q, b->threads, &ot);
      if (r1 == 0 && r2 == 0) {
        fat = (size_t)of;
        thin = (size_t)ot;
      }
    }
    const char *force = getenv("PX_VERTS_SMEM"); /* tuning switch, undocumented on purpose: 1 = keep them in shared memory, 0 = never */
    if (b->plan.total > (uint32_t)prop.sharedMemPerBlockOptin || (force ? force[0] == '0' : thin > fat)) b->plan = slim;
  }

  if (b->plan.total > (uint32_t)prop.sharedMemPerBlockOptin) {
    pxSetError(-2, "pxBatchCreate: a world needs %u bytes of shared memory, the device offers %zu per CTA; lower the capacities",
               b->plan.total, (size_t)prop.sharedMemPerBlockOptin);
    pxBatchFree(b);
    return nullptr;
  }
  const size_t mutBytes = (size_t)nWorlds * b->L.mutStride, immBytes = (size_t)nWorlds * b->L.immStride;
  if ((e = cudaMalloc(&b->dMut, mutBytes)) != cudaSuccess) return fail("cudaMalloc(mutable)", e);
  if ((e = cudaMalloc(&b->dImm, immBytes)) != cudaSuccess) return fail("cudaMalloc(immutable)", e);
  if ((e = cudaMemsetAsync(b->dMut, 0, mutBytes, b->stream)) != cudaSuccess) return fail("cudaMemset", e);
  if ((e = cudaMemsetAsync(b->dImm, 0, immBytes, b->stream)) != cudaSuccess) return fail("cudaMemset", e);
  if (b->flags & PX_BATCH_FLAG_TRACE) {
    const size_t traceBytes = (size_t)nWorlds * b->L.traceStride;
    if ((e = cudaMalloc(&b->dTrace, traceBytes)) != cudaSuccess) return fail("cudaMalloc(trace)", e);
    if ((e = cudaMemsetAsync(b->dTrace, 0, traceBytes, b->stream)) != cudaSuccess) return fail("cudaMemset", e);
  }
  if (b->flags & PX_BATCH_FLAG_PROFILE) {
    if ((e = cudaMalloc(&b->dProf, (size_t)nWorlds * PX_PROF_SLOTS * 8)) != cudaSuccess) return fail("cudaMalloc(prof)", e);
    if ((e = cudaMemsetAsync(b->dProf, 0, (size_t)nWorlds * PX_PROF_SLOTS * 8, b->stream)) != cudaSuccess) return fail("cudaMemset", e);
  }
  {
    /* the step kernel runs as ONE wave of persistent CTAs (px_step.cu); each owns a scratch slot */
    PxStepParams P;
    memset(&P, 0, sizeof(P));
    P.L = b->L;
    P.S = b->plan;
    P.mode = b->split ? PX_MODE_SWEEP : PX_MODE_FUSED;
    int occ = 0;
    int rc = pxStepOccupancy(&P, b->threads, &occ);
    if (rc != 0) return fail("occupancy query of the step kernel", (cudaError_t)rc);
    if (occ < 1) {
      pxSetError(-2, "pxBatchCreate: the step kernel does not fit an SM with %u bytes of shared memory per world", b->plan.total);
      pxBatchFree(b);
      return nullptr;
    }
    b->ctasPerSm = occ;
    const uint64_t capacity = (uint64_t)occ * (uint64_t)b->numSms;
    b->gridMax = (uint32_t)(capacity < nWorlds ? capacity : nWorlds);
    if (!b->split || b->staticSched) {
      const size_t scratchBytes = (size_t)PX_REGIONS * b->gridMax * b->L.maxManifolds * 3 * sizeof(float4);
      if ((e = cudaMalloc(&b->dScratch, scratchBytes)) != cudaSuccess) return fail("cudaMalloc(scratch)", e);
    }
    if (b->split) {
      pxPlanHandover(&b->L, &b->plan, &b->hand);
      pxPlanSolveSmem(&b->L, &b->plan, b->staticSched ? 1 : 0, &b->solvePlan);
      P.solveS = b->solvePlan;
      P.solver = b->staticSched ? PX_SOLVER_STATIC : PX_SOLVER_DYNAMIC;
      int socc = 0;
      rc = pxSolveOccupancy(&P, &socc);
      if (rc != 0 || socc < 1) return fail("occupancy query of the solver kernel", (cudaError_t)(rc ? rc : (int)cudaErrorLaunchOutOfResources));
      b->solveCtasPerSm = socc;
      const uint64_t scap = (uint64_t)socc * (uint64_t)b->numSms;
      b->gridSolve = (uint32_t)(scap < nWorlds ? scap : nWorlds);
      const size_t handBytes = (size_t)nWorlds * b->hand.stride;
      if ((e = cudaMalloc(&b->dHand, handBytes)) != cudaSuccess) return fail("cudaMalloc(hand-over records)", e);
      if ((e = cudaMemsetAsync(b->dHand, 0, handBytes, b->stream)) != cudaSuccess) return fail("cudaMemset", e);
    }
    if ((e = cudaMalloc(&b->dCounters, PX_COUNTER_RING * sizeof(uint32_t))) != cudaSuccess) return fail("cudaMalloc(counters)", e);
  }
  {
    /* the control path's buffers, sized so that ordinary per-step control (a few commands per
     * world) never allocates after creation */
    b->hCmdCap = b->dCmdCap = std::max<size_t>(4096, (size_t)nWorlds * 4);
    b->dGlobalCap = 256;
    b->dPayloadCap = std::max<size_t>(256, (size_t)nWorlds / 4);
    if ((e = cudaMallocHost(&b->hCmds, b->hCmdCap * sizeof(PxCommand))) != cudaSuccess) return fail("cudaMallocHost(commands)", e);
    if ((e = cudaMallocHost(&b->hWorldStart, (size_t)(nWorlds + 1) * sizeof(uint32_t))) != cudaSuccess) return fail("cudaMallocHost(worldStart)", e);
    if ((e = cudaMalloc(&b->dCmds, b->dCmdCap * sizeof(PxCommand))) != cudaSuccess) return fail("cudaMalloc(commands)", e);
    if ((e = cudaMalloc(&b->dWorldStart, (size_t)(nWorlds + 1) * sizeof(uint32_t))) != cudaSuccess) return fail("cudaMalloc(worldStart)", e);
    if ((e = cudaMalloc(&b->dGlobal, b->dGlobalCap * sizeof(PxCommand))) != cudaSuccess) return fail("cudaMalloc(global commands)", e);
    if ((e = cudaMalloc(&b->dPayloads, b->dPayloadCap * sizeof(PxBodyPayload))) != cudaSuccess) return fail("cudaMalloc(payloads)", e);
  }
  if ((e = cudaMallocHost(&b->hMut, mutBytes)) != cudaSuccess) return fail("cudaMallocHost(mutable)", e);
  if ((e = cudaMallocHost(&b->hImm, immBytes)) != cudaSuccess) return fail("cudaMallocHost(immutable)", e);
  memset(b->hMut, 0, mutBytes);
  memset(b->hImm, 0, immBytes);

  /* the sine table is the host's: pxInitSinTable, reference px_math.h:213-218 */
  if (desc->sinTable) {
    memcpy(b->sinTable, desc->sinTable, sizeof(b->sinTable));
  } else {
    for (int i = 0; i < 256; i++) b->sinTable[i] = sinf((i * 3.141592741f * 2.0f) / 256);
  }
  if ((e = cudaMalloc(&b->dSin, sizeof(b->sinTable))) != cudaSuccess) return fail("cudaMalloc(sin)", e);
  if ((e = cudaMemcpyAsync(b->dSin, b->sinTable, sizeof(b->sinTable), cudaMemcpyHostToDevice, b->stream)) != cudaSuccess)
    return fail("cudaMemcpy(sin)", e);

  if ((e = cudaStreamSynchronize(b->stream)) != cudaSuccess) return fail("cudaStreamSynchronize", e);
  return b;
}

extern "C" PxBatch *pxBatchNew(struct PxWorld *const *worlds, uint32_t nWorlds, const PxBatchDesc *desc) {
  if (!worlds || !desc || nWorlds == 0) {
    pxSetError(-1, "pxBatchNew: need worlds and a descriptor");
    return nullptr;
  }
  PxBatchDesc d = *desc;
  if (pxBatchFitDesc(worlds, nWorlds, &d) != 0) return nullptr;
  PxBatch *b = pxBatchCreate(nWorlds, &d);
  if (!b) return nullptr;
  if (pxBatchUpload(b, worlds, 0, nWorlds) != 0) {
    pxBatchFree(b);
    return nullptr;
  }
  return b;
}

extern "C" void pxBatchFree(PxBatch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->dMut);
  cudaFree(b->dImm);
  cudaFree(b->dTrace);
  cudaFree(b->dSin);
  cudaFree(b->dProf);
  cudaFree(b->dScratch);
  cudaFree(b->dHand);
  if (b->auxReady) {
    for (int i = 0; i < PX_AUX; i++) {
      cudaStreamDestroy(b->aux[i]);
      cudaEventDestroy(b->evJoin[i]);
    }
    cudaEventDestroy(b->evFork);
  }
  if (b->timerReady) {
    cudaEventDestroy(b->evTimer[0]);
    cudaEventDestroy(b->evTimer[1]);
  }
  cudaFree(b->dCounters);
  cudaFree(b->dCmds);
  cudaFree(b->dPayloads);
  cudaFree(b->dGlobal);
  cudaFree(b->dWorldStart);
  cudaFreeHost(b->hMut);
  cudaFreeHost(b->hImm);
  cudaFreeHost(b->hCmds);
  cudaFreeHost(b->hWorldStart);
  if (b->ownStream && b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

static int flushCommands(PxBatch *b);

static int checkRange(const PxBatch *b, uint32_t first, uint32_t n, const char *who) {
  if (!b) return pxSetError(-1, "%s: NULL batch", who);
  if (first > b->L.nWorlds || n > b->L.nWorlds - first) return pxSetError(-1, "%s: worlds [%u, %u) out of range (%u)", who, first, first + n, b->L.nWorlds);
  return 0;
}

extern "C" int pxBatchMutableH2D(PxBatch *b, const void *host, uint32_t first, uint32_t n) {
  if (checkRange(b, first, n, "pxBatchMutableH2D")) return -1;
  if (useDevice(b)) return -1;
  const uint8_t *src = host ? (const uint8_t *)host : b->hMut + (size_t)first * b->L.mutStride;
  CU(cudaMemcpyAsync(b->dMut + (size_t)first * b->L.mutStride, src, (size_t)n * b->L.mutStride, cudaMemcpyHostToDevice, b->stream));
  return 0;
}

extern "C" int pxBatchMutableD2H(PxBatch *b, void *host, uint32_t first, uint32_t n) {
  if (checkRange(b, first, n, "pxBatchMutableD2H")) return -1;
  if (useDevice(b)) return -1;
  if (flushCommands(b)) return -1; /* what the host reads reflects every call made so far */
  uint8_t *dst = host ? (uint8_t *)host : b->hMut + (size_t)first * b->L.mutStride;
  CU(cudaMemcpyAsync(dst, b->dMut + (size_t)first * b->L.mutStride, (size_t)n * b->L.mutStride, cudaMemcpyDeviceToHost, b->stream));
  return 0;
}

extern "C" int pxBatchImmutableH2D(PxBatch *b, const void *host, uint32_t first, uint32_t n) {
  if (checkRange(b, first, n, "pxBatchImmutableH2D")) return -1;
  if (useDevice(b)) return -1;
  const uint8_t *src = host ? (const uint8_t *)host : b->hImm + (size_t)first * b->L.immStride;
  CU(cudaMemcpyAsync(b->dImm + (size_t)first * b->L.immStride, src, (size_t)n * b->L.immStride, cudaMemcpyHostToDevice, b->stream));
  return 0;
}

extern "C" int pxBatchTraceD2H(PxBatch *b, void *host, uint32_t first, uint32_t n) {
  if (checkRange(b, first, n, "pxBatchTraceD2H")) return -1;
  if (!b->dTrace) return pxSetError(-1, "pxBatchTraceD2H: the batch was created without PX_BATCH_FLAG_TRACE");
  if (useDevice(b)) return -1;
  CU(cudaMemcpyAsync(host, b->dTrace + (size_t)first * b->L.traceStride, (size_t)n * b->L.traceStride, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  return 0;
}

extern "C" int pxBatchUpload(PxBatch *b, struct PxWorld *const *worlds, uint32_t first, uint32_t n) {
  if (checkRange(b, first, n, "pxBatchUpload")) return -1;
  if (useDevice(b)) return -1;
  CU(cudaStreamSynchronize(b->stream)); /* staging may still be the source of an earlier copy */
  int rc = pxPackWorlds(&b->L, worlds, n, b->hMut + (size_t)first * b->L.mutStride, b->hImm + (size_t)first * b->L.immStride);
  if (rc != 0) return rc;
  if (pxBatchMutableH2D(b, nullptr, first, n)) return -1;
  if (pxBatchImmutableH2D(b, nullptr, first, n)) return -1;
  return 0;
}

extern "C" int pxBatchSync(PxBatch *b) {
  if (!b) return pxSetError(-1, "pxBatchSync: NULL batch");
  if (useDevice(b)) return -1;
  cudaError_t e = cudaStreamSynchronize(b->stream);
  if (e != cudaSuccess) return pxSetError(-(int)e - 1000, "pxBatchSync: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int pxBatchReadback(PxBatch *b, struct PxWorld *const *worlds, uint32_t first, uint32_t n, uint32_t flags) {
  if (checkRange(b, first, n, "pxBatchReadback")) return -1;
  if (pxBatchMutableD2H(b, nullptr, first, n)) return -1;
  CU(cudaStreamSynchronize(b->stream));
  return pxUnpackWorlds(&b->L, worlds, n, b->hMut + (size_t)first * b->L.mutStride, flags);
}

/* ------------------------------------------------------------------------- commands */
static int pushCommand(PxBatch *b, uint32_t world, uint32_t op, uint32_t body, float a, float c0, float c1, float c2, const char *who) {
  if (!b) return pxSetError(-1, "%s: NULL batch", who);
  if (world != PX_ALL_WORLDS && world >= b->L.nWorlds) return pxSetError(-1, "%s: world %u out of range", who, world);
  const bool isStatic = (body & PX_BATCH_STATIC) != 0;
  body &= ~PX_BATCH_STATIC;
  if (body > 255u || (!isStatic && body >= b->L.maxDynamic)) return pxSetError(-1, "%s: body slot %u out of range", who, body);
  if (isStatic && (op < PX_CMD_SET_POSITION || op > PX_CMD_ROTATE)) return pxSetError(-1, "%s: static bodies only take pose commands", who);
  PxCommand c;
  c.world = world;
  c.op = op | (body << 8) | (isStatic ? PX_CMD_STATIC_BIT : 0u);
  c.seq = b->cmdSeq++;
  c.a = a;
  c.b = c0;
  c.c = c1;
  c.d = c2;
  if (world == PX_ALL_WORLDS) {
    b->globalCmds.push_back(c);
  } else {
    b->cmds.push_back(c);
  }
  return 0;
}

extern "C" int pxBatchApplyImpulse(PxBatch *b, uint32_t world, uint32_t body, float ix, float iy, float cx, float cy) {
  if (world == PX_ALL_WORLDS) return pxSetError(-1, "pxBatchApplyImpulse: needs a world index");
  return pushCommand(b, world, PX_CMD_IMPULSE, body, ix, iy, cx, cy, "pxBatchApplyImpulse");
}
extern "C" int pxBatchApplyForce(PxBatch *b, uint32_t world, uint32_t body, float fx, float fy, float cx, float cy) {
  if (world == PX_ALL_WORLDS) return pxSetError(-1, "pxBatchApplyForce: needs a world index");
  return pushCommand(b, world, PX_CMD_FORCE, body, fx, fy, cx, cy, "pxBatchApplyForce");
}
extern "C" int pxBatchApplyImpulses(PxBatch *b, uint32_t n, const uint32_t *worlds, const uint32_t *bodies, const float *impulseXY,
                                    const float *contactXY) {
  if (!worlds || !bodies || !impulseXY) return pxSetError(-1, "pxBatchApplyImpulses: NULL argument");
  for (uint32_t i = 0; i < n; i++) {
    if (worlds[i] == PX_ALL_WORLDS) return pxSetError(-1, "pxBatchApplyImpulses: element %u needs a world index", i);
    int rc = pushCommand(b, worlds[i], PX_CMD_IMPULSE, bodies[i], impulseXY[2 * i], impulseXY[2 * i + 1],
                         contactXY ? contactXY[2 * i] : 0.0f, contactXY ? contactXY[2 * i + 1] : 0.0f, "pxBatchApplyImpulses");
    if (rc) return rc;
  }
  return 0;
}
extern "C" int pxBatchSetPosition(PxBatch *b, uint32_t world, uint32_t body, float x, float y) {
  if (world == PX_ALL_WORLDS) return pxSetError(-1, "pxBatchSetPosition: needs a world index");
  return pushCommand(b, world, PX_CMD_SET_POSITION, body, x, y, 0, 0, "pxBatchSetPosition");
}
extern "C" int pxBatchMoveBy(PxBatch *b, uint32_t world, uint32_t body, float dx, float dy) {
  if (world == PX_ALL_WORLDS) return pxSetError(-1, "pxBatchMoveBy: needs a world index");
  return pushCommand(b, world, PX_CMD_MOVE_BY, body, dx, dy, 0, 0, "pxBatchMoveBy");
}
extern "C" int pxBatchSetOrientation(PxBatch *b, uint32_t world, uint32_t body, float radians) {
  return pushCommand(b, world, PX_CMD_SET_ORIENTATION, body, radians, 0, 0, 0, "pxBatchSetOrientation");
}
extern "C" int pxBatchRotate(PxBatch *b, uint32_t world, uint32_t body, float radians) {
  return pushCommand(b, world, PX_CMD_ROTATE, body, radians, 0, 0, 0, "pxBatchRotate");
}

/* ---- body lifecycle on a resident batch ---- */
static int copyImmWorld(PxBatch *b, uint32_t world) {
  const size_t off = (size_t)world * b->L.immStride;
  CU(cudaMemcpyAsync(b->dImm + off, b->hImm + off, b->L.immStride, cudaMemcpyHostToDevice, b->stream));
  return 0;
}

extern "C" int pxBatchAddBody(PxBatch *b, uint32_t world, const struct PxWorld *hostWorld, const void *body) {
  if (!b || !hostWorld || !body) return pxSetError(-1, "pxBatchAddBody: NULL argument");
  if (world >= b->L.nWorlds) return pxSetError(-1, "pxBatchAddBody: world %u out of range", world);
  if (useDevice(b)) return -1;
  uint8_t *immW = b->hImm + (size_t)world * b->L.immStride;
  if (pxPackBodyIsDynamic(body)) {
    /* earlier copies out of the mirror must have left before it is edited */
    CU(cudaStreamSynchronize(b->stream));
    PxBodyPayload pl;
    uint32_t slot = 0;
    int rc = pxPackAddDynamic(&b->L, hostWorld, body, immW, pl.dyn, &pl.flags, &slot);
    if (rc != 0) return rc;
    if (copyImmWorld(b, world)) return -1;
    b->payloads.push_back(pl);
    return pushCommand(b, world, PX_CMD_ADD_DYNAMIC, slot, (float)(b->payloads.size() - 1), 0, 0, 0, "pxBatchAddBody");
  }
  return pxBatchSyncStatics(b, world, hostWorld);
}

extern "C" int pxBatchFreeBody(PxBatch *b, uint32_t world, uint32_t body, const struct PxWorld *hostWorld) {
  if (!b) return pxSetError(-1, "pxBatchFreeBody: NULL batch");
  if (world >= b->L.nWorlds) return pxSetError(-1, "pxBatchFreeBody: world %u out of range", world);
  if (body & PX_BATCH_STATIC) {
    if (!hostWorld) return pxSetError(-1, "pxBatchFreeBody: freeing a static body needs the host world (its static list is authoritative)");
    return pxBatchSyncStatics(b, world, hostWorld);
  }
  return pushCommand(b, world, PX_CMD_FREE_DYNAMIC, body, 0, 0, 0, 0, "pxBatchFreeBody");
}

extern "C" int pxBatchSyncStatics(PxBatch *b, uint32_t world, const struct PxWorld *hostWorld) {
  if (!b || !hostWorld) return pxSetError(-1, "pxBatchSyncStatics: NULL argument");
  if (world >= b->L.nWorlds) return pxSetError(-1, "pxBatchSyncStatics: world %u out of range", world);
  if (useDevice(b)) return -1;
  if (flushCommands(b)) return -1; /* queued pose commands address statics by their old positions */
  CU(cudaStreamSynchronize(b->stream));
  const PxBatchLayout &L = b->L;
  uint8_t *mutW = b->hMut + (size_t)world * L.mutStride, *immW = b->hImm + (size_t)world * L.immStride;
  int rc = pxPackStatics(&L, hostWorld, mutW, immW);
  if (rc != 0) return rc;
  if (copyImmWorld(b, world)) return -1;
  uint8_t *dMutW = b->dMut + (size_t)world * L.mutStride;
  CU(cudaMemcpyAsync(dMutW + L.offStatMut, mutW + L.offStatMut, (size_t)PX_STATM_ROWS * L.maxStatic * 4, cudaMemcpyHostToDevice, b->stream));
  const size_t offN = L.offHeader + offsetof(PxWorldHeader, nStatic);
  CU(cudaMemcpyAsync(dMutW + offN, mutW + offN, sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
  return 0;
}

extern "C" uint64_t pxBatchAllocCount(const PxBatch *b) { return b ? b->allocs : 0; }

extern "C" int pxBatchSetGravity(PxBatch *b, uint32_t world, float x, float y) {
  return pushCommand(b, world, PX_CMD_GRAVITY, 0, x, y, 0, 0, "pxBatchSetGravity");
}
extern "C" int pxBatchSetIterations(PxBatch *b, uint32_t world, uint32_t iterations) {
  return pushCommand(b, world, PX_CMD_ITERATIONS, 0, (float)(iterations & 255u), 0, 0, 0, "pxBatchSetIterations");
}

static int flushCommands(PxBatch *b) {
  const size_t n = b->cmds.size(), ng = b->globalCmds.size();
  if (n == 0 && ng == 0) return 0;
  const uint32_t W = b->L.nWorlds;
  if (n > 0) {
    /* counting sort by world keeps submission order inside a world */
    CU(cudaStreamSynchronize(b->stream)); /* pinned buffers may be in flight */
    if (b->hCmdCap < n) {
      cudaFreeHost(b->hCmds);
      b->hCmds = nullptr;
      b->hCmdCap = std::max(n, b->hCmdCap * 2);
      CU(cudaMallocHost(&b->hCmds, b->hCmdCap * sizeof(PxCommand)));
      b->allocs++;
    }
    memset(b->hWorldStart, 0, (size_t)(W + 1) * sizeof(uint32_t));
    for (const PxCommand &c : b->cmds) b->hWorldStart[c.world + 1]++;
    for (uint32_t w = 0; w < W; w++) b->hWorldStart[w + 1] += b->hWorldStart[w];
    std::vector<uint32_t> at(b->hWorldStart, b->hWorldStart + W);
    for (const PxCommand &c : b->cmds) b->hCmds[at[c.world]++] = c;
    if (b->dCmdCap < n) {
      cudaFree(b->dCmds);
      b->dCmds = nullptr;
      b->dCmdCap = b->hCmdCap;
      CU(cudaMalloc(&b->dCmds, b->dCmdCap * sizeof(PxCommand)));
      b->allocs++;
    }
    CU(cudaMemcpyAsync(b->dCmds, b->hCmds, n * sizeof(PxCommand), cudaMemcpyHostToDevice, b->stream));
    CU(cudaMemcpyAsync(b->dWorldStart, b->hWorldStart, (size_t)(W + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
  }
  if (ng > 0) {
    if (b->dGlobalCap < ng) {
      cudaFree(b->dGlobal);
      b->dGlobal = nullptr;
      b->dGlobalCap = std::max(ng, b->dGlobalCap * 2);
      CU(cudaMalloc(&b->dGlobal, b->dGlobalCap * sizeof(PxCommand)));
      b->allocs++;
    }
    /* pageable source: the copy is staged by the runtime before the call returns */
    CU(cudaMemcpyAsync(b->dGlobal, b->globalCmds.data(), ng * sizeof(PxCommand), cudaMemcpyHostToDevice, b->stream));
  }
  const size_t np = b->payloads.size();
  if (np > 0) {
    if (b->dPayloadCap < np) {
      CU(cudaStreamSynchronize(b->stream));
      cudaFree(b->dPayloads);
      b->dPayloads = nullptr;
      b->dPayloadCap = std::max(np, b->dPayloadCap * 2);
      CU(cudaMalloc(&b->dPayloads, b->dPayloadCap * sizeof(PxBodyPayload)));
      b->allocs++;
    }
    /* pageable source: staged by the runtime before the call returns */
    CU(cudaMemcpyAsync(b->dPayloads, b->payloads.data(), np * sizeof(PxBodyPayload), cudaMemcpyHostToDevice, b->stream));
  }
  int rc = pxLaunchCommands(&b->L, b->dMut, b->dImm, n ? b->dCmds : nullptr, n ? b->dWorldStart : nullptr,
                            ng ? b->dGlobal : nullptr, (uint32_t)ng, b->dPayloads, b->dSin, b->stream);
  if (rc != 0) return pxSetError(-rc - 1000, "command kernel: %s", cudaGetErrorString((cudaError_t)rc));
  b->launches++;
  b->cmds.clear();
  b->globalCmds.clear();
  b->payloads.clear();
  return 0;
}

/* --------------------------------------------------------------------------- stepping */
static void baseParams(PxBatch *b, uint32_t kSubsteps, uint32_t first, uint32_t n, PxStepParams *P) {
  memset(P, 0, sizeof(*P));
  P->L = b->L;
  P->L.nWorlds = n;
  P->S = b->plan;
  P->mut = b->dMut + (size_t)first * b->L.mutStride;
  P->imm = b->dImm + (size_t)first * b->L.immStride;
  P->trace = b->dTrace ? b->dTrace + (size_t)first * b->L.traceStride : nullptr;
  P->sinTable = b->dSin;
  P->prof = b->dProf ? b->dProf + (size_t)first * PX_PROF_SLOTS : nullptr;
  P->kSubsteps = kSubsteps;
  P->flags = b->flags;
  P->mode = PX_MODE_FUSED;
  static const char *ns = getenv("PX_LAG_SLEEP_NS"); /* tuning switch (500 ns: throughput is flat from 64 to 2000) */
  P->lagSleepNs = ns ? (uint32_t)atoi(ns) : 500u;
}

static void splitParams(PxBatch *b, uint32_t first, PxStepParams *P) {
  P->mode = PX_MODE_SWEEP;
  P->solver = b->staticSched ? PX_SOLVER_STATIC : PX_SOLVER_DYNAMIC;
  if (!b->staticSched) P->scratch = nullptr; /* static solver: the sweep kernel's manifold records live in the launch's scratch region */
  P->hand = b->dHand + (size_t)first * b->hand.stride;
  P->H = b->hand;
  P->solveS = b->solvePlan;
}

/* launch s of the split pipeline's sweep kernel: phase F of substep s - 1 (s > 0), phases A-D of substep s (s < K) */
static int launchSweep(PxBatch *b, PxStepParams *P, uint32_t s, uint32_t kSubsteps, uint32_t grid, cudaStream_t stream) {
  P->doFinish = s > 0 ? 1u : 0u;
  P->doSweep = s < kSubsteps ? 1u : 0u;
  P->traceNow = (s + 1 == kSubsteps) ? 1u : 0u;
  P->worldCounter = b->dCounters + (b->launches % PX_COUNTER_RING);
  CU(cudaMemsetAsync(P->worldCounter, 0, sizeof(uint32_t), stream));
  int rc = pxLaunchStep(P, b->threads, grid, stream);
  if (rc != 0) return pxSetError(-rc - 1000, "sweep kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
  b->launches++;
  return 0;
}

static int launchSolve(PxBatch *b, PxStepParams *P, uint32_t grid, cudaStream_t stream) {
  P->worldCounter = b->dCounters + (b->launches % PX_COUNTER_RING);
  CU(cudaMemsetAsync(P->worldCounter, 0, sizeof(uint32_t), stream));
  int rc = pxLaunchSolve(P, grid, stream);
  if (rc != 0) return pxSetError(-rc - 1000, "solver kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
  b->launches++;
  return 0;
}

/* static solver: phase F of a step's last substep */
static int launchFinish(PxBatch *b, PxStepParams *P, cudaStream_t stream) {
  int rc = pxLaunchFinish(P, stream);
  if (rc != 0) return pxSetError(-rc - 1000, "finish kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
  b->launches++;
  return 0;
}

/* K substeps for worlds [first, first + n) on `stream`; launches that may overlap in time use
 * different scratch regions */
static int launchRange(PxBatch *b, uint32_t kSubsteps, uint32_t first, uint32_t n, cudaStream_t stream, uint32_t region) {
  PxStepParams P;
  baseParams(b, kSubsteps, first, n, &P);
  P.scratch = b->dScratch ? b->dScratch + (size_t)region * b->gridMax * b->L.maxManifolds * 3 : nullptr;
  P.worldCounter = b->dCounters + (b->launches % PX_COUNTER_RING);
  if (!b->split) {
    P.mode = PX_MODE_FUSED;
    CU(cudaMemsetAsync(P.worldCounter, 0, sizeof(uint32_t), stream));
    const uint32_t grid = n < b->gridMax ? n : b->gridMax;
    int rc = pxLaunchStep(&P, b->threads, grid, stream);
    if (rc != 0) return pxSetError(-rc - 1000, "step kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
    b->launches++;
    return 0;
  }
  /* split pipeline: sweep(A-D) | solve(E) | sweep(F, A-D) | solve(E) | ... | F */
  splitParams(b, first, &P);
  const uint32_t gridSweep = n < b->gridMax ? n : b->gridMax, gridSolve = n < b->gridSolve ? n : b->gridSolve;
  /* the last substep's phase F: static solver -- the streaming finish kernel; dataflow solver -- a sweep launch that
   * only finishes */
  const uint32_t nSweeps = b->staticSched ? kSubsteps : kSubsteps + 1u;
  for (uint32_t s = 0; s < nSweeps; s++) {
    if (launchSweep(b, &P, s, kSubsteps, gridSweep, stream)) return -1;
    if (s < kSubsteps && launchSolve(b, &P, gridSolve, stream)) return -1;
  }
  if (b->staticSched && launchFinish(b, &P, stream)) return -1;
  return 0;
}

/* the two helper streams (host-buffer steps pipeline their world ranges over them) */
static int ensureAux(PxBatch *b) {
  if (b->auxReady) return 0;
  for (int i = 0; i < PX_AUX; i++) {
    CU(cudaStreamCreateWithFlags(&b->aux[i], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&b->evJoin[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&b->evFork, cudaEventDisableTiming));
  b->auxReady = true;
  return 0;
}

extern "C" int pxBatchStep(PxBatch *b, uint32_t kSubsteps) {
  if (!b) return pxSetError(-1, "pxBatchStep: NULL batch");
  if (useDevice(b)) return -1;
  if (flushCommands(b)) return -1;
  if (kSubsteps == 0) return 0;
  /* Split pipeline, large batches: the two halves of the batch are stepped on the two helper streams.  Each kernel of
   * a half ends in a partial wave of resident CTAs (8 192 worlds are 2.8 waves of the solver kernel); with two
   * independent chains of launches the other half's next kernel fills those SMs.  PX_OVERLAP=<n>: n ranges (1 = off). */
  static const char *ov = getenv("PX_OVERLAP");
  const uint32_t nRanges = (ov && atoi(ov) >= 1 && atoi(ov) <= PX_AUX) ? (uint32_t)atoi(ov) : 2u;
  if (b->split && b->L.nWorlds >= 12u * (uint32_t)b->numSms && nRanges > 1) {
    if (ensureAux(b)) return -1;
    CU(cudaEventRecord(b->evFork, b->stream));
    const uint32_t W = b->L.nWorlds;
    for (uint32_t i = 0; i < nRanges; i++) {
      const uint32_t first = (uint32_t)((uint64_t)W * i / nRanges), last = (uint32_t)((uint64_t)W * (i + 1) / nRanges);
      CU(cudaStreamWaitEvent(b->aux[i], b->evFork, 0));
      if (launchRange(b, kSubsteps, first, last - first, b->aux[i], 1u + i)) return -1;
      CU(cudaEventRecord(b->evJoin[i], b->aux[i]));
      CU(cudaStreamWaitEvent(b->stream, b->evJoin[i], 0));
    }
    return 0;
  }
  return launchRange(b, kSubsteps, 0, b->L.nWorlds, b->stream, 0);
}

extern "C" int pxBatchStepHost(PxBatch *b, uint32_t kSubsteps, void *hostMut, uint32_t chunks) {
  if (!b) return pxSetError(-1, "pxBatchStepHost: NULL batch");
  if (useDevice(b)) return -1;
  if (flushCommands(b)) return -1;
  uint8_t *host = hostMut ? (uint8_t *)hostMut : b->hMut;
  const uint32_t W = b->L.nWorlds;
  if (chunks == 0) chunks = 5;
  if (chunks > W) chunks = W;
  /* world ranges: the first and the last get half the worlds of an inner one -- nothing overlaps the first range's
   * H2D copy or the last range's D2H copy, so those two are kept short */
  const uint32_t units = chunks >= 3 ? 2u * (chunks - 1u) : chunks;
  auto edge = [&](uint32_t c) -> uint32_t { /* first world of range c; c == chunks: W */
    if (c == 0) return 0;
    if (c >= chunks) return W;
    const uint32_t u = chunks >= 3 ? 2u * c - 1u : c;
    return (uint32_t)((uint64_t)W * u / units);
  };
  if (ensureAux(b)) return -1;
  /* fork: the helper streams start after everything already queued on the batch's stream */
  CU(cudaEventRecord(b->evFork, b->stream));
  static const char *hs = getenv("PX_HOST_STREAMS"); /* tuning switch */
  const uint32_t nStreams = (hs && atoi(hs) >= 1 && atoi(hs) <= PX_AUX) ? (uint32_t)atoi(hs) : (uint32_t)PX_AUX;
  for (uint32_t i = 0; i < nStreams; i++) CU(cudaStreamWaitEvent(b->aux[i], b->evFork, 0));
  const size_t stride = b->L.mutStride;
  for (uint32_t c = 0; c < chunks; c++) {
    const uint32_t first = edge(c), last = edge(c + 1);
    const uint32_t n = last - first;
    if (!n) continue;
    cudaStream_t st = b->aux[c % nStreams];
    CU(cudaMemcpyAsync(b->dMut + first * stride, host + first * stride, n * stride, cudaMemcpyHostToDevice, st));
    if (kSubsteps && launchRange(b, kSubsteps, first, n, st, 1u + c % nStreams)) return -1;
    CU(cudaMemcpyAsync(host + first * stride, b->dMut + first * stride, n * stride, cudaMemcpyDeviceToHost, st));
  }
  /* join: later work on the batch's stream (and pxBatchSync) waits for the helpers */
  for (uint32_t i = 0; i < nStreams; i++) {
    CU(cudaEventRecord(b->evJoin[i], b->aux[i]));
    CU(cudaStreamWaitEvent(b->stream, b->evJoin[i], 0));
  }
  return 0;
}

extern "C" int pxBatchStepTimed(PxBatch *b, uint32_t kSubsteps, float *ms) {
  if (!b) return pxSetError(-1, "pxBatchStepTimed: NULL batch");
  if (useDevice(b)) return -1;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaStreamSynchronize(b->stream));
  CU(cudaEventRecord(e0, b->stream));
  int rc = pxBatchStep(b, kSubsteps);
  CU(cudaEventRecord(e1, b->stream));
  cudaError_t err = cudaEventSynchronize(e1);
  float t = 0;
  if (err == cudaSuccess) err = cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != 0) return rc;
  if (err != cudaSuccess) return pxSetError(-(int)err - 1000, "pxBatchStepTimed: %s", cudaGetErrorString(err));
  if (ms) *ms = t;
  return 0;
}

/* device time of whatever is queued on the batch's stream between the two calls */
extern "C" int pxBatchTimerStart(PxBatch *b) {
  if (!b) return pxSetError(-1, "pxBatchTimerStart: NULL batch");
  if (useDevice(b)) return -1;
  if (!b->timerReady) {
    CU(cudaEventCreate(&b->evTimer[0]));
    CU(cudaEventCreate(&b->evTimer[1]));
    b->timerReady = true;
  }
  CU(cudaEventRecord(b->evTimer[0], b->stream));
  return 0;
}
extern "C" int pxBatchTimerStop(PxBatch *b, float *ms) {
  if (!b || !b->timerReady) return pxSetError(-1, "pxBatchTimerStop: no timer was started");
  if (useDevice(b)) return -1;
  CU(cudaEventRecord(b->evTimer[1], b->stream));
  CU(cudaEventSynchronize(b->evTimer[1]));
  float t = 0;
  CU(cudaEventElapsedTime(&t, b->evTimer[0], b->evTimer[1]));
  if (ms) *ms = t;
  return 0;
}

/* One step with every kernel launch bracketed by CUDA events: device time per kernel class.
 * ms[c] / n[c] = summed durations and number of the launches of class c: 0 sweep (or the one fused launch),
 * 1 finish kernel (the last substep's phase F, static solver only), 2 solver (dataflow or schedule + replay).  Blocks. */
extern "C" int pxBatchStepTimedKernels(PxBatch *b, uint32_t kSubsteps, float ms[3], uint32_t n[3]) {
  if (!b || !ms || !n) return pxSetError(-1, "pxBatchStepTimedKernels: NULL argument");
  if (useDevice(b)) return -1;
  if (flushCommands(b)) return -1;
  for (int c = 0; c < 3; c++) {
    ms[c] = 0;
    n[c] = 0;
  }
  if (!kSubsteps) return 0;
  const uint32_t W = b->L.nWorlds;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  auto mark = [&](int c) -> int {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    CU(cudaEventRecord(e, b->stream));
    ev.push_back(e);
    cls.push_back(c);
    return 0;
  };
  PxStepParams P;
  baseParams(b, kSubsteps, 0, W, &P);
  CU(cudaStreamSynchronize(b->stream));
  if (mark(-1)) return -1;
  if (!b->split) {
    if (launchRange(b, kSubsteps, 0, W, b->stream, 0)) return -1;
    if (mark(0)) return -1;
  } else {
    P.scratch = b->dScratch; /* region 0 */
    splitParams(b, 0, &P);
    const uint32_t gridSweep = W < b->gridMax ? W : b->gridMax, gridSolve = W < b->gridSolve ? W : b->gridSolve;
    const uint32_t nSweeps = b->staticSched ? kSubsteps : kSubsteps + 1u;
    for (uint32_t s = 0; s < nSweeps; s++) {
      if (launchSweep(b, &P, s, kSubsteps, gridSweep, b->stream) || mark(0)) return -1;
      if (s < kSubsteps && (launchSolve(b, &P, gridSolve, b->stream) || mark(2))) return -1;
    }
    if (b->staticSched && (launchFinish(b, &P, b->stream) || mark(1))) return -1;
  }
  CU(cudaEventSynchronize(ev.back()));
  for (size_t k = 1; k < ev.size(); k++) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, ev[k - 1], ev[k]));
    ms[cls[k]] += t;
    n[cls[k]]++;
  }
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

/* 1 if the batch runs the split pipeline (sweep kernel + solver kernel per substep), 0 if fused;
 * *solveCtasPerSm = residency of the solver kernel */
extern "C" int pxBatchPipelineInfo(const PxBatch *b, int *split, int *solveCtasPerSm, int *schedCtasPerSm) {
  if (!b) return pxSetError(-1, "pxBatchPipelineInfo: NULL batch");
  if (split) *split = b->split ? (b->staticSched ? 2 : 1) : 0;
  if (solveCtasPerSm) *solveCtasPerSm = b->solveCtasPerSm;
  if (schedCtasPerSm) *schedCtasPerSm = 0;
  return 0;
}

extern "C" int pxBatchProfileD2H(PxBatch *b, uint64_t *host) {
  if (!b || !host) return pxSetError(-1, "pxBatchProfileD2H: NULL argument");
  if (!b->dProf) return pxSetError(-1, "pxBatchProfileD2H: the batch was created without PX_BATCH_FLAG_PROFILE");
  if (useDevice(b)) return -1;
  CU(cudaMemcpyAsync(host, b->dProf, (size_t)b->L.nWorlds * PX_PROF_SLOTS * 8, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  return 0;
}

/* ---------------------------------------------------------------------------- self-tests */
extern "C" int pxSelfTestMath(int op, const float *a, const float *b, const float *sinTable256, float *out, uint32_t n) {
  if (!a || !out || n == 0 || op < 0 || op > 5) return pxSetError(-1, "pxSelfTestMath: bad argument");
  float *dA = nullptr, *dB = nullptr, *dT = nullptr, *dO = nullptr;
  float table[256];
  if (sinTable256) {
    memcpy(table, sinTable256, sizeof(table));
  } else {
    for (int i = 0; i < 256; i++) table[i] = sinf((i * 3.141592741f * 2.0f) / 256);
  }
  int rc = 0;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMalloc(&dA, n * sizeof(float))) != cudaSuccess) break;
    if ((e = cudaMalloc(&dB, n * sizeof(float))) != cudaSuccess) break;
    if ((e = cudaMalloc(&dO, n * sizeof(float))) != cudaSuccess) break;
    if ((e = cudaMalloc(&dT, sizeof(table))) != cudaSuccess) break;
    if ((e = cudaMemcpy(dA, a, n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = cudaMemcpy(dB, b ? b : a, n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = cudaMemcpy(dT, table, sizeof(table), cudaMemcpyHostToDevice)) != cudaSuccess) break;
    if ((e = (cudaError_t)pxLaunchMath(op, dA, dB, dT, dO, n, nullptr)) != cudaSuccess) break;
    if ((e = cudaMemcpy(out, dO, n * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
  } while (0);
  if (e != cudaSuccess) rc = pxSetError(-(int)e - 1000, "pxSelfTestMath: %s", cudaGetErrorString(e));
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dO);
  cudaFree(dT);
  return rc;
}

extern "C" int pxSelfTestRsqrtSweep(uint64_t *sums4096) {
  if (!sums4096) return pxSetError(-1, "pxSelfTestRsqrtSweep: NULL argument");
  unsigned long long *d = nullptr;
  cudaError_t e = cudaMalloc(&d, 4096 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = (cudaError_t)pxLaunchRsqrtSweep(d, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(sums4096, d, 4096 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return pxSetError(-(int)e - 1000, "pxSelfTestRsqrtSweep: %s", cudaGetErrorString(e));
  return 0;
}

/* ------------------------------------------------------------------------------ queries */
extern "C" int pxBatchGetLayout(const PxBatch *b, PxBatchLayout *out) {
  if (!b || !out) return pxSetError(-1, "pxBatchGetLayout: NULL argument");
  *out = b->L;
  return 0;
}
extern "C" void *pxBatchHostMutable(PxBatch *b) { return b ? b->hMut : nullptr; }
extern "C" void *pxBatchHostImmutable(PxBatch *b) { return b ? b->hImm : nullptr; }
extern "C" uint64_t pxBatchDeviceMutable(PxBatch *b) { return b ? (uint64_t)(uintptr_t)b->dMut : 0; }
extern "C" uint64_t pxBatchDeviceImmutable(PxBatch *b) { return b ? (uint64_t)(uintptr_t)b->dImm : 0; }
extern "C" uint64_t pxBatchLaunchCount(const PxBatch *b) { return b ? b->launches : 0; }
extern "C" int pxBatchKernelInfo(const PxBatch *b, int *ctasPerSm, int *smemBytes, int *threads, int *numSms) {
  if (!b) return pxSetError(-1, "pxBatchKernelInfo: NULL batch");
  if (ctasPerSm) *ctasPerSm = b->ctasPerSm;
  if (smemBytes) *smemBytes = (int)b->plan.total;
  if (threads) *threads = (int)b->threads;
  if (numSms) *numSms = b->numSms;
  return 0;
}
