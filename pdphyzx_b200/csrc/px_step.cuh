/*
 * px_step.cuh -- parameter block and shared-memory plan of the world-step kernel
 * (px_step.cu), shared with the C-ABI layer (px_batch.cu).
 */
#ifndef PX_STEP_CUH
#define PX_STEP_CUH

#include <stdint.h>

#include "px_batch.h"

/* Byte offsets of every shared-memory array of one world-CTA.  Computed once on the host
 * (pxPlanSmem) from the batch layout and handed to the kernel by value.
 * Body-id space: dynamic slot s -> id s, static position q -> id D + q (NB = D + S). */
struct PxSmemPlan {
  /* persistent across the K fused substeps: float4 records by body id */
  uint32_t bodyP, bodyV, bodyR;
  uint32_t shape;                                                   /* u32[NB] PX_SHAPE_PACK word by body id */
  uint32_t flags, order, order2, posOf, restFlag, wakeFlag, sSlot; /* u8[D] */
  uint32_t hdr;                                                     /* PxWorldHeader */
  uint32_t misc;                                                    /* u32[32] counters + round descriptors */
  uint32_t scan;                                                    /* u64[8] block-scan scratch */
  /* manifolds of the current substep: one id word per slot here; the three float4 records per
   * slot live in the CTA's L2-resident scratch slot (PxStepParams.scratch) */
  uint32_t mids;
  /* sweep -> list build */
  uint32_t validSlot; /* u16[MP] manifold slot + 1 by pair sequence number */
  uint32_t base;      /* u16[D+1] first pair sequence number by sorted position */
  uint32_t listStart; /* u16[D+1] by slot */
  uint32_t cnt, cntB; /* u16[D] manifolds per body / of which as bodyB, by slot */
  uint32_t fillB;     /* u32[D] atomic counters */
  /* region X (sweep + narrowphase), aliased with region Y (solver) */
  uint32_t sBox;  /* float4[D] AABB by sorted position */
  uint32_t sInfo; /* u32[D] slot | sleeping << 8 | polygon << 9 by sorted position */
  uint32_t stBox; /* float4[S] */
  uint32_t bins;  /* u32[max(MP, D)] pairs binned by shape-pair type (sort keys in phase A) */
  uint32_t verts; /* float4[V] (only if vertsInSmem) */
  /* region Y */
  uint32_t list;  /* u16[2*MM] per-body manifold lists (CSR payload), persists to phase F */
  uint32_t nextw; /* u16[2*MM] successor of manifold ms on side s: slot | successor has a dynamic B << 14 | side << 15 */
  uint32_t sig;   /* u32[MM]   signals received: A side low half, B side high half */
  uint32_t queue; /* u32[QCAP] ready ring of rich manifold words */
  /* region Y' (static schedule, split pipeline): overlays nextw / sig / queue, which only the dynamic solver uses */
  uint32_t predA, predB; /* u16[MM] BY POOL RANK: predecessor of the manifold in its A / B body's ring: rank | closes the ring << 10 |
                            contacts of the predecessor << 11 | own contacts << 13 (0xffff: static B) */
  uint32_t rankOf;       /* u16[MM] pool rank by manifold slot */
  uint32_t cm;           /* float4[MM] positional-correction moves of a manifold {A's.x, A's.y, B's.x, B's.y} (phase C4 -> D) */
  /* pxReplayKernel's own plan (pxPlanSolveSmem, static solver).  While the schedule is built: */
  uint32_t tlev;         /* u16[MM] by pool rank: start round of the manifold's first contact inside one iteration | contacts << 14 */
  uint32_t khist;        /* u16[PX_KSM + 2] visits per (slot, phase) key, then their exclusive prefix = segment starts / scatter cursors */
  uint32_t vkeys;        /* u16[2 * MM] key of every contact visit (0xffff: a manifold's absent second contact) */
  /* ... and, in the same bytes, while it is replayed: bodyV above = float4[NB] {vel.x, vel.y, angularVelocity, inverse mass} */
  uint32_t bodyI;        /* f32[NB] inverse inertia */
  uint32_t segS;         /* u16[PX_SEG_SMEM] the segment table's copy */
  uint32_t total; /* bytes */
  uint32_t qcap;  /* ring capacity (power of two >= MM) */
  uint32_t vertsInSmem;
};
/* static schedule: contact visits per world <= 2 * maxManifolds; (slot, phase) keys <= depth + period <= 4 * maxManifolds + margin */
#define PX_VCAP(MM) (2u * (MM))
#define PX_KCAP(MM) (4u * (MM) + 16u) /* entries of the segment table in the hand-over record */
#define PX_KSM(MM) (2u * (MM) + 64u) /* (slot, phase) keys pxReplayKernel's counting sort has room for in shared memory; a period whose keys
                                        do not fit is skipped (the sequential schedule's always fit) */
#define PX_SEG_SMEM 512u /* segment-table entries pxReplayKernel keeps in shared memory (longer tables are read from global memory) */
#define PX_PLANE 2048u /* entries between the operand planes of the visit records: >= PX_VCAP of the largest maxManifolds (1024), a
                          constant so that one address register serves the three loads of a visit */
#define PX_II_MARGIN 3u /* smallest margin: the first period tried for a world is its longest body ring + max(this, the margin that
                           worked last substep).  profiles/sched_sim.py: on settled 252-body piles the least feasible period is
                           ring + 0..5, ring + 2 on average; every extra round of period costs iterations - 1 replay rounds, a
                           failed attempt costs a relaxation run to its pass cap */

/* A world's hand-over record in global memory (split pipeline): byte offsets of what the sweep
 * kernel leaves for the solver kernel and for the phase F that follows it.  Laid out by
 * pxPlanHandover. */
struct PxHandLayout {
  uint32_t mids, nextw, sig, queue;                              /* solver inputs */
  uint32_t list, listStart, cnt, cntB, posOf, restFlag, wakeFlag; /* phase F inputs */
  uint32_t scalars;                                               /* u32[8], see PX_HS_* */
  uint32_t records;                                               /* float4[3][maxManifolds] manifold records */
  /* static schedule: the contact visits of ONE iteration sorted by (round mod period, round div period), as three
   * float4 planes of PX_VCAP entries {n, ra | rb, rcpDenom, e | sf, df, ids, -}, and the segment starts by key */
  uint32_t stream, seg;
  uint32_t predA, predB; /* u16[maxManifolds] rings by pool rank (PxSmemPlan.predA) */
  uint32_t vrec;         /* the visit operands by pool rank (entry 2 * rank + contact), same three planes as `stream` */
  /* static solver: phase F runs in pxReplayKernel.  Per dynamic slot: first entry of its positional-correction moves in
   * cvec | how many << 16 | its sorted position's wake flag << 30 | rest flag << 31 (px_world.c:179-219 tests them by position) */
  uint32_t bodyF;        /* u32[maxDynamic] */
  uint32_t cvec;         /* float2[2 * maxManifolds] the moves, per body in pool order (px_world.c:288-292) */
  uint32_t stride;
};
enum { PX_MODE_FUSED = 0, PX_MODE_SWEEP = 1 };
enum { PX_SOLVER_DYNAMIC = 0, PX_SOLVER_STATIC = 1 };
/* hand-over scalars: manifolds kept | dataflow solver: visits ready at the start (bit 31: the replay kernel's "schedule
 * did not converge", which cannot happen) | 2-5 unused | longest ring, in contact visits | written by pxReplayKernel
 * only and kept from substep to substep: the period margin that worked last time */
enum { PX_HS_M = 0, PX_HS_READY = 1, PX_HS_RING = 6, PX_HS_HINT = 7 };

#define PX_PROF_SLOTS 16 /* 0-5 phases A-F, 6 solver rounds, 7 polygon pairs << 32 | after quick reject, 8-11 narrowphase stages,
                            12 static schedule (end of phase D), 13 relaxation votes << 32 | period, 14 sort + stream build cycles,
                            15 relaxation cycles */

struct PxStepParams {
  PxBatchLayout L;
  PxSmemPlan S;
  uint8_t *mut;
  const uint8_t *imm;
  uint8_t *trace; /* NULL unless PX_BATCH_FLAG_TRACE */
  const float *sinTable;
  unsigned long long *prof; /* NULL, or u64[nWorlds][PX_PROF_SLOTS] SM-clock cycles per phase (PX_BATCH_FLAG_PROFILE) */
  /* manifold records {normal, e, sf | df, pen, c0 | c1, rcpDenom0, rcpDenom1}: float4[3] per manifold,
   * maxManifolds per slot, one slot per CTA of the (persistent, one-wave) grid: CTA i owns slot i of
   * this launch's region for its lifetime, so the same few MB are rewritten every substep and stay
   * in L2.  Launches that may run concurrently get different regions. */
  float4 *scratch;
  uint32_t *worldCounter; /* zeroed before the launch: the CTAs take worlds off it */
  uint32_t kSubsteps;
  uint32_t flags;
  uint32_t lagSleepNs; /* scheduler warp: nanosleep while it is PX_LAG rounds ahead of the math warp */
  /* split pipeline: pxStepKernel in PX_MODE_SWEEP does [phase F of the previous substep] + [phases
   * A-D of the next], pxSolveKernel does phase E in between */
  uint32_t mode;
  uint32_t solver; /* split pipeline: PX_SOLVER_DYNAMIC (pxSolveKernel) or PX_SOLVER_STATIC (schedule built in phase D, pxReplayKernel) */
  uint32_t doFinish, doSweep, traceNow;
  uint8_t *hand; /* hand-over records, one per world, H.stride apart */
  PxHandLayout H;
  PxSmemPlan solveS; /* shared memory of pxSolveKernel (bodyP, bodyV, mids, nextw, sig, queue, misc, hdr) or of pxReplayKernel */
};


#ifdef __cplusplus
extern "C" {
#endif
/* vertsInSmem: 1 keeps the polygon vertex pool in shared memory, 0 reads it through L1 */
void pxPlanSmem(const PxBatchLayout *L, int vertsInSmem, int staticSolver, PxSmemPlan *S);
/* launches the step kernel: `grid` persistent CTAs (<= resident capacity, <= scratch slots) work
 * through all worlds of the layout; returns cudaError_t as int */
int pxLaunchStep(const PxStepParams *P, uint32_t threads, uint32_t grid, void *stream);
/* occupancy query: resident CTAs per SM for this plan */
int pxStepOccupancy(const PxStepParams *P, uint32_t threads, int *ctasPerSm);
/* split pipeline: hand-over layout, the solver kernel's shared-memory plan, its launch and occupancy */
void pxPlanHandover(const PxBatchLayout *L, const PxSmemPlan *S, PxHandLayout *H);
void pxPlanSolveSmem(const PxBatchLayout *L, const PxSmemPlan *S, int staticSolver, PxSmemPlan *out);
int pxLaunchSolve(const PxStepParams *P, uint32_t grid, void *stream);
/* static solver: phase F of a step's last substep as a streaming kernel of its own, one CTA per world */
int pxLaunchFinish(const PxStepParams *P, void *stream);
int pxSolveOccupancy(const PxStepParams *P, int *ctasPerSm);
#define PX_SOLVE_THREADS 64u  /* dynamic solver: scheduler warp + math warp */
#define PX_REPLAY_THREADS 32u /* static solver: one warp derives the world's schedule and replays it */


/* control path: one command = one API call (px_batch.h "control path") */
struct PxCommand {
  uint32_t world; /* index or PX_ALL_WORLDS */
  uint32_t op;    /* PX_CMD_* | body << 8 */
  uint32_t seq;   /* submission order */
  float a, b, c, d;
};
enum {
  PX_CMD_IMPULSE = 1, PX_CMD_FORCE = 2, PX_CMD_GRAVITY = 3, PX_CMD_ITERATIONS = 4,
  PX_CMD_SET_POSITION = 5, PX_CMD_MOVE_BY = 6, PX_CMD_SET_ORIENTATION = 7, PX_CMD_ROTATE = 8, /* body may be static */
  PX_CMD_ADD_DYNAMIC = 9,  /* a = index into the payload array; the slot is the body field */
  PX_CMD_FREE_DYNAMIC = 10
};
#define PX_CMD_STATIC_BIT 0x10000u /* op word: PX_CMD_* | slot << 8 | PX_CMD_STATIC_BIT if the slot is a static body's */
/* mutable state of a body added to a resident world (PX_CMD_ADD_DYNAMIC) */
struct PxBodyPayload {
  float dyn[PX_DYN_ROWS];
  uint32_t flags;
};
/* cmds sorted by (world, seq); worldStart[w]..worldStart[w+1] is world w's range; the
 * `global` list (PX_ALL_WORLDS, in seq order) is merged in by sequence number */
int pxLaunchCommands(const PxBatchLayout *L, uint8_t *mut, const uint8_t *imm, const struct PxCommand *cmds,
                     const uint32_t *worldStart, const struct PxCommand *global, uint32_t nGlobal,
                     const struct PxBodyPayload *payloads, const float *sinTable, void *stream);
/* arithmetic self-tests (device pointers) */
int pxLaunchMath(int op, const float *a, const float *b, const float *sinTable, float *out, uint32_t n, void *stream);
int pxLaunchRsqrtSweep(unsigned long long *sums4096, void *stream);
#ifdef __cplusplus
}
#endif

#endif
