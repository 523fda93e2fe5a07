/* px_oracle.h -- C interface of the CPU restatement oracle.  TEST INFRASTRUCTURE ONLY
 * (see the header of px_oracle.c for who may call it and how it is pinned). */
#ifndef PX_ORACLE_H
#define PX_ORACLE_H

#include <stdint.h>

#include "../include/px_batch.h"

#define ORACLE_API __attribute__((visibility("default")))

/* k substeps (pxWorldStepWithDt, reference src/px_world.c:410-420) for worlds
 * [first, first + n) of the block pair (mut, imm) described by L.  `trace`, if not NULL,
 * receives n trace blocks (L->traceStride each) describing the LAST substep. */
ORACLE_API void pxOracleSubsteps(const PxBatchLayout *L, void *mut, const void *imm, const float *sinTable,
                                 uint32_t first, uint32_t n, uint32_t k, void *trace);

/* wall seconds for k substeps of nWorlds worlds on nThreads pthreads (CPU "port" baseline) */
ORACLE_API double pxOracleTimeSubsteps(const PxBatchLayout *L, void *mut, const void *imm, const float *sinTable,
                                       uint32_t nWorlds, uint32_t k, uint32_t nThreads);

/* op: 0 rsqrt 1 rcp 2 sqrt 3 div 4 sin 5 cos */
ORACLE_API void pxOracleMath(int op, const float *a, const float *b, const float *sinTable, float *out, uint32_t n);

ORACLE_API void pxOracleRsqrtSweep(uint64_t *sums, uint32_t first, uint32_t n);

#endif
