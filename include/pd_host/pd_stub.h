/*
 * pd_stub.h -- the concrete PlaydateAPI host stub instance.
 *
 * pdStubAPI() returns a process-wide PlaydateAPI whose allocator is libc realloc/free,
 * whose clock is a settable millisecond counter (deterministic frames for
 * px->world->step, reference src/px_clock.c:25-42), and whose graphics calls only
 * count invocations (reference src/px_world.c:451-499 is debug drawing).
 */
#ifndef PD_STUB_H
#define PD_STUB_H

#include "pd_api.h"

#ifdef __cplusplus
extern "C" {
#endif

PlaydateAPI *pdStubAPI(void);
void pdStubSetTimeMs(uint32_t ms);
void pdStubAdvanceTimeMs(uint32_t ms);
uint32_t pdStubGetTimeMs(void);
/* number of graphics calls since the last reset; last drawText string (may be NULL) */
unsigned pdStubDrawCalls(void);
const char *pdStubLastText(void);
void pdStubResetDraw(void);

#ifdef __cplusplus
}
#endif
#endif
