/*
 * pd_api.h -- PlaydateAPI HOST STUB (not the Playdate SDK).
 *
 * The reference (ctotheameron/pdphyzx) includes "pd_api.h" unconditionally
 * (src/includes/px_platform.h:4, src/includes/px_api.h:4) and touches exactly this
 * much of the SDK:
 *   pd->system->realloc                      src/includes/px_platform.h:9-21
 *   pd->system->logToConsole                 src/includes/px_platform.h:27 (macro, unused)
 *   pd->system->getCurrentTimeMilliseconds   src/px_clock.c:26
 *   pd->system->formatString                 src/px_world.c:495
 *   pd->graphics->drawRect/drawLine/drawText src/px_world.c:459-498
 *   kColorXOR, kColorBlack, kASCIIEncoding   src/px_world.c:460,489,498
 *
 * This file declares only that surface so the unmodified reference sources compile
 * and run on the host cores.  The concrete PlaydateAPI instance (with a settable
 * millisecond clock) lives in pd_stub.h (this directory) / pdphyzx_b200/csrc/pd_stub.c.
 */
#ifndef PD_API_HOST_STUB_H
#define PD_API_HOST_STUB_H

#include <stddef.h>
#include <stdint.h>

typedef enum { kColorBlack = 0, kColorWhite = 1, kColorClear = 2, kColorXOR = 3 } LCDSolidColor;
typedef uintptr_t LCDColor;
typedef enum { kASCIIEncoding = 0, kUTF8Encoding = 1, k16BitLEEncoding = 2 } PDStringEncoding;

struct playdate_sys {
  void *(*realloc)(void *ptr, size_t size); /* ptr==NULL: malloc; size==0: free */
  int (*formatString)(char **ret, const char *fmt, ...);
  void (*logToConsole)(const char *fmt, ...);
  unsigned int (*getCurrentTimeMilliseconds)(void);
};

struct playdate_graphics {
  void (*drawRect)(int x, int y, int width, int height, LCDColor color);
  void (*drawLine)(int x1, int y1, int x2, int y2, int width, LCDColor color);
  int (*drawText)(const void *text, size_t len, PDStringEncoding encoding, int x, int y);
};

typedef struct PlaydateAPI {
  const struct playdate_sys *system;
  const struct playdate_graphics *graphics;
} PlaydateAPI;

#endif /* PD_API_HOST_STUB_H */
