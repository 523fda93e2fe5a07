/* pd_stub.c -- PlaydateAPI host stub instance; see include/pd_host/pd_stub.h. */
#include "pd_stub.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

static uint32_t g_time_ms = 0;
static unsigned g_draw_calls = 0;
static char *g_last_text = NULL;

static void *stub_realloc(void *ptr, size_t size) {
  if (size == 0) {
    free(ptr);
    return NULL;
  }
  return realloc(ptr, size);
}

static int stub_format_string(char **ret, const char *fmt, ...) {
  va_list ap, ap2;
  va_start(ap, fmt);
  va_copy(ap2, ap);
  int n = vsnprintf(NULL, 0, fmt, ap);
  va_end(ap);
  char *buf = (char *)malloc((size_t)n + 1);
  vsnprintf(buf, (size_t)n + 1, fmt, ap2);
  va_end(ap2);
  *ret = buf;
  return n;
}

static void stub_log(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}

static unsigned int stub_time_ms(void) { return g_time_ms; }

static void stub_draw_rect(int x, int y, int w, int h, LCDColor c) {
  (void)x; (void)y; (void)w; (void)h; (void)c;
  g_draw_calls++;
}

static void stub_draw_line(int x1, int y1, int x2, int y2, int w, LCDColor c) {
  (void)x1; (void)y1; (void)x2; (void)y2; (void)w; (void)c;
  g_draw_calls++;
}

static int stub_draw_text(const void *text, size_t len, PDStringEncoding enc, int x, int y) {
  (void)len; (void)enc; (void)x; (void)y;
  g_draw_calls++;
  free(g_last_text); /* the reference hands over a formatString buffer it never frees */
  g_last_text = (char *)text;
  return 0;
}

static const struct playdate_sys g_sys = {
    .realloc = stub_realloc,
    .formatString = stub_format_string,
    .logToConsole = stub_log,
    .getCurrentTimeMilliseconds = stub_time_ms,
};

static const struct playdate_graphics g_gfx = {
    .drawRect = stub_draw_rect,
    .drawLine = stub_draw_line,
    .drawText = stub_draw_text,
};

static PlaydateAPI g_api = {.system = &g_sys, .graphics = &g_gfx};

PlaydateAPI *pdStubAPI(void) { return &g_api; }
void pdStubSetTimeMs(uint32_t ms) { g_time_ms = ms; }
void pdStubAdvanceTimeMs(uint32_t ms) { g_time_ms += ms; }
uint32_t pdStubGetTimeMs(void) { return g_time_ms; }
unsigned pdStubDrawCalls(void) { return g_draw_calls; }
const char *pdStubLastText(void) { return g_last_text; }
void pdStubResetDraw(void) { g_draw_calls = 0; }
