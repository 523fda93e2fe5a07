/* px_error.c -- last-error text of the batch layer (thread-local). */
#include <stdarg.h>
#include <stdio.h>

#include "px_batch.h"
#include "px_internal.h"

static __thread char g_error[512] = "";

int pxSetError(int rc, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return rc;
}

const char *pxBatchLastError(void) { return g_error; }
const char *pxBatchVersion(void) { return "pdphyzx-b200 0.1 (sm_100a)"; }
