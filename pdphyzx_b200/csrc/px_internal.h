/* px_internal.h -- shared by the host-side C and CUDA translation units of the batch layer. */
#ifndef PX_INTERNAL_H
#define PX_INTERNAL_H

#ifdef __cplusplus
extern "C" {
#endif

/* records a printf-style message for pxBatchLastError() and returns `rc` */
int pxSetError(int rc, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

#ifdef __cplusplus
}
#endif
#endif
