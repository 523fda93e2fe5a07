/* px_internal.h -- shared by the host-side C and CUDA translation units of the batch layer. */
#ifndef PX_INTERNAL_H
#define PX_INTERNAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* records a printf-style message for pxBatchLastError() and returns `rc` */
int pxSetError(int rc, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

/* body lifecycle on a resident world, host half (px_pack.c over include/px_pack_inl.h): the CUDA
 * translation unit cannot read pdphyzx.h (`new` is a member name there) */
struct PxBatchLayout;
struct PxWorld;
int pxPackAddDynamic(const struct PxBatchLayout *L, const struct PxWorld *world, const void *body /* const PxBody * */, void *immWorld,
                     float *dynRows, uint32_t *flags, uint32_t *slot);
int pxPackStatics(const struct PxBatchLayout *L, const struct PxWorld *world, void *mutWorld, void *immWorld);
int pxPackBodyIsDynamic(const void *body /* const PxBody * */);

#ifdef __cplusplus
}
#endif
#endif
