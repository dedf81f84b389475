/*
 * jabd_b200_selftest.h -- C-ABI of libjabd_b200_selftest.so: TEST AND BENCH HOOKS, not product.
 *
 * Kept out of libjabd_b200.so / jabd_b200.h on purpose: a maintainer of the reference binds jabd_b200.h only.
 * tests/ use the division self-test, bench.py uses the FP32 probe as the measured roofline denominator of the
 * matching kernel.  Same conventions as jabd_b200.h (device pointers, caller's stream, 0 / negative code).
 */
#ifndef JABD_B200_SELFTEST_H
#define JABD_B200_SELFTEST_H

#include "jabd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Compares the library's shared-reciprocal IEEE division (common.cuh: rcp_refined + fdiv_shared, what the matching,
 * encode and NMS kernels inline) with the compiler's div.rn on n pseudo-random operand pairs.
 * out[0] = number of mismatches (out[1] scratch), first_bad[4] = (a, d, got, expected). */
JABD_API int jabd_selftest_div(uint64_t n, uint64_t seed, unsigned long long *out, float *first_bad, jabd_stream_t stream);

/* Dependency-free FMUL/FADD chains (no FMA, no memory traffic) on `ctas` CTAs of 256 threads;
 * ctas * 256 * 32 * iters fp32 operations.  The caller times it: the measured FP32-pipe peak. */
JABD_API int jabd_selftest_fp32_probe(int ctas, int iters, float *sink_dev, jabd_stream_t stream);

JABD_API const char *jabd_selftest_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* JABD_B200_SELFTEST_H */
