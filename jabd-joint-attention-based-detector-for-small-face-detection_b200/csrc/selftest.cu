// selftest.cu -- libjabd_b200_selftest.so: test and bench hooks that are NOT part of the product ABI
// (include/jabd_b200_selftest.h).  Built from the same device helpers (common.cuh) as libjabd_b200.so, so the
// division self-test exercises exactly the code the kernels inline.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "jabd_b200_selftest.h"

namespace jabd {

// this library carries its own copy of the error plumbing (hidden visibility: no clash with libjabd_b200.so)
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return JABD_ECUDA;
}

// ---- device self-test of fdiv_shared() against the compiler's IEEE division ----------------------------------
__device__ __forceinline__ uint32_t mix32(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

__global__ void __launch_bounds__(256) selftest_div_kernel(unsigned long long n, unsigned long long seed,
                                                           unsigned long long *mismatches, float *first_bad)
{
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t r0 = mix32(seed + 3 * i), r1 = mix32(seed + 3 * i + 1), r2 = mix32(seed + 3 * i + 2);
        // divisor: any sign, exponent within the documented safe range 2^-60 .. 2^60, random mantissa
        const uint32_t de = 127u - 60u + (r2 % 120u);
        const float d = __uint_as_float((r0 & 0x807fffffu) | (de << 23));
        // numerator: arbitrary bit pattern (all exponents, subnormals, inf, NaN); every 16th is an exact zero
        float a = __uint_as_float(r1);
        if ((r2 >> 28) == 0u) a = 0.0f;
        if ((r2 >> 28) == 1u) a = __uint_as_float((r1 & 0x807fffffu) | (((r2 >> 8) % 8u) << 23)); // tiny / subnormal
        if (!divisor_safe(d)) continue;
        const float got = fdiv_shared(a, d, rcp_refined(d));
        const float ref = __fdiv_rn(a, d);
        const bool same = (__float_as_uint(got) == __float_as_uint(ref)) || (got != got && ref != ref) ||
                          (got == 0.0f && ref == 0.0f);
        if (!same) {
            if (bad == 0 && atomicAdd(mismatches + 1, 1ull) == 0ull) { first_bad[0] = a; first_bad[1] = d; first_bad[2] = got; first_bad[3] = ref; }
            ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---- FP32-pipe throughput probe: the roofline denominator of the matching kernel ----------------------------
// 16 independent chains per thread, alternating FMUL and FADD (never fused: --fmad=false and _rn intrinsics), i.e. the
// instruction mix of the IoU arithmetic without any memory traffic.  2 * 16 * iters operations per thread.
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float seed, float *sink)
{
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = seed + (float)(threadIdx.x + k);
    const float m = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = __fadd_rn(__fmul_rn(x[k], m), c);
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += x[k];
    if (acc == 123.456f) sink[0] = acc; // never true; keeps the chains alive
}

} // namespace jabd

extern "C" {

/* Launches the FP32 probe on `ctas` CTAs of 256 threads; the caller times it.  Operations executed:
 * ctas * 256 * 32 * iters (half FMUL, half FADD). */
int jabd_selftest_fp32_probe(int ctas, int iters, float *sink_dev, jabd_stream_t stream)
{
    JABD_REQUIRE(ctas > 0 && iters > 0 && sink_dev, JABD_EINVAL, "fp32_probe: bad argument");
    jabd::fp32_probe_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, 1.0f, sink_dev);
    JABD_LAUNCH_CHECK("fp32_probe_kernel");
    return JABD_OK;
}

/* Test hook: compares fdiv_shared() with __fdiv_rn() on n pseudo-random operand pairs.  out_dev[0] receives the
 * number of mismatching results (out_dev[1] is scratch), first_bad_dev[4] = (a, d, got, expected) of one of them. */
int jabd_selftest_div(uint64_t n, uint64_t seed, unsigned long long *out_dev, float *first_bad_dev, jabd_stream_t stream)
{
    JABD_REQUIRE(out_dev && first_bad_dev, JABD_EINVAL, "selftest_div: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    JABD_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(unsigned long long), st));
    jabd::selftest_div_kernel<<<148 * 8, 256, 0, st>>>(n, seed, out_dev, first_bad_dev);
    JABD_LAUNCH_CHECK("selftest_div_kernel");
    return JABD_OK;
}

const char *jabd_selftest_last_error(void) { return jabd::g_err; }

} // extern "C"
