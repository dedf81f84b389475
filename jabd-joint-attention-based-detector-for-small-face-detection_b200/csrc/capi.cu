// capi.cu -- error reporting, version and device query of the C-ABI (include/jabd_b200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace jabd {

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

} // namespace jabd

extern "C" {

int jabd_version(void) { return 100; /* 0.1.0 */ }

const char *jabd_last_error(void) { return jabd::g_err; }

int jabd_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    JABD_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    JABD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    JABD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    JABD_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    JABD_REQUIRE(major == 10, JABD_ENODEVICE, "device is sm_%d%d; this library is built for sm_100a only", major, minor);
    return JABD_OK;
}

} // extern "C"
