// C-ABI plumbing: error string, version, device properties.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "idee_b200.h"

static thread_local char g_err[512] = "";

void idee_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int idee_num_sms() {
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached_sms;
    if (dev != cached_dev) {
        int n = 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
        cached_dev = dev;
    }
    return cached_sms;
}

extern "C" const char* idee_last_error(void) { return g_err; }
extern "C" int idee_version(void) { return IDEE_B200_VERSION; }

// 0 if the current device can run the library (compute capability 10.x), else 1 with a message
extern "C" int idee_check_device(void) {
    int dev = 0, major = 0, minor = 0;
    IDEE_CUDA(cudaGetDevice(&dev), "check_device");
    IDEE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "check_device");
    IDEE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev), "check_device");
    IDEE_REQUIRE(major == 10, "idee_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
}
