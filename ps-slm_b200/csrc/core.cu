// Error reporting and device queries of libtasu_bridge.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace tasu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace tasu

extern "C" int tasu_abi_version(void) { return TASU_ABI_VERSION; }

extern "C" const char* tasu_last_error(void) { return tasu::g_err; }

extern "C" int tasu_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
    int dev = 0;
    TASU_CHECK_CUDA(cudaGetDevice(&dev));
    int sm = 0, maj = 0, min = 0;
    TASU_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    TASU_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    TASU_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count_host) *sm_count_host = sm;
    if (cc_major_host) *cc_major_host = maj;
    if (cc_minor_host) *cc_minor_host = min;
    return TASU_OK;
}
