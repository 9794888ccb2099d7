// Error reporting and device queries of libtasu_bridge.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>

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

// run-time options (tasu_set_option); the initial value of option X comes from the environment variable named below
static const char* const kOptionEnv[TASU_OPT_COUNT] = {"TASU_GEMM_PAIR"};
static const int kOptionDefault[TASU_OPT_COUNT] = {1};
static std::atomic<int> g_options[TASU_OPT_COUNT];
static std::once_flag g_options_once;

static void init_options() {
    std::call_once(g_options_once, [] {
        for (int i = 0; i < TASU_OPT_COUNT; ++i) {
            const char* e = getenv(kOptionEnv[i]);
            g_options[i].store(e != nullptr ? atoi(e) : kOptionDefault[i]);
        }
    });
}

int option(int id) {
    if (id < 0 || id >= TASU_OPT_COUNT) return 0;
    init_options();
    return g_options[id].load(std::memory_order_relaxed);
}

}  // namespace tasu

extern "C" int tasu_set_option(int option, int value) {
    TASU_CHECK_ARG(option >= 0 && option < TASU_OPT_COUNT, "unknown option");
    tasu::init_options();
    tasu::g_options[option].store(value);
    return TASU_OK;
}

extern "C" int tasu_get_option(int option) {
    TASU_CHECK_ARG(option >= 0 && option < TASU_OPT_COUNT, "unknown option");
    return tasu::option(option);
}

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
