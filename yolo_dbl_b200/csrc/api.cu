// Library-level entry points: error reporting, device probe, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "dbl_common.cuh"

namespace dbl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
int g_pdl_enabled = 1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DBL_E_CUDA;
    }
    return DBL_OK;
}

}  // namespace dbl

extern "C" {

const char* dbl_last_error(void) { return dbl::g_err; }

int dbl_version(void) { return 100; }

int dbl_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dbl::set_error("no CUDA device");
        return 0;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (major != 10) dbl::set_error("libdbl_sm100 needs compute capability 10.x, found %d.x", major);
    return major == 10;
}

int dbl_set_pdl(int enabled) {
    const int old = dbl::g_pdl_enabled;
    dbl::g_pdl_enabled = enabled ? 1 : 0;
    return old;
}

int64_t dbl_launch_count(void) { return dbl::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
