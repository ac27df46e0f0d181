// Library-level entry points: error reporting, device probe, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>

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

const DevInfo& dev_info() {
    static DevInfo table[64];
    static std::atomic<unsigned long long> known{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError(), dev = 0;
    DevInfo& d = table[dev & 63];
    if (!((known.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        DevInfo t{dev, 0, 0};
        cudaDeviceGetAttribute(&t.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&t.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        d = t;  // racing writers store identical values
        known.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    return d;
}

static std::mutex g_cfg_mu;
static std::map<std::pair<const void*, int>, int> g_cfg;

int func_config(const void* fn) {
    const int dev = dev_info().ordinal;
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    auto it = g_cfg.find({fn, dev});
    return it == g_cfg.end() ? -1 : it->second;
}

void func_set_config(const void* fn, int value) {
    const int dev = dev_info().ordinal;
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    g_cfg[{fn, dev}] = value;
}

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
