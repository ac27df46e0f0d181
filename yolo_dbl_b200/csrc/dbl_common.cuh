// Shared device/host helpers for libdbl_sm100 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dbl_sm100.h"

namespace dbl {

// ---- error plumbing (no exceptions across the C ABI) -------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> DBL_E_CUDA

#define DBL_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            ::dbl::set_error(__VA_ARGS__); \
            return (code);                \
        }                                 \
    } while (0)

// ---- per-device state ---------------------------------------------------------------------------------------------
// SM count / shared-memory limit and every cudaFuncSetAttribute opt-in are properties of ONE device: a process that drives
// several GPUs (DBLBackend.to('cuda:1')) needs them per ordinal, never in a process-wide static.
struct DevInfo {
    int ordinal, sms, max_smem;
};
const DevInfo& dev_info();  // of the CURRENT device, cached per ordinal (api.cu)
// "has this kernel been configured on the current device": one bit per device ordinal.  Usage:
//   static DevOnce once;  if (once.need()) { cudaFuncSetAttribute(...); once.mark(); }
struct DevOnce {
    std::atomic<unsigned long long> done{0};
    bool need() const { return !((done.load(std::memory_order_acquire) >> (dev_info().ordinal & 63)) & 1ull); }
    void mark() { done.fetch_or(1ull << (dev_info().ordinal & 63), std::memory_order_release); }
};

// Table keyed by (kernel function, current device) for launchers that configure many instantiations through one code path:
// func_config() returns -1 until func_set_config(fn, value >= 0) has been called on the current device (value: e.g. occupancy).
int func_config(const void* fn);
void func_set_config(const void* fn, int value);

// ---- launches: programmatic dependent launch (PDL) ----------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute: the next kernel of the
// stream (or the next node of the captured graph) may START -- block scheduling, shared-memory carve-up, barrier / TMEM
// set-up, weight staging -- while its predecessor is still draining, and blocks in pdl_wait() until the predecessor has
// completed and its writes are visible.  Rules every kernel follows: pdl_launch_dependents() first; nothing written by
// an earlier kernel is read, and nothing is written, before pdl_wait(); every kernel executes pdl_wait() (completion is
// then transitive along the stream).  dbl_set_pdl(0) turns the attribute off (the device-side calls become no-ops).
extern int g_pdl_enabled;
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = g_pdl_enabled ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through check_launch()
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

static inline cudaStream_t as_stream(dbl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline bool valid_tensor(const dbl_tensor* t) {
    return t && t->data && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->pitch >= t->c &&
           (t->dtype == DBL_BF16 || t->dtype == DBL_F32);
}
static inline bool same_shape(const dbl_tensor* a, const dbl_tensor* b) {
    return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

// ---- division of a tile index by a run-time constant without the ~25-instruction integer division ---------------------
// q = umulhi(x, m) with m = floor(2^32 / d) + 1 is exact while x * d < 2^32; the host leaves m = 0 (plain division) otherwise.
struct fastdiv {
    uint32_t d, m;
};
static inline fastdiv make_fastdiv(uint32_t d, uint64_t max_x) {
    fastdiv f{d, 0};
    if (d > 1 && max_x * d < (1ull << 32)) f.m = (uint32_t)((1ull << 32) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, fastdiv f) { return f.m ? __umulhi(x, f.m) : x / f.d; }

// ---- device math ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float sigmoid_f(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float apply_act(float v, int act) {
    return act == DBL_ACT_SILU ? silu_f(v) : (act == DBL_ACT_GELU ? gelu_f(v) : v);
}

// 8 bf16 <-> 8 floats through ONE 16-byte access.  The transport type is a plain uint4: a struct of __nv_bfloat162
// members is copied member-wise by nvcc (four 32-bit LDG/STG), which quarters the achieved bandwidth.
struct bf16x8 {
    uint4 u;
};
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
    unpack2(p.u.x, f[0], f[1]);
    unpack2(p.u.y, f[2], f[3]);
    unpack2(p.u.z, f[4], f[5]);
    unpack2(p.u.w, f[6], f[7]);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
    bf16x8 p;
    p.u = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    return p;
}
// packed fp32 FMA over an aligned register pair (SASS FFMA2): half the issue slots of two FFMAs, IEEE fma per element
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long bf16x2_to_f32x2(uint32_t w) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xffff0000u));
    return r;
}
__device__ __forceinline__ void f32x2_split(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// read-only path (ld.global.nc): only for tensors no thread of the running kernel writes
__device__ __forceinline__ bf16x8 ldg8(const __nv_bfloat16* p) {
    bf16x8 v;
    v.u = __ldg(reinterpret_cast<const uint4*>(p));
    return v;
}
// coherent load: for tensors that may alias the kernel's output (in-place residuals)
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) {
    bf16x8 v;
    v.u = *reinterpret_cast<const uint4*>(p);
    return v;
}
__device__ __forceinline__ void stg8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<uint4*>(p) = v.u; }

// generic element access for views whose channel count / pitch / base is not 16-byte friendly
__device__ __forceinline__ float ld_elem(const void* base, int dtype, size_t idx) {
    return dtype == DBL_F32 ? reinterpret_cast<const float*>(base)[idx]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}
__device__ __forceinline__ void st_elem(void* base, int dtype, size_t idx, float v) {
    if (dtype == DBL_F32)
        reinterpret_cast<float*>(base)[idx] = v;
    else
        reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

static inline bool vec8_ok(const dbl_tensor* t) {  // 16-byte vector access along channels is legal
    return t->dtype == DBL_BF16 && (t->c % 8 == 0) && (t->pitch % 8 == 0) && ((reinterpret_cast<uintptr_t>(t->data) & 15) == 0);
}

}  // namespace dbl
