// Inline-PTX wrappers shared by the tcgen05 kernels: mbarrier, TMA, tcgen05.{mma,commit,ld,fence}, cp.async.
#pragma once
#include <cuda.h>

#include "dbl_common.cuh"

namespace dbl {

constexpr uint32_t SPIN_LIMIT = 1u << 22;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > SPIN_LIMIT) __trap();  // a protocol bug must fail loudly, never hang the GPU
    }
}
// Same wait for roles that expect to sleep (a producer waiting for a free stage, an epilogue waiting for the next accumulator): the
// suspend-time hint lets the hardware park the warp until the phase completes or the hint expires, instead of re-issuing
// try_wait / compare / branch every ~80 cycles.  ncu on dsconv_tc_kernel: ten such warps per CTA spent 23 % of all issued
// instructions in this loop, slots the depthwise warps needed.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 4000) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(hint_ns)
            : "memory");
        if (done) break;
        if (++spins > (1u << 20)) __trap();  // a protocol bug must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// shared -> global tile store (bulk async group; rows / columns outside the tensor are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// one lane of a fully converged warp (elect.sync): the form ptxas recognises for single-thread tcgen05 / TMA issue; a plain
// `lane == 0` test makes it wrap every UTCHMMA in an ELECT / BRA.U.ANY loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cp.async (LDGSTS) with zero fill + mbarrier completion ---------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;  // src-size 0: nothing is read, the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
// the barrier receives one arrival once all cp.async issued so far by this thread have landed (count pre-armed: .noinc)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Stage a [rows][cols][G8] tile of 16-byte chunks (NHWC bf16, 8 channels per chunk) with cp.async, zero-filling everything
// outside the image (= the convolution's padding).  A thread keeps its (column, channel group) and walks down the rows: per
// chunk one row bounds test, one pointer add and the cp.async.  (ncu source view of the flat `for (i = tid; i < chunks; i += nt)`
// form: ~68 instructions per chunk on index decomposition and 64-bit address arithmetic -- 58 % of the 3x3 depthwise kernel.)
//   src_img: channel 0 of this CTA's channel block at pixel (0,0) of the image;  dummy: any valid global address (never read)
__device__ __forceinline__ void stage_tile16(uint32_t dst_smem, const __nv_bfloat16* src_img, const void* dummy, int rows, int cols, int G8,
                                             int g8_valid, int iy0, int ix0, int h, int w, int xpitch, int tid, int nt) {
    const int row_items = cols * G8;
    const int per_pass = row_items < nt ? row_items : nt;  // columns x groups handled per pass
    const int rpp = nt / per_pass;                          // rows per pass (>= 1)
    const int r0 = __float2int_rd(((float)tid + 0.5f) * (1.0f / (float)per_pass));  // exact: (q + 0.5) / d is never an integer
    const int e_in = tid - r0 * per_pass;
    if (r0 >= rpp) return;
    for (int e = e_in; e < row_items; e += per_pass) {  // one pass unless a tile row is longer than the CTA
        const int lx = G8 == 1 ? e : (G8 == 2 ? e >> 1 : e / G8), g = e - lx * G8;
        const int gx = ix0 + lx;
        const bool colok = g < g8_valid && gx >= 0 && gx < w;
        const __nv_bfloat16* src = src_img + ((long long)(iy0 + r0) * w + gx) * xpitch + g * 8;
        const long long rstep = (long long)rpp * w * xpitch;
        uint32_t dst = dst_smem + (uint32_t)(r0 * row_items + e) * 16;
        for (int ly = r0; ly < rows; ly += rpp) {
            const bool ok = colok && (unsigned)(iy0 + ly) < (unsigned)h;
            cp_async16(dst, ok ? static_cast<const void*>(src) : dummy, ok);
            src += rstep, dst += (uint32_t)(rpp * row_items) * 16;
        }
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// SiLU through one MUFU: x*sigmoid(x) = 0.5x(1 + tanh(0.5x)); tanh.approx is ~2^-11 accurate, far below bf16 rounding
__device__ __forceinline__ float silu_fast(float v) {
    float t;
    const float h = 0.5f * v;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
// SiLU(z + b) with the bias pre-scaled by 0.5: h = 0.5 z + hb (one FMA), result h + h * tanh(h)
__device__ __forceinline__ float silu_half_bias(float z, float hb) {
    float t;
    const float h = fmaf(z, 0.5f, hb);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    if (ACT == DBL_ACT_SILU) return silu_fast(v);
    if (ACT == DBL_ACT_GELU) return gelu_f(v);
    return v;
}

}  // namespace dbl
