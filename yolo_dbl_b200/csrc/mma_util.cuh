// Warp-level MMA building blocks shared by the thin-layer kernels (conv_mma.cu) and the fused Detect head (head_fused.cu):
// ldmatrix / mma.sync wrappers, cp.async groups, the planar shared-memory tile helpers and the packed-weight geometry.
#pragma once
#include "tc_ptx.cuh"

namespace dbl {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// float2 from a shared-memory float array through an explicit ld.shared (a `float*` carved out of the 128-byte-aligned dynamic
// shared buffer has lost its address space: ptxas emits generic LD.E there)
__device__ __forceinline__ float2 lds_f2(const float* p) {
    const uint2 v = lds64(smem_u32(p));
    return make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- packed-weight geometry (shared by the pack kernel and the consumers) -------------------------------------------
// k-steps: cin >= 16: tap-major, 16 channels per step;  cin == 8: two taps per step (an odd tap count pads with zeros)
__host__ __device__ constexpr int mma_ksteps(int cin, int taps) { return cin == 8 ? (taps + 1) / 2 : taps * (cin / 16); }
// GEMM column n (0..7) of n-block j  ->  output channel: thread t = n/2 owns channels [t*2NB, (t+1)*2NB)
__host__ __device__ constexpr int mma_chan(int nb_total, int j, int n) { return (n >> 1) * 2 * nb_total + 2 * j + (n & 1); }

// ---- tap offsets (bytes) inside a planar tile ---------------------------------------------------------------------
template <int PITCH>
struct OffS1 {  // stride 1: tile pixels linearised with row pitch PITCH
    static __device__ __forceinline__ constexpr uint32_t at(int tap) { return (uint32_t)(((tap / 3) * PITCH + tap % 3) * 16); }
};
template <int PITCH, int PARB>
struct OffS2 {  // stride 2: four parity sub-tiles of PARB bytes each, sub-tile pixels linearised with row pitch PITCH
    static __device__ __forceinline__ constexpr uint32_t at(int tap) {
        const int ky = tap / 3, kx = tap % 3;
        return (uint32_t)(((ky & 1) * 2 + (kx & 1)) * PARB + ((ky >> 1) * PITCH + (kx >> 1)) * 16);
    }
};

// One conv "phase": acc[mb][j] += A(tile rows m0*16.. , taps) x B for MB consecutive 16-row blocks.
//   a_addr: shared byte address of GEMM row (lane&7)+8*((lane>>3)&1) of the first block, tap 0, plane (CIN>8 ? lane>>4 : 0)
//   b_addr: shared byte address of this lane's slot in the packed weights
template <int CIN, int NB, int MB, int PS, class Off>
__device__ __forceinline__ void conv_phase(float (&acc)[MB][NB][4], uint32_t a_addr, uint32_t b_addr, bool hi) {
    constexpr int KS = mma_ksteps(CIN, 9);
    // software pipeline (everything is unrolled): A fragments one k-step ahead, B fragment one n-block ahead
    auto a_off = [&](int ks) -> uint32_t {
        if (CIN == 8) return hi ? Off::at(2 * ks + 1 < 9 ? 2 * ks + 1 : 8) : Off::at(2 * ks);
        return Off::at(ks / (CIN >= 16 ? CIN / 16 : 1)) + (uint32_t)(2 * (ks % (CIN >= 16 ? CIN / 16 : 1)) * PS);
    };
    uint32_t a[2][MB][4];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) ldsm_x4(a_addr + a_off(0) + mb * 256, a[0][mb]);
    uint2 b_next = lds64(b_addr);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        if (ks + 1 < KS) {
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) ldsm_x4(a_addr + a_off(ks + 1) + mb * 256, a[(ks + 1) & 1][mb]);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const uint2 b = b_next;
            if (ks * NB + j + 1 < KS * NB) b_next = lds64(b_addr + (uint32_t)((ks * NB + j + 1) * 256));
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) mma16816(acc[mb][j], a[ks & 1][mb], b.x, b.y);
        }
    }
}

// plane stride (in pixels) >= need, chosen so the 16-byte cp.async stores of one quarter-warp (G consecutive planes of
// the same pixel, then the next pixel) spread over all banks
__host__ __device__ constexpr int plane_px(int need, int g) {
    const int gg = g > 8 ? 8 : g;
    const int want = gg <= 1 ? 0 : (8 / gg) % 8;
    int v = need;
    if (gg > 1)
        while (v % 8 != want) ++v;
    return v;
}
__host__ __device__ constexpr int round16(int v) { return (v + 15) / 16 * 16; }

// store 2*NB contiguous bf16 channels (thread t's share of one pixel) into a planar tile / read them back
template <int NB>
__device__ __forceinline__ void planar_store(uint32_t tile, int ps_bytes, int pix, int t, const uint32_t* v) {
    if (NB == 1) sts32(tile + pix * 16 + t * 4, v[0]);
    else if (NB == 2) sts64(tile + (t >> 1) * ps_bytes + pix * 16 + (t & 1) * 8, make_uint2(v[0], v[1]));
    else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q)
            sts128(tile + (t * (NB / 4) + q) * ps_bytes + pix * 16, make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    }
}
template <int NB>
__device__ __forceinline__ void planar_load(uint32_t tile, int ps_bytes, int pix, int t, uint32_t* v) {
    if (NB == 1) v[0] = lds32(tile + pix * 16 + t * 4);
    else if (NB == 2) {
        const uint2 u = lds64(tile + (t >> 1) * ps_bytes + pix * 16 + (t & 1) * 8);
        v[0] = u.x, v[1] = u.y;
    } else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) {
            const uint4 u = lds128(tile + (t * (NB / 4) + q) * ps_bytes + pix * 16);
            v[4 * q] = u.x, v[4 * q + 1] = u.y, v[4 * q + 2] = u.z, v[4 * q + 3] = u.w;
        }
    }
}
// 2*NB contiguous bf16 channels to / from global memory (p is 4*NB-byte aligned)
template <int NB>
__device__ __forceinline__ void global_store(__nv_bfloat16* p, const uint32_t* v) {
    if (NB == 1) *reinterpret_cast<uint32_t*>(p) = v[0];
    else if (NB == 2) *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]);
    else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) reinterpret_cast<uint4*>(p)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}
template <int NB>
__device__ __forceinline__ void global_load(const __nv_bfloat16* p, uint32_t* v) {
    if (NB == 1) v[0] = *reinterpret_cast<const uint32_t*>(p);
    else if (NB == 2) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        v[0] = u.x, v[1] = u.y;
    } else {
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) {
            const uint4 u = reinterpret_cast<const uint4*>(p)[q];
            v[4 * q] = u.x, v[4 * q + 1] = u.y, v[4 * q + 2] = u.z, v[4 * q + 3] = u.w;
        }
    }
}


__device__ __forceinline__ float4 lds_f4(const float* p) {
    const uint4 v = lds128(smem_u32(p));
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

}  // namespace dbl
