// 3x3 stride-1 implicit-GEMM convolution on tcgen05 for LARGE reduction depth (Cin >= 128: the neck / head convolutions
// that hold most of the network's FLOPs).  Third generation: every activation byte is fetched from L2 once (halo reuse, as
// conv_tc3.cu) AND the weights are streamed once per SUPER-TILE of up to 16 output tiles whose fp32 accumulators all live
// in TMEM at the same time.
//
// Why: with K = 9*Cin in the thousands the weights (147..442 KB) cannot stay in shared memory.  The first-generation
// kernel (conv_tc.cu) re-reads both operands for each of the nine taps of each 128-pixel tile -> ~300 B/clk/SM of L2
// traffic against a measured cap of ~42 B/clk/SM: L2-bound at 160..270 TF/s.  Here the loop nest is
//
//     for super-tile (T tiles = T TMEM accumulators of N columns, T*N <= 512):
//         for channel chunk cc (64 channels):             <- weights of the chunk: 9 TMA boxes, once per super-tile
//             for tile t in super-tile:                    <- halo of the chunk: ONE TMA box (zero-filled padding)
//                 9 taps x 4 k-steps of tcgen05.mma into accumulator t
//
// which needs ~20 KB (halo) + 74 KB / T (weights) per 36 MMAs: ~25 B/clk/SM.  The epilogue of tile t overlaps the MMAs of
// tiles t+1.. (accumulators are released one by one), so no TMEM double buffering is needed.
//
// Operand layout: the halo tile lands pixel-major ([halo row][halo column] x 64 channels = one 128-byte swizzled row per
// pixel, exactly what one 4-D TMA box writes); GEMM row m of filter tap (dy,dx) is halo pixel m + dy*WL + dx, so a tap is
// the same buffer with the descriptor start advanced by whole 128-byte rows.  (Measured on B200: the swizzle is a function
// of the absolute shared-memory address bits, so an arbitrary row offset needs NO matrix-base-offset field; setting it to
// (start >> 7) & 7 gives wrong results.  Stage buffers are 1024-byte aligned so TMA and UMMA agree on the pattern.)
// GEMM rows that fall on the two halo columns of each tile row are junk and never stored (conv_tc3.cu's trick).
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int T4_EPI_WARPS = 16;
constexpr int T4_THREADS = 64 + 32 * T4_EPI_WARPS;  // warp 0 TMA producer, warp 1 MMA issuer, the rest epilogue
constexpr int T4_MAX_STAGES = 6;
constexpr int T4_MAX_ACC = 16;

struct T4Params {
    int n, h, w, cin, cout;
    int cchunk, nchunks;  // channels per K chunk (64 / 32 / 16), cin / cchunk
    int tw, th, wl, rows_h;
    int tiles_x, tiles_y, m_tiles;
    int npad, T, tmem_cols, n_super;
    int stages;
    uint32_t a_bytes, a_tx, b_bytes, b_tx;
    int base_off_mode;  // experiment switch: 1 sets the descriptor base-offset field = (start >> 7) & 7 (wrong on B200), 0 leaves it zero
    const float* bias;
    const __nv_bfloat16* r;
    void* y;
    int rpitch, ypitch, ydtype, epi, y_vec, r_vec;
    int b_single;          // 1: the whole filter is ONE chunk (cin == cchunk): staged once per CTA, stays resident
    int tma_out;           // 1: bf16 output tiles staged in shared memory (128B-swizzled rows) and written by TMA bulk stores
    int wl8;               // staging row pitch in pixels (wl rounded up to 8: every tile row starts a 1024-byte swizzle atom)
    uint32_t o_bytes;      // staging bytes per epilogue quad
};

struct T4Maps {
    CUtensorMap a, b;
    CUtensorMap y;  // output [N][H][W][cout], box = 64 channels x tw pixels of one row (TMA-store epilogue)
};

__device__ __forceinline__ uint64_t make_sw_desc(uint32_t saddr, uint32_t row_bytes, int base_off_mode) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    if (base_off_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
    d |= layout << 61;
    return d;
}

// tcgen05.mma with the descriptors given as (lo, hi) words: only the low word changes from MMA to MMA (start address)
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}

#ifdef DBL_TRACE  // developer aid (never in the shipped build): per-CTA time stamps, read back with dbl_debug_trace4_read
__device__ unsigned long long g_t4_trace[160 * 64];
__device__ __forceinline__ void t4_trace(int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x < 160 && slot < 64) g_t4_trace[blockIdx.x * 64 + slot] = t;
}
#define T4_TRACE(slot, cond) do { if (cond) t4_trace(slot); } while (0)
#else
#define T4_TRACE(slot, cond) do { } while (0)
#endif

template <int ACT, int KSTEPS>
__global__ void __launch_bounds__(T4_THREADS, 1) conv3x3_tc4_kernel(const __grid_constant__ T4Maps maps, const T4Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                  // [stages][a_bytes]
    uint8_t* sB = sA + (size_t)p.stages * p.a_bytes;    // [2][b_bytes] ([1] when the filter is resident)
    float* sBias = reinterpret_cast<float*>(sB + (p.b_single ? 1 : 2) * (size_t)p.b_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + ((p.npad + 3) & ~3));
    uint64_t* afull = bars;
    uint64_t* aempty = afull + T4_MAX_STAGES;
    uint64_t* bfull = aempty + T4_MAX_STAGES;  // [2]
    uint64_t* bempty = bfull + 2;              // [2]
    uint64_t* tfull = bempty + 2;              // [T4_MAX_ACC]
    uint64_t* tempty = tfull + T4_MAX_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + T4_MAX_ACC);
    uint8_t* sOut = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 1023) & ~(uintptr_t)1023);  // [quads][o_bytes] when p.tma_out

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T4_TRACE(0, threadIdx.x == 0);
    for (int i = threadIdx.x; i < p.npad; i += T4_THREADS) sBias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a);
        prefetch_tmap(&maps.b);
        if (p.tma_out) prefetch_tmap(&maps.y);
        for (int s = 0; s < p.stages; ++s) mbar_init(&afull[s], 1), mbar_init(&aempty[s], 1);
        for (int s = 0; s < 2; ++s) mbar_init(&bfull[s], 1), mbar_init(&bempty[s], 1);
        for (int a = 0; a < p.T; ++a) mbar_init(&tfull[a], 1), mbar_init(&tempty[a], 4);  // one quad of epilogue warps drains a tile
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();  // after the TMEM allocation (see conv_tc.cu)
    pdl_wait();
    T4_TRACE(1, threadIdx.x == 0);
    int tr_n = 0;  // trace counter (dead code unless DBL_TRACE)
    (void)tr_n;

    const int per_img = p.tiles_x * p.tiles_y;
    const uint32_t row_bytes = (uint32_t)p.cchunk * 2;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
        uint32_t stage = 0, aphase = 0, bbuf = 0, bphase = 0;
        for (int s = blockIdx.x; s < p.n_super; s += gridDim.x) {
            const int t0 = s * p.T, t1 = min(t0 + p.T, p.m_tiles);
            for (int cc = 0; cc < p.nchunks; ++cc) {
                if (!p.b_single || s == (int)blockIdx.x) {  // single-chunk filter: staged once, resident for the life of the CTA
                    mbar_wait(&bempty[bbuf], bphase ^ 1);
                    uint8_t* b_dst = sB + (size_t)bbuf * p.b_bytes;
                    if (elect_one()) {
                        mbar_expect_tx(&bfull[bbuf], p.b_tx);
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_2d(&maps.b, &bfull[bbuf], b_dst + (size_t)tap * p.npad * row_bytes, tap * p.cin + cc * p.cchunk, 0);
                    }
                    __syncwarp();
                    if (!p.b_single && ++bbuf == 2) bbuf = 0, bphase ^= 1;
                }
                for (int mt = t0; mt < t1; ++mt) {
                    const int img = mt / per_img, t = mt - img * per_img;
                    const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
                    mbar_wait(&aempty[stage], aphase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&afull[stage], p.a_tx);
                        tma_load_4d(&maps.a, &afull[stage], sA + (size_t)stage * p.a_bytes, cc * p.cchunk, x0 - 1, y0 - 1, img);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)p.stages) stage = 0, aphase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // One thread issues every MMA, so its instruction count per MMA is the kernel's pace when N is small (an N=32 MMA
        // occupies the tensor pipe for only 16 cycles).  The descriptors of a (tile, chunk) unit differ only in the start
        // address field of their low word: nine tap offsets live in registers, the k-step offset is an immediate, the
        // whole unit is unrolled -> two integer adds per MMA.
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dtempl = make_sw_desc(0, row_bytes, 0);
        const uint32_t d_lo = (uint32_t)dtempl, d_hi = (uint32_t)(dtempl >> 32);
        const uint32_t px16 = row_bytes >> 4, row16 = ((uint32_t)p.wl * row_bytes) >> 4;  // descriptor steps: one halo pixel / one halo row
        const uint32_t tap_b = ((uint32_t)p.npad * row_bytes) >> 4;
        // Control flow and operands stay warp-uniform (all 32 lanes walk the loops and poll the barriers); only the issue
        // itself is predicated on lane 0.  With the loops inside `if (lane == 0)` ptxas cannot prove the descriptors
        // uniform and wraps every UTCHMMA in an R2UR waterfall loop (~15 instructions per MMA).
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);  // read from shared memory: tell ptxas it is uniform
        uint32_t stage = 0, aphase = 0, bbuf = 0, bphase = 0, it = 0;
        for (int s = blockIdx.x; s < p.n_super; s += gridDim.x, ++it) {
            const int ntile = min(p.T, p.m_tiles - s * p.T);
            for (int cc = 0; cc < p.nchunks; ++cc) {
                if (!p.b_single || s == (int)blockIdx.x) mbar_wait(&bfull[bbuf], bphase);
                const uint32_t b_lo = d_lo + (smem_u32(sB + (size_t)bbuf * p.b_bytes) >> 4);
                for (int t = 0; t < ntile; ++t) {
                    if (cc == 0) mbar_wait(&tempty[t], (it & 1) ^ 1);  // the epilogue has drained this accumulator
                    mbar_wait(&afull[stage], aphase);
                    tc_fence_after();
                    T4_TRACE(2 + tr_n, lane == 0 && tr_n < 30);
                    ++tr_n;
                    const uint32_t a_lo = d_lo + (smem_u32(sA + (size_t)stage * p.a_bytes) >> 4);
                    const uint32_t d_tmem = tmem_u + (uint32_t)(t * p.npad);
                    const uint32_t first = cc ? 1u : 0u;
                    if (elect_one()) {
                        // rolled tap loops, k-steps unrolled: two descriptor low words stepped by adds, nothing else live.
                        // (Fully unrolled, the 72 descriptor pairs of a unit exceeded the uniform register file: ptxas spilled
                        // them through MOV.SPILL / R2UR.FILL / LDL and the single issuing thread needed ~70-90 cycles per MMA,
                        // 2.5x the 32 cycles an N <= 64 MMA occupies the tensor pipe -- measured with tools/tc4_trace.py.)
                        uint32_t acc = first, a_row = a_lo, b_t = b_lo;
#pragma unroll 1
                        for (int dy = 0; dy < 3; ++dy) {
#ifdef DBL_T4_ROLL_DX
#pragma unroll 1
#else
#pragma unroll
#endif
                            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                                for (int ks = 0; ks < KSTEPS; ++ks) {
                                    umma_bf16_lohi(d_tmem, a_row + dx * px16 + 2 * ks, d_hi, b_t + 2 * ks, d_hi, idesc, acc);
                                    acc = 1u;
                                }
                                b_t += tap_b;
                            }
                            a_row += row16;
                        }
                        umma_commit(&aempty[stage]);
                        if (cc == p.nchunks - 1) umma_commit(&tfull[t]);
                        if (t == ntile - 1 && !p.b_single) umma_commit(&bempty[bbuf]);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)p.stages) stage = 0, aphase ^= 1;
                }
                if (!p.b_single && ++bbuf == 2) bbuf = 0, bphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue =====================
        // The 16 warps form FOUR quads (one warp per TMEM lane group); quad q drains tiles q, q+4, ... of the super-tile,
        // all of a tile's columns.  Four tiles are therefore in their epilogue at the same time: the per-tile latency chain
        // (barrier -> tcgen05.ld -> activation -> residual load -> stores, ~1.5 us) overlaps instead of adding up, which is
        // what paced this kernel when all 16 warps walked the tiles in lock step.  The residual of a chunk is requested
        // before the accumulator is waited for.
        const int ew = warp - 2;
        const int lg = warp & 3;   // TMEM lane group this warp may access
        const int quad = ew >> 2;  // 0..3
        constexpr int NQUAD = T4_EPI_WARPS / 4;
        const int row = lg * 32 + lane;
        const int ty = row / p.wl, tx = row - ty * p.wl;
        const int ncols = min(p.npad, (p.cout + 15) & ~15);
        const bool r_fast = p.epi != DBL_EPI_NONE && p.r_vec;
        uint32_t it = 0;
        for (int s = blockIdx.x; s < p.n_super; s += gridDim.x, ++it) {
            const int t0 = s * p.T, ntile = min(p.T, p.m_tiles - t0);
            for (int t = quad; t < ntile; t += NQUAD) {
                const int mt = t0 + t;
                const int img = mt / per_img, tt = mt - img * per_img;
                const int oy = (tt / p.tiles_x) * p.th + ty, ox = (tt % p.tiles_x) * p.tw + tx;
                long long opix = -1;
                if (ty < p.th && tx < p.tw && oy < p.h && ox < p.w) opix = ((long long)img * p.h + oy) * p.w + ox;
                const __nv_bfloat16* rrow = p.r + (size_t)(opix < 0 ? 0 : opix) * p.rpitch;
                bf16x8 rn0, rn1;  // residual of the NEXT chunk (coherent loads: r may be the tensor this kernel overwrites in place)
                rn0.u = rn1.u = make_uint4(0u, 0u, 0u, 0u);
                if (r_fast && opix >= 0 && 16 <= p.cout) rn0 = ld8(rrow), rn1 = ld8(rrow + 8);
                mbar_wait(&tfull[t], it & 1);
                tc_fence_after();
                T4_TRACE(32 + 2 * tr_n, warp == 2 && lane == 0 && tr_n < 15);
                if (p.tma_out) {
                    // ---- staged epilogue (see conv_tc.cu): the quad's 128 rows go to a 128B-swizzled staging tile whose tile
                    // rows start on 1024-byte atoms; the quad's first lane writes every valid tile row with one TMA bulk store
                    // (full lines, columns / rows past the map clipped) instead of 32-line-scattered 16-byte warp stores.
                    uint8_t* obuf = sOut + (size_t)quad * p.o_bytes;
                    const bool issuer = (lg == 0 && lane == 0);  // exactly one thread per quad (the quad's warps have distinct lane groups)
                    const int sr = ty * p.wl8 + tx;              // staging row of this thread's pixel
                    const uint32_t srow = smem_u32(obuf) + (uint32_t)sr * 128u;
                    const int y0 = (tt / p.tiles_x) * p.th, x0 = (tt % p.tiles_x) * p.tw;
                    for (int h0 = 0; h0 < ncols; h0 += 64) {
                        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stores that read this buffer are done
                        asm volatile("bar.sync %0, 128;" ::"r"(quad + 1) : "memory");
                        const int hend = min(h0 + 64, ncols);
                        for (int c0 = h0; c0 < hend; c0 += 16) {
                            uint32_t v[16];
                            tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * p.npad + c0), v);
                            const bf16x8 rc0 = rn0, rc1 = rn1;
                            if (r_fast && opix >= 0 && c0 + 32 <= p.cout) rn0 = ld8(rrow + c0 + 16), rn1 = ld8(rrow + c0 + 24);
                            tmem_ld_wait();
                            if (ty >= p.th) continue;  // junk GEMM rows past the tile have no staging row
                            float f[16];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 bq = lds_f4(sBias + c0 + 4 * q);
                                f[4 * q + 0] = act_t<ACT>(__uint_as_float(v[4 * q + 0]) + bq.x);
                                f[4 * q + 1] = act_t<ACT>(__uint_as_float(v[4 * q + 1]) + bq.y);
                                f[4 * q + 2] = act_t<ACT>(__uint_as_float(v[4 * q + 2]) + bq.z);
                                f[4 * q + 3] = act_t<ACT>(__uint_as_float(v[4 * q + 3]) + bq.w);
                            }
                            if (p.epi != DBL_EPI_NONE && opix >= 0) {
                                float rr[16];
                                if (c0 + 16 <= p.cout && p.r_vec) {
                                    unpack8(rc0, rr);
                                    unpack8(rc1, rr + 8);
                                } else {
                                    const __nv_bfloat16* rp = rrow + c0;
#pragma unroll
                                    for (int i = 0; i < 16; ++i) rr[i] = (c0 + i < p.cout) ? __bfloat162float(rp[i]) : 0.f;
                                }
#pragma unroll
                                for (int i = 0; i < 16; ++i) f[i] = p.epi == DBL_EPI_ADD ? f[i] + rr[i] : f[i] * rr[i];
                            }
                            const uint32_t j = (uint32_t)((c0 - h0) >> 3);
                            sts128(srow + ((j ^ ((uint32_t)sr & 7u)) << 4), pack8(f).u);
                            sts128(srow + (((j + 1) ^ ((uint32_t)sr & 7u)) << 4), pack8(f + 8).u);
                        }
                        fence_proxy_async();
                        asm volatile("bar.sync %0, 128;" ::"r"(quad + 1) : "memory");
                        if (issuer) {
                            for (int r = 0; r < p.th; ++r)
                                if (y0 + r < p.h) tma_store_4d(&maps.y, obuf + (size_t)r * p.wl8 * 128, h0, x0, y0 + r, img);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[t]);
                    T4_TRACE(33 + 2 * tr_n, warp == 2 && lane == 0 && tr_n < 15);
                    ++tr_n;
                    continue;
                }
                for (int c0 = 0; c0 < ncols; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * p.npad + c0), v);
                    const bf16x8 rc0 = rn0, rc1 = rn1;
                    if (r_fast && opix >= 0 && c0 + 32 <= p.cout) rn0 = ld8(rrow + c0 + 16), rn1 = ld8(rrow + c0 + 24);
                    tmem_ld_wait();
                    if (opix < 0) continue;
                    float f[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bq = lds_f4(sBias + c0 + 4 * q);
                        f[4 * q + 0] = act_t<ACT>(__uint_as_float(v[4 * q + 0]) + bq.x);
                        f[4 * q + 1] = act_t<ACT>(__uint_as_float(v[4 * q + 1]) + bq.y);
                        f[4 * q + 2] = act_t<ACT>(__uint_as_float(v[4 * q + 2]) + bq.z);
                        f[4 * q + 3] = act_t<ACT>(__uint_as_float(v[4 * q + 3]) + bq.w);
                    }
                    const bool full16 = c0 + 16 <= p.cout;
                    if (p.epi != DBL_EPI_NONE) {
                        float rr[16];
                        if (full16 && p.r_vec) {
                            unpack8(rc0, rr);
                            unpack8(rc1, rr + 8);
                        } else {
                            const __nv_bfloat16* rp = rrow + c0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) rr[i] = (c0 + i < p.cout) ? __bfloat162float(rp[i]) : 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = p.epi == DBL_EPI_ADD ? f[i] + rr[i] : f[i] * rr[i];
                    }
                    if (p.ydtype == DBL_BF16) {
                        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)opix * p.ypitch + c0;
                        if (full16 && p.y_vec) {
                            stg8(yp, pack8(f));
                            stg8(yp + 8, pack8(f + 8));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.cout) yp[i] = __float2bfloat16_rn(f[i]);
                        }
                    } else {
                        float* yp = reinterpret_cast<float*>(p.y) + (size_t)opix * p.ypitch + c0;
                        if (full16 && p.y_vec) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                reinterpret_cast<float4*>(yp)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.cout) yp[i] = f[i];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[t]);
                T4_TRACE(33 + 2 * tr_n, warp == 2 && lane == 0 && tr_n < 15);
                ++tr_n;
            }
        }
    }
    if (p.tma_out && warp >= 2 && (warp & 3) == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the bulk stores read shared memory
    tc_fence_before();
    __syncthreads();
    T4_TRACE(63, threadIdx.x == 0);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static EncodeTiledFn4 get_encode4() {
    static EncodeTiledFn4 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn4>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool encode4(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes) {
    EncodeTiledFn4 enc = get_encode4();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is unavailable");
        return false;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides[i];
    const CUtensorMapSwizzle sw = sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv3x3_tc4: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return false;
    }
    return true;
}

bool conv3x3_tc4_eligible(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil) {
    if (k != 3 || stride != 1 || pad != 1 || dil != 1) return false;
    if (x->dtype != DBL_BF16 || x->c % 16 != 0 || x->pitch % 8 != 0 || (reinterpret_cast<uintptr_t>(x->data) & 15)) return false;
    if (y->c > 256 || x->w < 2) return false;
    return true;
}

// the multi-accumulator kernel pays off when the filter does not fit shared memory next to the activation ring
bool conv3x3_tc4_preferred(const dbl_tensor* x, const dbl_tensor* y) {
    static const int min_cin = getenv("DBL_TC4_MIN_CIN") ? atoi(getenv("DBL_TC4_MIN_CIN")) : 32;  // tuning aid
    return x->c >= min_cin && y->c >= 16;
}

int conv3x3_tc4(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int act, int epi, cudaStream_t st) {
    const int num_sms = dev_info().sms, max_smem = dev_info().max_smem;
    static DevOnce once;  // shared-memory opt-in is per device
    if (once.need()) {
        cudaError_t e = cudaSuccess;
#define DBL_T4_ATTR(A, K) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc4_kernel<A, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)
        DBL_T4_ATTR(0, 1); DBL_T4_ATTR(0, 2); DBL_T4_ATTR(0, 4);
        DBL_T4_ATTR(1, 1); DBL_T4_ATTR(1, 2); DBL_T4_ATTR(1, 4);
        DBL_T4_ATTR(2, 1); DBL_T4_ATTR(2, 2); DBL_T4_ATTR(2, 4);
#undef DBL_T4_ATTR
        if (e != cudaSuccess) {
            set_error("conv3x3_tc4: cannot opt in to %d bytes of shared memory: %s", max_smem, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, DBL_E_ARG, "conv3x3_tc4: weights must be 16-byte aligned");
    T4Params p;
    memset(&p, 0, sizeof(p));
    p.n = x->n, p.h = x->h, p.w = x->w, p.cin = x->c, p.cout = y->c;
    p.npad = ((y->c + 15) / 16) * 16;
    // ---- tile: th rows x tw columns, GEMM rows = th*(tw+2) <= 128 (fewest tiles, then least halo)
    long long best = -1;
    for (int tw = 2; tw <= (p.w < 126 ? p.w : 126); ++tw) {
        const int wl = tw + 2;
        int th = 128 / wl;
        if (th > p.h) th = p.h;
        if (th < 1) continue;
        const long long tiles = (long long)cdiv(p.w, tw) * cdiv(p.h, th);
        const long long cost = tiles * ((long long)(th + 2) * wl + 96);
        if (best < 0 || cost < best) best = cost, p.tw = tw, p.th = th, p.wl = wl;
    }
    p.rows_h = p.th + 2;
    p.tiles_x = cdiv(p.w, p.tw), p.tiles_y = cdiv(p.h, p.th);
    p.m_tiles = p.n * p.tiles_x * p.tiles_y;
    // ---- K chunking + shared-memory budget: weight buffers (one resident buffer when the filter is a single chunk, else
    // two) + >= 3 activation stages (+ the staging tiles of the TMA-store epilogue when they fit without shrinking the chunk)
    const size_t fixed = 1024 + (((size_t)p.npad + 3) & ~(size_t)3) * 4 + (2 * T4_MAX_STAGES + 4 + 2 * T4_MAX_ACC) * 8 + 16;
    const int halo_px = p.rows_h * p.wl;
    const int need_px = 128 + 2 * p.wl + 2 > halo_px ? 128 + 2 * p.wl + 2 : halo_px;  // junk rows read a little past the box
    static const int tma_out_env = getenv("DBL_TC4_TMA_OUT") ? atoi(getenv("DBL_TC4_TMA_OUT")) : 0;  // opt-in: measured SLOWER here (39 vs 37 us on 64->32 @80^2): this kernel is paced by the MMAs' A-operand reads, and the staging tiles cost activation stages
    const bool out_ok = tma_out_env && y->dtype == DBL_BF16 && ((size_t)y->pitch * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(y->data) & 15) == 0;
    p.wl8 = (p.wl + 7) & ~7;
    p.o_bytes = (uint32_t)(((size_t)p.th * p.wl8 * 128 + 1023) & ~(size_t)1023);
    const size_t out_bytes = 1024 + (size_t)(T4_EPI_WARPS / 4) * p.o_bytes;
    int cchunk = 0;
    for (int pass = out_ok ? 0 : 1; pass < 2 && !cchunk; ++pass) {  // pass 0: with output staging (only at the largest chunk), pass 1: without
        const size_t extra = pass == 0 ? out_bytes : 0;
        for (int c = 64; c >= 16; c >>= 1) {
            if (x->c % c) continue;
            static const int bsingle_env = getenv("DBL_TC4_BSINGLE") ? atoi(getenv("DBL_TC4_BSINGLE")) : 1;  // tuning aid
            const int nbuf = (x->c == c && c == 64 && bsingle_env) ? 1 : 2;  // measured: resident filter helps 64->128 @40^2 (31 vs 35.5 us), hurts 32->64 @80^2 (58.6 vs 51.4 us)
            const size_t a_b = (((size_t)need_px * c * 2) + 1023) & ~(size_t)1023;
            const size_t b_b = (((size_t)9 * p.npad * c * 2) + 1023) & ~(size_t)1023;
            if (fixed + extra + nbuf * b_b + 3 * a_b <= (size_t)max_smem) {
                cchunk = c;
                p.a_bytes = (uint32_t)a_b, p.b_bytes = (uint32_t)b_b;
                p.b_single = nbuf == 1, p.tma_out = pass == 0;
                int stages = (int)(((size_t)max_smem - fixed - extra - nbuf * b_b) / a_b);
                p.stages = stages > T4_MAX_STAGES ? T4_MAX_STAGES : stages;
                break;
            }
            if (pass == 0) break;  // staging must not cost chunk width
        }
    }
    DBL_REQUIRE(cchunk > 0, DBL_E_UNSUPPORTED, "conv3x3_tc4: cin %d / cout %d does not fit shared memory", x->c, y->c);
    p.cchunk = cchunk, p.nchunks = x->c / cchunk;
    const uint32_t row_bytes = (uint32_t)cchunk * 2;
    p.a_tx = (uint32_t)halo_px * row_bytes;
    p.b_tx = (uint32_t)9 * p.npad * row_bytes;
    // ---- accumulators per super-tile: as many as TMEM holds, then evened out over the waves of the grid
    int tmax = 512 / p.npad;
    if (tmax > T4_MAX_ACC) tmax = T4_MAX_ACC;
    const int waves = cdiv(p.m_tiles, num_sms * tmax);
    p.T = cdiv(p.m_tiles, num_sms * waves);
    if (p.T > tmax) p.T = tmax;
    if (p.T < 1) p.T = 1;
    p.n_super = cdiv(p.m_tiles, p.T);
    int cols = 32;
    while (cols < p.T * p.npad) cols <<= 1;
    p.tmem_cols = cols;
    static const int bo_mode = getenv("DBL_TC4_BASEOFF") ? atoi(getenv("DBL_TC4_BASEOFF")) : 0;  // measured: 0 is right
    p.base_off_mode = bo_mode;
    p.bias = bias;
    p.r = r ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr;
    p.rpitch = r ? r->pitch : 0;
    p.r_vec = r ? (int)(r->pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(r->data) & 15) == 0) : 0;
    p.y = y->data, p.ypitch = y->pitch, p.ydtype = y->dtype, p.epi = r ? epi : DBL_EPI_NONE;
    const int yelem = y->dtype == DBL_BF16 ? 2 : 4;
    p.y_vec = (int)(((size_t)y->pitch * yelem) % 16 == 0 && (reinterpret_cast<uintptr_t>(y->data) & 15) == 0);
    T4Maps maps;
    memset(&maps, 0, sizeof(maps));
    {
        const uint64_t pitchB = (uint64_t)x->pitch * 2;
        const uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        const uint64_t strides[3] = {pitchB, pitchB * x->w, pitchB * x->w * x->h};
        const uint32_t box[4] = {(uint32_t)cchunk, (uint32_t)p.wl, (uint32_t)p.rows_h, 1};
        if (!encode4(&maps.a, x->data, 4, dims, strides, box, row_bytes)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)9 * x->c, (uint64_t)y->c};
        const uint64_t strides[1] = {(uint64_t)9 * x->c * 2};
        const uint32_t box[2] = {(uint32_t)cchunk, (uint32_t)p.npad};
        if (!encode4(&maps.b, w, 2, dims, strides, box, row_bytes)) return DBL_E_CUDA;
    }
    if (p.tma_out) {
        const uint64_t pitchB = (uint64_t)y->pitch * 2;
        const uint64_t dims[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
        const uint64_t strides[3] = {pitchB, pitchB * y->w, pitchB * y->w * y->h};
        const uint32_t box[4] = {64, (uint32_t)p.tw, 1, 1};
        if (!encode4(&maps.y, y->data, 4, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    const size_t smem = fixed + (size_t)p.stages * p.a_bytes + (p.b_single ? 1 : 2) * (size_t)p.b_bytes + (p.tma_out ? out_bytes : 0);
    const int grid = p.n_super < num_sms ? p.n_super : num_sms;
    const int ks = cchunk / 16;
#define DBL_T4_GO(A, K) launch_k(conv3x3_tc4_kernel<A, K>, grid, T4_THREADS, smem, st, maps, p)
#define DBL_T4_ACT(A)            \
    do {                         \
        if (ks == 4) DBL_T4_GO(A, 4); \
        else if (ks == 2) DBL_T4_GO(A, 2); \
        else DBL_T4_GO(A, 1);    \
    } while (0)
    if (act == DBL_ACT_SILU) DBL_T4_ACT(DBL_ACT_SILU);
    else if (act == DBL_ACT_GELU) DBL_T4_ACT(DBL_ACT_GELU);
    else DBL_T4_ACT(DBL_ACT_NONE);
#undef DBL_T4_ACT
#undef DBL_T4_GO
    count_launch();
    return check_launch("dbl_conv2d(tcgen05 3x3 multi-accumulator)");
}

}  // namespace dbl

#ifdef DBL_TRACE
extern "C" int dbl_debug_trace4_read(void* dst, size_t bytes) {
    return cudaMemcpyFromSymbol(dst, dbl::g_t4_trace, bytes < sizeof(dbl::g_t4_trace) ? bytes : sizeof(dbl::g_t4_trace)) == cudaSuccess ? 0 : 1;
}
#endif
