// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) + tap, c] * W[n, tap, c]        m: 128 output pixels, n: <= 256 output channels
//
// * A tiles: NHWC bf16 activations fetched straight from the feature map by TMA, one box per (tap, 64-channel chunk).
//   3x3 taps are boxes shifted by (dy, dx); rows/columns that fall outside the image are zero-filled by the TMA unit,
//   which IS the convolution's zero padding.  Stride-2 convs use four "parity" tensor maps (even/odd rows x columns)
//   so each tap is still a dense box.  1x1 convs view the map as a flat [pixels][C] matrix.
// * B tiles: weights bf16 [Cout][taps*Cin] (K-major), boxes of [<=256 rows][64] per K block; rows past Cout zero-fill.
// * Both land in shared memory in the canonical K-major swizzled layout (128/64/32-byte swizzle chosen by the channel
//   chunk) that tcgen05.mma consumes through shared-memory descriptors; fp32 accumulators live in TMEM, double buffered.
// * Warp-specialised persistent CTAs: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 =
//   epilogue (tcgen05.ld -> +bias -> SiLU/GELU -> (+|*) residual -> bf16/f32 -> global), mbarrier pipelines in between.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int TC_EPI_WARPS = 8;                      // 2 per TMEM lane group: one warp per scheduler cannot hide its own latencies
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;  // warp 0 TMA, warp 1 MMA, the rest epilogue
constexpr int TC_MAX_STAGES = 8;
constexpr int MAX_ACC = 8;
constexpr int TILE_M = 128;

struct TcParams {
    // problem
    int n, h, w, cin;          // input
    int ho, wo, cout;          // output
    int k, stride, pad;        // 1 or 3; 1 or 2
    int cchunk;                // channels per K block: 64 / 32 / 16
    int nchunks;               // cin / cchunk
    int taps;                  // k*k
    // tiling
    int flat;                  // 1: 1x1 conv on the flat pixel matrix
    int tw, th;                // spatial tile (tw*th <= 128) when !flat
    int tiles_x, tiles_y;      // tiles per image
    int m_tiles, n_tiles;      // totals
    int npad;                  // N of the MMA (multiple of 16, <= 256)
    int tmem_cols;             // power of two >= 2*npad, >= 32
    int stages, nacc;
    uint32_t a_bytes, b_bytes; // per stage (a_bytes: 128 rows reserved)
    uint32_t a_tx, b_tx;       // bytes one A / B box actually delivers
    // epilogue
    const float* bias;
    const __nv_bfloat16* r;
    void* y;
    int rpitch, ypitch, ydtype, act, epi;
    int y_vec;                 // 16-byte stores legal
    int r_vec;
    int tma_out;               // 1: bf16 output tile staged in shared memory (128B-swizzled) and written by TMA bulk stores
};

struct TcMaps {
    CUtensorMap a[4];  // [0] stride-1 / flat; [ph*2+pw] parity maps for stride 2
    CUtensorMap b;
    CUtensorMap y;     // output as [pixels][cout] (flat) for the TMA-store epilogue
};

constexpr uint32_t TC_OUT_BYTES = 128 * 128;  // one staged output sub-tile: 128 rows x 64 bf16 channels (128-byte swizzled rows)

// K-major shared-memory matrix descriptor (tcgen05 "version 1"): rows of `row_bytes` (= swizzle span), 8-row groups
// `8*row_bytes` apart.  layout_type: 2 = 128B swizzle, 4 = 64B, 6 = 32B.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);             // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset  bits [16,30)  (unused with swizzle)
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;          // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version   bits [46,48)
    d |= layout << 61;                                    // swizzle mode         bits [61,64)
    return d;
}

#ifdef DBL_TRACE  // developer aid (never in the shipped build): per-CTA phase time stamps, read back with dbl_debug_trace_read
__device__ unsigned long long g_tc_trace[512 * 16];
__device__ __forceinline__ void tc_trace(int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x < 512) g_tc_trace[blockIdx.x * 16 + slot] = t;
}
#define TC_TRACE(slot, cond) do { if (cond) tc_trace(slot); } while (0)
#else
#define TC_TRACE(slot, cond) do { } while (0)
#endif

template <int ACT>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x A][stages x B][bias f32 cout_pad][barriers][tmem ptr]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)p.stages * p.a_bytes;
    float* sBias = reinterpret_cast<float*>(sB + (size_t)p.stages * p.b_bytes);
    const int bias_n = p.n_tiles * p.npad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + ((bias_n + 3) & ~3));
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_MAX_STAGES;
    uint64_t* tfull = bars + 2 * TC_MAX_STAGES;
    uint64_t* tempty = tfull + MAX_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + MAX_ACC);
    uint8_t* sOut = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 1023) & ~(uintptr_t)1023);  // [2][TC_OUT_BYTES] when p.tma_out

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TC_TRACE(0, threadIdx.x == 0);
    for (int i = threadIdx.x; i < bias_n; i += TC_THREADS) sBias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a[0]);
        prefetch_tmap(&maps.b);
        if (p.tma_out) prefetch_tmap(&maps.y);
        if (p.stride == 2) {
            prefetch_tmap(&maps.a[1]);
            prefetch_tmap(&maps.a[2]);
            prefetch_tmap(&maps.a[3]);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < p.nacc; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TC_TRACE(1, threadIdx.x == 0);
    // PDL: dependents may start only once every CTA of this grid HOLDS its TMEM columns (a dependent CTA that allocated
    // first and then blocked in its own pdl_wait() would starve us); then wait for our own predecessor
    pdl_launch_dependents();
    pdl_wait();
    TC_TRACE(2, threadIdx.x == 0);

    const int total_tiles = p.m_tiles * p.n_tiles;
    const int kblocks = p.taps * p.nchunks;
    const uint32_t row_bytes = (uint32_t)p.cchunk * 2;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            int img = 0, y0 = 0, x0 = 0;
            if (!p.flat) {
                const int per_img = p.tiles_x * p.tiles_y;
                img = mt / per_img;
                const int t = mt % per_img;
                y0 = (t / p.tiles_x) * p.th;
                x0 = (t % p.tiles_x) * p.tw;
            }
            for (int kb = 0; kb < kblocks; ++kb) {
                const int tap = kb / p.nchunks, ch = kb % p.nchunks;
                mbar_wait(&empty[stage], phase ^ 1);
                void* a_dst = sA + (size_t)stage * p.a_bytes;
                void* b_dst = sB + (size_t)stage * p.b_bytes;
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], p.a_tx + p.b_tx);
                    if (p.flat) {
                        tma_load_2d(&maps.a[0], &full[stage], a_dst, ch * p.cchunk, mt * TILE_M);
                    } else if (p.stride == 1) {
                        const int dy = tap / p.k - p.pad, dx = tap % p.k - p.pad;
                        tma_load_4d(&maps.a[0], &full[stage], a_dst, ch * p.cchunk, x0 + dx, y0 + dy, img);
                    } else {
                        const int ty = tap / p.k - p.pad, tx = tap % p.k - p.pad;  // input row = 2*oy + ty
                        const int ph = ty & 1, pw = tx & 1;
                        const int oy = (ty - ph) / 2, ox = (tx - pw) / 2;          // exact: ty-ph is even
                        tma_load_4d(&maps.a[ph * 2 + pw], &full[stage], a_dst, ch * p.cchunk, x0 + ox, y0 + oy, img);
                    }
                    tma_load_2d(&maps.b, &full[stage], b_dst, kb * p.cchunk, nt * p.npad);
                    TC_TRACE(3, tile == (int)blockIdx.x && kb == 0);
                }
                __syncwarp();
                if (++stage == (uint32_t)p.stages) stage = 0, phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        // instruction descriptor: D=f32, A=B=bf16, both K-major, N = npad, M = 128.  Everything the MMA consumes is kept
        // provably warp-uniform (loops outside the election, TMEM base re-broadcast): otherwise ptxas wraps each UTCHMMA
        // in an R2UR / ELECT waterfall loop of ~15 instructions, which paces the whole kernel when K is small.
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dtempl = make_kmajor_desc(0, row_bytes);
        const int ksteps = p.cchunk / 16;
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1;
            mbar_wait(&tempty[acc], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + acc * (uint32_t)p.npad;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                TC_TRACE(4, it == 0 && kb == 0 && lane == 0);
                TC_TRACE(5, it == 0 && kb == kblocks - 1 && lane == 0);
                const uint64_t ad0 = dtempl + (uint64_t)(smem_u32(sA + (size_t)stage * p.a_bytes) >> 4);
                const uint64_t bd0 = dtempl + (uint64_t)(smem_u32(sB + (size_t)stage * p.b_bytes) >> 4);
                if (elect_one()) {
                    for (int ks = 0; ks < ksteps; ++ks) umma_bf16(d_tmem, ad0 + 2 * ks, bd0 + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
                    umma_commit(&empty[stage]);                       // frees the smem slot when these MMAs retire
                    if (kb == kblocks - 1) umma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == (uint32_t)p.stages) stage = 0, phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int lg = warp & 3;          // TMEM lane group this warp may access
        const int chalf = (warp - 2) >> 2;  // which interleaved share of the column chunks this warp converts
        const int row = lg * 32 + lane;
        uint32_t it = 0, sub = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1;
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            // which output pixel does this thread own?
            long long opix = -1;
            if (p.flat) {
                const long long m = (long long)mt * TILE_M + row;
                if (m < (long long)p.n * p.ho * p.wo) opix = m;
            } else {
                const int per_img = p.tiles_x * p.tiles_y;
                const int img = mt / per_img, t = mt % per_img;
                const int ty = row / p.tw, tx = row % p.tw;
                const int oy = (t / p.tiles_x) * p.th + ty, ox = (t % p.tiles_x) * p.tw + tx;
                if (ty < p.th && oy < p.ho && ox < p.wo) opix = ((long long)img * p.ho + oy) * p.wo + ox;
            }
            mbar_wait(&tfull[acc], aphase);
            tc_fence_after();
            TC_TRACE(6, it == 0 && warp == 2 && lane == 0);
            const int n0 = nt * p.npad;
            const int ncols = min(p.npad, ((p.cout - n0) + 15) & ~15);
            if (p.tma_out) {
                // ---- staged epilogue: a thread's row-wise 16-byte pieces go to a 128B-swizzled shared-memory sub-tile
                // (128 rows x 64 channels), one elected thread writes it with ONE TMA bulk store (full 128-byte lines, rows
                // past the tensor clipped by the unit).  The direct form below issues, per 16 channels, a warp-wide store
                // that touches 32 different lines with 16 bytes each: ~2000 LSU line-cycles per 128x128 tile, which paced
                // every multi-tile 1x1 GEMM of the network.
                const bool issuer = (warp == 2 && lane == 0);
                for (int h0 = 0; h0 < ncols; h0 += 64, ++sub) {
                    uint8_t* obuf = sOut + (sub & 1) * TC_OUT_BYTES;
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store that last read this buffer is done
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                    const int hcols = min(64, ncols - h0);
                    for (int c0 = h0 + chalf * 16; c0 < h0 + hcols; c0 += 16 * (TC_EPI_WARPS / 4)) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.npad + c0, v);
                        tmem_ld_wait();
                        const int cg = n0 + c0;
                        float f[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 bq = lds_f4(sBias + cg + 4 * q);
                            f[4 * q + 0] = act_t<ACT>(__uint_as_float(v[4 * q + 0]) + bq.x);
                            f[4 * q + 1] = act_t<ACT>(__uint_as_float(v[4 * q + 1]) + bq.y);
                            f[4 * q + 2] = act_t<ACT>(__uint_as_float(v[4 * q + 2]) + bq.z);
                            f[4 * q + 3] = act_t<ACT>(__uint_as_float(v[4 * q + 3]) + bq.w);
                        }
                        if (p.epi != DBL_EPI_NONE && opix >= 0) {
                            const __nv_bfloat16* rp = p.r + (size_t)opix * p.rpitch + cg;
                            float rr[16];
                            if (cg + 16 <= p.cout && p.r_vec) {
                                unpack8(ld8(rp), rr);
                                unpack8(ld8(rp + 8), rr + 8);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) rr[i] = (cg + i < p.cout) ? __bfloat162float(rp[i]) : 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) f[i] = p.epi == DBL_EPI_ADD ? f[i] + rr[i] : f[i] * rr[i];
                        }
                        const uint32_t j = (uint32_t)((c0 - h0) >> 3);  // 16-byte chunk of the 128-byte row
                        const uint32_t base = smem_u32(obuf) + (uint32_t)row * 128u;
                        sts128(base + (((j) ^ ((uint32_t)row & 7u)) << 4), pack8(f).u);
                        sts128(base + (((j + 1) ^ ((uint32_t)row & 7u)) << 4), pack8(f + 8).u);
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                    if (issuer) {
                        tma_store_2d(&maps.y, obuf, n0 + h0, mt * TILE_M);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                TC_TRACE(7, it == 0 && warp == 2 && lane == 0);
                TC_TRACE(8, warp == 2 && lane == 0);
                continue;
            }
            for (int c0 = chalf * 16; c0 < ncols; c0 += 16 * (TC_EPI_WARPS / 4)) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.npad + c0, v);
                tmem_ld_wait();
                if (opix < 0) continue;
                const int cg = n0 + c0;  // first global output channel of this chunk
                float f[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 bq = lds_f4(sBias + cg + 4 * q);
                    f[4 * q + 0] = act_t<ACT>(__uint_as_float(v[4 * q + 0]) + bq.x);
                    f[4 * q + 1] = act_t<ACT>(__uint_as_float(v[4 * q + 1]) + bq.y);
                    f[4 * q + 2] = act_t<ACT>(__uint_as_float(v[4 * q + 2]) + bq.z);
                    f[4 * q + 3] = act_t<ACT>(__uint_as_float(v[4 * q + 3]) + bq.w);
                }
                const bool full16 = cg + 16 <= p.cout;
                if (p.epi != DBL_EPI_NONE) {
                    const __nv_bfloat16* rp = p.r + (size_t)opix * p.rpitch + cg;
                    float rr[16];
                    if (full16 && p.r_vec) {  // coherent loads: r may be the very tensor this kernel overwrites (in-place residual)
                        unpack8(ld8(rp), rr);
                        unpack8(ld8(rp + 8), rr + 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) rr[i] = (cg + i < p.cout) ? __bfloat162float(rp[i]) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = p.epi == DBL_EPI_ADD ? f[i] + rr[i] : f[i] * rr[i];
                }
                if (p.ydtype == DBL_BF16) {
                    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)opix * p.ypitch + cg;
                    if (full16 && p.y_vec) {
                        stg8(yp, pack8(f));
                        stg8(yp + 8, pack8(f + 8));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cg + i < p.cout) yp[i] = __float2bfloat16_rn(f[i]);
                    }
                } else {
                    float* yp = reinterpret_cast<float*>(p.y) + (size_t)opix * p.ypitch + cg;
                    if (full16 && p.y_vec) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            reinterpret_cast<float4*>(yp)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cg + i < p.cout) yp[i] = f[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            TC_TRACE(7, it == 0 && warp == 2 && lane == 0);
            TC_TRACE(8, warp == 2 && lane == 0);
        }
    }
    // teardown (the bulk stores read shared memory: they must have finished before the CTA exits)
    if (p.tma_out && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
    TC_TRACE(9, threadIdx.x == 0);
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

static bool encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
        return false;
    }
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                        : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return false;
    }
    return true;
}

bool conv2d_tcgen05_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad, int dil) {
    (void)r;
    if (x->dtype != DBL_BF16 || x->c % 16 != 0 || x->pitch % 8 != 0 || (reinterpret_cast<uintptr_t>(x->data) & 15)) return false;
    if (dil != 1) return false;
    if (k == 1) return stride == 1 && pad == 0;
    if (k == 3) return pad == 1 && (stride == 1 || (stride == 2 && x->h % 2 == 0 && x->w % 2 == 0));
    (void)y;
    return false;
}

static void pick_tile(int h, int w, int* tw_out, int* th_out) {
    long long best = -1;
    int btw = 1, bth = 1;
    for (int tw = 1; tw <= (w < TILE_M ? w : TILE_M); ++tw) {
        int th = TILE_M / tw;
        if (th > h) th = h;
        if (th > 256) th = 256;
        const long long tiles = (long long)cdiv(w, tw) * cdiv(h, th);
        // fewer tiles first, then wider rows (longer contiguous TMA runs)
        const long long score = -(tiles * 1024) + tw;
        if (best == -1 || score > best) best = score, btw = tw, bth = th;
    }
    *tw_out = btw, *th_out = bth;
}

int conv2d_tcgen05(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int k, int stride,
                   int pad, int act, int epi, cudaStream_t st) {
    const int num_sms = dev_info().sms, max_smem = dev_info().max_smem;
    static DevOnce once;  // shared-memory opt-in is per device
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) {
            set_error("conv_tc: cannot opt in to %d bytes of shared memory: %s", max_smem, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, DBL_E_ARG, "conv_tc: weights must be 16-byte aligned");
    TcParams p{};
    p.n = x->n, p.h = x->h, p.w = x->w, p.cin = x->c;
    p.ho = y->h, p.wo = y->w, p.cout = y->c;
    p.k = k, p.stride = stride, p.pad = pad, p.taps = k * k;
    p.cchunk = (x->c % 64 == 0) ? 64 : ((x->c % 32 == 0) ? 32 : 16);
    p.nchunks = x->c / p.cchunk;
    p.flat = (k == 1);
    p.n_tiles = cdiv(y->c, 256);
    p.npad = ((cdiv(y->c, p.n_tiles) + 15) / 16) * 16;
    p.n_tiles = cdiv(y->c, p.npad);
    // accumulator ring in TMEM: as many buffers as fit (<= MAX_ACC); two are not enough to hide the MMA -> epilogue ->
    // MMA latency chain when tiles are small
    // Two CTAs per SM when the reduction is short (1x1 convs with K <= 128: a CTA owns only a handful of tiles, so its
    // prologue / pipeline fill / drain are most of its life and a second resident CTA hides them): each then takes half of
    // TMEM and a small stage ring.
    const int kblocks_total = k * k * (x->c / ((x->c % 64 == 0) ? 64 : ((x->c % 32 == 0) ? 32 : 16)));
    static const int force_two = getenv("DBL_TC_TWO") ? atoi(getenv("DBL_TC_TWO")) : 0;  // tuning aid: K-block limit of the two-CTA mode
    const bool two_ctas = kblocks_total <= (force_two > 0 ? force_two : 2);
    p.nacc = (two_ctas ? 256 : 512) / p.npad;
    if (p.nacc > MAX_ACC) p.nacc = MAX_ACC;
    if (p.nacc < 1) p.nacc = 1;
    int cols = 32;
    while (cols < p.nacc * p.npad) cols <<= 1;
    p.tmem_cols = cols;
    if (p.flat) {
        p.tw = TILE_M, p.th = 1, p.tiles_x = p.tiles_y = 0;
        p.m_tiles = (int)(((long long)x->n * y->h * y->w + TILE_M - 1) / TILE_M);
    } else {
        pick_tile(y->h, y->w, &p.tw, &p.th);
        p.tiles_x = cdiv(y->w, p.tw), p.tiles_y = cdiv(y->h, p.th);
        p.m_tiles = x->n * p.tiles_x * p.tiles_y;
    }
    const uint32_t row_bytes = (uint32_t)p.cchunk * 2;
    p.a_bytes = TILE_M * row_bytes;
    p.b_bytes = (uint32_t)p.npad * row_bytes;
    p.b_bytes = (p.b_bytes + 1023u) & ~1023u;  // keep every stage buffer 1024-byte aligned
    p.a_tx = (uint32_t)(p.flat ? TILE_M : p.tw * p.th) * row_bytes;
    p.b_tx = (uint32_t)p.npad * row_bytes;
    static const int tma_out_env = getenv("DBL_TC_TMA_OUT") ? atoi(getenv("DBL_TC_TMA_OUT")) : 1;  // tuning aid
    // (one N tile only: the 64-channel store box of a partial last sub-tile must not reach into the next tile's channels)
    p.tma_out = (tma_out_env && p.flat && p.n_tiles == 1 && y->dtype == DBL_BF16 && ((size_t)y->pitch * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(y->data) & 15) == 0) ? 1 : 0;
    const size_t fixed = 1024 /*align slack*/ + (((size_t)p.n_tiles * p.npad + 3) & ~(size_t)3) * 4 + (2 * TC_MAX_STAGES + 2 * MAX_ACC) * 8 + 16 +
                         (p.tma_out ? 1024 + 2 * (size_t)TC_OUT_BYTES : 0);
    int stages = (int)(((size_t)max_smem - fixed) / (p.a_bytes + p.b_bytes));
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (two_ctas) {
        const int fit = (int)(((size_t)110 * 1024 - fixed) / (p.a_bytes + p.b_bytes));  // two resident CTAs: <= 110 KB each
        if (stages > 4) stages = 4;
        if (fit >= 2 && stages > fit) stages = fit;
    }
    DBL_REQUIRE(stages >= 2, DBL_E_UNSUPPORTED, "conv_tc: tile does not fit shared memory");
    p.stages = stages;
    p.bias = bias;
    p.r = r ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr;
    p.rpitch = r ? r->pitch : 0;
    p.r_vec = r ? (int)(r->pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(r->data) & 15) == 0) : 0;
    p.y = y->data, p.ypitch = y->pitch, p.ydtype = y->dtype, p.act = act, p.epi = r ? epi : DBL_EPI_NONE;
    const int yelem = y->dtype == DBL_BF16 ? 2 : 4;
    p.y_vec = (int)(((size_t)y->pitch * yelem) % 16 == 0 && (reinterpret_cast<uintptr_t>(y->data) & 15) == 0);
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    const uint64_t pitchB = (uint64_t)x->pitch * 2;
    if (p.flat) {
        const uint64_t dims[2] = {(uint64_t)x->c, (uint64_t)x->n * x->h * x->w};
        const uint64_t strides[1] = {pitchB};
        const uint32_t box[2] = {(uint32_t)p.cchunk, TILE_M};
        if (!encode_map(&maps.a[0], x->data, 2, dims, strides, box, row_bytes)) return DBL_E_CUDA;
    } else if (stride == 1) {
        const uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
        const uint64_t strides[3] = {pitchB, pitchB * x->w, pitchB * x->w * x->h};
        const uint32_t box[4] = {(uint32_t)p.cchunk, (uint32_t)p.tw, (uint32_t)p.th, 1};
        if (!encode_map(&maps.a[0], x->data, 4, dims, strides, box, row_bytes)) return DBL_E_CUDA;
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)((x->w - pw + 1) / 2), (uint64_t)((x->h - ph + 1) / 2), (uint64_t)x->n};
                const uint64_t strides[3] = {2 * pitchB, 2 * pitchB * x->w, pitchB * x->w * x->h};
                const uint32_t box[4] = {(uint32_t)p.cchunk, (uint32_t)p.tw, (uint32_t)p.th, 1};
                const uint8_t* base = reinterpret_cast<const uint8_t*>(x->data) + ((size_t)ph * x->w + pw) * pitchB;
                if (!encode_map(&maps.a[ph * 2 + pw], base, 4, dims, strides, box, row_bytes)) return DBL_E_CUDA;
            }
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.taps * x->c, (uint64_t)y->c};
        const uint64_t strides[1] = {(uint64_t)p.taps * x->c * 2};
        const uint32_t box[2] = {(uint32_t)p.cchunk, (uint32_t)p.npad};
        if (!encode_map(&maps.b, w, 2, dims, strides, box, row_bytes)) return DBL_E_CUDA;
    }
    if (p.tma_out) {
        const uint64_t dims[2] = {(uint64_t)y->c, (uint64_t)y->n * y->h * y->w};
        const uint64_t strides[1] = {(uint64_t)y->pitch * 2};
        const uint32_t box[2] = {64, TILE_M};
        if (!encode_map(&maps.y, y->data, 2, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    const size_t smem = fixed + (size_t)stages * (p.a_bytes + p.b_bytes);
    const int total = p.m_tiles * p.n_tiles;
    const int cap = (two_ctas && smem <= (size_t)110 * 1024) ? 2 * num_sms : num_sms;
    const int grid = total < cap ? total : cap;
    if (act == DBL_ACT_SILU)
        launch_k(conv_tc_kernel<DBL_ACT_SILU>, grid, TC_THREADS, smem, st, maps, p);
    else if (act == DBL_ACT_GELU)
        launch_k(conv_tc_kernel<DBL_ACT_GELU>, grid, TC_THREADS, smem, st, maps, p);
    else
        launch_k(conv_tc_kernel<DBL_ACT_NONE>, grid, TC_THREADS, smem, st, maps, p);
    count_launch();
    return check_launch("dbl_conv2d(tcgen05)");
}

}  // namespace dbl

#ifdef DBL_TRACE
extern "C" int dbl_debug_trace_read(void* dst, size_t bytes) {
    return cudaMemcpyFromSymbol(dst, dbl::g_tc_trace, bytes < sizeof(dbl::g_tc_trace) ? bytes : sizeof(dbl::g_tc_trace)) == cudaSuccess ? 0 : 1;
}
#endif
