// Depthwise-separable convolution as ONE persistent, warp-specialised kernel on the Blackwell data path:
//   TMA halo tile -> depthwise k x k (fp32 FFMA2) -> A operand written in place in the UMMA K-major swizzled layout ->
//   pointwise 1x1 on tcgen05 (accumulator in TMEM) -> bias / SiLU / residual epilogue -> TMA store.
// Reference: DSConv.forward (nn/modules/conv.py:91-108: dw -> pw -> BN -> SiLU, BN folded into the pointwise weights), both halves of
// DSBottleneck (nn/modules/block.py:1408-1444) and the [DWConv+BN+SiLU, Conv1x1+BN+SiLU] pairs of Detect's class branch
// (nn/modules/head.py:93-100, DW_ACT = 1).
//
// Why a second implementation next to dsconv_mma.cu: that kernel runs its three phases (tile fill, depthwise, pointwise + epilogue)
// one after the other in every CTA with 1-2 CTAs per SM -- ncu: issue 48 %, 27 % of the warp slots, a 7x7 layer takes 40 us against
// an FMA-pipe floor of 13 us.  Here a CTA is persistent and the phases of CONSECUTIVE tiles overlap:
//   warp 0        TMA producer: one 4-D box per (tile, 64-channel chunk) -- rows / columns outside the image are zero-filled by the
//                 unit (= the convolution's padding, no per-thread index math) -- plus the residual tile;
//   warps 2..11   depthwise: thread = two output columns x 2 channels x 5 rows, packed fp32 FMAs (one FFMA2 = one tap of a channel pair),
//                 result rounded to bf16 straight into the 128-byte-swizzled A tile (pixel = GEMM row);
//   warp 1        one elected lane issues the tcgen05.mma k-steps (M = 128 rows of which 100 are pixels, N = Cout), accumulator
//                 double-buffered in TMEM, tcgen05.commit releases the A tile / publishes the accumulator;
//   warps 12..19  epilogue: tcgen05.ld -> +bias -> SiLU -> (+|*) residual (from the TMA-loaded tile) -> bf16 -> swizzled staging
//                 tile -> ONE TMA store per 64-channel chunk (clipped at the image border by the unit).
// mbarrier rings connect the roles; weights (pointwise by TMA, depthwise by plain loads) are resident for the CTA's whole life and are
// fetched before griddepcontrol.wait.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int DT_TW = 20, DT_TH = 5, DT_PIX = DT_TW * DT_TH;  // output tile: 100 pixels = the first 100 rows of a 128-row MMA tile
constexpr int DT_NDW = 320;                                   // depthwise threads (10 warps): 16 four-channel groups x 20 columns
constexpr int DT_DW_WARPS = DT_NDW / 32;
constexpr int DT_EPI_WARPS = 8;                               // two per TMEM lane group: one warp per scheduler cannot hide its own latencies
constexpr int DT_THREADS = 64 + DT_NDW + 32 * DT_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..11 depthwise, warps 12..19 epilogue
constexpr int DT_EPI_WARP0 = 2 + DT_DW_WARPS;                 // = 12: a multiple of 4, so epilogue warp w owns TMEM lanes 32*(w%4)..
constexpr uint32_t DT_OCHUNK_BYTES = 13312;                   // 100 rows x 128 B, rounded up to 1024
constexpr int DT_SMEM_LIMIT = 227 * 1024;

__host__ __device__ constexpr int dt_round1k(int v) { return (v + 1023) / 1024 * 1024; }
__host__ __device__ constexpr int dt_cw(int c, int chunk) { return c - 64 * chunk < 64 ? c - 64 * chunk : 64; }  // channels of a 64-channel chunk

template <int K, int CIN, int COUT, bool RES, int S = 1>
struct DtCfg {
    static constexpr int PAD = (K - 1) / 2, IC = (DT_TW - 1) * S + K, IR = (DT_TH - 1) * S + K;
    static constexpr int NCH = (CIN + 63) / 64;     // K chunks of the pointwise GEMM (64 channels = one 128-byte swizzle row)
    // channels per staged input box: the whole CIN when two such stages fit next to everything else (one stage round per tile),
    // else 64-channel chunks
    static constexpr int CW = (CIN <= 64 || 2 * dt_round1k(IR * IC * CIN * 2) > 96 * 1024) ? (CIN < 64 ? CIN : 64) : CIN;
    static constexpr int NST = (CIN + CW - 1) / CW;  // stage rounds per tile
    static constexpr int NOC = (COUT + 63) / 64;    // output-channel chunks (staging / store granularity)
    static constexpr int IN_TX = IR * IC * CW * 2;  // bytes one input box delivers
    static constexpr int IN_BYTES = dt_round1k(IN_TX);
    static constexpr int A_BYTES = NCH * 128 * 128;  // per buffer: a 128-row x 128-byte slab per chunk
    static constexpr int B_CHUNK = dt_round1k(COUT * 128);
    static constexpr int B_BYTES = NCH * B_CHUNK;
    static constexpr int O_BYTES = NOC * (int)DT_OCHUNK_BYTES;
    static constexpr int WD_BYTES = K * K * CIN * 4;
    static constexpr int MISC = dt_round1k(WD_BYTES + (COUT + CIN) * 4 + 512);
    static constexpr int total(int nbuf, int nin) { return 1024 + nin * IN_BYTES + nbuf * A_BYTES + B_BYTES + nbuf * O_BYTES + (RES ? nbuf * O_BYTES : 0) + MISC; }
    static constexpr int NBUF = total(2, 2) <= DT_SMEM_LIMIT ? 2 : 1;  // A / staging / residual buffers: double when they fit
    static constexpr int NIN = total(NBUF, 3) <= DT_SMEM_LIMIT ? 3 : 2; // input stages: a third one hides the TMA round trip of the short 3x3 tiles
    static constexpr int SMEM = total(NBUF, NIN);
    // epilogue groups: with two staging buffers the eight epilogue warps split into two groups of four that drain ALTERNATE tiles
    // concurrently (group g owns accumulator g, staging buffer g, residual slot g); the per-tile chain tcgen05.ld -> SiLU -> staging ->
    // barrier -> TMA store is latency-bound (1.2 us), and it paced the 3x3 tiles
    static constexpr int EG = 1;  // measured: two groups of four warps (alternate tiles) hold each accumulator twice as long -> 3x3 64 ch 12.7 -> 18.0 us; one group of eight is kept
    static constexpr int TMEM_COLS = 2 * COUT <= 64 ? 64 : (2 * COUT <= 128 ? 128 : 256);
    static_assert(total(1, 2) <= DT_SMEM_LIMIT, "tile does not fit shared memory");
    static_assert(CIN % 16 == 0 && COUT % 16 == 0 && COUT <= 128 && CIN <= 256, "channel counts");
};

struct DtMaps {
    CUtensorMap x;     // input [C, W, H, N], box [CW, IC, IR, 1], no swizzle
    CUtensorMap b[4];  // pointwise weights [Cin, Cout], box [chunk width, Cout], swizzle = chunk row bytes
    CUtensorMap y[2];  // output [Cout, W, H, N], box [chunk width, TW, TH, 1], swizzle = chunk row bytes
    CUtensorMap r[2];  // residual, like y
};

struct DtParams {
    const float* wdw;  // [K*K][CIN]
    const float* bdw;  // [CIN] (DW_ACT) or null
    const float* bpw;  // [COUT]
    int n, h, w, tiles_x, tiles_y, ntiles, epi;
    unsigned long long* trace;  // developer aid (dbl_dsconv_tc_debug_trace): CTA 0 writes [role][tile iteration][event] globaltimer stamps
};

constexpr int DT_TRACE_TILES = 16, DT_TRACE_EVENTS = 4;
__device__ __forceinline__ void dt_stamp(const DtParams& p, int role, uint32_t it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < (uint32_t)DT_TRACE_TILES) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(role * DT_TRACE_TILES + it) * DT_TRACE_EVENTS + ev] = t;
    }
}

// position of 16-byte chunk j of row p inside a swizzled tile whose rows are `rowb` bytes (TMA / UMMA 128 / 64 / 32-byte swizzle)
__device__ __forceinline__ uint32_t dt_swz(uint32_t j, uint32_t p, int rowb) {
    return rowb == 128 ? (j ^ (p & 7u)) : (rowb == 64 ? (j ^ ((p >> 1) & 3u)) : (j ^ ((p >> 2) & 1u)));
}

// K-major shared-memory matrix descriptor (same encoding as conv_tc.cu): rows of `row_bytes` (= swizzle span), 8-row groups 8*row_bytes apart
__device__ __forceinline__ uint64_t dt_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int K, int CIN, int COUT, int DW_ACT, bool RES, int S>
__global__ void __launch_bounds__(DT_THREADS, 1) dsconv_tc_kernel(const __grid_constant__ DtMaps maps, const DtParams p) {
    using C = DtCfg<K, CIN, COUT, RES, S>;
    constexpr int NBUF = C::NBUF;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // carve (every region 1024-byte aligned): [NIN x input][NBUF x A][B][NBUF x out][NBUF x res][dw weights | biases | barriers]
    uint8_t* sIn = smem;
    uint8_t* sA = sIn + C::NIN * C::IN_BYTES;
    uint8_t* sB = sA + NBUF * C::A_BYTES;
    uint8_t* sOut = sB + C::B_BYTES;
    uint8_t* sRes = sOut + NBUF * C::O_BYTES;
    uint8_t* sMisc = sRes + (RES ? NBUF * C::O_BYTES : 0);
    float* sWd = reinterpret_cast<float*>(sMisc);
    float* sBp = sWd + K * K * CIN;
    float* sBd = sBp + COUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sBd + CIN) + 15) & ~(uintptr_t)15);
    uint64_t* in_full = bars;         // [3]  TMA -> depthwise
    uint64_t* in_empty = bars + 3;    // [3]  depthwise -> TMA           (count = depthwise warps)
    uint64_t* a_full = bars + 6;      // [2]  depthwise -> MMA           (count = depthwise warps)
    uint64_t* a_empty = bars + 8;     // [2]  MMA (tcgen05.commit) -> depthwise
    uint64_t* t_full = bars + 10;     // [2]  MMA (tcgen05.commit) -> epilogue
    uint64_t* t_empty = bars + 12;    // [2]  epilogue -> MMA            (count = warps of one epilogue group)
    uint64_t* r_full = bars + 14;     // [2]  TMA -> epilogue
    uint64_t* r_empty = bars + 16;    // [2]  epilogue -> TMA            (count = warps of one epilogue group)
    uint64_t* w_full = bars + 18;     // pointwise weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.x);
        prefetch_tmap(&maps.b[0]);
        prefetch_tmap(&maps.y[0]);
        for (int s = 0; s < 3; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], DT_DW_WARPS);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], DT_DW_WARPS);
            mbar_init(&a_empty[s], 1);
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], DT_EPI_WARPS / C::EG);
            mbar_init(&r_full[s], 1);
            mbar_init(&r_empty[s], DT_EPI_WARPS / C::EG);
        }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // constant data (weights, biases): staged before griddepcontrol.wait
    for (int i = threadIdx.x; i < K * K * CIN / 4; i += DT_THREADS)  // 16-byte loads / explicit shared stores (the generic form was 6 % of the kernel)
        sts128(smem_u32(sWd) + (uint32_t)i * 16, __ldg(reinterpret_cast<const uint4*>(p.wdw) + i));
    for (int i = threadIdx.x; i < COUT; i += DT_THREADS) sBp[i] = p.bpw ? p.bpw[i] : 0.f;
    for (int i = threadIdx.x; i < CIN; i += DT_THREADS) sBd[i] = (DW_ACT && p.bdw) ? p.bdw[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {  // pointwise weights: one box per K chunk, resident for the whole kernel
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) bytes += (uint32_t)(COUT * dt_cw(CIN, c) * 2);
        mbar_expect_tx(w_full, bytes);
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) tma_load_2d(&maps.b[c], w_full, sB + c * C::B_CHUNK, 64 * c, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_x * p.tiles_y;
    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t q = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int img = tile / per_img, t = tile - img * per_img;
            const int y0 = (t / p.tiles_x) * DT_TH, x0 = (t % p.tiles_x) * DT_TW;
#pragma unroll 1
            for (int c = 0; c < C::NST; ++c, ++q) {
                const uint32_t s = q % C::NIN, ph = (q / C::NIN) & 1;
                mbar_wait_parked(&in_empty[s], ph ^ 1);
                if (c == 0 && lane == 0) dt_stamp(p, 0, it, 0);
                if (elect_one()) {
                    mbar_expect_tx(&in_full[s], (uint32_t)C::IN_TX);
                    tma_load_4d(&maps.x, &in_full[s], sIn + s * C::IN_BYTES, C::CW * c, x0 * S - C::PAD, y0 * S - C::PAD, img);
                }
                __syncwarp();
            }
            if (RES) {
                const uint32_t rs = it % NBUF, rph = (it / NBUF) & 1;
                mbar_wait_parked(&r_empty[rs], rph ^ 1);
                if (elect_one()) {
                    uint32_t bytes = 0;
#pragma unroll
                    for (int oc = 0; oc < C::NOC; ++oc) bytes += (uint32_t)(DT_PIX * dt_cw(COUT, oc) * 2);
                    mbar_expect_tx(&r_full[rs], bytes);
#pragma unroll
                    for (int oc = 0; oc < C::NOC; ++oc)
                        tma_load_4d(&maps.r[oc], &r_full[rs], sRes + rs * C::O_BYTES + oc * DT_OCHUNK_BYTES, 64 * oc, x0, y0, img);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        mbar_wait_parked(w_full, 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t a = it % NBUF, aph = (it / NBUF) & 1;
            const uint32_t acc = it & 1, tph = (it >> 1) & 1;
            mbar_wait_parked(&t_empty[acc], tph ^ 1);
            if (lane == 0) dt_stamp(p, 1, it, 0);
            mbar_wait_parked(&a_full[a], aph);
            if (lane == 0) dt_stamp(p, 1, it, 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + acc * (uint32_t)COUT;
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < C::NCH; ++c) {
                    const uint32_t rowb = (uint32_t)dt_cw(CIN, c) * 2;
                    const uint64_t ad = dt_desc(smem_u32(sA + a * C::A_BYTES + c * 128 * 128), rowb);
                    const uint64_t bd = dt_desc(smem_u32(sB + c * C::B_CHUNK), rowb);
#pragma unroll
                    for (int ks = 0; ks < dt_cw(CIN, c) / 16; ++ks) umma_bf16(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc, (c | ks) ? 1u : 0u);
                }
                umma_commit(&a_empty[a]);   // the A tile may be overwritten once these MMAs have read it
                umma_commit(&t_full[acc]);  // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp < DT_EPI_WARP0) {
        // ===================== depthwise warps =====================
        const int dtid = threadIdx.x - 64;
        const uint32_t wd_s = smem_u32(sWd), bd_s = smem_u32(sBd);
        uint32_t q = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t a = it % NBUF, aph = (it / NBUF) & 1;
            mbar_wait_parked(&a_empty[a], aph ^ 1);
            if (dtid == 0) dt_stamp(p, 2, it, 0);
#pragma unroll 1
            for (int c = 0; c < C::NST; ++c, ++q) {
                const uint32_t s = q % C::NIN, ph = (q / C::NIN) & 1;
                const int cwid = CIN - C::CW * c < C::CW ? CIN - C::CW * c : C::CW;  // channels staged in this round
                const int np = cwid >> 1;                                            // channel pairs
                const uint32_t in_s = smem_u32(sIn + s * C::IN_BYTES);
                const uint32_t a_buf = smem_u32(sA + a * C::A_BYTES);
                mbar_wait_parked(&in_full[s], ph);
                if (dtid == 0 && c == 0) dt_stamp(p, 2, it, 1);
                // work item = (channel PAIR, column PAIR) x 5 rows: a warp holds 32 channel pairs of one column pair, so every input
                // load is one 128-byte wavefront and every weight / input value feeds two output columns.  (First version: 4 channels x
                // 1 column per thread -- 448 KB of shared-memory reads per 7x7 tile, 250 KB of them the same filter taps re-read by all
                // 20 columns.  Now 230 KB.)
                for (int item = dtid; item < np * (DT_TW / 2); item += DT_NDW) {
                    const int xp = (np & (np - 1)) == 0 ? item >> (31 - __clz(np)) : item / np;
                    const int cp = item - xp * np;
                    const int ch0 = C::CW * c + 2 * cp;            // first of this item's two channels
                    const int slab = ch0 >> 6, cs = ch0 & 63;     // 64-channel K chunk of the pointwise GEMM and the channel inside it
                    const int rowb = dt_cw(CIN, slab) * 2;        // bytes per GEMM row of that chunk's A slab
                    const uint32_t a_s = a_buf + (uint32_t)slab * 128 * 128;
                    unsigned long long acc[DT_TH][2];
                    {
                        unsigned long long b2 = 0ull;
                        if (DW_ACT) {
                            const uint2 bb = lds64(bd_s + (uint32_t)ch0 * 4);
                            b2 = ((unsigned long long)bb.y << 32) | bb.x;
                        }
#pragma unroll
                        for (int r = 0; r < DT_TH; ++r) acc[r][0] = acc[r][1] = b2;
                    }
                    const uint32_t xbase = in_s + (uint32_t)(2 * S * xp * C::CW + 2 * cp) * 2;  // staged box: CW channels per pixel
                    unsigned long long wprev[K], wcur[K];
#pragma unroll
                    for (int kxp = 0; kxp < K + S; ++kxp) {  // input column 2*S*xp + kxp feeds output column 0 with filter column kxp and output column 1 with kxp - S
                        if (S == 1) {
                            if (kxp < K) {
#pragma unroll
                                for (int ky = 0; ky < K; ++ky) {
                                    const uint2 w0 = lds64(wd_s + (uint32_t)(((ky * K + kxp) * CIN + ch0) * 4));
                                    wcur[ky] = ((unsigned long long)w0.y << 32) | w0.x;
                                }
                            }
                        } else {  // stride 2: the two filter columns in flight are not neighbours; k = 3, so reloading is cheap
#pragma unroll
                            for (int ky = 0; ky < K; ++ky) {
                                if (kxp < K) {
                                    const uint2 w0 = lds64(wd_s + (uint32_t)(((ky * K + kxp) * CIN + ch0) * 4));
                                    wcur[ky] = ((unsigned long long)w0.y << 32) | w0.x;
                                }
                                if (kxp >= S && kxp - S < K) {
                                    const uint2 w1 = lds64(wd_s + (uint32_t)(((ky * K + kxp - S) * CIN + ch0) * 4));
                                    wprev[ky] = ((unsigned long long)w1.y << 32) | w1.x;
                                }
                            }
                        }
#pragma unroll
                        for (int iy = 0; iy < C::IR; ++iy) {
                            bool used = false;
#pragma unroll
                            for (int r = 0; r < DT_TH; ++r) used = used || (iy - S * r >= 0 && iy - S * r < K);
                            if (!used) continue;
                            const unsigned long long xv = bf16x2_to_f32x2(lds32(xbase + (uint32_t)((iy * C::IC + kxp) * C::CW) * 2));
#pragma unroll
                            for (int r = 0; r < DT_TH; ++r) {
                                const int d = iy - S * r;
                                if (d >= 0 && d < K) {
                                    if (kxp < K) ffma2(acc[r][0], xv, wcur[d]);
                                    if (kxp >= S && kxp - S < K) ffma2(acc[r][1], xv, wprev[d]);
                                }
                            }
                        }
                        if (S == 1) {
#pragma unroll
                            for (int ky = 0; ky < K; ++ky) wprev[ky] = wcur[ky];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < DT_TH; ++r) {
#pragma unroll
                        for (int cx = 0; cx < 2; ++cx) {
                            float f0, f1;
                            f32x2_split(acc[r][cx], f0, f1);
                            if (DW_ACT) f0 = silu_fast(f0), f1 = silu_fast(f1);
                            const uint32_t pix = (uint32_t)(r * DT_TW + 2 * xp + cx);
                            sts32(a_s + pix * (uint32_t)rowb + (dt_swz((uint32_t)(cs >> 3), pix, rowb) << 4) + (uint32_t)(cs & 7) * 2, pack2(f0, f1));
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&in_empty[s]);  // this warp has finished reading the staged chunk
            }
            fence_proxy_async();  // the A tile was written through the generic proxy; tcgen05.mma reads it through the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[a]);
            if (dtid == 0) dt_stamp(p, 2, it, 2);
        }
    } else {
        // ===================== epilogue warps =====================
        constexpr int EG = C::EG, GW = DT_EPI_WARPS / EG;  // epilogue groups, warps per group
        constexpr int NPASS = EG;                          // a group of four warps converts its columns in two passes (register budget)
        constexpr int NG = (COUT / 16 + 1) / 2;            // 16-column groups per warp and pass
        const int wi = warp - DT_EPI_WARP0;
        const int lg = warp & 3;                           // TMEM lane group this warp may access
        const int grp = EG == 2 ? (wi >> 2) : 0;           // EG = 2: group g drains the tiles with (iteration & 1) == g
        const int half0 = EG == 2 ? 0 : (wi >> 2);         // EG = 1: the two warps of a lane group share the columns
        const uint32_t row = (uint32_t)(lg * 32 + lane);
        const bool valid = row < (uint32_t)DT_PIX;
        const bool issuer = (wi == grp * 4 && lane == 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            if (EG == 2 && (int)(it & 1) != grp) continue;
            const uint32_t acc = it & 1, tph = (it >> 1) & 1;
            const uint32_t ob = it % NBUF, oph = (it / NBUF) & 1;
            if (RES) mbar_wait_parked(&r_full[ob], oph);
            mbar_wait_parked(&t_full[acc], tph);
            tc_fence_after();
            if (issuer) dt_stamp(p, 3, it, 0);
            const uint32_t o_s = smem_u32(sOut + ob * C::O_BYTES), r_s = smem_u32(sRes + (RES ? ob * C::O_BYTES : 0));
#pragma unroll
            for (int pass = 0; pass < NPASS; ++pass) {
                const int half = half0 + pass;
                // this pass's accumulator columns with ONE wait (dependent tcgen05.ld round trips cost ~0.25 us each)
                uint32_t v[NG][16];
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) {
                    const int c0 = (2 * gi + half) * 16;
                    if (c0 < COUT) tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)COUT + c0, v[gi]);
                }
                if (pass == 0) {
                    if (issuer) {  // the bulk store that last read this staging buffer has finished reading it
                        if (EG == 1 && NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    if (EG == 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(32 * GW) : "memory");
                    else asm volatile("bar.sync 1, %0;" ::"n"(32 * GW) : "memory");
                    if (issuer) dt_stamp(p, 3, it, 1);
                }
                tmem_ld_wait();
                if (pass == NPASS - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[acc]);  // the accumulator is in registers: the MMA warp may reuse these TMEM columns
                }
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) {
                    const int c0 = (2 * gi + half) * 16;
                    if (c0 >= COUT) continue;
                    const int oc = c0 >> 6, rowb = dt_cw(COUT, oc) * 2;
                    const uint32_t j = (uint32_t)((c0 & 63) >> 3);
                    float f[16];
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float4 bq = lds_f4(sBp + c0 + 4 * qd);
                        f[4 * qd + 0] = silu_fast(__uint_as_float(v[gi][4 * qd + 0]) + bq.x);
                        f[4 * qd + 1] = silu_fast(__uint_as_float(v[gi][4 * qd + 1]) + bq.y);
                        f[4 * qd + 2] = silu_fast(__uint_as_float(v[gi][4 * qd + 2]) + bq.z);
                        f[4 * qd + 3] = silu_fast(__uint_as_float(v[gi][4 * qd + 3]) + bq.w);
                    }
                    if (valid) {
                        const uint32_t off0 = (uint32_t)oc * DT_OCHUNK_BYTES + row * (uint32_t)rowb + (dt_swz(j, row, rowb) << 4);
                        const uint32_t off1 = (uint32_t)oc * DT_OCHUNK_BYTES + row * (uint32_t)rowb + (dt_swz(j + 1, row, rowb) << 4);
                        if (RES) {
                            float rr[16];
                            bf16x8 r0, r1;
                            r0.u = lds128(r_s + off0), r1.u = lds128(r_s + off1);
                            unpack8(r0, rr);
                            unpack8(r1, rr + 8);
#pragma unroll
                            for (int i = 0; i < 16; ++i) f[i] = p.epi == DBL_EPI_MUL ? f[i] * rr[i] : f[i] + rr[i];
                        }
                        sts128(o_s + off0, pack8(f).u);
                        sts128(o_s + off1, pack8(f + 8).u);
                    }
                }
            }
            __syncwarp();
            if (RES && lane == 0) mbar_arrive(&r_empty[ob]);
            fence_proxy_async();
            if (EG == 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(32 * GW) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(32 * GW) : "memory");
            if (issuer) {
                const int img = tile / per_img, t = tile - img * per_img;
                const int y0 = (t / p.tiles_x) * DT_TH, x0 = (t % p.tiles_x) * DT_TW;
#pragma unroll
                for (int oc = 0; oc < C::NOC; ++oc) tma_store_4d(&maps.y[oc], sOut + ob * C::O_BYTES + oc * DT_OCHUNK_BYTES, 64 * oc, x0, y0, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                dt_stamp(p, 3, it, 2);
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores read shared memory: finish before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*DtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DtEncodeFn dt_get_encode() {
    static DtEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<DtEncodeFn>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool dt_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes) {
    DtEncodeFn enc = dt_get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
        return false;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides[i];
    const CUtensorMapSwizzle sw = sw_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : sw_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                   : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("dsconv_tc: cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] swizzle %d", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, sw_bytes);
        return false;
    }
    return true;
}

static bool dt_map_nhwc(CUtensorMap* m, const dbl_tensor* t, int c0, int cw, int bw, int bh, int sw_bytes) {
    const uint64_t pitchB = (uint64_t)t->pitch * 2;
    const uint64_t dims[4] = {(uint64_t)t->c, (uint64_t)t->w, (uint64_t)t->h, (uint64_t)t->n};
    const uint64_t strides[3] = {pitchB, pitchB * t->w, pitchB * t->w * t->h};
    const uint32_t box[4] = {(uint32_t)cw, (uint32_t)bw, (uint32_t)bh, 1};
    (void)c0;
    return dt_encode(m, t->data, 4, dims, strides, box, sw_bytes);
}

static unsigned long long* g_dt_trace = nullptr;  // developer aid, see dbl_dsconv_tc_debug_trace
unsigned long long* debug_trace_buffer() { return g_dt_trace; }  // conv_tc5.cu stamps into the same buffer

template <int K, int CIN, int COUT, int DW_ACT, bool RES, int S = 1>
static int launch_dt(const dbl_tensor* x, const float* wdw, const float* bdw, const void* wpw, const float* bpw, const dbl_tensor* r,
                     const dbl_tensor* y, int epi, cudaStream_t st) {
    using C = DtCfg<K, CIN, COUT, RES, S>;
    auto kern = dsconv_tc_kernel<K, CIN, COUT, DW_ACT, RES, S>;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("dbl_dsconv_tc: cannot opt in to %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DtMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (!dt_map_nhwc(&maps.x, x, 0, C::CW, C::IC, C::IR, 0)) return DBL_E_CUDA;
    for (int c = 0; c < C::NCH; ++c) {
        const int cw = dt_cw(CIN, c);
        const uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)COUT};
        const uint64_t strides[1] = {(uint64_t)CIN * 2};
        const uint32_t box[2] = {(uint32_t)cw, (uint32_t)COUT};
        if (!dt_encode(&maps.b[c], wpw, 2, dims, strides, box, cw * 2)) return DBL_E_CUDA;
    }
    for (int oc = 0; oc < C::NOC; ++oc) {
        const int cw = dt_cw(COUT, oc);
        if (!dt_map_nhwc(&maps.y[oc], y, 64 * oc, cw, DT_TW, DT_TH, cw * 2)) return DBL_E_CUDA;
        if (RES && !dt_map_nhwc(&maps.r[oc], r, 64 * oc, cw, DT_TW, DT_TH, cw * 2)) return DBL_E_CUDA;
    }
    DtParams p;
    memset(&p, 0, sizeof(p));
    p.wdw = wdw, p.bdw = bdw, p.bpw = bpw;
    p.n = x->n, p.h = x->h, p.w = x->w, p.epi = epi, p.trace = g_dt_trace;
    p.tiles_x = cdiv(y->w, DT_TW), p.tiles_y = cdiv(y->h, DT_TH);
    const long long nt = (long long)x->n * p.tiles_x * p.tiles_y;
    DBL_REQUIRE(nt < (1ll << 31), DBL_E_UNSUPPORTED, "dbl_dsconv_tc: too many tiles");
    p.ntiles = (int)nt;
    const int sms = dev_info().sms;
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch_k(kern, grid, DT_THREADS, (size_t)C::SMEM, st, maps, p);
    count_launch();
    return check_launch("dbl_dsconv_tc");
}

}  // namespace dbl

using namespace dbl;

// (k, cin, cout): DSBottleneck halves at n / s scale and the Detect class-branch pairs (nc = 80: c3 = 80; nc <= 64: c3 = 64)
#define DBL_DT_TABLE(X) \
    X(3, 32, 32) X(7, 32, 32) X(3, 64, 64) X(7, 64, 64) X(3, 128, 128) X(7, 128, 128) X(3, 64, 80) X(3, 80, 80) X(3, 128, 80) X(3, 128, 64) X(3, 256, 80) X(3, 256, 64)

extern "C" int dbl_dsconv_tc_debug_trace(void* device_buffer) {
    g_dt_trace = reinterpret_cast<unsigned long long*>(device_buffer);
    return 4 * DT_TRACE_TILES * DT_TRACE_EVENTS * (int)sizeof(unsigned long long);
}

extern "C" int dbl_dsconv_tc_supported(int k, int stride, int cin, int cout) {
    if (stride == 2) return k == 3 && cin == 128 && cout == 128;  // DSConv [512, 3, 2] of the backbone at n scale (layer 8)
    if (stride != 1) return 0;
#define X(K_, CI_, CO_) \
    if (k == K_ && cin == CI_ && cout == CO_) return 1;
    DBL_DT_TABLE(X)
#undef X
    return 0;
}

extern "C" int dbl_dsconv_tc(const dbl_tensor* x, const float* w_dw, const float* b_dw, int dw_act, const void* w_pw, const float* b_pw,
                             const dbl_tensor* r, const dbl_tensor* y, int k, int epi, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w_dw && w_pw, DBL_E_ARG, "dbl_dsconv_tc: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y), DBL_E_ARG, "dbl_dsconv_tc: x and y must be bf16 NHWC views, 16-byte aligned, pitch %% 8 == 0");
    DBL_REQUIRE(y->n == x->n, DBL_E_ARG, "dbl_dsconv_tc: batch mismatch");
    const bool s2 = y->h != x->h || y->w != x->w;  // stride 2 is recognised by the output shape: (h + 2*pad - k) / 2 + 1
    if (s2)
        DBL_REQUIRE(y->n == x->n && y->h == (x->h + 2 * ((k - 1) / 2) - k) / 2 + 1 && y->w == (x->w + 2 * ((k - 1) / 2) - k) / 2 + 1, DBL_E_ARG,
                    "dbl_dsconv_tc: output shape matches neither stride 1 nor stride 2");
    DBL_REQUIRE(((reinterpret_cast<uintptr_t>(w_dw) | reinterpret_cast<uintptr_t>(w_pw)) & 15) == 0, DBL_E_ARG, "dbl_dsconv_tc: weights must be 16-byte aligned");
    DBL_REQUIRE(!dw_act || b_dw, DBL_E_ARG, "dbl_dsconv_tc: dw_act needs the depthwise bias");
    if (epi != DBL_EPI_NONE) DBL_REQUIRE(valid_tensor(r) && same_shape(r, y) && vec8_ok(r), DBL_E_ARG, "dbl_dsconv_tc: residual mismatch");
    DBL_REQUIRE(x->data != y->data, DBL_E_ARG, "dbl_dsconv_tc: x and y must not alias (tiles read their neighbours' halo)");
    DBL_REQUIRE(x->n <= 65535 && x->h < 32768 && x->w < 32768, DBL_E_UNSUPPORTED, "dbl_dsconv_tc: tensor too large");
    cudaStream_t st = as_stream(stream);
    const bool res = epi != DBL_EPI_NONE;
    DBL_REQUIRE(!(dw_act && res), DBL_E_UNSUPPORTED, "dbl_dsconv_tc: dw_act with a residual is not instantiated");
    if (s2) {
        DBL_REQUIRE(k == 3 && x->c == 128 && y->c == 128 && !dw_act && !res, DBL_E_UNSUPPORTED, "dbl_dsconv_tc: stride 2 is instantiated for k 3, 128 -> 128, no residual");
        return launch_dt<3, 128, 128, 0, false, 2>(x, w_dw, b_dw, w_pw, b_pw, nullptr, y, epi, st);
    }
#define X(K_, CI_, CO_)                                                                                                       \
    if (k == K_ && x->c == CI_ && y->c == CO_) {                                                                              \
        if (dw_act) return launch_dt<K_, CI_, CO_, 1, false>(x, w_dw, b_dw, w_pw, b_pw, nullptr, y, epi, st);                 \
        return res ? launch_dt<K_, CI_, CO_, 0, true>(x, w_dw, b_dw, w_pw, b_pw, r, y, epi, st)                               \
                   : launch_dt<K_, CI_, CO_, 0, false>(x, w_dw, b_dw, w_pw, b_pw, nullptr, y, epi, st);                       \
    }
    DBL_DT_TABLE(X)
#undef X
    set_error("dbl_dsconv_tc: no instantiation for k %d cin %d cout %d", k, x->c, y->c);
    return DBL_E_UNSUPPORTED;
}
