// Layers 0 and 1 of the graph (cfg/models/v13/yolov13_DBL.yaml:9-10 at n scale: Conv [8, 3, 1] then Conv [16, 3, 2], nn/modules/conv.py:39-63)
// together with the input cast / scale of BasePredictor.preprocess (engine/predictor.py:116-134) as ONE persistent, warp-specialised
// tcgen05 kernel; the 8 x 640 x 640 intermediate lives in shared memory.  Same formulation as bottleneck_tc.cu -- a GEMM row is a group of
// consecutive pixels, the filter is a band matrix over the group's window -- with 8-pixel groups in both stages:
//   stage 0 (3 -> 8, stride 1):  input pixels are {R, G, B, 0} bf16 (8 bytes), a row = 8 pixels = 64 bytes, a k-step = 4 pixels x 4 channels,
//                                N = 8 pixels x 8 channels = 64; the 10-pixel window of a group is its own row plus 32 bytes of the next:
//                                3 k-steps per filter row, 9 MMAs per 1024 pixels (the warp-MMA kernel it replaces: 64 m16n8k16 per 1024).
//   stage 1 (8 -> 16, stride 2): hidden pixels are 16 bytes, a row = 8 pixels = 128 bytes, a k-step = 2 pixels x 8 channels, N = 4 output
//                                pixels x 16 channels = 64; hidden rows are stored by row parity, so the three filter rows are plane offsets;
//                                5 k-steps per filter row (the 9-pixel window = the row plus 32 bytes of the next).
// The linear column of an image pixel is skewed by one per stage (input column = x - X0 + 2, hidden column = x - X0 + 1) so that every window
// starts at its row: all k-step slices are 32-byte aligned pieces inside one swizzle row, offsets only reach forward.
//
// Roles (32 warps): warp 0 is the TMA producer (band matrices once, then one 4-D box of the raw NCHW tile per tile: rows / columns outside the
// image are zero-filled = layer 0's padding), warp 1 issues the MMAs (stage 0 of tile i+1 before stage 1 of tile i), warps 4-7
// convert the raw f32 / u8 tile (now in shared memory) to the bf16 A tile (the /255 of the u8 path is folded into the band matrix), warps 8-23 turn stage-0
// accumulators into the hidden tile (bias, SiLU, zero outside the image = layer 1's padding), warps 24-31 turn stage-1 accumulators into
// 128-byte pixel-group rows of the output tile that ONE TMA store writes.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int ST_PROD_WARPS = 4, ST_E0_WARPS = 16, ST_EF_WARPS = 8;
constexpr int ST_THREADS = 128 + 32 * (ST_PROD_WARPS + ST_E0_WARPS + ST_EF_WARPS);  // 1024
constexpr int ST_NIN = 2;   // converted A tiles
constexpr int ST_NRAW = 2;  // raw NCHW tiles (TMA ring)
constexpr int ST_N = 64;  // accumulator columns of both stages
constexpr int ST_C0 = 8, ST_C1 = 16;
constexpr int ST_SMEM_LIMIT = 227 * 1024;

struct StMaps {
    CUtensorMap x;   // raw input [W, H, 3, N] (f32 or u8), box [8 GW, 2 TH + 3, 3, 1], no swizzle; outside the image = zero fill = layer 0's padding
    CUtensorMap w0;  // band matrix of layer 0 [9 * 16, 64], box [16, 64], 32-byte swizzle
    CUtensorMap w1;  // band matrix of layer 1 [15 * 16, 64], box [16, 64]
    CUtensorMap y;   // pixel-group view of the output [128 B = 4 pixels x 16 channels, W1/4, H1, N], box [128 B, GW - 2, TH, 1], 128-byte swizzle
};

struct StParams {
    const float* bias0;  // [64] per accumulator column
    const float* bias1;  // [64]
    int h, w, gw, th, tiles_x, tiles_y, ntiles;
    float inv_gw, inv_pairs;  // 1 / gw, 1 / (4 * gw) (pixel pairs per tile row)
    uint32_t in_bytes, plane_bytes, out_bytes, raw_bytes, raw_tx;
    uint32_t off_raw, off_hid, off_out, off_w0, off_w1, off_misc;
    unsigned long long* trace;
};

unsigned long long* debug_trace_buffer();  // dsconv_tc.cu
__device__ __forceinline__ void st_stamp(const StParams& p, int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 16) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(role * 16 + it) * 4 + ev] = t;
    }
}

__device__ __forceinline__ uint64_t st_desc(uint32_t saddr, uint32_t row_bytes) {  // K-major, rows of `row_bytes` = swizzle span
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <typename TIn>
__device__ __forceinline__ void st_lds_pair(uint32_t a, float& lo, float& hi);
template <>
__device__ __forceinline__ void st_lds_pair<float>(uint32_t a, float& lo, float& hi) {
    const uint2 v = lds64(a);
    lo = __uint_as_float(v.x), hi = __uint_as_float(v.y);
}
template <>
__device__ __forceinline__ void st_lds_pair<uint8_t>(uint32_t a, float& lo, float& hi) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    lo = (float)(v & 0xff), hi = (float)(v >> 8);  // 0..255 is exact in bf16; the 1/255 lives in the band matrix
}

template <typename TIn>
__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const __grid_constant__ StMaps maps, const StParams p) {
    constexpr int ROFF = sizeof(TIn) == 4 ? 2 : 6;    // pixels between the raw box start and input column 0
    constexpr int RPAD = sizeof(TIn) == 4 ? 4 : 16;   // raw row width = 8 GW + RPAD pixels (a multiple of 16 bytes)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sIn = smem;                  // [2] converted input tiles: rows of 8 pixels x 8 B, 64-byte swizzle
    uint8_t* sRaw = smem + p.off_raw;     // [2] raw tiles [3 planes][2 TH + 3 rows][8 GW + RPAD columns] of TIn, written by TMA
    uint8_t* sHid = smem + p.off_hid;     // [2][2 planes] hidden tiles: rows of 8 pixels x 16 B, 128-byte swizzle; plane 0 = even hidden rows
    uint8_t* sOut = smem + p.off_out;     // [2] staging tiles [TH][GW - 2] x 128 B
    uint8_t* sW0 = smem + p.off_w0;       // [9]  x (64 rows x 32 B)
    uint8_t* sW1 = smem + p.off_w1;       // [15] x (64 rows x 32 B)
    float* sB0 = reinterpret_cast<float*>(smem + p.off_misc);
    float* sB1 = sB0 + ST_N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + ST_N);
    uint64_t* in_full = bars;         // [3]     converters -> MMA                    count = 4 warps
    uint64_t* in_empty = bars + 3;    // [3]     MMA (commit after stage 0) -> converters
    uint64_t* t0_full = bars + 6;     // [2][2]  MMA -> E0, per M tile
    uint64_t* t0_empty = bars + 10;   // [2]     E0 -> MMA                            count = 16
    uint64_t* hid_full = bars + 12;   // [2]     E0 -> MMA (stage 1)                  count = 16
    uint64_t* hid_empty = bars + 14;  // [2]     MMA (commit after stage 1) -> E0
    uint64_t* t1_full = bars + 16;    // [2]     MMA -> EF
    uint64_t* t1_empty = bars + 18;   // [2]     EF -> MMA                            count = 8
    uint64_t* w_full = bars + 20;
    uint64_t* raw_full = bars + 21;   // [2]     TMA -> converters
    uint64_t* raw_empty = bars + 23;  // [2]     converters -> TMA                    count = 4 warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.x);
        prefetch_tmap(&maps.w0);
        prefetch_tmap(&maps.w1);
        prefetch_tmap(&maps.y);
        for (int s = 0; s < ST_NIN; ++s) {
            mbar_init(&in_full[s], ST_PROD_WARPS);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t0_empty[s], ST_E0_WARPS);
            mbar_init(&hid_full[s], ST_E0_WARPS);
            mbar_init(&hid_empty[s], 1);
            mbar_init(&t1_full[s], 1);
            mbar_init(&t1_empty[s], ST_EF_WARPS);
            mbar_init(&t0_full[s * 2], 1);
            mbar_init(&t0_full[s * 2 + 1], 1);
        }
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], ST_PROD_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // every buffer a window can reach holds finite values from the start (0 x NaN would poison a whole accumulator row)
    {
        const uint32_t z0 = smem_u32(smem), zn = p.off_w0;
        for (uint32_t o = threadIdx.x * 16u; o < zn; o += ST_THREADS * 16u) sts128(z0 + o, make_uint4(0, 0, 0, 0));
    }
    for (int i = threadIdx.x; i < 2 * ST_N; i += ST_THREADS) sB0[i] = i < ST_N ? p.bias0[i] : p.bias1[i - ST_N];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(w_full, (uint32_t)((9 + 15) * ST_N * 32));
        for (int s = 0; s < 9; ++s) tma_load_2d(&maps.w0, w_full, sW0 + s * (ST_N * 32), s * 16, 0);
        for (int s = 0; s < 15; ++s) tma_load_2d(&maps.w1, w_full, sW1 + s * (ST_N * 32), s * 16, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_x * p.tiles_y;
    const int n_my = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tw = 4 * (p.gw - 2);      // output columns per tile
    const int hrows = 2 * p.th + 1;     // hidden rows per tile
    if (warp == 0) {
        // ===================== TMA producer: raw NCHW tiles =====================
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int img = tile / per_img, t = tile - img * per_img;
            const int y1 = (t / p.tiles_x) * p.th, x1 = (t % p.tiles_x) * tw;
            const uint32_t s = (uint32_t)it % ST_NRAW, ph = ((uint32_t)it / ST_NRAW) & 1;
            mbar_wait_parked(&raw_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&raw_full[s], p.raw_tx);
                // input column 0 is image column 2 x1 - 10; the box starts ROFF pixels further left: TMA wants the innermost start on a
                // 16-byte boundary (a start of -10 pixels faults with "illegal instruction", tools/ubench/tma_nchw.cu)
                tma_load_4d(&maps.x, &raw_full[s], sRaw + s * p.raw_bytes, 2 * x1 - 10 - ROFF, 2 * y1 - 2, 0, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ST_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t w0_s = smem_u32(sW0), w1_s = smem_u32(sW1), in0_s = smem_u32(sIn), hid0_s = smem_u32(sHid);
        mbar_wait_parked(w_full, 0);
        auto stage0 = [&](int it) {  // hidden rows rh = 0..2*TH of tile `it` from input rows rh + dy
            const uint32_t s = (uint32_t)it % ST_NIN, ph = ((uint32_t)it / ST_NIN) & 1;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&t0_empty[a], aph ^ 1);
            mbar_wait_parked(&in_full[s], ph);
            tc_fence_after();
            if (lane == 0) st_stamp(p, 1, it, 0);
            if (elect_one()) {
                const uint64_t b0 = st_desc(w0_s, 32);
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const uint32_t d_tmem = tmem_u + (a * 2 + q) * (uint32_t)ST_N;
                    uint64_t ad = st_desc(in0_s + s * p.in_bytes + (uint32_t)(q * 128) * 64u, 64);
                    uint64_t bd = b0;
#pragma unroll 1
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int st = 0; st < 3; ++st) umma_bf16(d_tmem, ad + 2 * st, bd + ((ST_N * 32) >> 4) * st, idesc, (dy | st) ? 1u : 0u);
                        ad += (uint64_t)(((uint32_t)p.gw * 64u) >> 4);
                        bd += (uint64_t)(3 * ((ST_N * 32) >> 4));
                    }
                    umma_commit(&t0_full[a * 2 + q]);
                }
                umma_commit(&in_empty[s]);
            }
            __syncwarp();
            if (lane == 0) st_stamp(p, 1, it, 1);
        };
        auto stage1 = [&](int it) {  // output rows r = 0..TH-1: filter row 0 / 2 = even-plane rows r / r + 1, filter row 1 = odd-plane row r
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&t1_empty[a], aph ^ 1);
            mbar_wait_parked(&hid_full[a], aph);
            tc_fence_after();
            if (lane == 0) st_stamp(p, 1, it, 2);
            if (elect_one()) {
                const uint32_t hid_s = hid0_s + a * 2 * p.plane_bytes;
                const uint32_t d_tmem = tmem_u + 256u + a * (uint32_t)ST_N;
                uint64_t bd = st_desc(w1_s, 32);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const uint64_t ad = st_desc(hid_s + (dy == 1 ? p.plane_bytes : 0u) + (dy == 2 ? (uint32_t)p.gw * 128u : 0u), 128);
#pragma unroll
                    for (int st = 0; st < 5; ++st) umma_bf16(d_tmem, ad + 2 * st, bd + ((ST_N * 32) >> 4) * st, idesc, (dy | st) ? 1u : 0u);
                    bd += (uint64_t)(5 * ((ST_N * 32) >> 4));
                }
                umma_commit(&t1_full[a]);
                umma_commit(&hid_empty[a]);
            }
            __syncwarp();
            if (lane == 0) st_stamp(p, 1, it, 3);
        };
        if (n_my > 0) stage0(0);
        for (int it = 0; it < n_my; ++it) {
            if (it + 1 < n_my) stage0(it + 1);
            stage1(it);
        }
    } else if (warp >= 4 && warp < 4 + ST_PROD_WARPS) {
        // ===================== converters: raw tile in shared memory -> {R, G, B, 0} bf16 A tile =====================
        const int ptid = threadIdx.x - 128;  // 0..127
        const int pairs_row = 4 * p.gw, in_rows = hrows + 2;
        const int rw = 8 * p.gw + RPAD;  // raw row width in pixels
        const uint32_t plane_b = (uint32_t)(in_rows * rw * (int)sizeof(TIn));
        for (int it = 0; it < n_my; ++it) {
            const uint32_t s = (uint32_t)it % ST_NIN, ph = ((uint32_t)it / ST_NIN) & 1;
            const uint32_t rs = (uint32_t)it % ST_NRAW, rph = ((uint32_t)it / ST_NRAW) & 1;
            mbar_wait_parked(&in_empty[s], ph ^ 1);
            mbar_wait_parked(&raw_full[rs], rph);
            if (ptid == 0) st_stamp(p, 0, it, 0);
            const uint32_t in_s = smem_u32(sIn + s * p.in_bytes), raw_s = smem_u32(sRaw + rs * p.raw_bytes);
            const int total = in_rows * pairs_row;
#pragma unroll 2
            for (int idx = ptid; idx < total; idx += 32 * ST_PROD_WARPS) {  // idx = (row, pixel pair): consecutive lanes read consecutive pairs
                const int ri = __float2int_rd(((float)idx + 0.5f) * p.inv_pairs);
                const int cp = idx - ri * pairs_row;
                float r0, r1, g0, g1, b0, b1;
                const uint32_t src = raw_s + (uint32_t)(ri * rw + ROFF + 2 * cp) * (uint32_t)sizeof(TIn);
                st_lds_pair<TIn>(src, r0, r1);
                st_lds_pair<TIn>(src + plane_b, g0, g1);
                st_lds_pair<TIn>(src + 2 * plane_b, b0, b1);
                const uint4 o = make_uint4(pack2(r0, g0), pack2(b0, 0.f), pack2(r1, g1), pack2(b1, 0.f));
                const uint32_t row64 = (uint32_t)(ri * p.gw + (cp >> 2));
                sts128(in_s + row64 * 64u + ((((uint32_t)cp & 3u) ^ ((row64 >> 1) & 3u)) << 4), o);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&raw_empty[rs]);
                mbar_arrive(&in_full[s]);
            }
            if (ptid == 0) st_stamp(p, 0, it, 1);
        }
    } else if (warp >= 8 && warp < 8 + ST_E0_WARPS) {
        // ===================== E0: stage-0 accumulators -> hidden tile =====================
        const int wi = warp - 8;
        const int lg = wi & 3, q = (wi >> 2) & 1, half = wi >> 3;
        const uint32_t m = (uint32_t)(q * 128 + lg * 32 + lane);  // GEMM row = (hidden row, pixel group)
        const int rh = __float2int_rd(((float)m + 0.5f) * p.inv_gw);
        const int g = (int)m - rh * p.gw;
        const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16);
        const uint32_t hid0_s = smem_u32(sHid);
        const bool stamp = (wi == 0 && lane == 0);
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int t = tile % per_img;
            const int y1 = (t / p.tiles_x) * p.th, x1 = (t % p.tiles_x) * tw;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&hid_empty[a], aph ^ 1);
            mbar_wait_parked(&t0_full[a * 2 + q], aph);
            tc_fence_after();
            if (stamp) st_stamp(p, 2, it, 0);
            uint32_t v[2][16];
            tmem_ld16(lane_base + (a * 2 + q) * (uint32_t)ST_N + half * 32, v[0]);
            tmem_ld16(lane_base + (a * 2 + q) * (uint32_t)ST_N + half * 32 + 16, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t0_empty[a]);
            if (rh < hrows) {
                const int gy = 2 * y1 - 1 + rh;                        // image row of this hidden row
                const bool rowok = gy >= 0 && gy < p.h;
                const uint32_t prow = (uint32_t)((rh >> 1) * p.gw + g);  // row of the parity plane
                const uint32_t row_s = hid0_s + (a * 2 + (uint32_t)(rh & 1)) * p.plane_bytes + prow * 128u;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {  // 16-byte chunk = the 8 channels of hidden pixel c of the group
                    const int c = half * 4 + cc;
                    const int gx = 2 * x1 - 9 + 8 * g + c;             // image column: X0 - 1 + hidden column
                    float f[8];
                    const float4 ba = lds_f4(sB0 + c * 8), bb = lds_f4(sB0 + c * 8 + 4);
                    const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = silu_fast(__uint_as_float(v[cc >> 1][(cc & 1) * 8 + i]) + bs[i]);
                    uint4 o = pack8(f).u;
                    if (!(rowok && gx >= 0 && gx < p.w)) o = make_uint4(0, 0, 0, 0);  // layer 1's zero padding applies to the hidden tensor
                    sts128(row_s + (((uint32_t)c ^ (prow & 7u)) << 4), o);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hid_full[a]);
            if (stamp) st_stamp(p, 2, it, 1);
        }
    } else if (warp >= 8 + ST_E0_WARPS) {
        // ===================== EF: stage-1 accumulators -> output tile =====================
        const int wi = warp - 8 - ST_E0_WARPS;
        const int lg = wi & 3, half = wi >> 2;
        const uint32_t m = (uint32_t)(lg * 32 + lane);  // GEMM row = (output row, group of 4 output pixels)
        const int r = __float2int_rd(((float)m + 0.5f) * p.inv_gw);
        const int g = (int)m - r * p.gw;
        const bool out_ok = r < p.th && g >= 1 && g <= p.gw - 2;  // one group of padding on each side
        const uint32_t orow = (uint32_t)(r * (p.gw - 2) + g - 1);
        const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16);
        const bool issuer = (wi == 0 && lane == 0);
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int img = tile / per_img, t = tile - img * per_img;
            const int y1 = (t / p.tiles_x) * p.th, x1 = (t % p.tiles_x) * tw;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&t1_full[a], aph);
            tc_fence_after();
            if (issuer) st_stamp(p, 3, it, 0);
            uint32_t v[2][16];
            tmem_ld16(lane_base + 256u + a * (uint32_t)ST_N + half * 32, v[0]);
            tmem_ld16(lane_base + 256u + a * (uint32_t)ST_N + half * 32 + 16, v[1]);
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of two tiles ago has read this staging tile
            asm volatile("bar.sync 1, %0;" ::"n"(32 * ST_EF_WARPS) : "memory");
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t1_empty[a]);
            const uint32_t out_row = smem_u32(sOut) + a * p.out_bytes + orow * 128u;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {  // 16-byte chunk c8 of the group's 128-byte output row (4 pixels x 16 channels)
                const int c8 = half * 4 + cc;
                float f[8];
                const float4 ba = lds_f4(sB1 + c8 * 8), bb = lds_f4(sB1 + c8 * 8 + 4);
                const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = silu_fast(__uint_as_float(v[cc >> 1][(cc & 1) * 8 + i]) + bs[i]);
                if (out_ok) sts128(out_row + (((uint32_t)c8 ^ (orow & 7u)) << 4), pack8(f).u);
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * ST_EF_WARPS) : "memory");
            if (issuer) {
                tma_store_4d(&maps.y, sOut + a * p.out_bytes, 0, x1 / 4, y1, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                st_stamp(p, 3, it, 1);
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- band matrices ------------------------------------------------------------------------------------------------
// packed blob: [w0 band: 64 rows x 144 bf16][w1 band: 64 rows x 240 bf16][bias0: 64 f32][bias1: 64 f32]
__global__ void stem_tc_pack_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const __nv_bfloat16* __restrict__ w1, const float* __restrict__ b1,
                                    float scale, uint8_t* __restrict__ out) {
    constexpr int K0 = 9 * 16, K1 = 15 * 16;
    __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(out);
    __nv_bfloat16* o1 = o0 + ST_N * K0;
    float* ob = reinterpret_cast<float*>(o1 + ST_N * K1);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int e = tid; e < ST_N * K0; e += nt) {  // layer 0: row n = hidden pixel j x channel co; k = (dy, window pixel, input channel of {R,G,B,0})
        const int n = e / K0, k = e % K0;
        const int j = n / ST_C0, co = n % ST_C0;
        const int dy = k / 48, st = (k / 16) % 3, wpx = st * 4 + (k % 16) / 4, ci = k % 4;
        const int dx = wpx - j;
        o0[e] = __float2bfloat16((dx >= 0 && dx < 3 && ci < 3) ? w0[((co * 3 + ci) * 3 + dy) * 3 + dx] * scale : 0.f);
    }
    for (int e = tid; e < ST_N * K1; e += nt) {  // layer 1: row n = output pixel j x channel co; k = (dy, hidden pixel of the window, ci)
        const int n = e / K1, k = e % K1;
        const int j = n / ST_C1, co = n % ST_C1;
        const int dy = k / 80, st = (k / 16) % 5, hpx = st * 2 + (k % 16) / ST_C0, ci = k % ST_C0;
        const int dx = hpx - 2 * j;
        o1[e] = (dx >= 0 && dx < 3) ? w1[((co * 3 + dy) * 3 + dx) * ST_C0 + ci] : __float2bfloat16(0.f);
    }
    for (int e = tid; e < 2 * ST_N; e += nt) ob[e] = e < ST_N ? b0[e % ST_C0] : b1[(e - ST_N) % ST_C1];
}

static size_t st_packed_bytes() { return (size_t)ST_N * (144 + 240) * 2 + 2 * ST_N * 4; }

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*StEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static StEncodeFn st_get_encode() {
    static StEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<StEncodeFn>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool st_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes,
                      CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    StEncodeFn enc = st_get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
        return false;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides[i];
    const CUtensorMapSwizzle sw = sw_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : sw_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : sw_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                   : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("dbl_stem_pair_tc: cuTensorMapEncodeTiled failed (CUresult %d): rank %d box [%u,%u,%u,%u] swizzle %d", (int)r, rank, box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, sw_bytes);
        return false;
    }
    return true;
}

static uint32_t st_round1k(size_t v) { return (uint32_t)((v + 1023) / 1024 * 1024); }

template <typename TIn>
static int launch_stem_tc(const void* x, int n, int h, int w, const uint8_t* packed, const dbl_tensor* y, cudaStream_t st) {
    constexpr int RPAD = sizeof(TIn) == 4 ? 4 : 16;
    auto kern = stem_tc_kernel<TIn>;
    static DevOnce once;
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_LIMIT);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("dbl_stem_pair_tc: cannot opt in to %d bytes of shared memory: %s", ST_SMEM_LIMIT, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    const int sms = dev_info().sms;
    const uint32_t w_bytes = (9 + 15) * ST_N * 32;
    // ---- tile: GW pixel groups per row (8 input / hidden pixels = 4 output pixels each, one group of padding per side); hidden rows
    // (2 TH + 1) * GW <= 256 = two M tiles of stage 0, output rows TH * GW <= 128 = one M tile of stage 1
    StParams p;
    memset(&p, 0, sizeof(p));
    long long best = -1;
    size_t best_smem = 0;
    for (int gw = 3; gw <= 64; ++gw) {
        int th = (256 / gw - 1) / 2;
        if (th > 128 / gw) th = 128 / gw;
        if (th < 1) continue;
        const int tw = 4 * (gw - 2);
        if (sizeof(TIn) == 1 && (gw & 1)) continue;  // TMA: the box must start on a 16-byte boundary of the image row: 2 x1 = 8 (GW - 2) k pixels
        const uint32_t in_b = st_round1k((size_t)(257 + 2 * gw) * 64), plane_b = st_round1k((size_t)(130 + gw) * 128), out_b = st_round1k((size_t)th * (gw - 2) * 128);
        if (8 * gw + RPAD > 256) continue;  // TMA box extent
        const uint32_t raw_b = st_round1k((size_t)3 * (2 * th + 3) * (8 * gw + RPAD) * sizeof(TIn));
        const size_t smem = 1024 + (size_t)ST_NIN * in_b + (size_t)ST_NRAW * raw_b + 4 * (size_t)plane_b + 2 * (size_t)out_b + w_bytes + 1024;
        if (smem > (size_t)ST_SMEM_LIMIT) continue;
        const long long tiles = (long long)n * cdiv(y->w, tw) * cdiv(y->h, th);
        const long long cost = ((tiles + sms - 1) / sms) * (1ll << 32) + tiles;
        if (best < 0 || cost < best) {
            best = cost, best_smem = smem;
            p.gw = gw, p.th = th;
            p.in_bytes = in_b, p.plane_bytes = plane_b, p.out_bytes = out_b, p.raw_bytes = raw_b;
        }
    }
    DBL_REQUIRE(best >= 0, DBL_E_UNSUPPORTED, "dbl_stem_pair_tc: no tile fits");
    const int tw = 4 * (p.gw - 2);
    p.h = h, p.w = w;
    p.raw_tx = (uint32_t)(3 * (2 * p.th + 3) * (8 * p.gw + RPAD) * sizeof(TIn));
    p.tiles_x = cdiv(y->w, tw), p.tiles_y = cdiv(y->h, p.th);
    const long long nt = (long long)n * p.tiles_x * p.tiles_y;
    DBL_REQUIRE(nt < (1ll << 31), DBL_E_UNSUPPORTED, "dbl_stem_pair_tc: too many tiles");
    p.ntiles = (int)nt;
    p.inv_gw = 1.0f / (float)p.gw, p.inv_pairs = 1.0f / (float)(4 * p.gw);
    p.off_raw = ST_NIN * p.in_bytes;
    p.off_hid = p.off_raw + ST_NRAW * p.raw_bytes;
    p.off_out = p.off_hid + 4 * p.plane_bytes;
    p.off_w0 = p.off_out + 2 * p.out_bytes;
    p.off_w1 = p.off_w0 + 9 * ST_N * 32;
    p.off_misc = p.off_w1 + 15 * ST_N * 32;
    p.trace = debug_trace_buffer();
    const __nv_bfloat16* wb0 = reinterpret_cast<const __nv_bfloat16*>(packed);
    const __nv_bfloat16* wb1 = wb0 + ST_N * 144;
    const float* bias = reinterpret_cast<const float*>(wb1 + ST_N * 240);
    p.bias0 = bias, p.bias1 = bias + ST_N;
    StMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        const uint64_t es = sizeof(TIn);
        const uint64_t dims[4] = {(uint64_t)w, (uint64_t)h, 3, (uint64_t)n};
        const uint64_t strides[3] = {(uint64_t)w * es, (uint64_t)w * h * es, (uint64_t)3 * w * h * es};
        const uint32_t box[4] = {(uint32_t)(8 * p.gw + RPAD), (uint32_t)(2 * p.th + 3), 3, 1};
        if (!st_encode(&maps.x, x, 4, dims, strides, box, 0, sizeof(TIn) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {144, ST_N};
        const uint64_t strides[1] = {144 * 2};
        const uint32_t box[2] = {16, ST_N};
        if (!st_encode(&maps.w0, wb0, 2, dims, strides, box, 32)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {240, ST_N};
        const uint64_t strides[1] = {240 * 2};
        const uint32_t box[2] = {16, ST_N};
        if (!st_encode(&maps.w1, wb1, 2, dims, strides, box, 32)) return DBL_E_CUDA;
    }
    {
        const uint64_t yp = (uint64_t)y->pitch * 2;  // == 32 (dense)
        const uint64_t dims[4] = {64, (uint64_t)y->w / 4, (uint64_t)y->h, (uint64_t)y->n};
        const uint64_t strides[3] = {128, yp * y->w, yp * y->w * y->h};
        const uint32_t box[4] = {64, (uint32_t)(p.gw - 2), (uint32_t)p.th, 1};
        if (!st_encode(&maps.y, y->data, 4, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch_k(kern, grid, ST_THREADS, best_smem, st, maps, p);
    count_launch();
    return check_launch("dbl_stem_pair_tc");
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_stem_tc_supported(int c0, int c1, int h, int w, int ypitch) {  // w % 16: the u8 image rows must be 16-byte multiples for TMA
    return c0 == ST_C0 && c1 == ST_C1 && ypitch == ST_C1 && h % 2 == 0 && w % 16 == 0 && w >= 16 && h >= 2;
}

extern "C" size_t dbl_stem_tc_packed_bytes(void) { return st_packed_bytes(); }

extern "C" int dbl_stem_tc_pack(const float* w0, const float* b0, const void* w1, const float* b1, int x_is_u8, void* packed, dbl_stream_t stream) {
    DBL_REQUIRE(w0 && b0 && w1 && b1 && packed, DBL_E_ARG, "dbl_stem_tc_pack: null argument");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, DBL_E_ARG, "dbl_stem_tc_pack: packed must be 16-byte aligned");
    stem_tc_pack_kernel<<<32, 256, 0, as_stream(stream)>>>(w0, b0, reinterpret_cast<const __nv_bfloat16*>(w1), b1, x_is_u8 ? 1.0f / 255.0f : 1.0f,
                                                           reinterpret_cast<uint8_t*>(packed));
    return check_launch("dbl_stem_tc_pack");
}

extern "C" int dbl_stem_pair_tc(const void* x_nchw, int x_is_u8, int n, int h, int w, const void* packed, const dbl_tensor* y, dbl_stream_t stream) {
    DBL_REQUIRE(x_nchw && packed && valid_tensor(y), DBL_E_ARG, "dbl_stem_pair_tc: null/invalid argument");
    DBL_REQUIRE(vec8_ok(y) && y->n == n && y->h == h / 2 && y->w == w / 2, DBL_E_ARG, "dbl_stem_pair_tc: y must be bf16 NHWC (n, h/2, w/2, 16), 16-byte aligned");
    DBL_REQUIRE(dbl_stem_tc_supported(ST_C0, y->c, h, w, y->pitch), DBL_E_UNSUPPORTED,
                "dbl_stem_pair_tc: needs (c0, c1) = (8, 16), a dense output (pitch == 16), h even and w %% 16 == 0");
    DBL_REQUIRE(((reinterpret_cast<uintptr_t>(x_nchw) | reinterpret_cast<uintptr_t>(packed)) & 15) == 0, DBL_E_ARG, "dbl_stem_pair_tc: x and packed must be 16-byte aligned");
    DBL_REQUIRE(n <= 65535, DBL_E_UNSUPPORTED, "dbl_stem_pair_tc: batch too large");
    cudaStream_t st = as_stream(stream);
    return x_is_u8 ? launch_stem_tc<uint8_t>(x_nchw, n, h, w, reinterpret_cast<const uint8_t*>(packed), y, st)
                   : launch_stem_tc<float>(x_nchw, n, h, w, reinterpret_cast<const uint8_t*>(packed), y, st);
}
