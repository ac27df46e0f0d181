// Bottleneck.forward (nn/modules/block.py:344-357: y = x + SiLU(cv2(SiLU(cv1(x)))), 3x3 / 3x3, e = 0.5) for the THIN high-resolution
// layers (C = 16 @320^2 and C = 32 @160^2 at n scale: layers 2 and 4 of cfg/models/v13/yolov13_DBL.yaml) as ONE persistent,
// warp-specialised tcgen05 kernel; the hidden tensor lives in shared memory.
//
// Why a new formulation.  tools/ubench/umma_rate.cu: an M=128, K=16 tcgen05.mma with both operands in shared memory costs 32 + N/4
// cycles -- the 4 KB A tile is read at 128 B/clk whatever N is -- so a 16-channel convolution written as nine taps of N = 8..16 pays
// ~40 cycles per tap per 128 pixels and is bound by the A-operand reads (the warp-MMA kernel it replaces, conv_mma.cu, pays the same
// bytes through ldmatrix).  Here a GEMM row is a GROUP of G = 64/C consecutive pixels (one 128-byte row of the NHWC tensor) and the
// N side holds the G output pixels of the group: N = G * Cout.  A k-step is one pixel (16 channels) of the (G+2)-pixel window of the
// group, taken from the same pixel-major tile with the descriptor start shifted by whole pixels (window pixel -1 is the tail of the
// previous 128-byte row, pixel G the head of the next), and the B operand is the filter laid out as a band matrix (zero where
// |input pixel - output pixel| > 1).  Every input byte is read 3 times instead of 9, the accumulator holds complete sums (no
// shifted adds in the epilogue, no wider TMEM drain), and all N columns are useful:
//     C = 16:  cv1 18 MMAs (N = 32) + cv2 12 MMAs (N = 64) per 512 pixels        (tap form: 36 + 36 with half of N padding)
//     C = 32:  cv1 24 MMAs (N = 32) + cv2 12 MMAs (N = 64) per 256 pixels        (tap form: 36 + 18)
//
// Pipeline.  warp 0: TMA producer (one 4-D box per tile: the halo tile as 128-byte pixel-group rows, zero fill = padding);
// warp 1: MMA issuer; warps 20-23: hidden-tile epilogue E1 (one warp per TMEM lane group, both M tiles), warps 4-19: output epilogue E2 in TWO groups of eight warps that drain ALTERNATE tiles (group = accumulator set = staging buffer).
// The MMA warp issues cv1(i+1) before cv2(i): while the hidden tile of one tile is converted
// (TMEM -> bias -> SiLU -> zero outside the image -> bf16 -> 64-byte-swizzled rows in shared memory), the tensor pipe works on the next.
// E2 adds the residual (TMA-loaded into the staging tile itself, added in place), stages the bf16 result as 128-byte pixel-group rows (a lane
// owns one row: conflict-free with the 128-byte swizzle; per-pixel rows were 32-way bank conflicts, 1.7 us per tile) and ONE TMA store writes it.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int BT_E2_WARPS = 16, BT_E1_WARPS = 4;
constexpr int BT_THREADS = 128 + 32 * (BT_E2_WARPS + BT_E1_WARPS);  // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-19 output epilogue (E2, two groups of 8), warps 20-23 hidden-tile epilogue (E1)
constexpr int BT_NIN = 3;                 // input stages (released by cv1's tcgen05.commit)
constexpr int BT_N1 = 32, BT_N2 = 64;     // accumulator columns per M tile: G * C/2 and G * C
constexpr int BT_SMEM_LIMIT = 227 * 1024;

struct BtMaps {
    CUtensorMap x;   // pixel-group view of the input [128 B, W/G, H, N], box [128 B, GW, TH + 4, 1], 128-byte swizzle
    CUtensorMap w1;  // band matrix of cv1 [K1, 32], box [16, 32], 32-byte swizzle
    CUtensorMap w2;  // band matrix of cv2 [192, 64], box [16, 64]
    CUtensorMap r;   // pixel-group view of the input with the OUTPUT tile's box: the residual, loaded straight into the staging tile
    CUtensorMap y;   // pixel-group view of the output [128 B, W/G, H, N], box [128 B, TW/G, TH, 1], 128-byte swizzle
};

struct BtParams {
    const float* bias1;  // [32]  cv1 bias per accumulator column
    const float* bias2;  // [64]
    int h, w, gw, th, tw, tiles_x, tiles_y, ntiles, shortcut;
    float inv_gw;
    uint32_t in_bytes, in_tx, out_bytes;                   // per input stage / TMA bytes per tile / staging tile
    uint32_t off_hid, off_out, off_w1, off_w2, off_misc;   // carve-up of the 1024-byte-aligned dynamic shared memory
    unsigned long long* trace;  // developer aid (dbl_dsconv_tc_debug_trace): CTA 0 writes [role][tile iteration][event] globaltimer stamps
};

unsigned long long* debug_trace_buffer();  // dsconv_tc.cu
__device__ __forceinline__ void bt_stamp(const BtParams& p, int role, int it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 16) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(role * 16 + it) * 4 + ev] = t;
    }
}

__device__ __forceinline__ uint64_t bt_desc(uint32_t saddr, uint32_t row_bytes) {  // K-major, rows of `row_bytes` = swizzle span
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int C>
__global__ void __launch_bounds__(BT_THREADS, 1) bottleneck_tc_kernel(const __grid_constant__ BtMaps maps, const BtParams p) {
    constexpr int CH = C / 2, G = 64 / C, P = G, KC = C / 16;
    constexpr int S1 = (G + 2) * KC;  // cv1 k-steps per filter row
    constexpr int S2 = 4;             // cv2 k-steps per filter row: 64 hidden bytes = the window [-32 B, +96 B) around the group
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sIn = smem + 1024;                    // [3] halo tiles behind a zeroed guard (window pixel -1 of the first group)
    uint8_t* sHid = smem + p.off_hid;              // [2] hidden tiles: 256 rows x 64 B, behind a zeroed guard
    uint8_t* sOut = smem + p.off_out;              // [2] staging tiles [TH][TW/G] x 128 B, one per E2 group
    uint8_t* sW1 = smem + p.off_w1;                // [3 * S1] x (32 rows x 32 B)
    uint8_t* sW2 = smem + p.off_w2;                // [12] x (64 rows x 32 B)
    float* sB1 = reinterpret_cast<float*>(smem + p.off_misc);
    float* sB2 = sB1 + BT_N1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + BT_N2);
    uint64_t* in_full = bars;         // [3]     TMA -> MMA (cv1)
    uint64_t* in_empty = bars + 3;    // [3]     MMA (tcgen05.commit after cv1) -> TMA
    uint64_t* t1_full = bars + 6;     // [2][2]  MMA -> E1, per M tile
    uint64_t* t1_empty = bars + 10;   // [2]     E1 -> MMA                                count = 8
    uint64_t* hid_full = bars + 12;   // [2]     E1 -> MMA (cv2)                          count = 8
    uint64_t* hid_empty = bars + 14;  // [2]     MMA (tcgen05.commit after cv2) -> E1
    uint64_t* t2_full = bars + 16;    // [2][2]  MMA -> E2, per M tile
    uint64_t* t2_empty = bars + 20;   // [2]     E2 -> MMA                                count = 8
    uint64_t* w_full = bars + 22;
    uint64_t* r_full = bars + 23;     // [2]     TMA (residual tile -> staging buffer of E2 group) -> E2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.x);
        prefetch_tmap(&maps.w1);
        prefetch_tmap(&maps.w2);
        prefetch_tmap(&maps.y);
        prefetch_tmap(&maps.r);
        for (int s = 0; s < 3; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t1_empty[s], BT_E1_WARPS);
            mbar_init(&hid_full[s], BT_E1_WARPS);
            mbar_init(&hid_empty[s], 1);
            mbar_init(&t2_empty[s], BT_E2_WARPS / 2);
            for (int q = 0; q < 2; ++q) {
                mbar_init(&t1_full[s * 2 + q], 1);
                mbar_init(&t2_full[s * 2 + q], 1);
            }
        }
        mbar_init(w_full, 1);
        mbar_init(&r_full[0], 1);
        mbar_init(&r_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Zero everything a window can touch before its first real fill (guards, tile tails, hidden tiles): an MMA row whose window
    // reaches a NaN bit pattern would be NaN in ALL its columns, also where the band matrix is zero.
    {
        const uint32_t z0 = smem_u32(smem), zn = p.off_out;
        for (uint32_t o = threadIdx.x * 16u; o < zn; o += BT_THREADS * 16u) sts128(z0 + o, make_uint4(0, 0, 0, 0));
    }
    for (int i = threadIdx.x; i < BT_N1 + BT_N2; i += BT_THREADS) sB1[i] = i < BT_N1 ? p.bias1[i] : p.bias2[i - BT_N1];
    fence_proxy_async();  // the zeroes were written through the generic proxy; TMA and tcgen05.mma use the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {  // both band matrices: resident for the CTA's whole life, fetched before griddepcontrol.wait
        mbar_expect_tx(w_full, (uint32_t)(3 * S1 * BT_N1 * 32 + 3 * S2 * BT_N2 * 32));
        for (int s = 0; s < 3 * S1; ++s) tma_load_2d(&maps.w1, w_full, sW1 + s * (BT_N1 * 32), s * 16, 0);
        for (int s = 0; s < 3 * S2; ++s) tma_load_2d(&maps.w2, w_full, sW2 + s * (BT_N2 * 32), s * 16, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_x * p.tiles_y;
    const int n_my = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        // ===================== TMA producer =====================
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int img = tile / per_img, t = tile - img * per_img;
            const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
            const uint32_t s = (uint32_t)it % BT_NIN, ph = ((uint32_t)it / BT_NIN) & 1;
            mbar_wait_parked(&in_empty[s], ph ^ 1);
            if (lane == 0) bt_stamp(p, 0, it, 0);
            if (elect_one()) {
                mbar_expect_tx(&in_full[s], p.in_tx);
                tma_load_4d(&maps.x, &in_full[s], sIn + s * p.in_bytes, 0, x0 / G - 1, y0 - 2, img);  // P = G pixels = one group of left halo
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT_N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT_N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t w1_s = smem_u32(sW1), w2_s = smem_u32(sW2), in0_s = smem_u32(sIn), hid0_s = smem_u32(sHid);
        const uint32_t row_step1 = (uint32_t)p.gw * 128u, row_step2 = (uint32_t)p.gw * 64u;
        mbar_wait_parked(w_full, 0);
        auto cv1 = [&](int it) {  // hidden rows r = 0..TH+1 of tile `it` from input rows r + dy
            const uint32_t s = (uint32_t)it % BT_NIN, ph = ((uint32_t)it / BT_NIN) & 1;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&t1_empty[a], aph ^ 1);
            mbar_wait_parked(&in_full[s], ph);
            tc_fence_after();
            if (lane == 0) bt_stamp(p, 1, it, 0);
            if (elect_one()) {
                const uint32_t in_s = in0_s + s * p.in_bytes;
                // running descriptors: only the start-address field changes from MMA to MMA (a fully unrolled form with 36 precomputed
                // descriptor pairs overflowed the uniform register file: R2UR.FILL chains, 59-65 cycles per MMA)
                const uint64_t b0 = bt_desc(w1_s, 32);
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const uint32_t d_tmem = tmem_u + (a * 2 + q) * (uint32_t)BT_N1;
                    uint64_t ad = bt_desc(in_s + (uint32_t)(q * 128) * 128u - 2 * C, 128);  // window pixel -1 of filter row 0
                    uint64_t bd = b0;
#pragma unroll 1
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int st = 0; st < S1; ++st)  // st = (window pixel + 1) * KC + 16-channel slice: consecutive 32-byte pieces, constant offsets
                            umma_bf16(d_tmem, ad + 2 * st, bd + ((BT_N1 * 32) >> 4) * st, idesc1, (dy | st) ? 1u : 0u);
                        ad += (uint64_t)(row_step1 >> 4);
                        bd += (uint64_t)(S1 * ((BT_N1 * 32) >> 4));
                    }
                    umma_commit(&t1_full[a * 2 + q]);
                }
                umma_commit(&in_empty[s]);  // the halo tile may be refilled once these MMAs have read it
            }
            __syncwarp();
            if (lane == 0) bt_stamp(p, 1, it, 1);
        };
        auto cv2 = [&](int it) {  // output rows r = 0..TH-1 from hidden rows r + dy
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&t2_empty[a], aph ^ 1);
            mbar_wait_parked(&hid_full[a], aph);
            tc_fence_after();
            if (lane == 0) bt_stamp(p, 1, it, 2);
            if (elect_one()) {
                const uint32_t hid_s = hid0_s + a * (256u * 64u + 1024u);
                const uint64_t b0 = bt_desc(w2_s, 32);
#pragma unroll 1
                for (int q = 0; q < 2; ++q) {
                    const uint32_t d_tmem = tmem_u + 128u + (a * 2 + q) * (uint32_t)BT_N2;
                    uint64_t ad = bt_desc(hid_s + (uint32_t)(q * 128) * 64u - 32u, 64);  // the window starts 32 bytes before the group
                    uint64_t bd = b0;
#pragma unroll 1
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int st = 0; st < S2; ++st) umma_bf16(d_tmem, ad + 2 * st, bd + ((BT_N2 * 32) >> 4) * st, idesc2, (dy | st) ? 1u : 0u);
                        ad += (uint64_t)(row_step2 >> 4);
                        bd += (uint64_t)(S2 * ((BT_N2 * 32) >> 4));
                    }
                    umma_commit(&t2_full[a * 2 + q]);
                }
                umma_commit(&hid_empty[a]);  // the hidden tile may be overwritten once these MMAs have read it
            }
            __syncwarp();
            if (lane == 0) bt_stamp(p, 1, it, 3);
        };
        if (n_my > 0) cv1(0);
        for (int it = 0; it < n_my; ++it) {
            if (it + 1 < n_my) cv1(it + 1);
            cv2(it);
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const bool is_e1 = warp >= 4 + BT_E2_WARPS;
        const int wi = is_e1 ? warp - 4 - BT_E2_WARPS : warp - 4;
        const int lg = wi & 3;               // TMEM lane group (= warp % 4)
        const int q2 = (wi >> 2) & 1;        // E2: M tile
        const int grp = wi >> 3;             // E2: group = parity of the tiles it drains
        const bool issuer = ((wi & 7) == 0 && lane == 0);  // of its E2 group (the TMA store) / of E1 (the trace)
        const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16);
        const uint32_t in0_s = smem_u32(sIn), hid0_s = smem_u32(sHid), out_s = smem_u32(sOut);
        auto e1 = [&](int it) {  // accumulators of cv1 -> hidden tile (both M tiles of the warp's lane group)
            const int tile = blockIdx.x + it * gridDim.x;
            const int t = tile % per_img;
            const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;
            mbar_wait_parked(&hid_empty[a], aph ^ 1);
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                const uint32_t m = (uint32_t)(q * 128 + lg * 32 + lane);     // GEMM row = pixel group
                const int r = __float2int_rd(((float)m + 0.5f) * p.inv_gw);  // exact: (m + 0.5) / gw is never an integer
                const int g = (int)m - r * p.gw;
                mbar_wait_parked(&t1_full[a * 2 + q], aph);
                tc_fence_after();
                if (issuer && q == 0) bt_stamp(p, 3, it, 0);
                uint32_t v[2][16];
                tmem_ld16(lane_base + (a * 2 + q) * (uint32_t)BT_N1, v[0]);
                tmem_ld16(lane_base + (a * 2 + q) * (uint32_t)BT_N1 + 16, v[1]);
                tmem_ld_wait();
                if (q == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t1_empty[a]);
                }
                if (r < p.th + 2) {
                    const int gy = y0 - 1 + r;
                    const bool rowok = gy >= 0 && gy < p.h;
                    const uint32_t row_s = hid0_s + a * (256u * 64u + 1024u) + m * 64u;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {  // 16-byte chunk = 8 hidden channels of pixel c4 / (CH/8)
                        const int gx = x0 - P + g * G + c4 / (CH / 8);
                        float f[8];
                        const float4 ba = lds_f4(sB1 + c4 * 8), bb = lds_f4(sB1 + c4 * 8 + 4);
                        const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = silu_fast(__uint_as_float(v[c4 >> 1][(c4 & 1) * 8 + i]) + bs[i]);
                        uint4 o = pack8(f).u;
                        if (!(rowok && gx >= 0 && gx < p.w)) o = make_uint4(0, 0, 0, 0);  // cv2's zero padding applies to the hidden tensor
                        sts128(row_s + (((uint32_t)c4 ^ ((m >> 1) & 3u)) << 4), o);
                    }
                }
            }
            fence_proxy_async();  // written through the generic proxy, read by tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(&hid_full[a]);
            if (issuer) bt_stamp(p, 3, it, 1);
        };
        auto e2 = [&](int it) {  // accumulator of cv2 -> + bias, SiLU, + x -> staging -> TMA store; this warp: M tile q2, 32 rows, all 64 columns
            const int tile = blockIdx.x + it * gridDim.x;
            const int img = tile / per_img, t = tile - img * per_img;
            const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
            const uint32_t a = (uint32_t)it & 1, aph = ((uint32_t)it >> 1) & 1;  // a == grp
            const uint32_t m = (uint32_t)(q2 * 128 + lg * 32 + lane);
            const int r = __float2int_rd(((float)m + 0.5f) * p.inv_gw);
            const int g = (int)m - r * p.gw;
            const bool out_ok = r < p.th && g >= 1 && g <= p.gw - 2;  // the halo tile has one group of padding on each side
            // residual: ONE TMA load of x's output tile into this group's staging tile (same swizzled pixel-group rows), issued as soon as
            // the store that last read the buffer has finished reading it; the conversion then adds in place.  (Reading it from the
            // halo tile kept that stage until here and starved the TMA ring: 1.3 us of cv1 waiting per tile; per-lane global loads
            // exposed their L2 latency inside the conversion.)
            if (issuer) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (p.shortcut) {
                    mbar_expect_tx(&r_full[a], (uint32_t)(p.th * (p.gw - 2) * 128));
                    tma_load_4d(&maps.r, &r_full[a], sOut + a * p.out_bytes, 0, x0 / G, y0, img);
                }
            }
            mbar_wait_parked(&t2_full[a * 2 + q2], aph);
            tc_fence_after();
            if (issuer && grp == 0) bt_stamp(p, 3, it >> 1, 2);
            const uint32_t orow = (uint32_t)(r * (p.gw - 2) + g - 1);  // row of the [TH][TW/G] staging tile
            const uint32_t out_row = out_s + a * p.out_bytes + orow * 128u, out_x = orow & 7u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[2][16];
                tmem_ld16(lane_base + 128u + (a * 2 + q2) * (uint32_t)BT_N2 + half * 32, v[0]);
                tmem_ld16(lane_base + 128u + (a * 2 + q2) * (uint32_t)BT_N2 + half * 32 + 16, v[1]);
                if (half == 0) {
                    if (p.shortcut) mbar_wait_parked(&r_full[a], aph);  // residual tile landed (also: the staging buffer is ours again)
                    else if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                    else asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (issuer && grp == 0) bt_stamp(p, 2, it >> 1, 0);
                }
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t2_empty[a]);
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {  // 16-byte chunk c8 of the group's 128-byte output row
                    const int c8 = half * 4 + cc;
                    float f[8];
                    const float4 ba = lds_f4(sB2 + c8 * 8), bb = lds_f4(sB2 + c8 * 8 + 4);  // explicit ld.shared (a carved float* is a generic pointer)
                    const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = silu_fast(__uint_as_float(v[cc >> 1][(cc & 1) * 8 + i]) + bs[i]);
                    if (out_ok) {
                        const uint32_t o = out_row + (((uint32_t)c8 ^ out_x) << 4);
                        if (p.shortcut) {
                            bf16x8 rv;
                            rv.u = lds128(o);
                            float rr[8];
                            unpack8(rv, rr);
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] += rr[i];
                        }
                        sts128(o, pack8(f).u);
                    }
                }
            }
            if (issuer && grp == 0) bt_stamp(p, 2, it >> 1, 1);
            fence_proxy_async();
            if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 2, 256;" ::: "memory");
            if (issuer) {
                tma_store_4d(&maps.y, sOut + a * p.out_bytes, 0, x0 / G, y0, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (grp == 0) bt_stamp(p, 3, it >> 1, 3);
            }
        };
        if (is_e1) {
            for (int it = 0; it < n_my; ++it) e1(it);
        } else {
            for (int it = grp; it < n_my; it += 2) e2(it);
            if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores read shared memory: finish before the CTA exits
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- band matrices ------------------------------------------------------------------------------------------------
// packed blob: [w1 band: 32 rows x K1 bf16][w2 band: 64 rows x 192 bf16][bias1: 32 f32][bias2: 64 f32], K1 = 3 * (G+2) * C
template <int C>
__global__ void bottleneck_tc_pack_kernel(const __nv_bfloat16* __restrict__ w1, const float* __restrict__ b1, const __nv_bfloat16* __restrict__ w2,
                                          const float* __restrict__ b2, uint8_t* __restrict__ out) {
    constexpr int CH = C / 2, G = 64 / C, KC = C / 16, S1 = (G + 2) * KC, K1 = 3 * S1 * 16, K2 = 192, PX2 = 16 / CH;
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(out);
    __nv_bfloat16* o2 = o1 + BT_N1 * K1;
    float* ob = reinterpret_cast<float*>(o2 + BT_N2 * K2);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int e = tid; e < BT_N1 * K1; e += nt) {  // cv1: row n = output pixel j x hidden channel co, k = (dy, window pixel w, channel ci)
        const int n = e / K1, k = e % K1;
        const int j = n / CH, co = n % CH;
        const int dy = k / (S1 * 16), st = (k / 16) % S1, ci = (st % KC) * 16 + k % 16, wpx = st / KC - 1;
        const int dx = wpx - j + 1;
        o1[e] = (dx >= 0 && dx < 3) ? w1[((co * 3 + dy) * 3 + dx) * C + ci] : __float2bfloat16(0.f);
    }
    for (int e = tid; e < BT_N2 * K2; e += nt) {  // cv2: row n = output pixel j x output channel co, k = (dy, hidden pixel of the 64-byte window, ci)
        const int n = e / K2, k = e % K2;
        const int j = n / C, co = n % C;
        const int dy = k / 64, st = (k / 16) % 4, ci = (k % 16) % CH, hpx = (st - 1) * PX2 + (k % 16) / CH;
        const int dx = hpx - j + 1;
        o2[e] = (dx >= 0 && dx < 3) ? w2[((co * 3 + dy) * 3 + dx) * CH + ci] : __float2bfloat16(0.f);
    }
    for (int e = tid; e < BT_N1 + BT_N2; e += nt) ob[e] = e < BT_N1 ? b1[e % CH] : b2[(e - BT_N1) % C];
}

static size_t bt_packed_bytes(int c) {
    const int g = 64 / c, k1 = 3 * (g + 2) * c;
    return (size_t)BT_N1 * k1 * 2 + (size_t)BT_N2 * 192 * 2 + (BT_N1 + BT_N2) * 4;
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*BtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static BtEncodeFn bt_get_encode() {
    static BtEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<BtEncodeFn>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool bt_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes) {
    BtEncodeFn enc = bt_get_encode();
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
        set_error("dbl_bottleneck3x3_tc: cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] swizzle %d", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, sw_bytes);
        return false;
    }
    return true;
}

static uint32_t bt_round1k(size_t v) { return (uint32_t)((v + 1023) / 1024 * 1024); }

template <int C>
static int launch_bt(const dbl_tensor* x, const uint8_t* packed, int shortcut, const dbl_tensor* y, cudaStream_t st) {
    constexpr int G = 64 / C, KC = C / 16, S1 = (G + 2) * KC, K1 = 3 * S1 * 16;
    auto kern = bottleneck_tc_kernel<C>;
    static DevOnce once;
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM_LIMIT);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("dbl_bottleneck3x3_tc: cannot opt in to %d bytes of shared memory: %s", BT_SMEM_LIMIT, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    const int sms = dev_info().sms;
    const uint32_t w1_bytes = 3 * S1 * BT_N1 * 32, w2_bytes = 12 * BT_N2 * 32, hid_bytes = 2 * (256 * 64 + 1024);
    // ---- tile: GW pixel groups per halo row; hidden rows (TH + 2) * GW <= 256 = the two M tiles; fewest rounds of the grid, then fewest tiles
    BtParams p;
    memset(&p, 0, sizeof(p));
    long long best = -1;
    size_t best_smem = 0;
    for (int gw = 3; gw <= 64; ++gw) {
        const int tw = (gw - 2) * G, th = 256 / gw - 2;
        if (th < 1 || tw < G) continue;
        const uint32_t in_b = bt_round1k((size_t)(th + 4) * gw * 128), out_b = bt_round1k((size_t)th * tw * 2 * C);
        const size_t smem = 1024 + 1024 + (size_t)BT_NIN * in_b + hid_bytes + 2 * (size_t)out_b + w1_bytes + w2_bytes + 1024;
        if (smem > (size_t)BT_SMEM_LIMIT) continue;
        const long long tiles = (long long)x->n * cdiv(x->w, tw) * cdiv(x->h, th);
        const long long cost = ((tiles + sms - 1) / sms) * (1ll << 32) + tiles;
        if (best < 0 || cost < best) {
            best = cost, best_smem = smem;
            p.gw = gw, p.tw = tw, p.th = th;
            p.in_bytes = in_b, p.out_bytes = out_b;
        }
    }
    DBL_REQUIRE(best >= 0, DBL_E_UNSUPPORTED, "dbl_bottleneck3x3_tc: no tile fits");
    p.h = x->h, p.w = x->w, p.shortcut = shortcut, p.trace = debug_trace_buffer();
    p.tiles_x = cdiv(x->w, p.tw), p.tiles_y = cdiv(x->h, p.th);
    const long long nt = (long long)x->n * p.tiles_x * p.tiles_y;
    DBL_REQUIRE(nt < (1ll << 31), DBL_E_UNSUPPORTED, "dbl_bottleneck3x3_tc: too many tiles");
    p.ntiles = (int)nt;
    p.inv_gw = 1.0f / (float)p.gw;
    p.in_tx = (uint32_t)((p.th + 4) * p.gw * 128);
    p.off_hid = 1024 + BT_NIN * p.in_bytes + 1024;  // hidden tile 0 behind its own zeroed guard (tile 1 follows at + 256*64 + 1024)
    p.off_out = p.off_hid + hid_bytes - 1024;
    p.off_w1 = p.off_out + 2 * p.out_bytes;
    p.off_w2 = p.off_w1 + w1_bytes;
    p.off_misc = p.off_w2 + w2_bytes;
    const __nv_bfloat16* wb1 = reinterpret_cast<const __nv_bfloat16*>(packed);
    const __nv_bfloat16* wb2 = wb1 + BT_N1 * K1;
    const float* bias = reinterpret_cast<const float*>(wb2 + BT_N2 * 192);
    p.bias1 = bias, p.bias2 = bias + BT_N1;
    BtMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        const uint64_t pitchB = (uint64_t)x->pitch * 2;  // == 2*C: the pixel-group view needs a dense tensor
        const uint64_t dims[4] = {64, (uint64_t)x->w / G, (uint64_t)x->h, (uint64_t)x->n};
        const uint64_t strides[3] = {128, pitchB * x->w, pitchB * x->w * x->h};
        const uint32_t box[4] = {64, (uint32_t)p.gw, (uint32_t)(p.th + 4), 1};
        if (!bt_encode(&maps.x, x->data, 4, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K1, BT_N1};
        const uint64_t strides[1] = {(uint64_t)K1 * 2};
        const uint32_t box[2] = {16, BT_N1};
        if (!bt_encode(&maps.w1, wb1, 2, dims, strides, box, 32)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {192, BT_N2};
        const uint64_t strides[1] = {192 * 2};
        const uint32_t box[2] = {16, BT_N2};
        if (!bt_encode(&maps.w2, wb2, 2, dims, strides, box, 32)) return DBL_E_CUDA;
    }
    {
        const uint64_t xp = (uint64_t)x->pitch * 2;
        const uint64_t dims[4] = {64, (uint64_t)x->w / G, (uint64_t)x->h, (uint64_t)x->n};
        const uint64_t strides[3] = {128, xp * x->w, xp * x->w * x->h};
        const uint32_t box[4] = {64, (uint32_t)(p.gw - 2), (uint32_t)p.th, 1};
        if (!bt_encode(&maps.r, x->data, 4, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    {
        const uint64_t yp = (uint64_t)y->pitch * 2;  // == 2*C (dense)
        const uint64_t dims[4] = {64, (uint64_t)y->w / G, (uint64_t)y->h, (uint64_t)y->n};
        const uint64_t strides[3] = {128, yp * y->w, yp * y->w * y->h};
        const uint32_t box[4] = {64, (uint32_t)(p.gw - 2), (uint32_t)p.th, 1};
        if (!bt_encode(&maps.y, y->data, 4, dims, strides, box, 128)) return DBL_E_CUDA;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch_k(kern, grid, BT_THREADS, best_smem, st, maps, p);
    count_launch();
    return check_launch("dbl_bottleneck3x3_tc");
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_bottleneck_tc_supported(int c, int w, int pitch) { return (c == 16 || c == 32) && pitch == c && w % (64 / c) == 0 && w >= 64 / c; }

extern "C" size_t dbl_bottleneck_tc_packed_bytes(int c) { return (c == 16 || c == 32) ? bt_packed_bytes(c) : 0; }

extern "C" int dbl_bottleneck_tc_pack(const void* w1, const float* b1, const void* w2, const float* b2, int c, void* packed, dbl_stream_t stream) {
    DBL_REQUIRE(w1 && b1 && w2 && b2 && packed, DBL_E_ARG, "dbl_bottleneck_tc_pack: null argument");
    DBL_REQUIRE(c == 16 || c == 32, DBL_E_UNSUPPORTED, "dbl_bottleneck_tc_pack: C must be 16 or 32");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, DBL_E_ARG, "dbl_bottleneck_tc_pack: packed must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    auto* a = reinterpret_cast<const __nv_bfloat16*>(w1);
    auto* b = reinterpret_cast<const __nv_bfloat16*>(w2);
    if (c == 16) bottleneck_tc_pack_kernel<16><<<32, 256, 0, st>>>(a, b1, b, b2, reinterpret_cast<uint8_t*>(packed));
    else bottleneck_tc_pack_kernel<32><<<32, 256, 0, st>>>(a, b1, b, b2, reinterpret_cast<uint8_t*>(packed));
    return check_launch("dbl_bottleneck_tc_pack");
}

extern "C" int dbl_bottleneck3x3_tc(const dbl_tensor* x, const void* packed, int shortcut, const dbl_tensor* y, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && packed, DBL_E_ARG, "dbl_bottleneck3x3_tc: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y) && same_shape(x, y), DBL_E_ARG, "dbl_bottleneck3x3_tc: x and y must be bf16 NHWC views of one shape, 16-byte aligned");
    DBL_REQUIRE(dbl_bottleneck_tc_supported(x->c, x->w, x->pitch) && y->pitch == y->c, DBL_E_UNSUPPORTED,
                "dbl_bottleneck3x3_tc: needs C in {16, 32}, dense tensors (pitch == C) and W %% (64 / C) == 0");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, DBL_E_ARG, "dbl_bottleneck3x3_tc: packed weights must be 16-byte aligned");
    DBL_REQUIRE(x->data != y->data, DBL_E_ARG, "dbl_bottleneck3x3_tc: x and y must not alias (tiles read their neighbours' halo)");
    DBL_REQUIRE(x->n <= 65535, DBL_E_UNSUPPORTED, "dbl_bottleneck3x3_tc: batch too large");
    cudaStream_t st = as_stream(stream);
    return x->c == 16 ? launch_bt<16>(x, reinterpret_cast<const uint8_t*>(packed), shortcut, y, st)
                      : launch_bt<32>(x, reinterpret_cast<const uint8_t*>(packed), shortcut, y, st);
}
