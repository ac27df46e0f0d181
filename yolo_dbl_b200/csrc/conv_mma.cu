// Thin, high-resolution 3x3 convolutions on warp-level tensor-core MMA (mma.sync m16n8k16 bf16, fp32 accumulate).
//
// Layers 0-6 of YOLO-DBL run at 640^2..80^2 with 3..64 channels: their arithmetic intensity is far below the B200 ridge,
// so they are HBM-bound and what matters is (a) touching every activation once and (b) spending few instructions per
// pixel.  tcgen05's 128-row, single-issuer, TMEM-epilogue machinery costs thousands of instructions per 128-pixel tile
// on such layers (ncu: issue-bound at 19% of HBM peak), so these layers get their own kernels:
//
//   * bottleneck_mma_kernel : Bottleneck.forward (block.py:344-357) = x + cv2(cv1(x)) in ONE kernel.  The input tile
//     (+2 pixel halo) is fetched once with cp.async into a planar shared-memory layout, cv1's output (+1 pixel halo,
//     recomputed) stays in shared memory as bf16, cv2 reads it from there, the residual comes from the input tile.
//     HBM traffic = read x once + write y once.
//   * stem_pair_kernel      : layer 0 (3 -> C0, stride 1, on the raw f32/uint8 NCHW image) + layer 1 (C0 -> C1, stride 2)
//     fused; the C0 x 640 x 640 intermediate never leaves shared memory.
//   * conv3x3_mma_kernel    : single 3x3 conv (stride 1 or 2) for thin layers (the stride-2 down-samplers).
//
// Shared-memory operand layout ("planar"): plane g holds channels [8g, 8g+8) of every tile pixel, 16 bytes per pixel,
// pixels linearised with the INPUT tile's row pitch.  With that pitch shared by producer and consumer tiles, the A
// operand of filter tap (ky,kx) for 16 consecutive GEMM rows is the same buffer advanced by (ky*pitch + kx) pixels: one
// ldmatrix.x4 per (16 pixels x 16 k) fragment, conflict-free, no im2col.  GEMM rows that fall on the 2 (or 4) halo
// columns of each tile row are junk and never stored.  8-channel inputs pair two taps per k-step.  Stride-2 convs keep
// four parity sub-tiles (row parity x column parity) so taps are unit-stride again.
// Weights are pre-packed once (dbl_mma_pack_weights) in B-fragment order: one 8-byte load per lane per (k-step, n-block).
// Output channels are permuted inside the packed weights so that every thread owns 2*NB CONTIGUOUS channels of its
// pixels -> 8/16-byte global stores without any shuffle.
#include <cstdlib>

#include "mma_util.cuh"

namespace dbl {

__global__ void mma_pack_kernel(const __nv_bfloat16* __restrict__ w, int cout, int cin, int taps, uint2* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int nb = cout / 8, ks_n = mma_ksteps(cin, taps);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ks_n * nb * 32) return;
    const int lane = i & 31, j = (i >> 5) % nb, ks = (i >> 5) / nb;
    const int ch = mma_chan(nb, j, lane >> 2), t = lane & 3;
    uint32_t v[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {  // k = 2t, 2t+1 (half 0) and 2t+8, 2t+9 (half 1) of this k-step
        int tap, c;
        if (cin == 8) {
            tap = 2 * ks + half, c = 2 * t;
        } else {
            const int cp = cin / 16;
            tap = ks / cp, c = (ks % cp) * 16 + half * 8 + 2 * t;
        }
        uint32_t pair = 0;
        if (tap < taps) pair = *reinterpret_cast<const uint32_t*>(w + ((size_t)ch * taps + tap) * cin + c);
        v[half] = pair;
    }
    out[i] = make_uint2(v[0], v[1]);
}

// =====================================================================================================================
// fused Bottleneck: y = x + SiLU(cv2(SiLU(cv1(x))))          x, y: bf16 NHWC with C channels; hidden = C/2 channels
// =====================================================================================================================
struct BnParams {
    const __nv_bfloat16* x;
    __nv_bfloat16* y;
    const uint2* w1;  // packed [ks][nb][lane]
    const uint2* w2;
    const float* b1;
    const float* b2;
    int n, h, w, xpitch, ypitch, add;
    int tiles_x, tiles_y, ntiles;
    fastdiv per_img, tx_div;  // tiles per image, tiles per row
};

template <int C, int TH, int TW>
struct BnCfg {
    static constexpr int H = C / 2;
    static constexpr int WI = TW + 4;                       // row pitch of every tile (pixels)
    static constexpr int M1 = round16((TH + 2) * WI) / 16;  // cv1 m-blocks (hidden tile incl. halo, pitch WI)
    static constexpr int M2 = round16(TH * WI) / 16;        // cv2 m-blocks
    static constexpr int IN_NEED = M1 * 16 + 2 * WI + 2;    // pixels a cv1 fragment load may touch
    static constexpr int GI = C / 8, GH = H / 8;
    static constexpr int PSI = plane_px(IN_NEED > (TH + 4) * WI ? IN_NEED : (TH + 4) * WI, GI) * 16;  // bytes
    static constexpr int HID_NEED = (M1 * 16 > M2 * 16 + 2 * WI + 2) ? M1 * 16 : M2 * 16 + 2 * WI + 2;
    static constexpr int PSH = plane_px(HID_NEED, GH) * 16;
    static constexpr int NB1 = H / 8, NB2 = C / 8;
    static constexpr int KS1 = mma_ksteps(C, 9), KS2 = mma_ksteps(H, 9);
    static constexpr int W1B = KS1 * NB1 * 256, W2B = KS2 * NB2 * 256;
    static constexpr int IN_BYTES = GI * PSI, HID_BYTES = GH * PSH;
    static constexpr int SMEM = 2 * IN_BYTES + HID_BYTES + W1B + W2B + (H + C) * 4 + 128;
    static_assert(SMEM <= 227 * 1024, "bottleneck tile does not fit shared memory");
};

template <int C, int TH, int TW, int NWARPS, int MB>
__global__ void __launch_bounds__(NWARPS * 32) bottleneck_mma_kernel(const BnParams p) {
    using K = BnCfg<C, TH, TW>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sIn = smem_u32(smem);
    const uint32_t sHid = sIn + 2 * K::IN_BYTES;
    const uint32_t sW1 = sHid + K::HID_BYTES;
    const uint32_t sW2 = sW1 + K::W1B;
    float* sB1 = reinterpret_cast<float*>(smem + 2 * K::IN_BYTES + K::HID_BYTES + K::W1B + K::W2B);
    float* sB2 = sB1 + K::H;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = NWARPS * 32;
    pdl_launch_dependents();

    auto prefetch = [&](int tile, int buf) {
        // tile -> (image, tile row, tile column) by multiply-high (four integer divisions per tile were ~100 of the ~630
        // instructions a warp spends on a tile of the thin stride-2 conv)
        const int img = (int)fdiv((uint32_t)tile, p.per_img), r = tile - img * (int)p.per_img.d;
        const int tyi = (int)fdiv((uint32_t)r, p.tx_div);
        const int oy0 = tyi * TH, ox0 = (r - tyi * p.tiles_x) * TW;
        const uint32_t dst0 = sIn + buf * K::IN_BYTES;
        // a thread keeps its (plane, column) and walks down the rows: per 16-byte item one bounds test, one pointer add and
        // the cp.async (the flat item loop spent ~125 instructions per item on div/mod and 64-bit address arithmetic)
        constexpr int ROW_ITEMS = K::WI * K::GI;
        constexpr int COL_PASSES = (ROW_ITEMS + NT - 1) / NT;
        constexpr int ROWS_PER_PASS = NT >= 2 * ROW_ITEMS ? NT / ROW_ITEMS : 1;
#pragma unroll
        for (int cp = 0; cp < COL_PASSES; ++cp) {
            const int e = ROWS_PER_PASS > 1 ? tid % ROW_ITEMS : tid + cp * NT;
            const int ly0 = ROWS_PER_PASS > 1 ? tid / ROW_ITEMS : 0;
            if (e >= ROW_ITEMS || ly0 >= ROWS_PER_PASS) continue;
            const int gi = e % K::GI, lx = e / K::GI;
            const int gx = ox0 - 2 + lx;
            const bool colok = gx >= 0 && gx < p.w;
            const __nv_bfloat16* src = p.x + (((size_t)img * p.h + (oy0 - 2 + ly0)) * p.w + gx) * p.xpitch + gi * 8;
            const size_t rstep = (size_t)ROWS_PER_PASS * p.w * p.xpitch;
            uint32_t dst = dst0 + gi * K::PSI + (ly0 * K::WI + lx) * 16;
            for (int ly = ly0; ly < TH + 4; ly += ROWS_PER_PASS) {
                const bool ok = colok && (unsigned)(oy0 - 2 + ly) < (unsigned)p.h;
                cp_async16(dst, ok ? src : p.x, ok);
                src += rstep, dst += ROWS_PER_PASS * K::WI * 16;
            }
        }
    };

    // weights + biases: once per CTA
    for (int i = tid; i < K::W1B / 16; i += NT) cp_async16(sW1 + i * 16, reinterpret_cast<const uint8_t*>(p.w1) + i * 16, true);
    for (int i = tid; i < K::W2B / 16; i += NT) cp_async16(sW2 + i * 16, reinterpret_cast<const uint8_t*>(p.w2) + i * 16, true);
    for (int i = tid; i < K::H; i += NT) sB1[i] = p.b1 ? p.b1[i] : 0.f;
    for (int i = tid; i < C; i += NT) sB2[i] = p.b2 ? p.b2[i] : 0.f;
    pdl_wait();  // weights / biases are constant; x is the previous kernel's output
    int tile = blockIdx.x, buf = 0;
    if (tile < p.ntiles) prefetch(tile, 0);
    cp_async_commit();

    // lane constants of the ldmatrix row addresses
    const uint32_t row16 = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 16);
    const bool hi = lane >= 16;
    const uint32_t a1_lane = row16 + (C > 8 && hi ? K::PSI : 0);
    const uint32_t a2_lane = row16 + (K::H > 8 && hi ? K::PSH : 0);

    for (; tile < p.ntiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (next < p.ntiles) prefetch(next, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        // tile -> (image, tile row, tile column) by multiply-high (four integer divisions per tile were ~100 of the ~630
        // instructions a warp spends on a tile of the thin stride-2 conv)
        const int img = (int)fdiv((uint32_t)tile, p.per_img), r = tile - img * (int)p.per_img.d;
        const int tyi = (int)fdiv((uint32_t)r, p.tx_div);
        const int oy0 = tyi * TH, ox0 = (r - tyi * p.tiles_x) * TW;
        const uint32_t in = sIn + buf * K::IN_BYTES;

        // ---------------- cv1: hidden tile (TH+2 rows, pitch WI) ----------------
        for (int m0 = warp * MB; m0 < K::M1; m0 += NWARPS * MB) {
            float acc[MB][K::NB1][4];
#pragma unroll
            for (int j = 0; j < K::NB1; ++j) {
                const float2 bb = lds_f2(sB1 + t * 2 * K::NB1 + 2 * j);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) acc[mb][j][0] = acc[mb][j][2] = bb.x, acc[mb][j][1] = acc[mb][j][3] = bb.y;
            }
            conv_phase<C, K::NB1, MB, K::PSI, OffS1<K::WI>>(acc, in + a1_lane + m0 * 256, sW1 + lane * 8, hi);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                if (m0 + mb >= K::M1) break;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int q = (m0 + mb) * 16 + g + 8 * rh;
                    const int hy = q / K::WI, hx = q - hy * K::WI;
                    const bool inside = (unsigned)(oy0 - 1 + hy) < (unsigned)p.h && (unsigned)(ox0 - 1 + hx) < (unsigned)p.w;  // cv2 zero-pads cv1's OUTPUT
                    uint32_t v[K::NB1];
#pragma unroll
                    for (int j = 0; j < K::NB1; ++j)
                        v[j] = inside ? pack2(silu_fast(acc[mb][j][2 * rh]), silu_fast(acc[mb][j][2 * rh + 1])) : 0u;
                    planar_store<K::NB1>(sHid, K::PSH, q, t, v);
                }
            }
        }
        __syncthreads();

        // ---------------- cv2 + residual -> global ----------------
        // 64-bit address of the tile origin once per tile; per pixel a 32-bit offset (the tensor is < 2^31 elements per tile row span)
        __nv_bfloat16* const ytile = p.y + (((size_t)img * p.h + oy0) * p.w + ox0) * p.ypitch + t * 2 * K::NB2;
        const int rows_ok = p.h - oy0, cols_ok = min(TW, p.w - ox0);
        for (int m0 = warp * MB; m0 < K::M2; m0 += NWARPS * MB) {
            float acc[MB][K::NB2][4];
#pragma unroll
            for (int j = 0; j < K::NB2; ++j) {
                const float2 bb = lds_f2(sB2 + t * 2 * K::NB2 + 2 * j);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) acc[mb][j][0] = acc[mb][j][2] = bb.x, acc[mb][j][1] = acc[mb][j][3] = bb.y;
            }
            conv_phase<K::H, K::NB2, MB, K::PSH, OffS1<K::WI>>(acc, sHid + a2_lane + m0 * 256, sW2 + lane * 8, hi);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                if (m0 + mb >= K::M2) break;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int o = (m0 + mb) * 16 + g + 8 * rh;
                    const int oy = o / K::WI, ox = o - oy * K::WI;
                    if (ox >= cols_ok || oy >= rows_ok) continue;
                    uint32_t res[K::NB2];
                    if (p.add) planar_load<K::NB2>(in, K::PSI, o + 2 * K::WI + 2, t, res);
                    uint32_t v[K::NB2];
#pragma unroll
                    for (int j = 0; j < K::NB2; ++j) {
                        float a = silu_fast(acc[mb][j][2 * rh]), b = silu_fast(acc[mb][j][2 * rh + 1]);
                        if (p.add) {
                            float r0, r1;
                            unpack2(res[j], r0, r1);
                            a += r0, b += r1;
                        }
                        v[j] = pack2(a, b);
                    }
                    global_store<K::NB2>(ytile + (uint32_t)((oy * p.w + ox) * p.ypitch), v);
                }
            }
        }
        __syncthreads();  // the next iteration's prefetch overwrites this buffer's twin; cv1 overwrites the hidden tile
    }
    cp_async_wait<0>();
}

// =====================================================================================================================
// single thin 3x3 conv, stride 1 or 2, SiLU / none, optional residual (add / mul) read from global
// =====================================================================================================================
struct CmParams {
    const __nv_bfloat16* x;
    __nv_bfloat16* y;
    const __nv_bfloat16* r;
    const uint2* w;
    const float* b;
    int n, h, wd, ho, wo, xpitch, ypitch, rpitch, act, epi;
    int tiles_x, tiles_y, ntiles;
    fastdiv per_img, tx_div;  // tiles per image, tiles per row
};

template <int CIN, int COUT, int S, int TH, int TW>
struct CmCfg {
    static constexpr int E = S == 1 ? 2 : 1;           // extra rows / columns per (sub-)tile
    static constexpr int WP = TW + E;                   // row pitch of a (sub-)tile
    static constexpr int M = round16(TH * WP) / 16;
    static constexpr int NPAR = S == 1 ? 1 : 4;
    static constexpr int SUB_NEED = M * 16 + E * WP + E;  // pixels per (sub-)tile a fragment load may touch
    static constexpr int SUB_PX = SUB_NEED > (TH + E) * WP ? SUB_NEED : (TH + E) * WP;
    static constexpr int GI = CIN / 8;
    static constexpr int PARB = SUB_PX * 16;                       // bytes of one parity sub-tile within a plane
    static constexpr int PS = plane_px(NPAR * SUB_PX, GI) * 16;    // plane stride (bytes)
    static constexpr int NB = COUT / 8;
    static constexpr int KS = mma_ksteps(CIN, 9);
    static constexpr int WB = KS * NB * 256;
    static constexpr int IN_BYTES = GI * PS;
    static constexpr int SMEM = 2 * IN_BYTES + WB + COUT * 4 + 128;
    static_assert(SMEM <= 227 * 1024, "conv tile does not fit shared memory");
};

template <int CIN, int COUT, int S, int TH, int TW, int NWARPS, int MB>
__global__ void __launch_bounds__(NWARPS * 32) conv3x3_mma_kernel(const CmParams p) {
    using K = CmCfg<CIN, COUT, S, TH, TW>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sIn = smem_u32(smem);
    const uint32_t sW = sIn + 2 * K::IN_BYTES;
    float* sB = reinterpret_cast<float*>(smem + 2 * K::IN_BYTES + K::WB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = NWARPS * 32;
    pdl_launch_dependents();

    auto prefetch = [&](int tile, int buf) {
        // tile -> (image, tile row, tile column) by multiply-high (four integer divisions per tile were ~100 of the ~630
        // instructions a warp spends on a tile of the thin stride-2 conv)
        const int img = (int)fdiv((uint32_t)tile, p.per_img), r = tile - img * (int)p.per_img.d;
        const int tyi = (int)fdiv((uint32_t)r, p.tx_div);
        const int oy0 = tyi * TH, ox0 = (r - tyi * p.tiles_x) * TW;
        const uint32_t dst0 = sIn + buf * K::IN_BYTES;
        constexpr int ROWS = S * TH + (S == 1 ? 2 : 1), COLS = S * TW + (S == 1 ? 2 : 1);  // input footprint of the tile
        // (plane, column) fixed per thread, rows walked with a pointer increment (see bottleneck_mma_kernel)
        constexpr int ROW_ITEMS = COLS * K::GI;
        constexpr int COL_PASSES = (ROW_ITEMS + NT - 1) / NT;
        constexpr int ROWS_PER_PASS = NT >= 2 * ROW_ITEMS ? NT / ROW_ITEMS : 1;
#pragma unroll
        for (int cp = 0; cp < COL_PASSES; ++cp) {
            const int e = ROWS_PER_PASS > 1 ? tid % ROW_ITEMS : tid + cp * NT;
            const int ly0 = ROWS_PER_PASS > 1 ? tid / ROW_ITEMS : 0;
            if (e >= ROW_ITEMS || ly0 >= ROWS_PER_PASS) continue;
            const int gi = e % K::GI, lx = e / K::GI;
            const int gx = S * ox0 - 1 + lx;
            const bool colok = gx >= 0 && gx < p.wd;
            const __nv_bfloat16* src = p.x + (((size_t)img * p.h + (S * oy0 - 1 + ly0)) * p.wd + gx) * p.xpitch + gi * 8;
            const size_t rstep = (size_t)ROWS_PER_PASS * p.wd * p.xpitch;
            for (int ly = ly0; ly < ROWS; ly += ROWS_PER_PASS) {
                const bool ok = colok && (unsigned)(S * oy0 - 1 + ly) < (unsigned)p.h;
                uint32_t d;
                if (S == 1) d = (uint32_t)((ly * K::WP + lx) * 16);
                else d = (uint32_t)(((ly & 1) * 2 + (lx & 1)) * K::PARB + ((ly >> 1) * K::WP + (lx >> 1)) * 16);
                cp_async16(dst0 + gi * K::PS + d, ok ? src : p.x, ok);
                src += rstep;
            }
        }
    };

    for (int i = tid; i < K::WB / 16; i += NT) cp_async16(sW + i * 16, reinterpret_cast<const uint8_t*>(p.w) + i * 16, true);
    for (int i = tid; i < COUT; i += NT) sB[i] = p.b ? p.b[i] : 0.f;
    pdl_wait();
    int tile = blockIdx.x, buf = 0;
    if (tile < p.ntiles) prefetch(tile, 0);
    cp_async_commit();
    const uint32_t row16 = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 16);
    const bool hi = lane >= 16;
    const uint32_t a_lane = row16 + (CIN > 8 && hi ? K::PS : 0);

    for (; tile < p.ntiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (next < p.ntiles) prefetch(next, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        // tile -> (image, tile row, tile column) by multiply-high (four integer divisions per tile were ~100 of the ~630
        // instructions a warp spends on a tile of the thin stride-2 conv)
        const int img = (int)fdiv((uint32_t)tile, p.per_img), r = tile - img * (int)p.per_img.d;
        const int tyi = (int)fdiv((uint32_t)r, p.tx_div);
        const int oy0 = tyi * TH, ox0 = (r - tyi * p.tiles_x) * TW;
        const uint32_t in = sIn + buf * K::IN_BYTES;
        const size_t tile0 = ((size_t)img * p.ho + oy0) * p.wo + ox0;  // 64-bit once per tile, 32-bit offsets per pixel
        __nv_bfloat16* const ytile = p.y + tile0 * p.ypitch + t * 2 * K::NB;
        const __nv_bfloat16* const rtile = p.r ? p.r + tile0 * p.rpitch + t * 2 * K::NB : nullptr;
        const int rows_ok = min(TH, p.ho - oy0), cols_ok = min(TW, p.wo - ox0);
        for (int m0 = warp * MB; m0 < K::M; m0 += NWARPS * MB) {
            float acc[MB][K::NB][4];
#pragma unroll
            for (int j = 0; j < K::NB; ++j) {
                const float2 bb = lds_f2(sB + t * 2 * K::NB + 2 * j);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) acc[mb][j][0] = acc[mb][j][2] = bb.x, acc[mb][j][1] = acc[mb][j][3] = bb.y;
            }
            if (S == 1) conv_phase<CIN, K::NB, MB, K::PS, OffS1<K::WP>>(acc, in + a_lane + m0 * 256, sW + lane * 8, hi);
            else conv_phase<CIN, K::NB, MB, K::PS, OffS2<K::WP, K::PARB>>(acc, in + a_lane + m0 * 256, sW + lane * 8, hi);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                if (m0 + mb >= K::M) break;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int o = (m0 + mb) * 16 + g + 8 * rh;
                    const int oy = o / K::WP, ox = o - oy * K::WP;
                    if (ox >= cols_ok || oy >= rows_ok) continue;
                    const uint32_t opix = (uint32_t)(oy * p.wo + ox);  // relative to the tile origin
                    uint32_t res[K::NB];
                    if (p.epi != DBL_EPI_NONE) global_load<K::NB>(rtile + opix * (uint32_t)p.rpitch, res);
                    uint32_t v[K::NB];
#pragma unroll
                    for (int j = 0; j < K::NB; ++j) {
                        float a = acc[mb][j][2 * rh], b = acc[mb][j][2 * rh + 1];
                        if (p.act == DBL_ACT_SILU) a = silu_fast(a), b = silu_fast(b);
                        if (p.epi != DBL_EPI_NONE) {
                            float r0, r1;
                            unpack2(res[j], r0, r1);
                            if (p.epi == DBL_EPI_ADD) a += r0, b += r1;
                            else a *= r0, b *= r1;
                        }
                        v[j] = pack2(a, b);
                    }
                    global_store<K::NB>(ytile + opix * (uint32_t)p.ypitch, v);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

// =====================================================================================================================
// fused stem: layer 0 (3 -> C0, 3x3 s1, SiLU) on the raw NCHW image + layer 1 (C0 -> C1, 3x3 s2, SiLU)
// =====================================================================================================================
struct StemParams {
    const void* x;      // f32 or u8 NCHW (n,3,h,w)
    const float* w0;    // f32 [C0][3][3][3]
    const float* b0;
    const uint2* w1;    // packed
    const float* b1;
    __nv_bfloat16* y;   // (n, h/2, w/2, C1)
    int n, h, w, ho, wo, ypitch;
    float in_scale;     // 1 (f32 input in 0..1) or 1/255 (u8), folded into layer 0's weights
    int tiles_x, tiles_y;
};

template <int C0, int C1, int TH, int TW>
struct StemCfg {
    static constexpr int R0 = 2 * TH + 1, W0 = 2 * TW + 1;  // layer-0 outputs layer 1 needs (rows, cols)
    static constexpr int RI = R0 + 2;                        // input rows
    static constexpr int PI = 2 * TW + 4;                    // input row pitch (pixels of 8 bytes: 4 x bf16, RGB0)
    static constexpr int M0 = round16(R0 * PI) / 16;         // layer-0 m-blocks (pitch PI -> tap shift = address offset)
    static constexpr int IN_PX = M0 * 16 + 2 * PI + 8;
    static constexpr int IN_BYTES = ((IN_PX > RI * PI ? IN_PX : RI * PI) * 8 + 127) / 128 * 128;
    static constexpr int WP = PI / 2;                        // layer-0 parity sub-tile pitch (junk columns keep their own slots)
    static constexpr int M1 = round16(TH * WP) / 16;
    static constexpr int SUB_NEED = M1 * 16 + WP + 2;
    static constexpr int SUB_PX = SUB_NEED > (TH + 1) * WP ? SUB_NEED : (TH + 1) * WP;
    static constexpr int PARB = SUB_PX * 16;
    static constexpr int G0 = C0 / 8;
    static constexpr int PS0 = plane_px(4 * SUB_PX, G0) * 16;
    static constexpr int L0_BYTES = G0 * PS0;
    static constexpr int NB0 = C0 / 8, NB1 = C1 / 8;
    static constexpr int KS1 = mma_ksteps(C0, 9);
    static constexpr int W0B = 3 * NB0 * 256, W1B = KS1 * NB1 * 256;
    static constexpr int SMEM = IN_BYTES + L0_BYTES + W0B + W1B + (C0 + C1) * 4 + 128;
    static_assert(SMEM <= 227 * 1024, "stem tile does not fit shared memory");
};

template <int C0, int C1, int TH, int TW, int NWARPS, typename TIN>
__global__ void __launch_bounds__(NWARPS * 32) stem_pair_kernel(const StemParams p) {
    using K = StemCfg<C0, C1, TH, TW>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sIn = smem_u32(smem);
    const uint32_t sL0 = sIn + K::IN_BYTES;
    const uint32_t sW0 = sL0 + K::L0_BYTES;
    const uint32_t sW1 = sW0 + K::W0B;
    float* sB0 = reinterpret_cast<float*>(smem + K::IN_BYTES + K::L0_BYTES + K::W0B + K::W1B);
    float* sB1 = sB0 + C0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = NWARPS * 32;

    const int img = blockIdx.z, oy0 = blockIdx.y * TH, ox0 = blockIdx.x * TW;
    pdl_launch_dependents();

    // ---- weights: layer 1 packed (async copy); layer 0 packed here from f32 [C0][ci][ky][kx]: k-step = ky, k = kx*4 + ci
    for (int i = tid; i < K::W1B / 16; i += NT) cp_async16(sW1 + i * 16, reinterpret_cast<const uint8_t*>(p.w1) + i * 16, true);
    cp_async_commit();
    for (int i = tid; i < 3 * K::NB0 * 32; i += NT) {
        const int ln = i & 31, j = (i >> 5) % K::NB0, ky = (i >> 5) / K::NB0;
        const int ch = mma_chan(K::NB0, j, ln >> 2), tt = ln & 3;
        uint32_t v[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float f[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * tt + 8 * half + e, kx = k >> 2, ci = k & 3;
                f[e] = (kx < 3 && ci < 3) ? p.w0[((ch * 3 + ci) * 3 + ky) * 3 + kx] * p.in_scale : 0.f;
            }
            v[half] = pack2(f[0], f[1]);
        }
        sts64(sW0 + i * 8, make_uint2(v[0], v[1]));
    }
    for (int i = tid; i < C0; i += NT) sB0[i] = p.b0 ? p.b0[i] : 0.f;
    for (int i = tid; i < C1; i += NT) sB1[i] = p.b1 ? p.b1[i] : 0.f;

    pdl_wait();  // the image batch may have been produced by an earlier kernel / copy of the stream
    // ---- stage the input tile: rows 2*oy0-2 .., cols 2*ox0-2 ..; pixel = {R,G,B,0} bf16 (8 bytes); zero outside the image
    {
        const TIN* xin = reinterpret_cast<const TIN*>(p.x) + (size_t)img * 3 * p.h * p.w;
        const size_t plane = (size_t)p.h * p.w;
        // a thread keeps a PAIR of columns and walks down the rows (2 * NT / PI rows per pass): no div/mod per pixel, one 8-byte
        // (f32) / 2-byte (u8) load per colour plane and one 16-byte shared store per two pixels.  Tile origins and the image width
        // are even, so a pair is either inside the image or outside it as a whole.
        static_assert(NT % (K::PI / 2) == 0, "threads must cover whole input rows");
        constexpr int RPP = NT / (K::PI / 2);
        const int lx = (tid % (K::PI / 2)) * 2, ly0 = tid / (K::PI / 2);
        const int gx = 2 * ox0 - 2 + lx;
        const bool colok = gx >= 0 && gx + 1 < p.w;
        const TIN* s = xin + (size_t)(2 * oy0 - 2 + ly0) * p.w + gx;
        uint32_t dst = sIn + (uint32_t)((ly0 * K::PI + lx) * 8);
        for (int ly = ly0; ly < K::RI; ly += RPP) {
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f;
            if (colok && (unsigned)(2 * oy0 - 2 + ly) < (unsigned)p.h) {
                if (sizeof(TIN) == 4) {
                    const float2 va = __ldg(reinterpret_cast<const float2*>(s)), vb = __ldg(reinterpret_cast<const float2*>(s + plane)),
                                 vc = __ldg(reinterpret_cast<const float2*>(s + 2 * plane));
                    a0 = va.x, a1 = va.y, b0 = vb.x, b1 = vb.y, c0 = vc.x, c1 = vc.y;
                } else {
                    const uchar2 va = __ldg(reinterpret_cast<const uchar2*>(s)), vb = __ldg(reinterpret_cast<const uchar2*>(s + plane)),
                                 vc = __ldg(reinterpret_cast<const uchar2*>(s + 2 * plane));
                    a0 = (float)va.x, a1 = (float)va.y, b0 = (float)vb.x, b1 = (float)vb.y, c0 = (float)vc.x, c1 = (float)vc.y;
                }
            }
            sts128(dst, make_uint4(pack2(a0, b0), pack2(c0, 0.f), pack2(a1, b1), pack2(c1, 0.f)));
            s += (size_t)RPP * p.w, dst += RPP * K::PI * 8;
        }
        // the few pixels past the last row that junk GEMM rows touch: keep them finite
        for (int i = K::RI * K::PI + tid; i < K::IN_BYTES / 8; i += NT) sts64(sIn + i * 8, make_uint2(0u, 0u));
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- layer 0: GEMM rows q = ly*PI + lx over the layer-0 tile; k-step ky reads 32 contiguous bytes of input row ly+ky
    {
        uint2 bw[3][K::NB0];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int j = 0; j < K::NB0; ++j) bw[ky][j] = lds64(sW0 + ((ky * K::NB0 + j) * 32 + lane) * 8);
        float2 bias[K::NB0];
#pragma unroll
        for (int j = 0; j < K::NB0; ++j) {
            bias[j] = lds_f2(sB0 + t * 2 * K::NB0 + 2 * j);
            bias[j].x *= 0.5f, bias[j].y *= 0.5f;  // SiLU(z + b) = h + h * tanh(h) with h = 0.5 z + 0.5 b: the bias rides on the 0.5 scaling
        }
        constexpr int MB = 2;
        for (int m0 = warp * MB; m0 < K::M0; m0 += NWARPS * MB) {
            float acc[MB][K::NB0][4];  // zero accumulators: the first HMMA takes RZ, no per-iteration register set-up
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int j = 0; j < K::NB0; ++j) acc[mb][j][0] = acc[mb][j][1] = acc[mb][j][2] = acc[mb][j][3] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    const uint32_t base = sIn + (uint32_t)((((m0 + mb) * 16 + g) + ky * K::PI) * 8 + t * 4);
                    uint32_t a[4];
                    a[0] = lds32(base), a[1] = lds32(base + 64), a[2] = lds32(base + 16), a[3] = lds32(base + 80);
#pragma unroll
                    for (int j = 0; j < K::NB0; ++j) mma16816(acc[mb][j], a, bw[ky][j].x, bw[ky][j].y);
                }
            }
            // The MB = 2 blocks of an iteration are 32 consecutive GEMM rows of ONE layer-0 row (PI = 64, m0 even): the row index,
            // its bounds test and the row / parity part of the parity-sub-tile address are computed once, the four (block, half)
            // stores differ by constants.  (ncu source view: this epilogue was ~70 of the loop's 113 instructions.)
            static_assert(K::PI == 64 && MB == 2, "layer-0 epilogue assumes 4 m-blocks per row and block pairs");
            const int ly = m0 >> 2, lx0 = (m0 & 3) * 16 + g;
            if (ly < K::R0) {
                const bool rowin = (unsigned)(2 * oy0 - 1 + ly) < (unsigned)p.h;
                const int sub0 = ((ly & 1) * 2 + (g & 1)) * K::SUB_PX + (ly >> 1) * K::WP + (lx0 >> 1);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    if (m0 + mb >= K::M0) break;
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const int lx = lx0 + mb * 16 + rh * 8;
                        if (lx >= K::W0) continue;
                        const bool inside = rowin && (unsigned)(2 * ox0 - 1 + lx) < (unsigned)p.w;  // layer 1 zero-pads layer 0's OUTPUT
                        uint32_t v[K::NB0];
#pragma unroll
                        for (int j = 0; j < K::NB0; ++j)
                            v[j] = inside ? pack2(silu_half_bias(acc[mb][j][2 * rh], bias[j].x), silu_half_bias(acc[mb][j][2 * rh + 1], bias[j].y)) : 0u;
                        planar_store<K::NB0>(sL0, K::PS0, sub0 + mb * 8 + rh * 4, t, v);
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- layer 1: stride 2 over the parity sub-tiles
    {
        const uint32_t row16 = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 16);
        const bool hi = lane >= 16;
        const uint32_t a_lane = row16 + (C0 > 8 && hi ? K::PS0 : 0);
        constexpr int MB = 1;
        __nv_bfloat16* const ytile = p.y + (((size_t)img * p.ho + oy0) * p.wo + ox0) * p.ypitch + t * 2 * K::NB1;
        const int rows_ok = min(TH, p.ho - oy0), cols_ok = min(TW, p.wo - ox0);
        for (int m0 = warp * MB; m0 < K::M1; m0 += NWARPS * MB) {
            float acc[MB][K::NB1][4];
#pragma unroll
            for (int j = 0; j < K::NB1; ++j) {
                const float2 bb = lds_f2(sB1 + t * 2 * K::NB1 + 2 * j);
                acc[0][j][0] = acc[0][j][2] = bb.x, acc[0][j][1] = acc[0][j][3] = bb.y;
            }
            conv_phase<C0, K::NB1, MB, K::PS0, OffS2<K::WP, K::PARB>>(acc, sL0 + a_lane + m0 * 256, sW1 + lane * 8, hi);
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
                const int o = m0 * 16 + g + 8 * rh;
                const int oy = o / K::WP, ox = o - oy * K::WP;
                if (ox >= cols_ok || oy >= rows_ok) continue;
                uint32_t v[K::NB1];
#pragma unroll
                for (int j = 0; j < K::NB1; ++j) v[j] = pack2(silu_fast(acc[0][j][2 * rh]), silu_fast(acc[0][j][2 * rh + 1]));
                global_store<K::NB1>(ytile + (uint32_t)((oy * p.wo + ox) * p.ypitch), v);
            }
        }
    }
}

template <class Kern>
static int launch_persistent(Kern kern, size_t smem, int threads, int ntiles, cudaStream_t st, const char* what, int* grid_out) {
    int occ = func_config(reinterpret_cast<const void*>(kern));  // resident CTAs per SM on the current device, -1 = not configured there yet
    if (occ < 1) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess || occ < 1) {
            cudaGetLastError();
            set_error("%s: cannot configure kernel (%zu bytes of shared memory): %s", what, smem, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        func_set_config(reinterpret_cast<const void*>(kern), occ);
    }
    const int cap = dev_info().sms * occ;
    *grid_out = ntiles < cap ? ntiles : cap;
    return DBL_OK;
}

template <int C, int TH, int TW, int NWARPS, int MB>
static int launch_bottleneck(const dbl_tensor* x, const void* w1p, const float* b1, const void* w2p, const float* b2, int add,
                             const dbl_tensor* y, cudaStream_t st) {
    using K = BnCfg<C, TH, TW>;
    BnParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data), p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.w1 = reinterpret_cast<const uint2*>(w1p), p.w2 = reinterpret_cast<const uint2*>(w2p), p.b1 = b1, p.b2 = b2;
    p.n = x->n, p.h = x->h, p.w = x->w, p.xpitch = x->pitch, p.ypitch = y->pitch, p.add = add;
    p.tiles_x = cdiv(x->w, TW), p.tiles_y = cdiv(x->h, TH), p.ntiles = x->n * p.tiles_x * p.tiles_y;
    p.per_img = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y), (uint64_t)p.ntiles), p.tx_div = make_fastdiv((uint32_t)p.tiles_x, (uint64_t)p.tiles_x * p.tiles_y);
    auto kern = bottleneck_mma_kernel<C, TH, TW, NWARPS, MB>;
    int grid = 0;
    const int rc = launch_persistent(kern, K::SMEM, NWARPS * 32, p.ntiles, st, "dbl_bottleneck3x3", &grid);
    if (rc) return rc;
    launch_k(kern, grid, NWARPS * 32, K::SMEM, st, p);
    count_launch();
    return check_launch("dbl_bottleneck3x3");
}

template <int CIN, int COUT, int S, int TH, int TW, int NWARPS, int MB>
static int launch_conv_mma(const dbl_tensor* x, const void* wp, const float* b, const dbl_tensor* r, const dbl_tensor* y, int act, int epi,
                           cudaStream_t st) {
    using K = CmCfg<CIN, COUT, S, TH, TW>;
    CmParams p;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data), p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.r = r ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr;
    p.w = reinterpret_cast<const uint2*>(wp), p.b = b;
    p.n = x->n, p.h = x->h, p.wd = x->w, p.ho = y->h, p.wo = y->w, p.xpitch = x->pitch, p.ypitch = y->pitch, p.rpitch = r ? r->pitch : 0;
    p.act = act, p.epi = r ? epi : DBL_EPI_NONE;
    p.tiles_x = cdiv(y->w, TW), p.tiles_y = cdiv(y->h, TH), p.ntiles = x->n * p.tiles_x * p.tiles_y;
    p.per_img = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y), (uint64_t)p.ntiles), p.tx_div = make_fastdiv((uint32_t)p.tiles_x, (uint64_t)p.tiles_x * p.tiles_y);
    auto kern = conv3x3_mma_kernel<CIN, COUT, S, TH, TW, NWARPS, MB>;
    int grid = 0;
    const int rc = launch_persistent(kern, K::SMEM, NWARPS * 32, p.ntiles, st, "dbl_conv3x3_mma", &grid);
    if (rc) return rc;
    launch_k(kern, grid, NWARPS * 32, K::SMEM, st, p);
    count_launch();
    return check_launch("dbl_conv3x3_mma");
}

template <int C0, int C1, typename TIN, int TH = 16>
static int launch_stem_pair(const void* x, int n, int h, int w, const float* w0, const float* b0, const void* w1p, const float* b1,
                            const dbl_tensor* y, float scale, cudaStream_t st) {
    constexpr int TW = 30, NWARPS = 8;  // input row pitch 64 pixels, layer-0 sub-tile pitch 32: index math = shifts
    using K = StemCfg<C0, C1, TH, TW>;
    StemParams p;
    p.x = x, p.w0 = w0, p.b0 = b0, p.w1 = reinterpret_cast<const uint2*>(w1p), p.b1 = b1;
    p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.n = n, p.h = h, p.w = w, p.ho = y->h, p.wo = y->w, p.ypitch = y->pitch, p.in_scale = scale;
    p.tiles_x = cdiv(y->w, TW), p.tiles_y = cdiv(y->h, TH);
    auto kern = stem_pair_kernel<C0, C1, TH, TW, NWARPS, TIN>;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("dbl_stem_pair: cannot opt in to %d bytes of shared memory: %s", (int)K::SMEM, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DBL_REQUIRE(n <= 65535 && p.tiles_y <= 65535, DBL_E_UNSUPPORTED, "dbl_stem_pair: grid too large");
    launch_k(kern, dim3(p.tiles_x, p.tiles_y, n), NWARPS * 32, K::SMEM, st, p);
    count_launch();
    return check_launch("dbl_stem_pair");
}

}  // namespace dbl

using namespace dbl;

extern "C" size_t dbl_mma_packed_bytes(int cout, int cin, int taps) {
    if (cout <= 0 || cout % 8 || (cin != 8 && (cin <= 0 || cin % 16)) || taps <= 0) return 0;
    return (size_t)mma_ksteps(cin, taps) * (cout / 8) * 256;
}

extern "C" int dbl_mma_pack_weights(const void* w, int cout, int cin, int taps, void* packed, dbl_stream_t stream) {
    DBL_REQUIRE(w && packed, DBL_E_ARG, "dbl_mma_pack_weights: null pointer");
    DBL_REQUIRE(dbl_mma_packed_bytes(cout, cin, taps) > 0, DBL_E_UNSUPPORTED, "dbl_mma_pack_weights: cout %d (need %%8), cin %d (need 8 or %%16)", cout, cin);
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 3) == 0 && (reinterpret_cast<uintptr_t>(packed) & 15) == 0, DBL_E_ARG,
                "dbl_mma_pack_weights: misaligned pointer");
    const int total = mma_ksteps(cin, taps) * (cout / 8) * 32;
    launch_k(mma_pack_kernel, cdiv(total, 128), 128, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(w), cout, cin, taps,
                                                                      reinterpret_cast<uint2*>(packed));
    count_launch();
    return check_launch("dbl_mma_pack_weights");
}

extern "C" int dbl_bottleneck3x3(const dbl_tensor* x, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                                 int shortcut, const dbl_tensor* y, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w1_packed && w2_packed, DBL_E_ARG, "dbl_bottleneck3x3: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y) && same_shape(x, y), DBL_E_ARG, "dbl_bottleneck3x3: x and y must be bf16 NHWC of equal shape, 16-byte aligned");
    DBL_REQUIRE(x->data != y->data, DBL_E_ARG, "dbl_bottleneck3x3: in-place operation is not supported (tiles read their neighbours' halo)");
    DBL_REQUIRE(((reinterpret_cast<uintptr_t>(w1_packed) | reinterpret_cast<uintptr_t>(w2_packed)) & 15) == 0, DBL_E_ARG,
                "dbl_bottleneck3x3: packed weights must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    switch (x->c) {
        // tile width 28 -> row pitch 32: every 16-row GEMM block is half a tile row (index math = shifts) and the
        // m-block counts of both phases divide evenly over the warps
        // measured (layer 2 / layer 4 of YOLO-DBL-n at batch 64): C=16: 22 rows (3 resident CTAs) 266 us, 30 rows 285 us, 14 rows 269 us,
        // 46 rows 341 us; C=32: 14 rows 182 us, 22 rows 211 us, 6 rows 206 us; 12 or 16 warps per CTA are slower than 8 for both
        case 16: return launch_bottleneck<16, 22, 28, 8, 2>(x, w1_packed, b1, w2_packed, b2, shortcut, y, st);
        case 32: return launch_bottleneck<32, 14, 28, 8, 2>(x, w1_packed, b1, w2_packed, b2, shortcut, y, st);
        case 64: return launch_bottleneck<64, 10, 28, 12, 2>(x, w1_packed, b1, w2_packed, b2, shortcut, y, st);
        default: break;
    }
    set_error("dbl_bottleneck3x3: %d channels not in {16,32,64}", x->c);
    return DBL_E_UNSUPPORTED;
}

extern "C" int dbl_conv3x3_mma(const dbl_tensor* x, const void* w_packed, const float* bias, const dbl_tensor* r, const dbl_tensor* y,
                               int stride, int act, int epi, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w_packed, DBL_E_ARG, "dbl_conv3x3_mma: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y), DBL_E_ARG, "dbl_conv3x3_mma: x and y must be bf16 NHWC, 16-byte aligned");
    DBL_REQUIRE(stride == 1 || stride == 2, DBL_E_ARG, "dbl_conv3x3_mma: stride %d", stride);
    DBL_REQUIRE(act == DBL_ACT_NONE || act == DBL_ACT_SILU, DBL_E_UNSUPPORTED, "dbl_conv3x3_mma: activation %d", act);
    const int ho = (x->h - 1) / stride + 1, wo = (x->w - 1) / stride + 1;
    DBL_REQUIRE(y->n == x->n && y->h == ho && y->w == wo, DBL_E_ARG, "dbl_conv3x3_mma: output shape mismatch");
    if (epi != DBL_EPI_NONE) DBL_REQUIRE(valid_tensor(r) && same_shape(r, y) && vec8_ok(r), DBL_E_ARG, "dbl_conv3x3_mma: residual mismatch");
    DBL_REQUIRE(x->data != y->data, DBL_E_ARG, "dbl_conv3x3_mma: in-place operation is not supported");
    cudaStream_t st = as_stream(stream);
    const dbl_tensor* rr = epi != DBL_EPI_NONE ? r : nullptr;
#define DBL_CM(CI, CO, S, TH) /* tile width 30 (stride 1) / 31 (stride 2) -> (sub-)tile row pitch 32 */ \
    if (x->c == CI && y->c == CO && stride == S) return launch_conv_mma<CI, CO, S, TH, 32 - (S == 1 ? 2 : 1), 8, 2>(x, w_packed, bias, rr, y, act, epi, st)
    DBL_CM(8, 16, 2, 8);
    DBL_CM(16, 32, 2, 8);
    DBL_CM(32, 64, 2, 8);
    DBL_CM(16, 8, 1, 16);
    DBL_CM(8, 16, 1, 16);
    DBL_CM(32, 16, 1, 16);
    DBL_CM(16, 32, 1, 16);
    DBL_CM(64, 32, 1, 8);
    DBL_CM(32, 64, 1, 8);
#undef DBL_CM
    set_error("dbl_conv3x3_mma: no instantiation for cin %d cout %d stride %d", x->c, y->c, stride);
    return DBL_E_UNSUPPORTED;
}

extern "C" int dbl_stem_pair(const void* x_nchw, int x_is_u8, int n, int h, int w, const float* w0, const float* b0, int c0,
                             const void* w1_packed, const float* b1, const dbl_tensor* y, dbl_stream_t stream) {
    DBL_REQUIRE(x_nchw && w0 && w1_packed && valid_tensor(y), DBL_E_ARG, "dbl_stem_pair: null/invalid argument");
    DBL_REQUIRE(h % 2 == 0 && w % 2 == 0 && y->n == n && y->h == h / 2 && y->w == w / 2, DBL_E_ARG, "dbl_stem_pair: output shape mismatch");
    DBL_REQUIRE(vec8_ok(y), DBL_E_ARG, "dbl_stem_pair: y must be bf16 NHWC, 16-byte aligned");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(x_nchw) & (x_is_u8 ? 1 : 7)) == 0, DBL_E_ARG, "dbl_stem_pair: x must be %d-byte aligned (pixel pairs are loaded at once)",
                x_is_u8 ? 2 : 8);
    cudaStream_t st = as_stream(stream);
    const float scale = x_is_u8 ? 1.0f / 255.0f : 1.0f;
    // tile height 16 measured best (layer 0+1 at batch 64: 8 rows 349 us, 16 rows 337 us, 32 rows 503 us, 4 rows 420 us)
#define DBL_SP(A, B)                                                                                             \
    if (c0 == A && y->c == B)                                                                                    \
        return x_is_u8 ? launch_stem_pair<A, B, uint8_t>(x_nchw, n, h, w, w0, b0, w1_packed, b1, y, scale, st) \
                       : launch_stem_pair<A, B, float>(x_nchw, n, h, w, w0, b0, w1_packed, b1, y, scale, st)
    DBL_SP(8, 16);
    DBL_SP(16, 32);
#undef DBL_SP
    set_error("dbl_stem_pair: channels (%d, %d) not in {(8,16), (16,32)}", c0, y->c);
    return DBL_E_UNSUPPORTED;
}
