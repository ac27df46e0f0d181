// Thin stride-2 3x3 convolution (the backbone's down-samplers: Conv [c, 3, 2], layers 3 / 5 of cfg/models/v13/yolov13_DBL.yaml,
// nn/modules/conv.py:39-63) as ONE persistent, warp-specialised kernel on the Blackwell data path:
//   TMA (two row-parity planes per tile) -> nine filter taps as descriptor offsets into those planes -> tcgen05.mma, accumulators in TMEM
//   -> bias / SiLU epilogue -> swizzled staging tile -> ONE TMA store.
//
// Why: these layers are HBM-bound by their bytes (16 -> 32 @160^2: 315 MB per 64-image step), but the warp-MMA kernel that ran them
// (conv_mma.cu) sits at 2.5 TB/s with 25 % of the warp slots: its CTAs stage, multiply and store one after the other.  Here the three
// phases of CONSECUTIVE tiles overlap and no thread touches an input byte.
//
// Operand layout.  A stride-2 tap reads every other pixel.  Instead of four parity planes (rows x columns) the image is viewed as
// PIXEL PAIRS: [H][W/2][2*Cin] -- a pair is one contiguous 4*Cin-byte row (needs a dense tensor: pitch == Cin).  Two TMA boxes per tile
// fetch the even-row and the odd-row plane of that view ([rows][WL pairs], 64- or 128-byte swizzled rows, zero fill outside the image =
// the convolution's padding).  With GEMM row m = r * WL + c (output row r, column c of the tile; column TW is junk and never stored)
//   filter row    0: odd plane row r      1: even plane row r      2: odd plane row r + 1
//   filter column 0: pair c, SECOND half  1: pair c + 1, first half  2: pair c + 1, second half
// so a tap is the same plane buffer with the descriptor start advanced by whole pairs (+ half a row), and the K = Cin slice of a tap
// is Cin/16 k-steps inside the swizzled row.  (The swizzle is a function of the absolute shared-memory address bits -- see
// conv_tc4.cu -- so arbitrary row offsets need no base-offset field.)
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 4.. epilogue: four warps (one per TMEM lane group) per 128-row
// M tile, MT M tiles per CTA tile drained at once.  Accumulators (MT x Cout columns) are double-buffered in TMEM.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int S2_SMEM_LIMIT = 227 * 1024;
__host__ __device__ constexpr int s2_round1k(int v) { return (v + 1023) / 1024 * 1024; }

template <int CIN, int COUT, int MT, int WLMAX>
struct S2Cfg {
    static constexpr int ROWA = 4 * CIN;                      // bytes of one pixel pair = one GEMM row of a plane
    static constexpr int ROWB = 2 * CIN;                      // bytes of one weight row of a tap
    static constexpr int ROWO = 2 * COUT;                     // bytes of one output pixel
    static constexpr int PLANE_PIX = MT * 128 + WLMAX + 8;    // >= (TH + 1) * WL and >= the junk rows the last M tile reads
    static constexpr int PLANE_BYTES = s2_round1k(PLANE_PIX * ROWA);
    static constexpr int STAGE_BYTES = 2 * PLANE_BYTES;       // [even rows | odd rows]
    static constexpr int B_TAP = s2_round1k(COUT * ROWB);
    static constexpr int B_BYTES = 9 * B_TAP;
    static constexpr int O_BYTES = s2_round1k(MT * 128 * ROWO);
    static constexpr int MISC = 1024;                         // bias + barriers
    static constexpr int total(int nin, int nout) { return 1024 + nin * STAGE_BYTES + B_BYTES + nout * O_BYTES + MISC; }
    static constexpr int NIN = 2;
    static constexpr int NOUT = total(NIN, 2) <= S2_SMEM_LIMIT ? 2 : 1;
    static constexpr int SMEM = total(NIN, NOUT);
    static constexpr int EPI_WARPS = 4 * MT;
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;      // warp 0 TMA, warp 1 MMA, warps 2-3 idle (epilogue warp w must own TMEM lanes 32*(w%4)..)
    static constexpr int TMEM_COLS = 2 * MT * COUT <= 32 ? 32 : (2 * MT * COUT <= 64 ? 64 : (2 * MT * COUT <= 128 ? 128 : (2 * MT * COUT <= 256 ? 256 : 512)));
    static_assert(ROWA == 64 || ROWA == 128, "a pixel pair must be one 64- or 128-byte swizzle row");
    static_assert(COUT % 16 == 0 && COUT <= 64 && 2 * MT * COUT <= 512, "accumulators do not fit TMEM");
    static_assert(total(NIN, 1) <= S2_SMEM_LIMIT, "tile does not fit shared memory");
};

struct S2Maps {
    CUtensorMap x[2];  // pair view of the input, row parity 0 / 1: [2*Cin, W/2, H/2, N], box [2*Cin, WL, TH + parity, 1]
    CUtensorMap b;     // weights [9*Cin, Cout], box [Cin, Cout]
    CUtensorMap y;     // output [Cout, Wo, Ho, N], box [Cout, TW, TH, 1]
};

struct S2Params {
    const float* bias;
    int tw, th, wl, tiles_x, tiles_y, ntiles;
    float inv_wl;
    uint32_t tx_bytes;  // bytes both plane boxes deliver
};

// position of 16-byte chunk j of row p inside a swizzled tile whose rows are `rowb` bytes (TMA / UMMA 128 / 64 / 32-byte swizzle)
__device__ __forceinline__ uint32_t s2_swz(uint32_t j, uint32_t p, int rowb) {
    return rowb == 128 ? (j ^ (p & 7u)) : (rowb == 64 ? (j ^ ((p >> 1) & 3u)) : (j ^ ((p >> 2) & 1u)));
}
// K-major shared-memory matrix descriptor: rows of `row_bytes` (= swizzle span), 8-row groups 8*row_bytes apart
__device__ __forceinline__ uint64_t s2_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int CIN, int COUT, int MT, int WLMAX, int ACT>
__global__ void __launch_bounds__(S2Cfg<CIN, COUT, MT, WLMAX>::THREADS, 1) conv3x3s2_tc_kernel(const __grid_constant__ S2Maps maps, const S2Params p) {
    using C = S2Cfg<CIN, COUT, MT, WLMAX>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sIn = smem;                              // [NIN][even plane | odd plane]
    uint8_t* sB = sIn + C::NIN * C::STAGE_BYTES;      // [9 taps][Cout rows x Cin]
    uint8_t* sOut = sB + C::B_BYTES;                  // [NOUT][TH*TW rows x Cout]
    float* sBias = reinterpret_cast<float*>(sOut + C::NOUT * C::O_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 64);
    uint64_t* in_full = bars;        // [2]     TMA -> MMA
    uint64_t* in_empty = bars + 2;   // [2]     MMA (tcgen05.commit) -> TMA
    uint64_t* t_full = bars + 4;     // [2][4]  MMA (tcgen05.commit) -> epilogue quad of one M tile
    uint64_t* t_empty = bars + 12;   // [2]     epilogue -> MMA (count = epilogue warps)
    uint64_t* w_full = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.x[0]);
        prefetch_tmap(&maps.x[1]);
        prefetch_tmap(&maps.b);
        prefetch_tmap(&maps.y);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
            mbar_init(&t_empty[s], C::EPI_WARPS);
            for (int q = 0; q < 4; ++q) mbar_init(&t_full[s * 4 + q], 1);
        }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 64; i += C::THREADS) sBias[i] = (p.bias && i < COUT) ? p.bias[i] : 0.f;  // constant data: before griddepcontrol.wait
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {  // the filter: one box per tap, resident for the CTA's whole life
        mbar_expect_tx(w_full, (uint32_t)(9 * COUT * C::ROWB));
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(&maps.b, w_full, sB + tap * C::B_TAP, tap * CIN, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_x * p.tiles_y;
    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int img = tile / per_img, t = tile - img * per_img;
            const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
            const uint32_t s = it % C::NIN, ph = (it / C::NIN) & 1;
            mbar_wait_parked(&in_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&in_full[s], p.tx_bytes);
                // even rows 2*(y0 + i), odd rows 2*(y0 + i) - 1 = odd-plane row y0 - 1 + i; pairs x0 - 1 .. x0 + TW - 1
                tma_load_4d(&maps.x[0], &in_full[s], sIn + s * C::STAGE_BYTES, 0, x0 - 1, y0, img);
                tma_load_4d(&maps.x[1], &in_full[s], sIn + s * C::STAGE_BYTES + C::PLANE_BYTES, 0, x0 - 1, y0 - 1, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t b_s = smem_u32(sB);
        mbar_wait_parked(w_full, 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % C::NIN, ph = (it / C::NIN) & 1;
            const uint32_t acc = it & 1, tph = (it >> 1) & 1;
            mbar_wait_parked(&t_empty[acc], tph ^ 1);
            mbar_wait_parked(&in_full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t in_s = smem_u32(sIn + s * C::STAGE_BYTES);
#pragma unroll 1
                for (int q = 0; q < MT; ++q) {
                    const uint32_t d_tmem = tmem_u + (acc * MT + q) * (uint32_t)COUT;
                    const uint32_t a_q = in_s + (uint32_t)(q * 128) * C::ROWA;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t a_row = a_q + (dy == 1 ? 0u : (uint32_t)C::PLANE_BYTES) + (dy == 2 ? (uint32_t)p.wl * C::ROWA : 0u);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t a_tap = a_row + (dx >= 1 ? (uint32_t)C::ROWA : 0u) + (dx != 1 ? (uint32_t)C::ROWB : 0u);
                            const uint64_t ad = s2_desc(a_tap, C::ROWA);
                            const uint64_t bd = s2_desc(b_s + (uint32_t)(dy * 3 + dx) * C::B_TAP, C::ROWB);
#pragma unroll
                            for (int ks = 0; ks < CIN / 16; ++ks) umma_bf16(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc, (dy | dx | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&t_full[acc * 4 + q]);  // this M tile's accumulator is complete
                }
                umma_commit(&in_empty[s]);  // the planes may be overwritten once these MMAs have read them
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int wi = warp - 4;
        const int q = wi >> 2, lg = wi & 3;  // M tile, TMEM lane group (= warp % 4)
        const uint32_t m = (uint32_t)(q * 128 + lg * 32 + lane);
        const uint32_t r = (uint32_t)__float2int_rd(((float)m + 0.5f) * p.inv_wl);  // exact: (m + 0.5) / wl is never an integer
        const uint32_t c = m - r * (uint32_t)p.wl;
        const bool valid = r < (uint32_t)p.th && c < (uint32_t)p.tw;
        const uint32_t srow = r * (uint32_t)p.tw + c;  // row of the dense [TH][TW] staging tile
        const bool issuer = (wi == 0 && lane == 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, tph = (it >> 1) & 1;
            const uint32_t ob = it % C::NOUT;
            mbar_wait_parked(&t_full[acc * 4 + q], tph);
            tc_fence_after();
            uint32_t v[COUT / 16][16];
#pragma unroll
            for (int g = 0; g < COUT / 16; ++g) tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (acc * MT + q) * (uint32_t)COUT + g * 16, v[g]);
            if (C::NOUT == 1) {  // one staging buffer: the bulk store of the previous tile must have finished READING it
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);  // the accumulator is in registers: the MMA warp may reuse these TMEM columns
            const uint32_t o_s = smem_u32(sOut + ob * C::O_BYTES);
#pragma unroll
            for (int g = 0; g < COUT / 16; ++g) {
                float f[16];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const float4 bq = lds_f4(sBias + g * 16 + 4 * qd);
                    f[4 * qd + 0] = act_t<ACT>(__uint_as_float(v[g][4 * qd + 0]) + bq.x);
                    f[4 * qd + 1] = act_t<ACT>(__uint_as_float(v[g][4 * qd + 1]) + bq.y);
                    f[4 * qd + 2] = act_t<ACT>(__uint_as_float(v[g][4 * qd + 2]) + bq.z);
                    f[4 * qd + 3] = act_t<ACT>(__uint_as_float(v[g][4 * qd + 3]) + bq.w);
                }
                if (valid) {
                    sts128(o_s + srow * (uint32_t)C::ROWO + (s2_swz((uint32_t)(2 * g), srow, C::ROWO) << 4), pack8(f).u);
                    sts128(o_s + srow * (uint32_t)C::ROWO + (s2_swz((uint32_t)(2 * g + 1), srow, C::ROWO) << 4), pack8(f + 8).u);
                }
            }
            fence_proxy_async();
            if (C::NOUT == 2 && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // store of tile it-1 (issued a tile ago): its buffer is the next tile's
            asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
            if (issuer) {
                const int img = tile / per_img, t = tile - img * per_img;
                const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
                tma_store_4d(&maps.y, sOut + ob * C::O_BYTES, 0, x0, y0, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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
typedef CUresult (*S2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static S2EncodeFn s2_get_encode() {
    static S2EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<S2EncodeFn>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool s2_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes) {
    S2EncodeFn enc = s2_get_encode();
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
        set_error("conv3x3s2_tc: cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] swizzle %d", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, sw_bytes);
        return false;
    }
    return true;
}

template <int CIN, int COUT, int MT, int WLMAX>
static int launch_s2(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* y, int act, cudaStream_t st) {
    using C = S2Cfg<CIN, COUT, MT, WLMAX>;
    auto k_silu = conv3x3s2_tc_kernel<CIN, COUT, MT, WLMAX, DBL_ACT_SILU>;
    auto k_none = conv3x3s2_tc_kernel<CIN, COUT, MT, WLMAX, DBL_ACT_NONE>;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_silu, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_none, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("conv3x3s2_tc: cannot opt in to %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    const int sms = dev_info().sms;
    // ---- tile: TH x TW output pixels, GEMM rows TH * (TW + 1) <= MT * 128; fewest rounds of the persistent grid, then fewest tiles
    S2Params p;
    memset(&p, 0, sizeof(p));
    long long best = -1;
    for (int tw = 4; tw <= WLMAX - 1 && tw <= y->w + 3; ++tw) {
        const int wl = tw + 1;
        int th = (MT * 128) / wl;
        if (th > y->h) th = y->h;
        if (th < 1 || th + 1 > 256) continue;
        const long long tiles = (long long)x->n * cdiv(y->w, tw) * cdiv(y->h, th);
        const long long cost = ((tiles + sms - 1) / sms) * (1ll << 32) + tiles;
        if (best < 0 || cost < best) best = cost, p.tw = tw, p.th = th, p.wl = wl;
    }
    DBL_REQUIRE(best >= 0, DBL_E_UNSUPPORTED, "conv3x3s2_tc: no tile for a %dx%d output", y->h, y->w);
    p.tiles_x = cdiv(y->w, p.tw), p.tiles_y = cdiv(y->h, p.th);
    const long long nt = (long long)x->n * p.tiles_x * p.tiles_y;
    DBL_REQUIRE(nt < (1ll << 31), DBL_E_UNSUPPORTED, "conv3x3s2_tc: too many tiles");
    p.ntiles = (int)nt;
    p.inv_wl = 1.0f / (float)p.wl;
    p.bias = bias;
    p.tx_bytes = (uint32_t)((2 * p.th + 1) * p.wl * C::ROWA);
    S2Maps maps;
    memset(&maps, 0, sizeof(maps));
    const uint64_t pitchB = (uint64_t)x->pitch * 2;
    for (int par = 0; par < 2; ++par) {
        const uint64_t dims[4] = {(uint64_t)2 * CIN, (uint64_t)x->w / 2, (uint64_t)x->h / 2, (uint64_t)x->n};
        const uint64_t strides[3] = {2 * pitchB, 2 * pitchB * x->w, pitchB * x->w * x->h};
        const uint32_t box[4] = {(uint32_t)2 * CIN, (uint32_t)p.wl, (uint32_t)(p.th + par), 1};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(x->data) + (par ? pitchB * x->w : 0);
        if (!s2_encode(&maps.x[par], base, 4, dims, strides, box, C::ROWA)) return DBL_E_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)9 * CIN, (uint64_t)COUT};
        const uint64_t strides[1] = {(uint64_t)9 * CIN * 2};
        const uint32_t box[2] = {(uint32_t)CIN, (uint32_t)COUT};
        if (!s2_encode(&maps.b, w, 2, dims, strides, box, C::ROWB)) return DBL_E_CUDA;
    }
    {
        const uint64_t yp = (uint64_t)y->pitch * 2;
        const uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
        const uint64_t strides[3] = {yp, yp * y->w, yp * y->w * y->h};
        const uint32_t box[4] = {(uint32_t)COUT, (uint32_t)p.tw, (uint32_t)p.th, 1};
        if (!s2_encode(&maps.y, y->data, 4, dims, strides, box, C::ROWO)) return DBL_E_CUDA;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    if (act == DBL_ACT_SILU) launch_k(k_silu, grid, C::THREADS, (size_t)C::SMEM, st, maps, p);
    else launch_k(k_none, grid, C::THREADS, (size_t)C::SMEM, st, maps, p);
    count_launch();
    return check_launch("dbl_conv2d(tcgen05 thin stride-2)");
}

// (cin, cout, M tiles per CTA tile, widest plane row): the stride-2 down-samplers of the backbone at n and s scale
#define DBL_S2_TABLE(X) X(16, 32, 4, 64) X(32, 64, 2, 32)

bool conv3x3s2_tc_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad, int dil, int act) {
    static const int off = getenv("DBL_S2_TC") ? !atoi(getenv("DBL_S2_TC")) : 0;  // DBL_S2_TC=0: keep the round-1 kernels
    if (off || k != 3 || stride != 2 || pad != 1 || dil != 1 || r || (act != DBL_ACT_SILU && act != DBL_ACT_NONE)) return false;
    if (x->dtype != DBL_BF16 || y->dtype != DBL_BF16 || !vec8_ok(x) || !vec8_ok(y)) return false;
    if (x->pitch != x->c || (x->h & 1) || (x->w & 1) || x->w < 8 || x->n > 65535) return false;  // the pixel-pair view needs a dense, even-sized image
    if (x->data == y->data) return false;
#define X(CI_, CO_, MT_, WL_) \
    if (x->c == CI_ && y->c == CO_) return true;
    DBL_S2_TABLE(X)
#undef X
    return false;
}

int conv3x3s2_tc(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* y, int act, cudaStream_t st) {
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, DBL_E_ARG, "conv3x3s2_tc: weights must be 16-byte aligned");
#define X(CI_, CO_, MT_, WL_) \
    if (x->c == CI_ && y->c == CO_) return launch_s2<CI_, CO_, MT_, WL_>(x, w, bias, y, act, st);
    DBL_S2_TABLE(X)
#undef X
    set_error("conv3x3s2_tc: no instantiation for cin %d cout %d", x->c, y->c);
    return DBL_E_UNSUPPORTED;
}

}  // namespace dbl
