// Stride-1 3x3 convolution whose whole reduction depth fits ONE shared-memory tile (Cin <= 64, Cout <= 64: the Bottleneck halves of
// layer 6, the neck Bottleneck tails and Detect's box branch at P3, nn/modules/conv.py:39-63 / block.py:344-357 / head.py:86-92) as a
// persistent, warp-specialised kernel in the style of conv_s2_tc.cu:
//   ONE 4-D TMA box per tile (halo included, zero fill = padding) -> nine taps as descriptor row offsets of that buffer ->
//   tcgen05.mma, MT accumulators of a tile double-buffered in TMEM -> bias / SiLU / residual (TMA-loaded tile) epilogue ->
//   swizzled staging tile -> ONE TMA store.
// conv_tc4.cu (multi-accumulator super-tiles, weights streamed) stays the kernel for Cin >= 128.  For these small-K shapes its tiles
// hold only 18-36 MMAs (< 1 us of tensor work) and were paced by its row-per-thread epilogue stores: 64 -> 32 @80^2 took 37 us
// against 20 us of HBM time, 32 -> 64 (+residual) 45 us against 20.  tools/ubench/umma_rate.cu: an M=128 K=16 MMA with N <= 32 holds
// the tensor pipe 40 cycles (N = 64: 48) whatever the layout or the accumulator pattern, so these layers are HBM-bound by design.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mma_util.cuh"

namespace dbl {

constexpr int T5_SMEM_LIMIT = 227 * 1024;
__host__ __device__ constexpr int t5_round1k(int v) { return (v + 1023) / 1024 * 1024; }

template <int CIN, int COUT, int MT, int WLMAX, bool RES>
struct T5Cfg {
    static constexpr int ROWA = 2 * CIN;                        // bytes of one input pixel = one GEMM row (32 / 64 / 128: the swizzle span)
    static constexpr int ROWO = 2 * COUT;                       // bytes of one output pixel
    static constexpr int TILE_PIX = MT * 128 + 2 * WLMAX + 8;   // >= (TH + 2) * WL and >= the junk rows the last M tile reads
    static constexpr int STAGE_BYTES = t5_round1k(TILE_PIX * ROWA);
    static constexpr int B_TAP = t5_round1k(COUT * ROWA);
    static constexpr int B_BYTES = 9 * B_TAP;
    static constexpr int O_BYTES = t5_round1k(MT * 128 * ROWO);
    static constexpr int MISC = 1024;                           // bias + barriers
    static constexpr int total(int nin, int nout, int nres) { return 1024 + nin * STAGE_BYTES + B_BYTES + (nout + (RES ? nres : 0)) * O_BYTES + MISC; }
    static constexpr int NOUT = total(2, 2, 1) <= T5_SMEM_LIMIT ? 2 : 1;
    static constexpr int NRES = total(2, NOUT, 2) <= T5_SMEM_LIMIT ? 2 : 1;
    static constexpr int NIN = total(4, NOUT, NRES) <= T5_SMEM_LIMIT ? 4 : (total(3, NOUT, NRES) <= T5_SMEM_LIMIT ? 3 : 2);
    static constexpr int SMEM = total(NIN, NOUT, NRES);
    static constexpr int EPI_WARPS = 4 * MT;
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;        // warp 0 TMA, warp 1 MMA, warps 2-3 idle (epilogue warp w must own TMEM lanes 32*(w%4)..)
    static constexpr int TMEM_COLS = 2 * MT * COUT <= 32 ? 32 : (2 * MT * COUT <= 64 ? 64 : (2 * MT * COUT <= 128 ? 128 : (2 * MT * COUT <= 256 ? 256 : 512)));
    static_assert(ROWA == 32 || ROWA == 64 || ROWA == 128, "an input pixel must be one swizzle row");
    static_assert(COUT % 16 == 0 && COUT <= 64 && 2 * MT * COUT <= 512, "accumulators do not fit TMEM");
    static_assert(total(2, 1, 1) <= T5_SMEM_LIMIT, "tile does not fit shared memory");
};

struct T5Maps {
    CUtensorMap x;  // input [Cin, W, H, N], box [Cin, WL, TH + 2, 1]
    CUtensorMap b;  // weights [9*Cin, Cout], box [Cin, Cout]
    CUtensorMap y;  // output [Cout, W, H, N], box [Cout, TW, TH, 1]
    CUtensorMap r;  // residual, like y
};

struct T5Params {
    const float* bias;
    int tw, th, wl, tiles_x, tiles_y, ntiles, epi;
    float inv_wl;
    uint32_t tx_bytes;  // bytes of one halo box
    uint32_t r_bytes;   // bytes of one residual box
    unsigned long long* trace;  // developer aid (dbl_dsconv_tc_debug_trace): CTA 0 writes [role][tile iteration][event] globaltimer stamps
};

unsigned long long* debug_trace_buffer();  // dsconv_tc.cu
__device__ __forceinline__ void t5_stamp(const T5Params& p, int role, uint32_t it, int ev) {
    if (p.trace && blockIdx.x == 0 && it < 16u) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(role * 16 + it) * 4 + ev] = t;
    }
}

// position of 16-byte chunk j of row p inside a swizzled tile whose rows are `rowb` bytes (TMA / UMMA 128 / 64 / 32-byte swizzle)
__device__ __forceinline__ uint32_t t5_swz(uint32_t j, uint32_t p, int rowb) {
    return rowb == 128 ? (j ^ (p & 7u)) : (rowb == 64 ? (j ^ ((p >> 1) & 3u)) : (j ^ ((p >> 2) & 1u)));
}
// K-major shared-memory matrix descriptor: rows of `row_bytes` (= swizzle span), 8-row groups 8*row_bytes apart
__device__ __forceinline__ uint64_t t5_desc(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int CIN, int COUT, int MT, int WLMAX, bool RES, int ACT>
__global__ void __launch_bounds__(T5Cfg<CIN, COUT, MT, WLMAX, RES>::THREADS, 1) conv3x3_tc5_kernel(const __grid_constant__ T5Maps maps, const T5Params p) {
    using C = T5Cfg<CIN, COUT, MT, WLMAX, RES>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sIn = smem;                              // [NIN] halo tiles, pixel-major
    uint8_t* sB = sIn + C::NIN * C::STAGE_BYTES;      // [9 taps][Cout rows x Cin]
    uint8_t* sOut = sB + C::B_BYTES;                  // [NOUT][TH*TW rows x Cout]
    uint8_t* sRes = sOut + C::NOUT * C::O_BYTES;      // [NRES] residual tiles, same layout
    float* sBias = reinterpret_cast<float*>(sRes + (RES ? C::NRES * C::O_BYTES : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 64);
    uint64_t* in_full = bars;        // [4]     TMA -> MMA
    uint64_t* in_empty = bars + 4;   // [4]     MMA (tcgen05.commit) -> TMA
    uint64_t* t_full = bars + 8;     // [2][4]  MMA (tcgen05.commit) -> epilogue quad of one M tile
    uint64_t* t_empty = bars + 16;   // [2]     epilogue -> MMA (count = epilogue warps)
    uint64_t* r_full = bars + 18;    // [2]     TMA -> epilogue
    uint64_t* r_empty = bars + 20;   // [2]     epilogue -> TMA (count = epilogue warps)
    uint64_t* w_full = bars + 22;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.x);
        prefetch_tmap(&maps.b);
        prefetch_tmap(&maps.y);
        if (RES) prefetch_tmap(&maps.r);
        for (int s = 0; s < 4; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t_empty[s], C::EPI_WARPS);
            mbar_init(&r_full[s], 1);
            mbar_init(&r_empty[s], C::EPI_WARPS);
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
        mbar_expect_tx(w_full, (uint32_t)(9 * COUT * C::ROWA));
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
            if (lane == 0) t5_stamp(p, 0, it, 0);
            if (elect_one()) {
                mbar_expect_tx(&in_full[s], p.tx_bytes);
                tma_load_4d(&maps.x, &in_full[s], sIn + s * C::STAGE_BYTES, 0, x0 - 1, y0 - 1, img);
            }
            __syncwarp();
            if (RES) {
                const uint32_t rs = it % C::NRES, rph = (it / C::NRES) & 1;
                mbar_wait_parked(&r_empty[rs], rph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&r_full[rs], p.r_bytes);
                    tma_load_4d(&maps.r, &r_full[rs], sRes + rs * C::O_BYTES, 0, x0, y0, img);
                }
                __syncwarp();
            }
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
            if (lane == 0) t5_stamp(p, 1, it, 0);
            mbar_wait_parked(&in_full[s], ph);
            if (lane == 0) t5_stamp(p, 1, it, 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t in_s = smem_u32(sIn + s * C::STAGE_BYTES);
#pragma unroll 1
                for (int q = 0; q < MT; ++q) {
                    const uint32_t d_tmem = tmem_u + (acc * MT + q) * (uint32_t)COUT;
                    const uint32_t a_q = in_s + (uint32_t)(q * 128) * C::ROWA;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t a_row = a_q + (uint32_t)(dy * p.wl) * C::ROWA;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t ad = t5_desc(a_row + (uint32_t)dx * C::ROWA, C::ROWA);
                            const uint64_t bd = t5_desc(b_s + (uint32_t)(dy * 3 + dx) * C::B_TAP, C::ROWA);
#pragma unroll
                            for (int ks = 0; ks < CIN / 16; ++ks) umma_bf16(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc, (dy | dx | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&t_full[acc * 4 + q]);  // this M tile's accumulator is complete
                }
                umma_commit(&in_empty[s]);  // the halo tile may be overwritten once these MMAs have read it
            }
            __syncwarp();
            if (lane == 0) t5_stamp(p, 1, it, 2);
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int wi = warp - 4;
        const int q = wi >> 2, lg = wi & 3;  // M tile, TMEM lane group (= warp % 4)
        const uint32_t m = (uint32_t)(q * 128 + lg * 32 + lane);
        const uint32_t r = (uint32_t)__float2int_rd(((float)m + 0.5f) * p.inv_wl);  // exact: (m + 0.5) / wl is never an integer
        const uint32_t c = m - r * (uint32_t)p.wl;
        const bool valid = r < (uint32_t)p.th && c < (uint32_t)p.tw;
        const uint32_t srow = r * (uint32_t)p.tw + c;  // row of the dense [TH][TW] staging / residual tile
        const bool issuer = (wi == 0 && lane == 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, tph = (it >> 1) & 1;
            const uint32_t ob = it % C::NOUT;
            const uint32_t rs = it % C::NRES, rph = (it / C::NRES) & 1;
            if (RES) mbar_wait_parked(&r_full[rs], rph);
            mbar_wait_parked(&t_full[acc * 4 + q], tph);
            tc_fence_after();
            if (issuer) t5_stamp(p, 3, it, 0);
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
            if (issuer) t5_stamp(p, 3, it, 1);
            const uint32_t o_s = smem_u32(sOut + ob * C::O_BYTES), r_s = smem_u32(sRes + (RES ? rs * C::O_BYTES : 0));
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
                    const uint32_t off0 = srow * (uint32_t)C::ROWO + (t5_swz((uint32_t)(2 * g), srow, C::ROWO) << 4);
                    const uint32_t off1 = srow * (uint32_t)C::ROWO + (t5_swz((uint32_t)(2 * g + 1), srow, C::ROWO) << 4);
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
            if (RES) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&r_empty[rs]);
            }
            fence_proxy_async();
            if (C::NOUT == 2 && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // store of tile it-1 (issued a tile ago): its buffer is the next tile's
            asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
            if (issuer) {
                const int img = tile / per_img, t = tile - img * per_img;
                const int y0 = (t / p.tiles_x) * p.th, x0 = (t % p.tiles_x) * p.tw;
                tma_store_4d(&maps.y, sOut + ob * C::O_BYTES, 0, x0, y0, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                t5_stamp(p, 3, it, 2);
            }
            if (wi == 4 * (MT - 1) && lane == 0) t5_stamp(p, 2, it, 0);  // last M tile's quad: end of its tile
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
typedef CUresult (*T5EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static T5EncodeFn t5_get_encode() {
    static T5EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<T5EncodeFn>(q);
        else
            cudaGetLastError();
    });
    return fn;
}
static bool t5_encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int sw_bytes) {
    T5EncodeFn enc = t5_get_encode();
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
        set_error("conv3x3_tc5: cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] swizzle %d", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, sw_bytes);
        return false;
    }
    return true;
}


static bool t5_map_nhwc(CUtensorMap* m, const dbl_tensor* t, int cw, int bw, int bh, int sw_bytes) {
    const uint64_t pitchB = (uint64_t)t->pitch * 2;
    const uint64_t dims[4] = {(uint64_t)t->c, (uint64_t)t->w, (uint64_t)t->h, (uint64_t)t->n};
    const uint64_t strides[3] = {pitchB, pitchB * t->w, pitchB * t->w * t->h};
    const uint32_t box[4] = {(uint32_t)cw, (uint32_t)bw, (uint32_t)bh, 1};
    return t5_encode(m, t->data, 4, dims, strides, box, sw_bytes);
}

template <int CIN, int COUT, int MT, int WLMAX, bool RES, int ACT>
static int launch_t5(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int epi, cudaStream_t st) {
    using C = T5Cfg<CIN, COUT, MT, WLMAX, RES>;
    auto kern = conv3x3_tc5_kernel<CIN, COUT, MT, WLMAX, RES, ACT>;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("conv3x3_tc5: cannot opt in to %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    const int sms = dev_info().sms;
    // ---- tile: TH x TW output pixels, GEMM rows TH * (TW + 2) <= MT * 128; fewest rounds of the persistent grid, then fewest tiles
    T5Params p;
    memset(&p, 0, sizeof(p));
    long long best = -1;
    for (int tw = 4; tw <= WLMAX - 2 && tw <= y->w + 3; ++tw) {
        const int wl = tw + 2;
        int th = (MT * 128) / wl;
        if (th > y->h) th = y->h;
        if (th < 1 || th + 2 > 256) continue;
        const long long tiles = (long long)x->n * cdiv(y->w, tw) * cdiv(y->h, th);
        const long long cost = ((tiles + sms - 1) / sms) * (1ll << 32) + tiles;
        if (best < 0 || cost < best) best = cost, p.tw = tw, p.th = th, p.wl = wl;
    }
    DBL_REQUIRE(best >= 0, DBL_E_UNSUPPORTED, "conv3x3_tc5: no tile for a %dx%d output", y->h, y->w);
    p.tiles_x = cdiv(y->w, p.tw), p.tiles_y = cdiv(y->h, p.th);
    const long long nt = (long long)x->n * p.tiles_x * p.tiles_y;
    DBL_REQUIRE(nt < (1ll << 31), DBL_E_UNSUPPORTED, "conv3x3_tc5: too many tiles");
    p.ntiles = (int)nt;
    p.inv_wl = 1.0f / (float)p.wl;
    p.bias = bias, p.epi = epi, p.trace = debug_trace_buffer();
    p.tx_bytes = (uint32_t)((p.th + 2) * p.wl * C::ROWA);
    p.r_bytes = (uint32_t)(p.th * p.tw * C::ROWO);
    T5Maps maps;
    memset(&maps, 0, sizeof(maps));
    if (!t5_map_nhwc(&maps.x, x, CIN, p.wl, p.th + 2, C::ROWA)) return DBL_E_CUDA;
    {
        const uint64_t dims[2] = {(uint64_t)9 * CIN, (uint64_t)COUT};
        const uint64_t strides[1] = {(uint64_t)9 * CIN * 2};
        const uint32_t box[2] = {(uint32_t)CIN, (uint32_t)COUT};
        if (!t5_encode(&maps.b, w, 2, dims, strides, box, C::ROWA)) return DBL_E_CUDA;
    }
    if (!t5_map_nhwc(&maps.y, y, COUT, p.tw, p.th, C::ROWO)) return DBL_E_CUDA;
    if (RES && !t5_map_nhwc(&maps.r, r, COUT, p.tw, p.th, C::ROWO)) return DBL_E_CUDA;
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch_k(kern, grid, C::THREADS, (size_t)C::SMEM, st, maps, p);
    count_launch();
    return check_launch("dbl_conv2d(tcgen05 3x3 single-chunk)");
}

// (cin, cout, M tiles per CTA tile, widest halo row)
#define DBL_T5_TABLE(X) X(64, 32, 2, 32) X(32, 64, 2, 32) X(64, 64, 2, 32) X(32, 32, 2, 32) X(32, 16, 4, 64) X(16, 32, 4, 64) X(16, 16, 4, 64)

bool conv3x3_tc5_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad, int dil, int act) {
    static const int off = getenv("DBL_TC5") ? !atoi(getenv("DBL_TC5")) : 0;  // DBL_TC5=0: keep conv_tc4 for these shapes
    if (off || k != 3 || stride != 1 || pad != 1 || dil != 1 || (act != DBL_ACT_SILU && act != DBL_ACT_NONE)) return false;
    if (x->dtype != DBL_BF16 || y->dtype != DBL_BF16 || !vec8_ok(x) || !vec8_ok(y) || (r && !vec8_ok(r))) return false;
    if (x->w < 4 || x->n > 65535 || x->data == y->data) return false;
#define X(CI_, CO_, MT_, WL_) \
    if (x->c == CI_ && y->c == CO_) return true;
    DBL_T5_TABLE(X)
#undef X
    return false;
}

int conv3x3_tc5(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int act, int epi, cudaStream_t st) {
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, DBL_E_ARG, "conv3x3_tc5: weights must be 16-byte aligned");
    const bool res = r && epi != DBL_EPI_NONE;
#define X(CI_, CO_, MT_, WL_)                                                                                                        \
    if (x->c == CI_ && y->c == CO_) {                                                                                                \
        if (act == DBL_ACT_SILU)                                                                                                     \
            return res ? launch_t5<CI_, CO_, MT_, WL_, true, DBL_ACT_SILU>(x, w, bias, r, y, epi, st)                                \
                       : launch_t5<CI_, CO_, MT_, WL_, false, DBL_ACT_SILU>(x, w, bias, nullptr, y, DBL_EPI_NONE, st);               \
        return res ? launch_t5<CI_, CO_, MT_, WL_, true, DBL_ACT_NONE>(x, w, bias, r, y, epi, st)                                    \
                   : launch_t5<CI_, CO_, MT_, WL_, false, DBL_ACT_NONE>(x, w, bias, nullptr, y, DBL_EPI_NONE, st);                   \
    }
    DBL_T5_TABLE(X)
#undef X
    set_error("conv3x3_tc5: no instantiation for cin %d cout %d", x->c, y->c);
    return DBL_E_UNSUPPORTED;
}

}  // namespace dbl
