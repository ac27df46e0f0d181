// Fused depthwise-separable convolution: depthwise k x k (CUDA cores, fp32) -> pointwise 1x1 (warp MMA) -> bias -> SiLU
// (-> + residual) in ONE kernel.  Reference: DSConv.forward (nn/modules/conv.py:91-108: dw -> pw -> BN -> SiLU, BN folded
// into the pointwise weights here), the two halves of DSBottleneck (nn/modules/block.py:1408-1444) and the
// [DWConv + BN + SiLU, Conv 1x1 + BN + SiLU] pairs of Detect's class branch (nn/modules/head.py:93-100; DW_ACT = 1 applies
// the depthwise bias + SiLU before the pointwise GEMM).
//
// Unfused, every DSConv was two launches with the depthwise result making a round trip through L2, and the pointwise GEMMs
// of these layers (K = N = 32..128 on 26 MB) are so small that their tcgen05 launches are all fixed cost (11-16 us each,
// 38 of them per step).  Here a CTA owns a TW x TH output tile:
//   1. cp.async the input halo tile (pixel-major, zero-filled = padding), the depthwise filter (f32) and the packed
//      pointwise weights into shared memory;
//   2. depthwise: thread = one output column x 8 channels x R rows (the inner loops of dwconv_tile_kernel), result rounded
//      to bf16 into a PLANAR tile (plane = 8 channels, 16 bytes per pixel) = the A operand layout of conv_mma.cu;
//   3. pointwise: each warp takes 16-pixel blocks x 64-channel chunks, ldmatrix + mma.sync m16n8k16, epilogue straight to
//      global memory (the packed weights' channel permutation gives every thread 16-byte contiguous output runs).
#include <cstdlib>
#include <cstring>

#include "mma_util.cuh"

namespace dbl {

// n 32-bit words (2 bf16 each) to / from global memory with VEC-byte accesses (p is VEC-byte aligned)
template <int N, int VEC>
__device__ __forceinline__ void store_words(__nv_bfloat16* p, const uint32_t* v) {
    if (VEC == 16) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) reinterpret_cast<uint4*>(p)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if (VEC == 8) {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) reinterpret_cast<uint2*>(p)[q] = make_uint2(v[2 * q], v[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) reinterpret_cast<uint32_t*>(p)[q] = v[q];
    }
}
template <int N, int VEC>
__device__ __forceinline__ void load_words(const __nv_bfloat16* p, uint32_t* v) {
    if (VEC == 16) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const uint4 u = reinterpret_cast<const uint4*>(p)[q];
            v[4 * q] = u.x, v[4 * q + 1] = u.y, v[4 * q + 2] = u.z, v[4 * q + 3] = u.w;
        }
    } else if (VEC == 8) {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const uint2 u = reinterpret_cast<const uint2*>(p)[q];
            v[2 * q] = u.x, v[2 * q + 1] = u.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = reinterpret_cast<const uint32_t*>(p)[q];
    }
}

struct DsParams {
    const __nv_bfloat16* x;
    const float* wdw;    // [K*K][CIN]
    const float* bdw;    // [CIN] or null (DW_ACT only)
    const uint2* wpw;    // packed (COUT x CIN, 1 tap)
    const float* bpw;    // [COUT]
    const __nv_bfloat16* r;
    __nv_bfloat16* y;
    int h, w, ho, wo, xpitch, ypitch, rpitch, epi;
    int tw, th, tiles_x, tiles_y;
};

template <int K, int S, int CIN, int COUT>
struct DsCfg {
    static constexpr int R = 4;                       // output rows per depthwise thread
    static constexpr int G = CIN / 8;                 // 8-channel groups
    static constexpr int NB = COUT / 8;               // n-blocks of the pointwise GEMM
    static constexpr int NBC = NB <= 10 ? NB : 8;     // n-blocks per warp work item (<= 80 output channels)
    static constexpr int NCHUNK = NB / NBC;
    static constexpr int VEC = (NB % 4 == 0) ? 16 : ((NB % 2 == 0) ? 8 : 4);  // alignment (bytes) of a thread's channel run
    static constexpr int KS = CIN / 16;
    static constexpr int WPB = KS * NB * 256;         // packed pointwise weights (bytes)
    static constexpr int WDB = K * K * CIN * 4;       // depthwise filter (bytes)
    static_assert(NB % NBC == 0, "COUT must be a multiple of 64 or at most 80");
};

// depthwise stage of the fused kernel = the inner loops of dwconv_tile3_kernel (conv_direct.cu): a work item is one output
// column x FOUR channels x R rows, packed fp32 FMAs over channel pairs, filter column in registers; the result is rounded to
// bf16 into the planar A tile (plane = 8 channels, 16 bytes per pixel; a 4-channel item fills one half of a 16-byte slot).
template <int K, int S, int CIN, int R, int DW_ACT>
__device__ __forceinline__ void ds_depthwise(uint32_t sX, uint32_t sA, const float* sWd, const float* sBd, int tw, int th, int ic, int psa,
                                             int tid, int NT) {
    constexpr int G4 = CIN / 4;
    constexpr int ROWS_T = (R - 1) * S + K;
    const int strips = th / R;
    const int items = G4 * tw * strips;
    const uint32_t wd_s = smem_u32(sWd);
    for (int it = tid; it < items; it += NT) {
        const int g = it % G4, rest = it / G4;
        const int xi = rest % tw, ys = rest / tw;
        unsigned long long acc[R][2];
        {
            ulonglong2 b2 = make_ulonglong2(0ull, 0ull);
            if (DW_ACT) {
                const uint4 bb = lds128(smem_u32(sBd) + (uint32_t)g * 16);
                b2 = make_ulonglong2(((unsigned long long)bb.y << 32) | bb.x, ((unsigned long long)bb.w << 32) | bb.z);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][0] = b2.x, acc[r][1] = b2.y;
        }
        const uint32_t xbase = sX + (uint32_t)(((ys * R * S) * ic + xi * S) * G4 + g) * 8;
#pragma unroll 1
        for (int kx = 0; kx < K; ++kx) {
            unsigned long long wk[K][2];
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                // explicit shared-memory load: through the generic `const float*` parameter ptxas emitted LD.E.128 (generic) here
                const uint4 w0 = lds128(wd_s + (uint32_t)(((ky * K + kx) * CIN + g * 4) * 4));
                wk[ky][0] = ((unsigned long long)w0.y << 32) | w0.x, wk[ky][1] = ((unsigned long long)w0.w << 32) | w0.z;
            }
#pragma unroll
            for (int iy = 0; iy < ROWS_T; ++iy) {
                bool used = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int d = iy - r * S;
                    if (d >= 0 && d < K) used = true;
                }
                if (!used) continue;
                const uint2 v = lds64(xbase + (uint32_t)((iy * ic + kx) * G4) * 8);
                const unsigned long long x0 = bf16x2_to_f32x2(v.x), x1 = bf16x2_to_f32x2(v.y);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int d = iy - r * S;
                    if (d >= 0 && d < K) {
                        ffma2(acc[r][0], x0, wk[d][0]);
                        ffma2(acc[r][1], x1, wk[d][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float f[4];
            f32x2_split(acc[r][0], f[0], f[1]);
            f32x2_split(acc[r][1], f[2], f[3]);
            if (DW_ACT) {
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = silu_fast(f[i]);
            }
            sts64(sA + (uint32_t)((g >> 1) * psa + ((ys * R + r) * tw + xi) * 16 + (g & 1) * 8), make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3])));
        }
    }
}

template <int K, int S, int CIN, int COUT, int DW_ACT>
__global__ void __launch_bounds__(512) dsconv_mma_kernel(const DsParams p) {
    using C = DsCfg<K, S, CIN, COUT>;
    constexpr int G = C::G, PAD = (K - 1) / 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int tw = p.tw, th = p.th, mpx = tw * th;
    const int ir = (th - 1) * S + K, ic = (tw - 1) * S + K;
    const int mpad = (mpx + 15) & ~15;
    const int psa = (mpad + 1) * 16;  // A-tile plane stride (bytes), = 16 mod 128 when mpad is a multiple of 8
    // carve: [packed pw weights][dw filter f32][pw bias][dw bias][A tile planar][input tile]
    const uint32_t sWp = smem_u32(smem);
    float* sWd = reinterpret_cast<float*>(smem + C::WPB);
    float* sBp = sWd + K * K * CIN;
    float* sBd = sBp + COUT;
    const uint32_t sA = smem_u32(reinterpret_cast<uint8_t*>(sBd + CIN));
    const uint32_t sX = sA + (uint32_t)(G * psa);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x;  // chosen by the host = depthwise work items of the tile (multiple of 32, <= 512)
    pdl_launch_dependents();

    const int img = blockIdx.z;
    const int oy0 = blockIdx.y * th, ox0 = blockIdx.x * tw;
    for (int i = tid; i < C::WPB / 16; i += NT) cp_async16(sWp + i * 16, reinterpret_cast<const uint8_t*>(p.wpw) + i * 16, true);
    for (int i = tid; i < C::WDB / 16; i += NT) cp_async16(smem_u32(sWd) + i * 16, reinterpret_cast<const uint8_t*>(p.wdw) + i * 16, true);
    for (int i = tid; i < COUT; i += NT) sBp[i] = p.bpw ? p.bpw[i] : 0.f;
    for (int i = tid; i < CIN; i += NT) sBd[i] = (DW_ACT && p.bdw) ? p.bdw[i] : 0.f;
    pdl_wait();
    {
        const __nv_bfloat16* ximg = p.x + (size_t)img * p.h * p.w * p.xpitch;
        const int iy0 = oy0 * S - PAD, ix0 = ox0 * S - PAD;
        stage_tile16(sX, ximg, p.x, ir, ic, G, G, iy0, ix0, p.h, p.w, p.xpitch, tid, NT);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---------------- depthwise: (4-channel group, column, strip of R rows) per work item ----------------
    if (th % 8 == 0)
        ds_depthwise<K, S, CIN, 8, DW_ACT>(sX, sA, sWd, sBd, tw, th, ic, psa, tid, NT);
    else
        ds_depthwise<K, S, CIN, 4, DW_ACT>(sX, sA, sWd, sBd, tw, th, ic, psa, tid, NT);
    __syncthreads();

    // ---------------- pointwise: work item = (16-pixel block, 64-channel chunk) ----------------
    {
        const int g = lane >> 2, t = lane & 3;
        const int mblocks = mpad / 16;
        const uint32_t row16 = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 16);
        const uint32_t a_lane = sA + row16 + (lane >= 16 ? (uint32_t)psa : 0u);
        for (int item = warp; item < mblocks * C::NCHUNK; item += (NT >> 5)) {
            const int mb = item / C::NCHUNK, ch = item - mb * C::NCHUNK;
            float acc[C::NBC][4];
#pragma unroll
            for (int j = 0; j < C::NBC; ++j) {
                const float2 bb = *reinterpret_cast<const float2*>(sBp + t * 2 * C::NB + 2 * (ch * C::NBC + j));  // generic load on purpose: the explicit (volatile) ld.shared form was measured 10 us slower here (8 loads pinned between the MMAs)
                acc[j][0] = acc[j][2] = bb.x, acc[j][1] = acc[j][3] = bb.y;
            }
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks) {
                uint32_t a[4];
                ldsm_x4(a_lane + (uint32_t)(2 * ks * psa + mb * 256), a);
#pragma unroll
                for (int j = 0; j < C::NBC; ++j) {
                    const uint2 bw = lds64(sWp + (uint32_t)((ks * C::NB + ch * C::NBC + j) * 256) + lane * 8);
                    mma16816(acc[j], a, bw.x, bw.y);
                }
            }
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
                const int pidx = mb * 16 + g + 8 * rh;
                const int ty = pidx / tw, tx = pidx - ty * tw;
                const int gy = oy0 + ty, gx = ox0 + tx;
                if (pidx >= mpx || gy >= p.ho || gx >= p.wo) continue;
                const size_t opix = ((size_t)img * p.ho + gy) * p.wo + gx;
                const int c0 = t * 2 * C::NB + 2 * ch * C::NBC;  // first of this thread's 2*NBC contiguous channels
                uint32_t res[C::NBC];
                if (p.epi != DBL_EPI_NONE) load_words<C::NBC, C::VEC>(p.r + opix * p.rpitch + c0, res);
                uint32_t v[C::NBC];
#pragma unroll
                for (int j = 0; j < C::NBC; ++j) {
                    float a0 = silu_fast(acc[j][2 * rh]), a1 = silu_fast(acc[j][2 * rh + 1]);
                    if (p.epi != DBL_EPI_NONE) {
                        float r0, r1;
                        unpack2(res[j], r0, r1);
                        if (p.epi == DBL_EPI_ADD) a0 += r0, a1 += r1;
                        else a0 *= r0, a1 *= r1;
                    }
                    v[j] = pack2(a0, a1);
                }
                store_words<C::NBC, C::VEC>(p.y + opix * p.ypitch + c0, v);
            }
        }
    }
}

static bool ds_pick_tile(int k, int s, int cin, int cout, int ho, int wo, int max_smem, int* tw_, int* th_, int* nt_, size_t* smem_) {
    // cost = depthwise passes x threads (incl. idle lanes of the last pass and padding of the map) + pointwise rounds + fill
    long long best = -1;
    const int G = cin / 8, nchunk = (cout / 8) <= 10 ? 1 : (cout / 64);
    for (int tw = 4; tw <= 40; tw += 2) {
        if (tw > wo && tw != 4) break;
        for (int th = 4; th <= 16; th += 4) {
            const int mpx = tw * th;
            if (mpx > 320) continue;
            const int ir = (th - 1) * s + k, ic = (tw - 1) * s + k;
            const int mpad = (mpx + 15) & ~15;
            const size_t smem = 256 + (size_t)(cin / 16) * (cout / 8) * 256 + (size_t)k * k * cin * 4 + (size_t)(cout + cin) * 4 +
                                (size_t)(cin / 8) * (mpad + 1) * 16 + (size_t)ir * ic * cin * 2;
            if (smem > (size_t)max_smem) continue;
            const int items = 2 * G * tw * (th / (th % 8 == 0 ? 8 : 4));
            int nt = (items + 31) / 32 * 32;
            if (nt > 512) nt = 512;
            if (nt < 64) nt = 64;
            const long long tiles = (long long)cdiv(wo, tw) * cdiv(ho, th);
            const long long dw_cost = (long long)cdiv(items, nt) * nt * 4 * k * k / 4;            // thread-slots x taps
            const long long pw_cost = (long long)cdiv((mpad / 16) * nchunk, nt / 32) * 32 * (cin / 16) * 10;
            const long long fill = (long long)ir * ic * G / 2;
            long long cost = tiles * (dw_cost + pw_cost + fill + 2000);
            if (smem > 112 * 1024) cost += cost / 2;  // a single resident CTA cannot overlap its fill with another CTA's math
            if (best < 0 || cost < best) best = cost, *tw_ = tw, *th_ = th, *nt_ = nt, *smem_ = smem;
        }
    }
    return best >= 0;
}

template <int K, int S, int CIN, int COUT, int DW_ACT>
static int launch_ds(const dbl_tensor* x, const float* wdw, const float* bdw, const void* wpw, const float* bpw, const dbl_tensor* r,
                     const dbl_tensor* y, int epi, cudaStream_t st) {
    const int max_smem = dev_info().max_smem;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        if (cudaFuncSetAttribute(dsconv_mma_kernel<K, S, CIN, COUT, DW_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
            set_error("dbl_dsconv: cannot opt in to %d bytes of shared memory: %s", max_smem, cudaGetErrorString(cudaGetLastError()));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DsParams p;
    memset(&p, 0, sizeof(p));
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data), p.wdw = wdw, p.bdw = bdw, p.wpw = reinterpret_cast<const uint2*>(wpw), p.bpw = bpw;
    p.r = r ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr, p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
    p.h = x->h, p.w = x->w, p.ho = y->h, p.wo = y->w, p.xpitch = x->pitch, p.ypitch = y->pitch, p.rpitch = r ? r->pitch : 0;
    p.epi = r ? epi : DBL_EPI_NONE;
    size_t smem = 0;
    int nt = 256;
    if (!ds_pick_tile(K, S, CIN, COUT, y->h, y->w, max_smem - 1024, &p.tw, &p.th, &nt, &smem)) {
        set_error("dbl_dsconv: no tile fits shared memory (cin %d cout %d k %d)", CIN, COUT, K);
        return DBL_E_UNSUPPORTED;
    }
    {
        static const char* force = getenv("DBL_DS_TILE");  // tuning aid: "tw,th" (th a multiple of 4)
        int ftw, fth;
        if (force && sscanf(force, "%d,%d", &ftw, &fth) == 2 && ftw >= 4 && fth >= 4 && fth % 4 == 0 && ftw * fth <= 320) {
            const int ir = (fth - 1) * S + K, ic = (ftw - 1) * S + K, mpad = (ftw * fth + 15) & ~15;
            const size_t sm = 256 + (size_t)(CIN / 16) * (COUT / 8) * 256 + (size_t)K * K * CIN * 4 + (size_t)(COUT + CIN) * 4 +
                              (size_t)(CIN / 8) * (mpad + 1) * 16 + (size_t)ir * ic * CIN * 2;
            if (sm <= (size_t)max_smem - 1024) {
                p.tw = ftw, p.th = fth, smem = sm;
                const int items = 2 * (CIN / 8) * ftw * (fth / (fth % 8 == 0 ? 8 : 4));
                nt = (items + 31) / 32 * 32;
                if (nt > 512) nt = 512;
                if (nt < 64) nt = 64;
            }
        }
        static const char* dbg = getenv("DBL_DS_DEBUG");
        if (dbg) fprintf(stderr, "dsconv k%d s%d %d->%d %dx%d: tw=%d th=%d threads=%d smem=%zu\n", K, S, CIN, COUT, y->h, y->w, p.tw, p.th, nt, smem);
    }
    p.tiles_x = cdiv(y->w, p.tw), p.tiles_y = cdiv(y->h, p.th);
    DBL_REQUIRE(p.tiles_y <= 65535 && x->n <= 65535, DBL_E_UNSUPPORTED, "dbl_dsconv: grid too large");
    launch_k(dsconv_mma_kernel<K, S, CIN, COUT, DW_ACT>, dim3(p.tiles_x, p.tiles_y, x->n), nt, smem, st, p);
    count_launch();
    return check_launch("dbl_dsconv");
}

}  // namespace dbl

using namespace dbl;

#define DBL_DS_TABLE(X) \
    X(3, 1, 32, 32) X(7, 1, 32, 32) X(3, 1, 64, 64) X(7, 1, 64, 64) X(3, 1, 128, 128) X(7, 1, 128, 128) X(3, 2, 128, 128) X(3, 2, 128, 256) \
    X(3, 1, 64, 80) X(3, 1, 80, 80) X(3, 1, 128, 80) X(3, 1, 256, 80) X(3, 1, 128, 64) X(3, 1, 256, 64) X(3, 2, 64, 64) X(3, 2, 256, 256)

extern "C" int dbl_dsconv_supported(int k, int stride, int cin, int cout) {
#define X(K_, S_, CI_, CO_) \
    if (k == K_ && stride == S_ && cin == CI_ && cout == CO_) return 1;
    DBL_DS_TABLE(X)
#undef X
    return 0;
}

extern "C" int dbl_dsconv(const dbl_tensor* x, const float* w_dw, const float* b_dw, int dw_act, const void* w_pw_packed, const float* b_pw,
                          const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int epi, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w_dw && w_pw_packed, DBL_E_ARG, "dbl_dsconv: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y), DBL_E_ARG, "dbl_dsconv: x and y must be bf16 NHWC, 16-byte aligned");
    const int pad = (k - 1) / 2;
    const int ho = (x->h + 2 * pad - k) / stride + 1, wo = (x->w + 2 * pad - k) / stride + 1;
    DBL_REQUIRE(y->n == x->n && y->h == ho && y->w == wo, DBL_E_ARG, "dbl_dsconv: output shape mismatch");
    DBL_REQUIRE(((reinterpret_cast<uintptr_t>(w_dw) | reinterpret_cast<uintptr_t>(w_pw_packed)) & 15) == 0, DBL_E_ARG, "dbl_dsconv: weights must be 16-byte aligned");
    if (epi != DBL_EPI_NONE) DBL_REQUIRE(valid_tensor(r) && same_shape(r, y) && vec8_ok(r), DBL_E_ARG, "dbl_dsconv: residual mismatch");
    DBL_REQUIRE(x->data != y->data, DBL_E_ARG, "dbl_dsconv: x and y must not alias (tiles read their neighbours' halo)");
    cudaStream_t st = as_stream(stream);
    const dbl_tensor* rr = epi != DBL_EPI_NONE ? r : nullptr;
#define X(K_, S_, CI_, CO_)                                                                                            \
    if (k == K_ && stride == S_ && x->c == CI_ && y->c == CO_)                                                         \
        return dw_act ? launch_ds<K_, S_, CI_, CO_, 1>(x, w_dw, b_dw, w_pw_packed, b_pw, rr, y, epi, st)               \
                      : launch_ds<K_, S_, CI_, CO_, 0>(x, w_dw, b_dw, w_pw_packed, b_pw, rr, y, epi, st);
    DBL_DS_TABLE(X)
#undef X
    set_error("dbl_dsconv: no instantiation for k %d stride %d cin %d cout %d", k, stride, x->c, y->c);
    return DBL_E_UNSUPPORTED;
}
