// 3x3 implicit-GEMM convolution on tcgen05, second generation: every input pixel is fetched from L2/HBM ONCE per tile.
//
// The first-generation kernel (conv_tc.cu) issues one TMA box per filter tap, i.e. it re-reads each activation tile nine
// times through L2 and pays the TMA unit's per-row cost on rows as short as 32 bytes.  Here the producer warps copy the
// tile's halo (tile + 1-pixel border, zero-filled outside the image) into shared memory exactly once with 16-byte
// cp.async, in a PLANAR layout: plane g holds channels [8g, 8g+8) of all halo pixels, 16 bytes per pixel, pixels
// contiguous.  That is precisely tcgen05's canonical K-major NO-SWIZZLE operand layout (core matrix = 8 rows x 16 B):
//   row stride inside a core matrix 16 B  (= next pixel),  SBO = 128 B (next 8 pixels),  LBO = plane stride (next 8 ch)
// so the A operand of filter tap (dy,dx) is the SAME buffer with its start address advanced by (dy*Wl + dx) pixels --
// a descriptor change, no data movement.  The GEMM's M index runs over the tile's rows INCLUDING the two halo columns
// (their outputs are junk and never stored), which is what makes the tap shift a pure address offset.
// Stride-2 convs split the halo into four parity planes (even/odd row x column) so each tap is again unit-stride.
// 8-channel inputs (one plane) pair two TAPS into one K=16 step: LBO is then the pixel distance between the two taps.
//
// Weights: bf16 [Cout][9][Cin] in global memory; the producers scatter them into the matching planar layout
// ([tap][8-channel chunk][Cout rows][8]) either once per CTA (they stay resident when the whole filter fits) or per
// channel chunk through the same pipeline.  Accumulators: fp32 in TMEM, double buffered; epilogue as in conv_tc.cu.
#include <cstring>

#include "mma_util.cuh"

namespace dbl {

constexpr int C3_PROD_WARPS = 4;
constexpr int C3_EPI_WARPS = 8;
constexpr int C3_THREADS = 32 * (1 + C3_PROD_WARPS + C3_EPI_WARPS);  // warp 0 MMA, 1..4 producers, 5..12 epilogue
constexpr int C3_MAX_STAGES = 8;
constexpr int MAX_ACC = 8;
#ifndef DBL_C3_NO_FENCE
#define DBL_C3_NO_FENCE 0
#endif

struct C3Params {
    const __nv_bfloat16* x;
    const __nv_bfloat16* w;
    int xpitch;
    int n, h, wd, cin, ho, wo, cout, stride;
    int G, Gc, ncc;           // 8-channel planes: total, per pipeline unit, units per tile
    int tw, th, wl;           // output tile and local row length (tw + 2, or tw + 1 for stride 2)
    int tiles_x, tiles_y, m_tiles, n_tiles, npad, tmem_cols;
    int npar, hq, halo_px;    // parity planes (1 or 4), pixels per parity array, pixels filled per plane
    int rows_h, tail_px, gc_log2;  // halo rows per parity array, overrun pixels after them, log2(Gc)
    int ps, psb;              // plane strides in bytes: activations / weights
    int a_bytes, b_bytes;     // per stage
    int stages, b_resident, nacc;
    int off[9];               // pixel offset of each tap inside a plane (sorted ascending)
    int tap[9];               // which filter tap (ky*3+kx) sits at sorted position j
    // epilogue
    const float* bias;
    const __nv_bfloat16* r;
    void* y;
    int rpitch, ypitch, ydtype, epi, y_vec, r_vec;
};

__device__ __forceinline__ uint64_t make_noswz_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address           bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset (K) bits [16,30)
    d |= (uint64_t)(128 >> 4) << 32;                // stride byte offset (8 rows) bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version
    return d;                                       // layout_type 0 = no swizzle
}

template <int ACT>
__global__ void __launch_bounds__(C3_THREADS, 1) conv3x3_tc_kernel(const C3Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int stage_bytes = p.a_bytes + (p.b_resident ? 0 : p.b_bytes);
    uint8_t* sRing = smem;
    uint8_t* sBres = sRing + (size_t)p.stages * stage_bytes;  // resident weights (b_resident only)
    float* sBias = reinterpret_cast<float*>(sBres + (p.b_resident ? p.b_bytes : 0));
    const int bias_n = p.n_tiles * p.npad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + ((bias_n + 3) & ~3));
    uint64_t* full = bars;
    uint64_t* empty = bars + C3_MAX_STAGES;
    uint64_t* tfull = bars + 2 * C3_MAX_STAGES;
    uint64_t* tempty = tfull + MAX_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + MAX_ACC);
    uint64_t* sDesc = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // [<=36][2] descriptor templates (A, B) of one unit

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < bias_n; i += C3_THREADS) sBias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 32 * C3_PROD_WARPS);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < p.nacc; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], C3_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: dependents may start only once every CTA of this grid HOLDS its TMEM columns (a dependent CTA that allocated
    // first and then blocked in its own pdl_wait() would starve us); then wait for our own predecessor
    pdl_launch_dependents();
    pdl_wait();
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        // One thread issues every MMA, so its instruction count per MMA is the kernel's pace for thin layers.  All operands
        // are kept provably warp-uniform (loops outside the election, descriptors computed from kernel parameters, TMEM base
        // re-broadcast), so ptxas emits plain uniform-datapath arithmetic + UTCHMMA instead of an R2UR waterfall per MMA.
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const int pairs = p.Gc >> 1;
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1;
            mbar_wait(&tempty[acc], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + acc * (uint32_t)p.npad;
            for (int cc = 0; cc < p.ncc; ++cc) {
                mbar_wait(&full[stage], phase);
                if (!DBL_C3_NO_FENCE) fence_proxy_async();  // cp.async wrote through the generic proxy; tcgen05 reads through the async proxy
                tc_fence_after();
                const uint32_t a_base = smem_u32(sRing + (size_t)stage * stage_bytes);
                const uint32_t b_base = p.b_resident ? smem_u32(sBres) : a_base + p.a_bytes;
                if (elect_one()) {
                    if (p.G == 1) {  // 8 input channels: two filter taps per K=16 step, LBO = pixel distance between them
#pragma unroll 1
                        for (int k = 0; k < 5; ++k) {
                            const int o0 = p.off[2 * k], o1 = (2 * k + 1 < 9) ? p.off[2 * k + 1] : o0 + 1;
                            umma_bf16(d_tmem, make_noswz_desc(a_base + o0 * 16, (uint32_t)(o1 - o0) * 16),
                                      make_noswz_desc(b_base + (2 * k) * p.psb, (uint32_t)p.psb), idesc, (cc | k) ? 1u : 0u);
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 9; ++j) {
                            const uint32_t a_tap = a_base + p.off[j] * 16, b_tap = b_base + (uint32_t)(j * p.Gc) * p.psb;
                            for (int pp = 0; pp < pairs; ++pp)
                                umma_bf16(d_tmem, make_noswz_desc(a_tap + (2 * pp) * p.ps, (uint32_t)p.ps),
                                          make_noswz_desc(b_tap + (2 * pp) * p.psb, (uint32_t)p.psb), idesc, (cc | j | pp) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (cc == p.ncc - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == (uint32_t)p.stages) stage = 0, phase ^= 1;
            }
        }
    } else if (warp <= C3_PROD_WARPS) {
        // ===================== producers: halo + weights -> planar shared memory =====================
        const int pt = threadIdx.x - 32, NP = 32 * C3_PROD_WARPS;
        uint32_t stage = 0, phase = 0;
        bool first = true;
        const int nchunks_b = (p.G == 1) ? 10 : 9 * p.Gc;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int img = mt / per_img, t = mt % per_img;
            const int oy0 = (t / p.tiles_x) * p.th, ox0 = (t % p.tiles_x) * p.tw;
            const __nv_bfloat16* ximg = p.x + (size_t)img * p.h * p.wd * p.xpitch;
            for (int cc = 0; cc < p.ncc; ++cc) {
                if (lane == 0) mbar_wait(&empty[stage], phase ^ 1);
                __syncwarp();
                const uint32_t a_dst = smem_u32(sRing + (size_t)stage * stage_bytes);
                // ---- activations: walk the halo row by row; (row, item) advance incrementally, no per-item division
                {
                    const int row_items = p.wl << p.gc_log2;          // 16-byte items per halo row
                    const int rows = p.npar * p.rows_h;               // halo rows over all parity arrays
                    const int dq = NP / row_items, dr = NP - dq * row_items;
                    int hr = pt / row_items, t = pt - hr * row_items;  // this thread's first (row, item)
                    const uint32_t chan0 = (uint32_t)cc << p.gc_log2;
                    while (hr < rows) {
                        const int par = (p.npar == 1) ? 0 : hr / p.rows_h;
                        const int ly = hr - par * p.rows_h;
                        const int lx = t >> p.gc_log2, g = t & (p.Gc - 1);
                        int iy, ix;
                        if (p.npar == 1) {
                            iy = oy0 - 1 + ly, ix = ox0 - 1 + lx;
                        } else {
                            iy = 2 * (oy0 - 1 + ly) + (par >> 1), ix = 2 * (ox0 - 1 + lx) + (par & 1);
                        }
                        const bool ok = iy >= 0 && iy < p.h && ix >= 0 && ix < p.wd;
                        const __nv_bfloat16* src = ok ? ximg + ((size_t)iy * p.wd + ix) * p.xpitch + (chan0 + g) * 8 : p.x;
                        cp_async16(a_dst + g * p.ps + (par * p.hq + ly * p.wl + lx) * 16, src, ok);
                        t += dr, hr += dq;
                        if (t >= row_items) t -= row_items, ++hr;
                    }
                    // the few pixels past the last halo row that junk GEMM rows may touch: keep them finite (zero)
                    for (int i = pt; i < p.npar * p.tail_px * p.Gc; i += NP) {
                        const int g = i & (p.Gc - 1), q = i >> p.gc_log2;
                        const int par = q / p.tail_px, e = q - par * p.tail_px;
                        cp_async16(a_dst + g * p.ps + (par * p.hq + p.rows_h * p.wl + e) * 16, p.x, false);
                    }
                }
                // ---- weights
                if (!p.b_resident || first) {
                    const uint32_t b_dst = p.b_resident ? smem_u32(sBres) : a_dst + p.a_bytes;
                    const int n0 = nt * p.npad;
                    const int bitems = nchunks_b * p.npad;
                    for (int i = pt; i < bitems; i += NP) {
                        int g, row, j;
                        if (p.G == 1) {
                            row = i % p.npad, j = i / p.npad, g = 0;
                        } else {
                            g = i % p.Gc;
                            const int rest = i / p.Gc;
                            row = rest % p.npad, j = rest / p.npad;
                        }
                        const bool ok = (j < 9) && (n0 + row < p.cout);
                        const __nv_bfloat16* src = ok ? p.w + ((size_t)(n0 + row) * 9 + p.tap[j]) * p.cin + (cc * p.Gc + g) * 8 : p.w;
                        cp_async16(b_dst + (j * (p.G == 1 ? 1 : p.Gc) + g) * p.psb + row * 16, src, ok);
                    }
                }
                first = false;
                cp_async_arrive_noinc(&full[stage]);
                if (++stage == (uint32_t)p.stages) stage = 0, phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 1 - C3_PROD_WARPS;  // 0..7
        const int lg = warp & 3;                  // TMEM lane group of this warp
        const int chalf = ew >> 2;
        const int row = lg * 32 + lane;
        const int ty = row / p.wl, tx = row - ty * p.wl;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it % (uint32_t)p.nacc, aphase = (it / (uint32_t)p.nacc) & 1;
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int img = mt / per_img, t = mt % per_img;
            const int oy = (t / p.tiles_x) * p.th + ty, ox = (t % p.tiles_x) * p.tw + tx;
            long long opix = -1;
            if (ty < p.th && tx < p.tw && oy < p.ho && ox < p.wo) opix = ((long long)img * p.ho + oy) * p.wo + ox;
            mbar_wait(&tfull[acc], aphase);
            tc_fence_after();
            const int n0 = nt * p.npad;
            const int ncols = min(p.npad, ((p.cout - n0) + 15) & ~15);
            for (int c0 = chalf * 16; c0 < ncols; c0 += 16 * (C3_EPI_WARPS / 4)) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.npad + c0, v);
                tmem_ld_wait();
                if (opix < 0) continue;
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
                const bool full16 = cg + 16 <= p.cout;
                if (p.epi != DBL_EPI_NONE) {
                    const __nv_bfloat16* rp = p.r + (size_t)opix * p.rpitch + cg;
                    float rr[16];
                    if (full16 && p.r_vec) {
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
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// plane stride (in 16-byte units) >= need whose residue makes the quarter-warp cp.async stores bank-conflict free
static int plane_units(int need, int gc) {
    if (gc <= 1) return need;
    const int want = (8 / (gc > 8 ? 8 : gc)) % 8;
    int v = need;
    while (v % 8 != want) ++v;
    return v;
}

bool conv3x3_tc3_eligible(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil) {
    if (k != 3 || pad != 1 || dil != 1 || (stride != 1 && stride != 2)) return false;
    if (x->dtype != DBL_BF16 || x->c % 8 != 0 || x->pitch % 8 != 0 || (reinterpret_cast<uintptr_t>(x->data) & 15)) return false;
    if (x->c != 8 && x->c % 16 != 0) return false;
    if (stride == 2 && (x->h % 2 || x->w % 2)) return false;
    (void)y;
    return true;
}

// Heuristic used by DBL_IMPL_AUTO: the halo-reuse kernel wins when the whole filter stays resident in shared memory
// (cin <= 64, modest cout) and, for stride 2, when the channel count is small; otherwise streaming the weights through
// cp.async costs more than the TMA kernel's repeated activation reads (measured, tools/conv_bench.py).
bool conv3x3_tc3_preferred(const dbl_tensor* x, const dbl_tensor* y, int stride) {
    const int G = x->c / 8;
    if (G > 8 || y->c > 256) return false;
    const int npad = ((y->c + 15) / 16) * 16;
    const size_t b_bytes = (size_t)(G == 1 ? 10 : 9 * G) * (npad * 16 + 16);
    if (b_bytes > 100 * 1024) return false;
    return stride == 1 || x->c <= 16;
}

int conv3x3_tc3(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int stride, int act,
                int epi, cudaStream_t st) {
    const int num_sms = dev_info().sms, max_smem = dev_info().max_smem;
    static DevOnce once;  // shared-memory opt-in is per device
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) {
            set_error("conv3x3_tc3: cannot opt in to %d bytes of shared memory: %s", max_smem, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, DBL_E_ARG, "conv3x3_tc3: weights must be 16-byte aligned");
    C3Params p;
    memset(&p, 0, sizeof(p));
    p.x = reinterpret_cast<const __nv_bfloat16*>(x->data), p.w = reinterpret_cast<const __nv_bfloat16*>(w), p.xpitch = x->pitch;
    p.n = x->n, p.h = x->h, p.wd = x->w, p.cin = x->c, p.ho = y->h, p.wo = y->w, p.cout = y->c, p.stride = stride;
    p.G = x->c / 8;
    p.n_tiles = cdiv(y->c, 256);
    p.npad = ((cdiv(y->c, p.n_tiles) + 15) / 16) * 16;
    p.n_tiles = cdiv(y->c, p.npad);
    // accumulator ring in TMEM: as many buffers as fit (<= MAX_ACC); two are not enough to hide the MMA -> epilogue ->
    // MMA latency chain when tiles are small
    p.nacc = 512 / p.npad;
    if (p.nacc > MAX_ACC) p.nacc = MAX_ACC;
    int cols = 32;
    while (cols < p.nacc * p.npad) cols <<= 1;
    p.tmem_cols = cols;
    p.npar = stride == 1 ? 1 : 4;
    const int extra = stride == 1 ? 2 : 1;  // halo columns / rows beyond the tile
    // ---- tile: th rows x tw columns with th*(tw+extra) <= 128; minimise (tiles x (halo pixels + fixed cost))
    long long best = -1;
    for (int tw = 2; tw <= (p.wo < 126 ? p.wo : 126); ++tw) {
        const int wl = tw + extra;
        int th = 128 / wl;
        if (th > p.ho) th = p.ho;
        if (th < 1) continue;
        const long long tiles = (long long)cdiv(p.wo, tw) * cdiv(p.ho, th);
        const long long halo = (long long)p.npar * ((th + extra) * wl + extra);
        const long long cost = tiles * (halo + 96);
        if (best < 0 || cost < best) best = cost, p.tw = tw, p.th = th, p.wl = wl;
    }
    p.tiles_x = cdiv(p.wo, p.tw), p.tiles_y = cdiv(p.ho, p.th);
    p.m_tiles = p.n * p.tiles_x * p.tiles_y;
    p.hq = (p.th + extra) * p.wl + extra;
    p.halo_px = p.npar * p.hq;
    p.rows_h = p.th + extra, p.tail_px = extra;
    // ---- tap offsets, sorted ascending (the 8-channel path pairs neighbours in this order)
    int off[9], tap[9];
    for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t % 3;
        if (stride == 1) {
            off[t] = ky * p.wl + kx;
        } else {
            const int ph = ky == 1 ? 0 : 1, la = ky == 0 ? 0 : 1, pw = kx == 1 ? 0 : 1, lb = kx == 0 ? 0 : 1;
            off[t] = (ph * 2 + pw) * p.hq + la * p.wl + lb;
        }
        tap[t] = t;
    }
    for (int i = 0; i < 9; ++i)
        for (int j = i + 1; j < 9; ++j)
            if (off[j] < off[i]) {
                int a = off[i]; off[i] = off[j]; off[j] = a;
                a = tap[i]; tap[i] = tap[j]; tap[j] = a;
            }
    for (int t = 0; t < 9; ++t) p.off[t] = off[t], p.tap[t] = tap[t];
    const int need_px = (p.halo_px > 128 + off[8] + 2) ? p.halo_px : 128 + off[8] + 2;
    // ---- channel chunking and stages
    const size_t fixed = 256 + (((size_t)p.n_tiles * p.npad + 3) & ~(size_t)3) * 4 + (2 * C3_MAX_STAGES + 2 * MAX_ACC) * 8 + 16 + 36 * 2 * 8;
    int gc = 8;
    while (p.G % gc) gc >>= 1;  // largest power of two <= 8 dividing the plane count
    for (;; gc >>= 1) {
        DBL_REQUIRE(gc >= 1 && (gc >= 2 || p.G == 1), DBL_E_UNSUPPORTED, "conv3x3_tc3: cannot chunk %d channels", x->c);
        p.Gc = gc, p.ncc = p.G / gc;
        p.gc_log2 = gc == 8 ? 3 : (gc == 4 ? 2 : (gc == 2 ? 1 : 0));
        p.ps = plane_units(need_px, gc) * 16;
        p.psb = plane_units(p.npad, gc) * 16;
        p.a_bytes = (gc * p.ps + 127) & ~127;
        const int nplanes_b = (p.G == 1) ? 10 : 9 * gc;
        p.b_bytes = (nplanes_b * p.psb + 127) & ~127;
        p.b_resident = (p.ncc == 1 && p.n_tiles == 1 && (size_t)p.b_bytes + 3 * (size_t)p.a_bytes + fixed <= (size_t)max_smem) ? 1 : 0;
        const size_t per_stage = (size_t)p.a_bytes + (p.b_resident ? 0 : p.b_bytes);
        const size_t avail = (size_t)max_smem - fixed - (p.b_resident ? p.b_bytes : 0);
        int stages = (int)(avail / per_stage);
        if (stages > C3_MAX_STAGES) stages = C3_MAX_STAGES;
        if (stages >= 2 || gc == 1) {
            p.stages = stages;
            break;
        }
    }
    DBL_REQUIRE(p.stages >= 1, DBL_E_UNSUPPORTED, "conv3x3_tc3: tile does not fit shared memory");
    p.bias = bias;
    p.r = r ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr;
    p.rpitch = r ? r->pitch : 0;
    p.r_vec = r ? (int)(r->pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(r->data) & 15) == 0) : 0;
    p.y = y->data, p.ypitch = y->pitch, p.ydtype = y->dtype, p.epi = r ? epi : DBL_EPI_NONE;
    const int yelem = y->dtype == DBL_BF16 ? 2 : 4;
    p.y_vec = (int)(((size_t)y->pitch * yelem) % 16 == 0 && (reinterpret_cast<uintptr_t>(y->data) & 15) == 0);
    const size_t smem = fixed + (size_t)p.stages * (p.a_bytes + (p.b_resident ? 0 : p.b_bytes)) + (p.b_resident ? p.b_bytes : 0);
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < num_sms ? total : num_sms;
    if (act == DBL_ACT_SILU)
        launch_k(conv3x3_tc_kernel<DBL_ACT_SILU>, grid, C3_THREADS, smem, st, p);
    else if (act == DBL_ACT_GELU)
        launch_k(conv3x3_tc_kernel<DBL_ACT_GELU>, grid, C3_THREADS, smem, st, p);
    else
        launch_k(conv3x3_tc_kernel<DBL_ACT_NONE>, grid, C3_THREADS, smem, st, p);
    count_launch();
    return check_launch("dbl_conv2d(tcgen05 3x3)");
}

}  // namespace dbl
