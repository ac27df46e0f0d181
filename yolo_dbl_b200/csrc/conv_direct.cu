// CUDA-core convolutions: the fp32 NCHW -> bf16 NHWC stem, a generic direct convolution (used where the tensor-core
// implicit GEMM does not apply: Cin % 16 != 0, dilation, forced by the caller) and the depthwise convolution.
#include <cstdlib>

#include "tc_ptx.cuh"

namespace dbl {
int conv2d_tcgen05(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int k,
                   int stride, int pad, int act, int epi, cudaStream_t st);
bool conv2d_tcgen05_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad,
                             int dil);
int conv3x3_tc3(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int stride, int act,
                int epi, cudaStream_t st);
bool conv3x3_tc3_eligible(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil);
bool conv3x3_tc3_preferred(const dbl_tensor* x, const dbl_tensor* y, int stride);
int conv3x3_tc4(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int act, int epi, cudaStream_t st);
bool conv3x3_tc4_eligible(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil);
int conv3x3_tc5(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y, int act, int epi, cudaStream_t st);
bool conv3x3_tc5_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad, int dil, int act);
int conv3x3s2_tc(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* y, int act, cudaStream_t st);
bool conv3x3s2_tc_eligible(const dbl_tensor* x, const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int pad, int dil, int act);
bool conv3x3_tc4_preferred(const dbl_tensor* x, const dbl_tensor* y);

// ------------------------------------------------------------------------------------------------------------------
// stem: x f32 NCHW (n,3,h,w) -> y bf16 NHWC (n,ho,wo,COUT), 3x3, pad 1, SiLU
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float stem_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float stem_load(const uint8_t* p) { return (float)__ldg(p) * (1.0f / 255.0f); }

template <int COUT, typename TIN>
__global__ void __launch_bounds__(256) stem_conv_kernel(const TIN* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                        int n, int h, int wd, int ho, int wo, int stride, int ypitch) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float sw[27 * COUT];  // [ci][ky][kx][co]
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
        int co = i % COUT, j = i / COUT;  // j = ci*9 + ky*3 + kx ; source layout [co][ci][ky][kx]
        sw[i] = w[co * 27 + j];
    }
    if (threadIdx.x < COUT) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    if (ox >= wo) return;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
    const int iy0 = oy * stride - 1, ix0 = ox * stride - 1;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const TIN* xp = x + ((size_t)b * 3 + ci) * h * wd;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ix0 + kx;
                if (ix < 0 || ix >= wd) continue;
                const float v = stem_load(xp + (size_t)iy * wd + ix);
                const float4* wp = reinterpret_cast<const float4*>(sw + ((ci * 3 + ky) * 3 + kx) * COUT);
#pragma unroll
                for (int c4 = 0; c4 < COUT / 4; ++c4) {
                    const float4 ww = wp[c4];
                    acc[4 * c4 + 0] = fmaf(v, ww.x, acc[4 * c4 + 0]);
                    acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
                    acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
                    acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
                }
            }
        }
    }
    __nv_bfloat16* yp = y + (((size_t)b * ho + oy) * wo + ox) * ypitch;
#pragma unroll
    for (int c8 = 0; c8 < COUT / 8; ++c8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = silu_f(acc[c8 * 8 + i]);
        stg8(yp + c8 * 8, pack8(o));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// generic direct conv.  One thread = one output pixel x 8 output channels; weights of the 8-channel group staged in
// shared memory as f32 [tap*cin][8] so a warp reads them as broadcasts.
// ------------------------------------------------------------------------------------------------------------------
constexpr int DC_THREADS = 128;
constexpr int DC_CO = 8;

__global__ void __launch_bounds__(DC_THREADS) conv_direct_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
    const __nv_bfloat16* __restrict__ r, void* __restrict__ y, int n, int h, int wd, int cin, int xpitch, int ho, int wo,
    int cout, int ypitch, int ydtype, int rpitch, int k, int stride, int pad, int dil, int act, int epi, int cin_chunk) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float swt[];  // [k*k][cin_chunk][8]
    const int co0 = blockIdx.y * DC_CO;
    const long long total = (long long)n * ho * wo;
    const long long pix = (long long)blockIdx.x * DC_THREADS + threadIdx.x;
    const bool live = pix < total;
    int b = 0, oy = 0, ox = 0;
    if (live) {
        ox = (int)(pix % wo);
        oy = (int)((pix / wo) % ho);
        b = (int)(pix / ((long long)wo * ho));
    }
    float acc[DC_CO];
#pragma unroll
    for (int i = 0; i < DC_CO; ++i) acc[i] = (bias && co0 + i < cout) ? bias[co0 + i] : 0.f;
    const int kk = k * k;
    for (int c0 = 0; c0 < cin; c0 += cin_chunk) {
        const int cc = min(cin_chunk, cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < kk * cc * DC_CO; i += DC_THREADS) {
            const int co = i % DC_CO, ci = (i / DC_CO) % cc, tap = i / (DC_CO * cc);
            swt[i] = (co0 + co < cout) ? __bfloat162float(w[((size_t)(co0 + co) * kk + tap) * cin + c0 + ci]) : 0.f;
        }
        __syncthreads();
        if (!live) continue;
        for (int tap = 0; tap < kk; ++tap) {
            const int iy = oy * stride - pad + (tap / k) * dil;
            const int ix = ox * stride - pad + (tap % k) * dil;
            if (iy < 0 || iy >= h || ix < 0 || ix >= wd) continue;
            const __nv_bfloat16* xp = x + (((size_t)b * h + iy) * wd + ix) * xpitch + c0;
            const float* wp = swt + (size_t)tap * cc * DC_CO;
            for (int ci = 0; ci < cc; ci += 8) {
                float xv[8];
                unpack8(ldg8(xp + ci), xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wp + (ci + j) * DC_CO);
                    const float4 w1 = *reinterpret_cast<const float4*>(wp + (ci + j) * DC_CO + 4);
                    acc[0] = fmaf(xv[j], w0.x, acc[0]);
                    acc[1] = fmaf(xv[j], w0.y, acc[1]);
                    acc[2] = fmaf(xv[j], w0.z, acc[2]);
                    acc[3] = fmaf(xv[j], w0.w, acc[3]);
                    acc[4] = fmaf(xv[j], w1.x, acc[4]);
                    acc[5] = fmaf(xv[j], w1.y, acc[5]);
                    acc[6] = fmaf(xv[j], w1.z, acc[6]);
                    acc[7] = fmaf(xv[j], w1.w, acc[7]);
                }
            }
        }
    }
    if (!live) return;
    const size_t opix = ((size_t)b * ho + oy) * wo + ox;
#pragma unroll
    for (int i = 0; i < DC_CO; ++i) acc[i] = apply_act(acc[i], act);
    const bool full = co0 + DC_CO <= cout;
    if (epi != DBL_EPI_NONE) {
        const __nv_bfloat16* rp = r + opix * rpitch + co0;
        float rr[DC_CO];
        if (full && (rpitch % 8 == 0) && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
            unpack8(ld8(rp), rr);
        } else {
#pragma unroll
            for (int i = 0; i < DC_CO; ++i) rr[i] = (co0 + i < cout) ? __bfloat162float(rp[i]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < DC_CO; ++i) acc[i] = (epi == DBL_EPI_ADD) ? acc[i] + rr[i] : acc[i] * rr[i];
    }
    if (ydtype == DBL_BF16 && full && (ypitch % 8 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
        stg8(reinterpret_cast<__nv_bfloat16*>(y) + opix * ypitch + co0, pack8(acc));
    } else {
#pragma unroll
        for (int i = 0; i < DC_CO; ++i)
            if (co0 + i < cout) st_elem(y, ydtype, opix * ypitch + co0 + i, acc[i]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// depthwise conv: one thread = one output pixel x 8 channels (one 16-byte load per tap)
// ------------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) dwconv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n,
                                                     int h, int wd, int c, int xpitch, int ho, int wo, int ypitch,
                                                     int stride, int dil, int act) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long total = (long long)n * ho * wo * cg;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), b = (int)(pix / ((long long)wo * ho));
    const int pad = dil * (K - 1) / 2;
    float acc[8];
    if (bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
        acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w, acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    }
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * stride - pad + ky * dil;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int ix = ox * stride - pad + kx * dil;
            if (ix < 0 || ix >= wd) continue;
            float xv[8];
            unpack8(ldg8(x + (((size_t)b * h + iy) * wd + ix) * xpitch + g * 8), xv);
            const float* wp = w + (size_t)(ky * K + kx) * c + g * 8;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
            acc[0] = fmaf(xv[0], w0.x, acc[0]);
            acc[1] = fmaf(xv[1], w0.y, acc[1]);
            acc[2] = fmaf(xv[2], w0.z, acc[2]);
            acc[3] = fmaf(xv[3], w0.w, acc[3]);
            acc[4] = fmaf(xv[4], w1.x, acc[4]);
            acc[5] = fmaf(xv[5], w1.y, acc[5]);
            acc[6] = fmaf(xv[6], w1.z, acc[6]);
            acc[7] = fmaf(xv[7], w1.w, acc[7]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = apply_act(acc[i], act);
    stg8(y + (((size_t)b * ho + oy) * wo + ox) * ypitch + g * 8, pack8(acc));
}

// depthwise conv, register-blocked: one thread = R consecutive output pixels of one row x 8 channels.  Each input row
// contributes (R-1)*S + K 16-byte loads that feed all R outputs (K=7, R=8: 12 loads per output instead of 49), the
// filter lives in shared memory.  Dilation 1 only.
template <int K, int S, int R>
__global__ void __launch_bounds__(128) dwconv_strip_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n,
                                                           int h, int wd, int c, int xpitch, int ho, int wo, int ypitch, int act) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sw[];  // [K*K][c]
    for (int i = threadIdx.x; i < K * K * c; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int cg = c / 8;
    const int strips = (wo + R - 1) / R;
    const long long total = (long long)n * ho * strips * cg;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % cg);
    long long rest = idx / cg;
    const int sx = (int)(rest % strips);
    rest /= strips;
    const int oy = (int)(rest % ho), b = (int)(rest / ho);
    const int ox0 = sx * R;
    constexpr int PAD = (K - 1) / 2;
    constexpr int WIN = (R - 1) * S + K;
    float acc[R][8];
    {
        float b8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
            b8[0] = b0.x, b8[1] = b0.y, b8[2] = b0.z, b8[3] = b0.w, b8[4] = b1.x, b8[5] = b1.y, b8[6] = b1.z, b8[7] = b1.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[r][i] = b8[i];
    }
    const int ix0 = ox0 * S - PAD;
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S - PAD + ky;
        if (iy < 0 || iy >= h) continue;
        const __nv_bfloat16* xrow = x + ((size_t)b * h + iy) * wd * xpitch + g * 8;
        float wk[K][8];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const float4 w0 = *reinterpret_cast<const float4*>(sw + (ky * K + kx) * c + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + (ky * K + kx) * c + g * 8 + 4);
            wk[kx][0] = w0.x, wk[kx][1] = w0.y, wk[kx][2] = w0.z, wk[kx][3] = w0.w;
            wk[kx][4] = w1.x, wk[kx][5] = w1.y, wk[kx][6] = w1.z, wk[kx][7] = w1.w;
        }
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const int ix = ix0 + j;
            float xv[8];
            if (ix >= 0 && ix < wd) {
                unpack8(ldg8(xrow + (size_t)ix * xpitch), xv);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[i] = 0.f;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int kx = j - r * S;  // compile-time after unrolling
                if (kx >= 0 && kx < K) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(xv[i], wk[kx][i], acc[r][i]);
                }
            }
        }
    }
    __nv_bfloat16* yrow = y + (((size_t)b * ho + oy) * wo) * ypitch + g * 8;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (ox0 + r < wo) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[r][i] = apply_act(acc[r][i], act);
            stg8(yrow + (size_t)(ox0 + r) * ypitch, pack8(acc[r]));
        }
    }
}


// Tiled depthwise kernel, third generation (the first -- 8 channels x 4 rows per thread, scalar FMAs, 128 registers -- ran
// its issue slots 47 % busy at 14 resident warps per SM).  One CTA = one (image, spatial tile, channel block), input tile + halo staged once with
// 16-byte cp.async (zero fill = padding), filter block next to it as f32 -- but
//  * a thread owns FOUR channels (one 8-byte shared-memory load per tap row) and R = 8 (or 4) output rows: the K filter
//    taps of the current filter column are 28 registers instead of 56, the whole kernel fits ~80 registers, so 4-5 CTAs
//    (20-25 warps) are resident per SM;
//  * packed fp32 FMAs (fma.rn.f32x2, SASS FFMA2) over channel pairs: a bf16x2 word unpacks straight into an aligned
//    register pair {w << 16, w & 0xffff0000}, the filter is read as float2.  Half the FMA issue slots, which leaves room
//    for the LDS / unpack instructions next to the FMA pipe that is the real bound (49 FMA per output for 7x7);
//  * staging index math through float reciprocals instead of two integer divisions per 16-byte chunk.
// Numerics are those of the scalar form (IEEE fma per element, same summation order).
// G4 = 4-channel groups per CTA (even: the staging copy moves 16 bytes = two groups); threads = G4 * tw * ty.
template <int K, int S, int D, int R>
__global__ void __launch_bounds__(256, 3) dwconv_tile3_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int h, int wd,
                                                           int c, int xpitch, int ho, int wo, int ypitch, int act, int G4, int tw, int ty,
                                                           int tiles_x) {
    extern __shared__ __align__(16) uint8_t dw_smem[];
    pdl_launch_dependents();
    constexpr int PAD = D * (K - 1) / 2;
    constexpr int ROWS_T = (R - 1) * S + (K - 1) * D + 1;  // input rows one thread walks
    const int th = ty * R;
    const int ir = (th - 1) * S + (K - 1) * D + 1, ic = (tw - 1) * S + (K - 1) * D + 1;
    const int img = blockIdx.z, cb = blockIdx.y;
    // small run-time divisions through exact float reciprocals ((q + 0.5) / d is never within float error of an integer for
    // these magnitudes): the integer form costs ~25 instructions each, and this kernel's prologue was half of its instructions
    const int tile_y = __float2int_rd(((float)blockIdx.x + 0.5f) * (1.0f / (float)tiles_x));
    const int oy0 = tile_y * th, ox0 = ((int)blockIdx.x - tile_y * tiles_x) * tw;
    const int c0 = cb * G4 * 4;
    const int c_here = min(G4 * 4, c - c0);  // channels of this block (the last block may be short; a multiple of 8)
    float* sw = reinterpret_cast<float*>(dw_smem);                        // [K*K][G4*4]
    const uint32_t sx = smem_u32(dw_smem) + (uint32_t)(K * K * G4 * 16);  // [ir][ic][G4] x 8 bytes
    const int nt = blockDim.x;
    {
        const int sh = G4 == 4 ? 4 : 3;  // G4 is 2 or 4: G4 * 4 floats per tap
        for (int i = threadIdx.x; i < K * K * G4 * 4; i += nt) {
            const int tap = i >> sh, cc = i - (tap << sh);
            sw[i] = (cc < c_here) ? w[(size_t)tap * c + c0 + cc] : 0.f;
        }
    }
    pdl_wait();  // the filter block above is constant data; the activations are the previous kernel's output
    {
        const __nv_bfloat16* ximg = x + (size_t)img * h * wd * xpitch + c0;
        const int iy0 = oy0 * S - PAD, ix0 = ox0 * S - PAD;
        stage_tile16(sx, ximg, x, ir, ic, G4 >> 1, c_here >> 3, iy0, ix0, h, wd, xpitch, threadIdx.x, nt);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int g = threadIdx.x & (G4 - 1), rest = threadIdx.x >> (G4 == 4 ? 2 : 1);  // G4 is 2 or 4
    const int ys = __float2int_rd(((float)rest + 0.5f) * (1.0f / (float)tw)), xi = rest - ys * tw;
    if (ys >= ty || g * 4 >= c_here) return;
    const int ox = ox0 + xi, oyb = oy0 + ys * R;
    if (ox >= wo || oyb >= ho) return;
    unsigned long long acc[R][2];
    {
        ulonglong2 b2 = make_ulonglong2(0ull, 0ull);
        if (bias) b2 = __ldg(reinterpret_cast<const ulonglong2*>(bias + c0 + g * 4));
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][0] = b2.x, acc[r][1] = b2.y;
    }
    const uint32_t xbase = sx + (uint32_t)(((ys * R * S) * ic + xi * S) * G4 + g) * 8;
    // the filter-column loop stays ROLLED (fully unrolled, the first generation was instruction-fetch bound)
#pragma unroll 1
    for (int kx = 0; kx < K; ++kx) {
        unsigned long long wk[K][2];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(sw + (ky * K + kx) * G4 * 4 + g * 4);
            wk[ky][0] = w0.x, wk[ky][1] = w0.y;
        }
#pragma unroll
        for (int iy = 0; iy < ROWS_T; ++iy) {
            bool used = false;  // compile-time: does any of the R outputs read input row iy?
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = iy - r * S;
                if (d >= 0 && d % D == 0 && d / D < K) used = true;
            }
            if (!used) continue;
            uint32_t v0, v1;
            asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(xbase + (uint32_t)((iy * ic + kx * D) * G4) * 8));
            const unsigned long long x0 = bf16x2_to_f32x2(v0), x1 = bf16x2_to_f32x2(v1);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = iy - r * S;
                if (d >= 0 && d % D == 0 && d / D < K) {
                    ffma2(acc[r][0], x0, wk[d / D][0]);
                    ffma2(acc[r][1], x1, wk[d / D][1]);
                }
            }
        }
    }
    __nv_bfloat16* yp = y + (((size_t)img * ho + oyb) * wo + ox) * ypitch + c0 + g * 4;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (oyb + r < ho) {
            float f[4];
            f32x2_split(acc[r][0], f[0], f[1]);
            f32x2_split(acc[r][1], f[2], f[3]);
            if (act == DBL_ACT_SILU) {  // a (uniform) branch, not a select: DSConv's depthwise half has no activation
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = silu_fast(f[i]);
            }
            *reinterpret_cast<uint2*>(yp + (size_t)r * wo * ypitch) = make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3]));
        }
    }
}

// host-side tile choice for dwconv_tile3_kernel: enumerate (G4, tw, ty), keep 96..256 threads and <= 44 KB of shared
// memory (>= 5 resident CTAs), minimise  padded outputs x K^2  (FMA work incl. tile padding)  +  4 x staged pixels
// (halo amplification); 16 channels per CTA (32-byte = full-sector rows) preferred over 8.
static bool dw_pick_tile3(int k, int s, int d, int R, int c, int ho, int wo, int* G4_, int* tw_, int* ty_, size_t* smem_) {
    double best = 1e300;
    bool found = false;
    for (int G4 = 2; G4 <= 4; G4 += 2) {
        if (G4 * 4 > c && G4 > 2) continue;  // a short last block is fine, a block wider than the tensor is not
        for (int tw = 4; tw <= (wo < 64 ? wo : 64); ++tw)
            for (int ty = 1; ty <= 8; ++ty) {
                const int threads = G4 * tw * ty;
                if (threads < 96 || threads > 256) continue;
                const int th = ty * R;
                const int ir = (th - 1) * s + (k - 1) * d + 1, ic = (tw - 1) * s + (k - 1) * d + 1;
                const size_t smem = (size_t)k * k * G4 * 16 + (size_t)ir * ic * G4 * 8;
                if (smem > 44 * 1024) continue;
                const int tx_n = cdiv(wo, tw), ty_n = cdiv(ho, th);
                const double padded = (double)tx_n * tw * ty_n * th, staged = (double)tx_n * ty_n * ir * ic;
                double cost = padded * k * k + 4.0 * staged * (G4 == 2 ? 1.5 : 1.0);
                cost *= 1.0 + 0.02 * (256 - threads) / 256.0;  // mild preference for fuller CTAs
                if (cost < best) best = cost, found = true, *G4_ = G4, *tw_ = tw, *ty_ = ty, *smem_ = smem;
            }
    }
    return found;
}

}  // namespace dbl

using namespace dbl;

template <typename TIN>
static int stem_launch(const TIN* x, int n, int h, int w_, const float* w, const float* bias, const dbl_tensor* y, int stride,
                       dbl_stream_t stream) {
    DBL_REQUIRE(x && w && valid_tensor(y), DBL_E_ARG, "dbl_stem_conv: null/invalid argument");
    DBL_REQUIRE(stride == 1 || stride == 2, DBL_E_ARG, "dbl_stem_conv: stride %d", stride);
    const int ho = (h + 2 - 3) / stride + 1, wo = (w_ + 2 - 3) / stride + 1;
    DBL_REQUIRE(y->n == n && y->h == ho && y->w == wo, DBL_E_ARG, "dbl_stem_conv: output shape mismatch");
    DBL_REQUIRE(vec8_ok(y), DBL_E_ARG, "dbl_stem_conv: y must be bf16 with 16-byte aligned channel groups");
    dim3 grid(cdiv(wo, 256), ho, n), block(256);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
    cudaStream_t st = as_stream(stream);
    switch (y->c) {
        case 8: launch_k(stem_conv_kernel<8, TIN>, grid, block, 0, st, x, w, bias, yp, n, h, w_, ho, wo, stride, y->pitch); break;
        case 16: launch_k(stem_conv_kernel<16, TIN>, grid, block, 0, st, x, w, bias, yp, n, h, w_, ho, wo, stride, y->pitch); break;
        case 32: launch_k(stem_conv_kernel<32, TIN>, grid, block, 0, st, x, w, bias, yp, n, h, w_, ho, wo, stride, y->pitch); break;
        case 64: launch_k(stem_conv_kernel<64, TIN>, grid, block, 0, st, x, w, bias, yp, n, h, w_, ho, wo, stride, y->pitch); break;   // yolov13l stem
        case 96: launch_k(stem_conv_kernel<96, TIN>, grid, block, 0, st, x, w, bias, yp, n, h, w_, ho, wo, stride, y->pitch); break;   // yolov13x stem
        default: DBL_REQUIRE(false, DBL_E_UNSUPPORTED, "dbl_stem_conv: cout %d not in {8,16,32,64,96}", y->c);
    }
    count_launch();
    return check_launch("dbl_stem_conv");
}

extern "C" int dbl_stem_conv(const float* x, int n, int h, int w_, const float* w, const float* bias, const dbl_tensor* y,
                             int stride, dbl_stream_t stream) {
    return stem_launch<float>(x, n, h, w_, w, bias, y, stride, stream);
}

extern "C" int dbl_stem_conv_u8(const uint8_t* x, int n, int h, int w_, const float* w, const float* bias, const dbl_tensor* y,
                                int stride, dbl_stream_t stream) {
    return stem_launch<uint8_t>(x, n, h, w_, w, bias, y, stride, stream);
}

extern "C" int dbl_conv2d(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y,
                          int k, int stride, int pad, int dil, int act, int epi, int impl, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w, DBL_E_ARG, "dbl_conv2d: null/invalid tensor");
    DBL_REQUIRE(x->dtype == DBL_BF16, DBL_E_ARG, "dbl_conv2d: x must be bf16");
    DBL_REQUIRE(k >= 1 && k <= 7 && stride >= 1 && dil >= 1 && pad >= 0, DBL_E_ARG, "dbl_conv2d: bad geometry");
    DBL_REQUIRE(act >= 0 && act <= 2 && epi >= 0 && epi <= 2, DBL_E_ARG, "dbl_conv2d: bad act/epi");
    const int ho = (x->h + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    const int wo = (x->w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    DBL_REQUIRE(y->n == x->n && y->h == ho && y->w == wo, DBL_E_ARG, "dbl_conv2d: output shape (%d,%d,%d) != expected (%d,%d,%d)",
                y->n, y->h, y->w, x->n, ho, wo);
    if (epi != DBL_EPI_NONE) {
        DBL_REQUIRE(valid_tensor(r) && same_shape(r, y) && r->dtype == DBL_BF16, DBL_E_ARG, "dbl_conv2d: residual mismatch");
    }
    cudaStream_t st = as_stream(stream);
    const bool tc_ok = conv2d_tcgen05_eligible(x, r, y, k, stride, pad, dil);
    const bool tc3_ok = conv3x3_tc3_eligible(x, y, k, stride, pad, dil);
    const bool tc4_ok = conv3x3_tc4_eligible(x, y, k, stride, pad, dil);
    if (impl == DBL_IMPL_AUTO && conv3x3s2_tc_eligible(x, epi ? r : nullptr, y, k, stride, pad, dil, act)) return conv3x3s2_tc(x, w, bias, y, act, st);
    if (impl == DBL_IMPL_AUTO && conv3x3_tc5_eligible(x, epi ? r : nullptr, y, k, stride, pad, dil, act)) return conv3x3_tc5(x, w, bias, epi ? r : nullptr, y, act, epi, st);
    if (impl == DBL_IMPL_TCGEN05_MULTI) DBL_REQUIRE(tc4_ok, DBL_E_UNSUPPORTED, "dbl_conv2d: shape not eligible for the multi-accumulator tcgen05 path");
    if (tc4_ok && (impl == DBL_IMPL_TCGEN05_MULTI || ((impl == DBL_IMPL_AUTO || impl == DBL_IMPL_TCGEN05) && conv3x3_tc4_preferred(x, y))))
        return conv3x3_tc4(x, w, bias, epi ? r : nullptr, y, act, epi, st);
    if (impl == DBL_IMPL_TCGEN05) DBL_REQUIRE(tc_ok || tc3_ok, DBL_E_UNSUPPORTED, "dbl_conv2d: shape not eligible for a tcgen05 path");
    if (impl == DBL_IMPL_TCGEN05_TMA) DBL_REQUIRE(tc_ok, DBL_E_UNSUPPORTED, "dbl_conv2d: shape not eligible for the TMA tcgen05 path");
    if (tc3_ok && (impl == DBL_IMPL_TCGEN05 || (impl == DBL_IMPL_AUTO && (!tc_ok || conv3x3_tc3_preferred(x, y, stride)))))
        return conv3x3_tc3(x, w, bias, epi ? r : nullptr, y, stride, act, epi, st);
    if (impl == DBL_IMPL_TCGEN05 || impl == DBL_IMPL_TCGEN05_TMA || (impl == DBL_IMPL_AUTO && tc_ok))
        return conv2d_tcgen05(x, w, bias, epi ? r : nullptr, y, k, stride, pad, act, epi, st);

    DBL_REQUIRE(vec8_ok(x), DBL_E_UNSUPPORTED, "dbl_conv2d(direct): cin/pitch must be multiples of 8 and 16-byte aligned");
    int chunk = (48 * 1024) / (k * k * DC_CO * 4);
    chunk = (chunk / 8) * 8;
    if (chunk > x->c) chunk = x->c;
    DBL_REQUIRE(chunk >= 8, DBL_E_UNSUPPORTED, "dbl_conv2d(direct): kernel too large");
    const size_t smem = (size_t)k * k * chunk * DC_CO * sizeof(float);
    const long long total = (long long)x->n * ho * wo;
    dim3 grid((unsigned)((total + DC_THREADS - 1) / DC_THREADS), cdiv(y->c, DC_CO));
    launch_k(conv_direct_kernel, grid, DC_THREADS, smem, st, 
        reinterpret_cast<const __nv_bfloat16*>(x->data), reinterpret_cast<const __nv_bfloat16*>(w), bias,
        epi ? reinterpret_cast<const __nv_bfloat16*>(r->data) : nullptr, y->data, x->n, x->h, x->w, x->c, x->pitch, ho, wo,
        y->c, y->pitch, y->dtype, epi ? r->pitch : 0, k, stride, pad, dil, act, epi, chunk);
    count_launch();
    return check_launch("dbl_conv2d(direct)");
}

/* which kernel dbl_conv2d would run for this problem: 0 direct (CUDA cores), 1 conv_tc (TMA per tap / 1x1 GEMM), 3 conv_tc3 (halo reuse,
 * resident weights), 4 conv_tc4 (halo reuse, multi-accumulator), 5 conv_s2_tc (thin stride-2, pixel-pair planes), 6 conv_tc5 (stride-1, whole K in one tile), -1 invalid.  Lets callers attribute time to the right kernel. */
extern "C" int dbl_conv2d_which(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil, int impl) {
    if (!valid_tensor(x) || !valid_tensor(y)) return -1;
    const bool tc_ok = conv2d_tcgen05_eligible(x, nullptr, y, k, stride, pad, dil);
    const bool tc3_ok = conv3x3_tc3_eligible(x, y, k, stride, pad, dil);
    const bool tc4_ok = conv3x3_tc4_eligible(x, y, k, stride, pad, dil);
    if (impl == DBL_IMPL_AUTO && conv3x3s2_tc_eligible(x, nullptr, y, k, stride, pad, dil, DBL_ACT_SILU)) return 5;
    if (impl == DBL_IMPL_AUTO && conv3x3_tc5_eligible(x, nullptr, y, k, stride, pad, dil, DBL_ACT_SILU)) return 6;
    if (tc4_ok && (impl == DBL_IMPL_TCGEN05_MULTI || ((impl == DBL_IMPL_AUTO || impl == DBL_IMPL_TCGEN05) && conv3x3_tc4_preferred(x, y)))) return 4;
    if (tc3_ok && (impl == DBL_IMPL_TCGEN05 || (impl == DBL_IMPL_AUTO && (!tc_ok || conv3x3_tc3_preferred(x, y, stride))))) return 3;
    if ((impl == DBL_IMPL_TCGEN05 || impl == DBL_IMPL_TCGEN05_TMA || impl == DBL_IMPL_AUTO) && tc_ok) return 1;
    return 0;
}

extern "C" int dbl_dwconv2d(const dbl_tensor* x, const float* w, const float* bias, const dbl_tensor* y, int k, int stride,
                            int dil, int act, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && w, DBL_E_ARG, "dbl_dwconv2d: null/invalid tensor");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y) && x->c == y->c, DBL_E_ARG, "dbl_dwconv2d: needs bf16, c %% 8 == 0, aligned, cin == cout");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0), DBL_E_ARG,
                "dbl_dwconv2d: weights must be 16-byte aligned");
    const int pad = dil * (k - 1) / 2;
    const int ho = (x->h + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    const int wo = (x->w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    DBL_REQUIRE(y->n == x->n && y->h == ho && y->w == wo, DBL_E_ARG, "dbl_dwconv2d: output shape mismatch");
    const long long total = (long long)x->n * ho * wo * (x->c / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    auto* xp = reinterpret_cast<const __nv_bfloat16*>(x->data);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
    cudaStream_t st = as_stream(stream);
    {
        // shared-memory tiled kernel for every (k, stride, dilation) the graph uses
        typedef void (*DwKern)(const __nv_bfloat16*, const float*, const float*, __nv_bfloat16*, int, int, int, int, int, int, int, int, int,
                               int, int, int);
        DwKern kern = nullptr;
        const int R = (ho % 8 == 0) ? 8 : 4;
#define DBL_DW3(KK, SS, DD) (R == 8 ? (DwKern)dwconv_tile3_kernel<KK, SS, DD, 8> : (DwKern)dwconv_tile3_kernel<KK, SS, DD, 4>)
        if (k == 3 && stride == 1 && dil == 1) kern = DBL_DW3(3, 1, 1);
        else if (k == 3 && stride == 2 && dil == 1) kern = DBL_DW3(3, 2, 1);
        else if (k == 5 && stride == 1 && dil == 1) kern = DBL_DW3(5, 1, 1);
        else if (k == 7 && stride == 1 && dil == 1) kern = DBL_DW3(7, 1, 1);
        else if (k == 7 && stride == 1 && dil == 3) kern = DBL_DW3(7, 1, 3);
#undef DBL_DW3
        int G4 = 0, tw = 0, ty = 0;
        size_t smem = 0;
        bool have_tile = kern && act != DBL_ACT_GELU && x->n <= 65535 && dw_pick_tile3(k, stride, dil, R, x->c, ho, wo, &G4, &tw, &ty, &smem);
        if (have_tile) {
            static const char* force = getenv("DBL_DW_TILE");  // tuning aid: "G4,tw,ty"
            int fg, ftw, fty;
            if (force && sscanf(force, "%d,%d,%d", &fg, &ftw, &fty) == 3 && fg * ftw * fty <= 256 && fg % 2 == 0) {
                const int ir = (fty * R - 1) * stride + (k - 1) * dil + 1, ic = (ftw - 1) * stride + (k - 1) * dil + 1;
                const size_t sm = (size_t)k * k * fg * 16 + (size_t)ir * ic * fg * 8;
                if (sm <= 100 * 1024) G4 = fg, tw = ftw, ty = fty, smem = sm;
            }
            static const char* dbg = getenv("DBL_DW_DEBUG");
            if (dbg) fprintf(stderr, "dw k%d s%d d%d c%d %dx%d: R=%d G4=%d tw=%d ty=%d smem=%zu\n", k, stride, dil, x->c, ho, wo, R, G4, tw, ty, smem);
        }
        if (have_tile) {
            if (func_config(reinterpret_cast<const void*>(kern)) < 0) {
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
                    set_error("dbl_dwconv2d: cannot opt in to 100 KB of shared memory: %s", cudaGetErrorString(cudaGetLastError()));
                    return DBL_E_CUDA;
                }
                func_set_config(reinterpret_cast<const void*>(kern), 1);
            }
            const int tiles_x = cdiv(wo, tw), tiles_y = cdiv(ho, ty * R);
            dim3 grid(tiles_x * tiles_y, cdiv(x->c, G4 * 4), x->n);
            launch_k(kern, grid, G4 * tw * ty, smem, st, xp, w, bias, yp, x->h, x->w, x->c, x->pitch, ho, wo, y->pitch, act, G4, tw, ty, tiles_x);
            count_launch();
            return check_launch("dbl_dwconv2d(tile)");
        }
    }
    if (dil == 1 && (stride == 1 || stride == 2) && (size_t)k * k * x->c * 4 <= 48 * 1024) {
        // register-blocked strips along W
        const int R = 4;
        const long long tot = (long long)x->n * ho * ((wo + R - 1) / R) * (x->c / 8);
        const unsigned gs = (unsigned)((tot + 127) / 128);
        const size_t sm = (size_t)k * k * x->c * 4;
#define DBL_DWS(K, S) launch_k(dwconv_strip_kernel<K, S, 4>, gs, 128, sm, st, xp, w, bias, yp, x->n, x->h, x->w, x->c, x->pitch, ho, wo, y->pitch, act)
        bool done = true;
        if (k == 3 && stride == 1) DBL_DWS(3, 1);
        else if (k == 3 && stride == 2) DBL_DWS(3, 2);
        else if (k == 5 && stride == 1) DBL_DWS(5, 1);
        else if (k == 7 && stride == 1) DBL_DWS(7, 1);
        else done = false;
#undef DBL_DWS
        if (done) {
            count_launch();
            return check_launch("dbl_dwconv2d(strip)");
        }
    }
#define DBL_DW(K) launch_k(dwconv_kernel<K>, grid, 256, 0, st, xp, w, bias, yp, x->n, x->h, x->w, x->c, x->pitch, ho, wo, y->pitch, stride, dil, act)
    switch (k) {
        case 3: DBL_DW(3); break;
        case 5: DBL_DW(5); break;
        case 7: DBL_DW(7); break;
        default: DBL_REQUIRE(false, DBL_E_UNSUPPORTED, "dbl_dwconv2d: k=%d not in {3,5,7}", k);
    }
#undef DBL_DW
    count_launch();
    return check_launch("dbl_dwconv2d");
}
