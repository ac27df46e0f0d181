// Bandwidth-bound glue kernels: pooling, nearest upsample, FullPAD gated add, copies, DySample, LSK squeeze / gate.
// All operate on bf16 NHWC views with 16-byte (8-channel) vector accesses.
#include "dbl_common.cuh"

namespace dbl {

// mode 0: copy, 1: avgpool 2x2, 2: nearest upsample x2
template <int MODE>
__global__ void __launch_bounds__(256) resample_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       int n, int ho, int wo, int c, int xpitch, int ypitch, int hi, int wi) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long total = (long long)n * ho * wo * cg;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), b = (int)(pix / ((long long)wo * ho));
    bf16x8 out;
    if (MODE == 0) {
        out = ldg8(x + (((size_t)b * hi + oy) * wi + ox) * xpitch + g * 8);
    } else if (MODE == 2) {
        out = ldg8(x + (((size_t)b * hi + (oy >> 1)) * wi + (ox >> 1)) * xpitch + g * 8);
    } else {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v[8];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                unpack8(ldg8(x + (((size_t)b * hi + 2 * oy + dy) * wi + 2 * ox + dx) * xpitch + g * 8), v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += v[i];
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= 0.25f;
        out = pack8(acc);
    }
    stg8(y + (((size_t)b * ho + oy) * wo + ox) * ypitch + g * 8, out);
}

__global__ void __launch_bounds__(256) gated_add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                        float gate, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y2,
                                                        long long npix, int c, int apitch, int bpitch, int ypitch, int y2pitch) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    float av[8], bv[8];
    unpack8(ldg8(a + pix * apitch + g * 8), av);
    unpack8(ldg8(b + pix * bpitch + g * 8), bv);
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = fmaf(gate, bv[i], av[i]);
    const bf16x8 o = pack8(av);
    stg8(y + pix * ypitch + g * 8, o);
    if (y2) stg8(y2 + pix * y2pitch + g * 8, o);
}

// y = a + gamma[c] * b with a per-channel scale (A2C2f's learnable residual, reference block.py:1403-1404)
__global__ void __launch_bounds__(256) channel_scale_add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                                const float* __restrict__ gamma, __nv_bfloat16* __restrict__ y, long long npix, int c,
                                                                int apitch, int bpitch, int ypitch) {
    pdl_launch_dependents();
    pdl_wait();
    const int cg = c / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    float av[8], bv[8];
    unpack8(ldg8(a + pix * apitch + g * 8), av);
    unpack8(ldg8(b + pix * bpitch + g * 8), bv);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + g * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + g * 8 + 4));
    av[0] = fmaf(g0.x, bv[0], av[0]), av[1] = fmaf(g0.y, bv[1], av[1]), av[2] = fmaf(g0.z, bv[2], av[2]), av[3] = fmaf(g0.w, bv[3], av[3]);
    av[4] = fmaf(g1.x, bv[4], av[4]), av[5] = fmaf(g1.y, bv[5], av[5]), av[6] = fmaf(g1.z, bv[6], av[6]), av[7] = fmaf(g1.w, bv[7], av[7]);
    stg8(y + pix * ypitch + g * 8, pack8(av));
}

// DySample: output pixel (2h+i, 2w+j) of channel group g samples the input of that group bilinearly at
// (w + off_x, h + off_y), off = conv_out*0.25 + init_pos, border clamped (grid_sample align_corners=False).
__global__ void __launch_bounds__(256) dysample_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ off,
                                                       __nv_bfloat16* __restrict__ y, int n, int h, int w, int c, int xpitch,
                                                       int opitch, int ypitch, int groups) {
    pdl_launch_dependents();
    pdl_wait();
    // grid (ceil(wo*cg/256), ho, n): the row and the image come from the block index, so the per-thread index math is one
    // 32-bit division (ncu: the flat 64-bit version of this kernel was ALU-bound on its div/mod chains)
    const int cg = c / 8;
    const int ho = 2 * h, wo = 2 * w;
    const int tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= wo * cg) return;
    const int ox = tix / cg, cgi = tix - ox * cg;
    const int oy = blockIdx.y, b = blockIdx.z;
    const int hh = oy >> 1, ww = ox >> 1, i = oy & 1, j = ox & 1;
    const int g = (cgi * 8) / (c / groups);
    const int m = g * 4 + i * 2 + j;
    const float* op = off + (((size_t)b * h + hh) * w + ww) * opitch;
    const float offx = __ldg(op + m) * 0.25f + (j ? 0.25f : -0.25f);
    const float offy = __ldg(op + groups * 4 + m) * 0.25f + (i ? 0.25f : -0.25f);
    float fx = fminf(fmaxf((float)ww + offx, 0.f), (float)(w - 1));
    float fy = fminf(fmaxf((float)hh + offy, 0.f), (float)(h - 1));
    const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    const float tx = fx - (float)x0, ty = fy - (float)y0;
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
    const __nv_bfloat16* base = x + (size_t)b * h * w * xpitch + cgi * 8;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    unpack8(ldg8(base + ((size_t)y0 * w + x0) * xpitch), v00);
    unpack8(ldg8(base + ((size_t)y0 * w + x1) * xpitch), v01);
    unpack8(ldg8(base + ((size_t)y1 * w + x0) * xpitch), v10);
    unpack8(ldg8(base + ((size_t)y1 * w + x1) * xpitch), v11);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = v00[q] * w00 + v01[q] * w01 + v10[q] * w10 + v11[q] * w11;
    stg8(y + (((size_t)b * ho + oy) * wo + ox) * ypitch + cgi * 8, pack8(o));
}

// DySample, 16 channels per thread (two consecutive 8-channel chunks of the same sampling group = one full 32-byte sector per
// access): the sampling coordinates and the four bilinear weights are computed once per 16 channels, and the blend runs on packed
// fp32 pairs (fma.rn.f32x2) -- the 8-channel form above was instruction bound (ncu: issue slots 77 % busy, ~120 instructions per
// 16 output bytes).  Same arithmetic per element: ((v00*w00 + v01*w01) + v10*w10) + v11*w11 in fp32.
__device__ __forceinline__ unsigned long long f32x2_bcast(float v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__global__ void __launch_bounds__(256) dysample16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ off,
                                                         __nv_bfloat16* __restrict__ y, int n, int h, int w, int c, int xpitch,
                                                         int opitch, int ypitch, int groups) {
    pdl_launch_dependents();
    pdl_wait();
    const int c16 = c / 16;
    const int ho = 2 * h, wo = 2 * w;
    const int tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= wo * c16) return;
    const int ox = tix / c16, ci = tix - ox * c16;
    const int oy = blockIdx.y, b = blockIdx.z;
    const int hh = oy >> 1, ww = ox >> 1, i = oy & 1, j = ox & 1;
    const int g = (ci * 16) / (c / groups);
    const int m = g * 4 + i * 2 + j;
    const float* op = off + (((size_t)b * h + hh) * w + ww) * opitch;
    const float offx = __ldg(op + m) * 0.25f + (j ? 0.25f : -0.25f);
    const float offy = __ldg(op + groups * 4 + m) * 0.25f + (i ? 0.25f : -0.25f);
    const float fx = fminf(fmaxf((float)ww + offx, 0.f), (float)(w - 1));
    const float fy = fminf(fmaxf((float)hh + offy, 0.f), (float)(h - 1));
    const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    const float tx = fx - (float)x0, ty = fy - (float)y0;
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const unsigned long long w00 = f32x2_bcast((1.f - tx) * (1.f - ty)), w01 = f32x2_bcast(tx * (1.f - ty)),
                             w10 = f32x2_bcast((1.f - tx) * ty), w11 = f32x2_bcast(tx * ty);
    const __nv_bfloat16* base = x + (size_t)b * h * w * xpitch + ci * 16;
    const __nv_bfloat16* p00 = base + ((size_t)y0 * w + x0) * xpitch;
    const __nv_bfloat16* p01 = base + ((size_t)y0 * w + x1) * xpitch;
    const __nv_bfloat16* p10 = base + ((size_t)y1 * w + x0) * xpitch;
    const __nv_bfloat16* p11 = base + ((size_t)y1 * w + x1) * xpitch;
    uint4 a[2], bq[2], cq[2], d[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) a[q] = ldg8(p00 + 8 * q).u, bq[q] = ldg8(p01 + 8 * q).u, cq[q] = ldg8(p10 + 8 * q).u, d[q] = ldg8(p11 + 8 * q).u;
    __nv_bfloat16* yp = y + (((size_t)b * ho + oy) * wo + ox) * ypitch + ci * 16;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t va[4] = {a[q].x, a[q].y, a[q].z, a[q].w}, vb[4] = {bq[q].x, bq[q].y, bq[q].z, bq[q].w};
        const uint32_t vc[4] = {cq[q].x, cq[q].y, cq[q].z, cq[q].w}, vd[4] = {d[q].x, d[q].y, d[q].z, d[q].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            unsigned long long acc = fmul2(bf16x2_to_f32x2(va[e]), w00);
            ffma2(acc, bf16x2_to_f32x2(vb[e]), w01);
            ffma2(acc, bf16x2_to_f32x2(vc[e]), w10);
            ffma2(acc, bf16x2_to_f32x2(vd[e]), w11);
            float lo, hi;
            f32x2_split(acc, lo, hi);
            o[e] = pack2(lo, hi);
        }
        *reinterpret_cast<uint4*>(yp + 8 * q) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// LSK squeeze: one warp per pixel -> (mean, max) over all channels of [attn1|attn2]
__global__ void __launch_bounds__(256) lsk_squeeze_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ agg,
                                                          long long npix, int c, int apitch, int gpitch) {
    pdl_launch_dependents();
    pdl_wait();
    const long long pix = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (pix >= npix) return;
    float s = 0.f, mx = -INFINITY, v[8];
    for (int c0 = lane * 8; c0 < c; c0 += 256) {
        unpack8(ldg8(a + pix * apitch + c0), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += v[i];
            mx = fmaxf(mx, v[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        agg[pix * gpitch + 0] = s / (float)c;
        agg[pix * gpitch + 1] = mx;
    }
}

// LSK gate: sig = sigmoid(conv7x7(agg)), y = attn1*sig0 + attn2*sig1.  One warp per pixel.
__global__ void __launch_bounds__(256) lsk_gate_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ agg,
                                                       const float* __restrict__ wsq, const float* __restrict__ bsq,
                                                       __nv_bfloat16* __restrict__ y, int n, int h, int w, int half, int apitch,
                                                       int gpitch, int ypitch) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sw[2 * 2 * 49];
    for (int i = threadIdx.x; i < 196; i += blockDim.x) sw[i] = wsq[i];
    __syncthreads();
    const long long npix = (long long)n * h * w;
    const long long pix = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (pix >= npix) return;
    const int px = (int)(pix % w), py = (int)((pix / w) % h);
    const long long img0 = pix - ((long long)py * w + px);
    float s0 = 0.f, s1 = 0.f;
    for (int t = lane; t < 98; t += 32) {  // t = ci*49 + ky*7 + kx
        const int ci = t / 49, ky = (t % 49) / 7, kx = t % 7;
        const int iy = py + ky - 3, ix = px + kx - 3;
        if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
        const float v = __ldg(agg + (img0 + (long long)iy * w + ix) * gpitch + ci);
        s0 = fmaf(v, sw[t], s0);
        s1 = fmaf(v, sw[98 + t], s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float g0 = sigmoid_f(s0 + bsq[0]), g1 = sigmoid_f(s1 + bsq[1]);
    float v1[8], v2[8];
    for (int c0 = lane * 8; c0 < half; c0 += 256) {
        unpack8(ldg8(a + pix * apitch + c0), v1);
        unpack8(ldg8(a + pix * apitch + half + c0), v2);
#pragma unroll
        for (int i = 0; i < 8; ++i) v1[i] = v1[i] * g0 + v2[i] * g1;
        stg8(y + pix * ypitch + c0, pack8(v1));
    }
}

static int launch_resample(int mode, const dbl_tensor* x, const dbl_tensor* y, cudaStream_t st, const char* name) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && vec8_ok(x) && vec8_ok(y) && x->c == y->c && x->n == y->n, DBL_E_ARG,
                "%s: needs bf16 views with c %% 8 == 0 and equal n, c", name);
    const int eh = mode == 1 ? x->h / 2 : (mode == 2 ? x->h * 2 : x->h), ew = mode == 1 ? x->w / 2 : (mode == 2 ? x->w * 2 : x->w);
    DBL_REQUIRE(y->h == eh && y->w == ew, DBL_E_ARG, "%s: output shape mismatch", name);
    const long long total = (long long)y->n * y->h * y->w * (y->c / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    auto* xp = reinterpret_cast<const __nv_bfloat16*>(x->data);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
    if (mode == 0) launch_k(resample_kernel<0>, grid, 256, 0, st, xp, yp, y->n, y->h, y->w, y->c, x->pitch, y->pitch, x->h, x->w);
    if (mode == 1) launch_k(resample_kernel<1>, grid, 256, 0, st, xp, yp, y->n, y->h, y->w, y->c, x->pitch, y->pitch, x->h, x->w);
    if (mode == 2) launch_k(resample_kernel<2>, grid, 256, 0, st, xp, yp, y->n, y->h, y->w, y->c, x->pitch, y->pitch, x->h, x->w);
    count_launch();
    return check_launch(name);
}


// ---- letterbox pre-process ---------------------------------------------------------------------------------------------
// LetterBox.__call__ (data/augment.py:1558-1600; auto=False, centre, pad 114) + the BGR->RGB / HWC->CHW of BasePredictor.preprocess
// (engine/predictor.py:127-129) for a batch of uint8 images of DIFFERENT sizes in one launch.  The bilinear resize repeats the
// fixed-point arithmetic of cv2.resize(INTER_LINEAR) for 8-bit images: 11-bit coefficients cvRound(f * 2048), horizontal pass in
// int32, vertical pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.  The two axes treat the image border
// differently (imgproc/src/resize.cpp): along x a sample outside the image gets the weights (2048, 0) on the edge pixel; along y
// the fractional weight is KEPT and only the row indices are clamped, so both taps read the edge row and truncate separately.
template <bool VERTICAL>
__device__ __forceinline__ void lb_coeff(int d, int n_out, int n_in, int& s0, int& s1, int& a0, int& a1) {
    const double scale = 1.0 / ((double)n_out / (double)n_in);  // cv2: inv_scale = dsize / ssize, scale = 1 / inv_scale
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (!VERTICAL) {
        if (s < 0) f = 0.f, s = 0;
        if (s >= n_in - 1) f = 0.f, s = n_in - 1;
    }
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    s0 = min(max(s, 0), n_in - 1), s1 = min(max(s + 1, 0), n_in - 1);
}

__global__ void __launch_bounds__(256) letterbox_kernel(const dbl_lb_image* __restrict__ tab, uint8_t* __restrict__ dst, int H, int W, int pad) {
    pdl_launch_dependents();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const dbl_lb_image im = tab[b];
    int r = pad, g = pad, bl = pad;
    const int ry = y - im.top, rx = x - im.left;
    if (ry >= 0 && ry < im.new_h && rx >= 0 && rx < im.new_w) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(im.src);
        if (im.new_h == im.h && im.new_w == im.w) {
            const uint8_t* px = src + (size_t)ry * im.pitch + rx * 3;
            bl = px[0], g = px[1], r = px[2];
        } else {
            int x0, x1, ax0, ax1, y0, y1, ay0, ay1;
            lb_coeff<false>(rx, im.new_w, im.w, x0, x1, ax0, ax1);
            lb_coeff<true>(ry, im.new_h, im.h, y0, y1, ay0, ay1);
            const uint8_t* p00 = src + (size_t)y0 * im.pitch + x0 * 3;
            const uint8_t* p01 = src + (size_t)y0 * im.pitch + x1 * 3;
            const uint8_t* p10 = src + (size_t)y1 * im.pitch + x0 * 3;
            const uint8_t* p11 = src + (size_t)y1 * im.pitch + x1 * 3;
            int out[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int r0 = (int)p00[c] * ax0 + (int)p01[c] * ax1, r1 = (int)p10[c] * ax0 + (int)p11[c] * ax1;
                const int v = (((ay0 * (r0 >> 4)) >> 16) + ((ay1 * (r1 >> 4)) >> 16) + 2) >> 2;
                out[c] = min(max(v, 0), 255);
            }
            bl = out[0], g = out[1], r = out[2];
        }
    }
    uint8_t* d = dst + ((size_t)b * 3 * H + y) * W + x;  // NCHW, RGB
    d[0] = (uint8_t)r;
    d[(size_t)H * W] = (uint8_t)g;
    d[(size_t)2 * H * W] = (uint8_t)bl;
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_copy(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t s) {
    return launch_resample(0, x, y, as_stream(s), "dbl_copy");
}
extern "C" int dbl_avgpool2(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t s) {
    DBL_REQUIRE(x && x->h % 2 == 0 && x->w % 2 == 0, DBL_E_ARG, "dbl_avgpool2: odd input size");
    return launch_resample(1, x, y, as_stream(s), "dbl_avgpool2");
}
extern "C" int dbl_upsample2_nearest(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t s) {
    return launch_resample(2, x, y, as_stream(s), "dbl_upsample2_nearest");
}

extern "C" int dbl_gated_add(const dbl_tensor* a, const dbl_tensor* b, float gate, const dbl_tensor* y, const dbl_tensor* y2,
                             dbl_stream_t s) {
    DBL_REQUIRE(valid_tensor(a) && valid_tensor(b) && valid_tensor(y) && vec8_ok(a) && vec8_ok(b) && vec8_ok(y), DBL_E_ARG,
                "dbl_gated_add: needs bf16 views with c %% 8 == 0");
    DBL_REQUIRE(same_shape(a, b) && same_shape(a, y), DBL_E_ARG, "dbl_gated_add: shape mismatch");
    if (y2) DBL_REQUIRE(valid_tensor(y2) && vec8_ok(y2) && same_shape(a, y2), DBL_E_ARG, "dbl_gated_add: y2 mismatch");
    const long long npix = (long long)a->n * a->h * a->w;
    const long long total = npix * (a->c / 8);
    launch_k(gated_add_kernel, (unsigned)((total + 255) / 256), 256, 0, as_stream(s), 
        reinterpret_cast<const __nv_bfloat16*>(a->data), reinterpret_cast<const __nv_bfloat16*>(b->data), gate,
        reinterpret_cast<__nv_bfloat16*>(y->data), y2 ? reinterpret_cast<__nv_bfloat16*>(y2->data) : nullptr, npix, a->c,
        a->pitch, b->pitch, y->pitch, y2 ? y2->pitch : 0);
    count_launch();
    return check_launch("dbl_gated_add");
}

extern "C" int dbl_channel_scale_add(const dbl_tensor* a, const dbl_tensor* b, const float* gamma, const dbl_tensor* y, dbl_stream_t s) {
    DBL_REQUIRE(valid_tensor(a) && valid_tensor(b) && valid_tensor(y) && gamma && vec8_ok(a) && vec8_ok(b) && vec8_ok(y), DBL_E_ARG,
                "dbl_channel_scale_add: needs bf16 views with c %% 8 == 0");
    DBL_REQUIRE(same_shape(a, b) && same_shape(a, y) && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0, DBL_E_ARG, "dbl_channel_scale_add: shape / alignment");
    const long long npix = (long long)a->n * a->h * a->w;
    const long long total = npix * (a->c / 8);
    launch_k(channel_scale_add_kernel, (unsigned)((total + 255) / 256), 256, 0, as_stream(s), reinterpret_cast<const __nv_bfloat16*>(a->data),
             reinterpret_cast<const __nv_bfloat16*>(b->data), gamma, reinterpret_cast<__nv_bfloat16*>(y->data), npix, a->c, a->pitch, b->pitch, y->pitch);
    count_launch();
    return check_launch("dbl_channel_scale_add");
}

extern "C" int dbl_dysample(const dbl_tensor* x, const dbl_tensor* off, const dbl_tensor* y, int groups, dbl_stream_t s) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(off) && valid_tensor(y) && vec8_ok(x) && vec8_ok(y), DBL_E_ARG,
                "dbl_dysample: needs bf16 x/y with c %% 8 == 0");
    DBL_REQUIRE(off->dtype == DBL_F32 && off->c == 2 * groups * 4 && off->n == x->n && off->h == x->h && off->w == x->w, DBL_E_ARG,
                "dbl_dysample: offsets must be f32 (n,h,w,%d)", 2 * groups * 4);
    DBL_REQUIRE(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == x->c, DBL_E_ARG, "dbl_dysample: output shape");
    DBL_REQUIRE(groups > 0 && x->c % groups == 0 && (x->c / groups) % 8 == 0, DBL_E_UNSUPPORTED,
                "dbl_dysample: channels per group must be a multiple of 8");
    DBL_REQUIRE(y->h <= 65535 && y->n <= 65535, DBL_E_UNSUPPORTED, "dbl_dysample: grid too large");
    if ((x->c / groups) % 16 == 0) {
        launch_k(dysample16_kernel, dim3((unsigned)cdiv(y->w * (y->c / 16), 256), (unsigned)y->h, (unsigned)y->n), 256, 0, as_stream(s),
                 reinterpret_cast<const __nv_bfloat16*>(x->data), reinterpret_cast<const float*>(off->data),
                 reinterpret_cast<__nv_bfloat16*>(y->data), x->n, x->h, x->w, x->c, x->pitch, off->pitch, y->pitch, groups);
        count_launch();
        return check_launch("dbl_dysample");
    }
    launch_k(dysample_kernel, dim3((unsigned)cdiv(y->w * (y->c / 8), 256), (unsigned)y->h, (unsigned)y->n), 256, 0, as_stream(s), 
        reinterpret_cast<const __nv_bfloat16*>(x->data), reinterpret_cast<const float*>(off->data),
        reinterpret_cast<__nv_bfloat16*>(y->data), x->n, x->h, x->w, x->c, x->pitch, off->pitch, y->pitch, groups);
    count_launch();
    return check_launch("dbl_dysample");
}

extern "C" int dbl_lsk_squeeze(const dbl_tensor* attn, const dbl_tensor* agg, dbl_stream_t s) {
    DBL_REQUIRE(valid_tensor(attn) && valid_tensor(agg) && vec8_ok(attn), DBL_E_ARG, "dbl_lsk_squeeze: bad attn");
    DBL_REQUIRE(agg->dtype == DBL_F32 && agg->c == 2 && agg->n == attn->n && agg->h == attn->h && agg->w == attn->w, DBL_E_ARG,
                "dbl_lsk_squeeze: agg must be f32 (n,h,w,2)");
    const long long npix = (long long)attn->n * attn->h * attn->w;
    launch_k(lsk_squeeze_kernel, (unsigned)((npix + 7) / 8), 256, 0, as_stream(s), 
        reinterpret_cast<const __nv_bfloat16*>(attn->data), reinterpret_cast<float*>(agg->data), npix, attn->c, attn->pitch,
        agg->pitch);
    count_launch();
    return check_launch("dbl_lsk_squeeze");
}

extern "C" int dbl_lsk_gate(const dbl_tensor* attn, const dbl_tensor* agg, const float* w_sq, const float* b_sq,
                            const dbl_tensor* y, dbl_stream_t s) {
    DBL_REQUIRE(valid_tensor(attn) && valid_tensor(agg) && valid_tensor(y) && w_sq && b_sq && vec8_ok(attn) && vec8_ok(y), DBL_E_ARG,
                "dbl_lsk_gate: bad argument");
    DBL_REQUIRE(attn->c == 2 * y->c && (y->c % 8) == 0 && agg->dtype == DBL_F32 && agg->c == 2, DBL_E_ARG, "dbl_lsk_gate: channels");
    DBL_REQUIRE(attn->n == y->n && attn->h == y->h && attn->w == y->w && agg->n == y->n && agg->h == y->h && agg->w == y->w, DBL_E_ARG,
                "dbl_lsk_gate: shape mismatch");
    const long long npix = (long long)y->n * y->h * y->w;
    launch_k(lsk_gate_kernel, (unsigned)((npix + 7) / 8), 256, 0, as_stream(s), 
        reinterpret_cast<const __nv_bfloat16*>(attn->data), reinterpret_cast<const float*>(agg->data), w_sq, b_sq,
        reinterpret_cast<__nv_bfloat16*>(y->data), y->n, y->h, y->w, y->c, attn->pitch, agg->pitch, y->pitch);
    count_launch();
    return check_launch("dbl_lsk_gate");
}

extern "C" int dbl_letterbox_u8(const dbl_lb_image* table_dev, int batch, uint8_t* dst_nchw, int H, int W, int pad_value, dbl_stream_t s) {
    DBL_REQUIRE(table_dev && dst_nchw && batch > 0 && H > 0 && W > 0 && pad_value >= 0 && pad_value <= 255, DBL_E_ARG,
                "dbl_letterbox_u8: null/invalid argument");
    DBL_REQUIRE(H <= 65535 && batch <= 65535, DBL_E_UNSUPPORTED, "dbl_letterbox_u8: grid too large");
    launch_k(letterbox_kernel, dim3((unsigned)cdiv(W, 256), (unsigned)H, (unsigned)batch), 256, 0, as_stream(s), table_dev, dst_nchw, H, W, pad_value);
    count_launch();
    return check_launch("dbl_letterbox_u8");
}
