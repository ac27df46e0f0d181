// Area attention of the stock YOLOv13 graph: the core of AAttn.forward (reference nn/modules/block.py:1236-1262) between the qkv
// convolution and the positional-encoding / projection convolutions.
//   qkv: NHWC bf16 (n, h, w, 3*heads*hd) exactly as the 1x1 qkv Conv writes it: per head the channels are [q | k | v], hd each
//        (qkv.view(B, N, heads, 3*hd), block.py:1250-1254);
//   tokens = pixels in row-major order, split into `area` contiguous bands that attend only within themselves (block.py:1247-1249);
//   o[token, head*hd + d] = sum_j softmax_j(q . k_j / sqrt(hd)) v_j[d]      v_out[token, head*hd + d] = v (input of the pe conv).
// Flash-attention forward on warp MMAs (m16n8k16 bf16, fp32 accumulation and softmax): a CTA owns 64 queries of one (image, band, head),
// streams the band's keys / values through shared memory in chunks of 64 with an online softmax; P is rounded to bf16 for the PV
// product.  hd = 32 (the graph fixes heads = dim // 32).  The problem sizes are tiny (400 - 1600 tokens per band, 0.02 % of the model's
// FLOPs): latency-bound, not a roofline candidate.
#include <math_constants.h>

#include "mma_util.cuh"

namespace dbl {

constexpr int AT_HD = 32, AT_QB = 64, AT_KB = 64, AT_PITCH = 40;  // bf16 per shared-memory row (80 bytes: conflict-free fragment loads)

__global__ void __launch_bounds__(128) area_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ o,
                                                             __nv_bfloat16* __restrict__ v_out, int N, int Nt, int heads, int area, int qpitch,
                                                             int opitch, int vpitch, float scale_log2e) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 sK[AT_KB * AT_PITCH];
    __shared__ __align__(16) __nv_bfloat16 sV[AT_KB * AT_PITCH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, ba = blockIdx.z, b = ba / area, band = ba - b * area;
    const int q0 = blockIdx.x * AT_QB;
    const size_t tok0 = (size_t)b * N + (size_t)band * Nt;  // first token (pixel) of this band
    const __nv_bfloat16* base = qkv + tok0 * qpitch + head * 3 * AT_HD;

    // this warp's 16 queries as A fragments (two k-steps of 16 dims); rows past the band read row 0 and are never stored
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const __nv_bfloat16* p0 = base + (size_t)(r0 < Nt ? r0 : 0) * qpitch + ks * 16 + 2 * t;
        const __nv_bfloat16* p1 = base + (size_t)(r1 < Nt ? r1 : 0) * qpitch + ks * 16 + 2 * t;
        qa[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
        qa[ks][1] = *reinterpret_cast<const uint32_t*>(p1);
        qa[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        qa[ks][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
    }
    // v_out: the v vectors of this CTA's own 64 tokens (64 bytes each = four 16-byte pieces)
    for (int i = tid; i < AT_QB * 4; i += 128) {
        const int tok = q0 + (i >> 2), piece = i & 3;
        if (tok < Nt)
            *reinterpret_cast<uint4*>(v_out + (tok0 + tok) * vpitch + head * AT_HD + piece * 8) =
                *reinterpret_cast<const uint4*>(base + (size_t)tok * qpitch + 2 * AT_HD + piece * 8);
    }

    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

    for (int k0 = 0; k0 < Nt; k0 += AT_KB) {
        __syncthreads();  // the previous chunk has been consumed
        for (int i = tid; i < AT_KB * 8; i += 128) {  // 64 keys x (4 pieces of k + 4 pieces of v)
            const int key = i >> 3, piece = i & 7;
            const int tok = k0 + key;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (tok < Nt) val = *reinterpret_cast<const uint4*>(base + (size_t)tok * qpitch + AT_HD + piece * 8);  // k (pieces 0-3) then v (4-7): contiguous
            __nv_bfloat16* dst = (piece < 4 ? sK : sV) + key * AT_PITCH + (piece & 3) * 8;
            *reinterpret_cast<uint4*>(dst) = val;
        }
        __syncthreads();
        // S = Q K^T for 16 queries x 64 keys: 8 n-tiles x 2 k-steps
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const __nv_bfloat16* kp = sK + (8 * j + g) * AT_PITCH + ks * 16 + 2 * t;
                mma16816(s[j], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
            }
        }
        // scale (log2 domain), mask keys past the band, online softmax: thread holds keys 8j + 2t, +1 of rows g (s[j][0..1]) and g+8 (s[j][2..3])
        float cm0 = -CUDART_INF_F, cm1 = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool ok = k0 + 8 * j + 2 * t + e < Nt;
                s[j][e] = ok ? s[j][e] * scale_log2e : -CUDART_INF_F;
                s[j][2 + e] = ok ? s[j][2 + e] * scale_log2e : -CUDART_INF_F;
                cm0 = fmaxf(cm0, s[j][e]);
                cm1 = fmaxf(cm1, s[j][2 + e]);
            }
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);  // finite: every chunk holds at least one valid key
        const float f0 = exp2f(m0 - nm0), f1 = exp2f(m1 - nm1);
        m0 = nm0, m1 = nm1;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pa[4][4];  // P as A fragments: k-step kk covers keys 16kk .. 16kk+15 = n-tiles 2kk, 2kk+1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p00 = exp2f(s[j][0] - nm0), p01 = exp2f(s[j][1] - nm0), p10 = exp2f(s[j][2] - nm1), p11 = exp2f(s[j][3] - nm1);
            ps0 += p00 + p01, ps1 += p10 + p11;
            pa[j >> 1][(j & 1) * 2 + 0] = pack2(p00, p01);
            pa[j >> 1][(j & 1) * 2 + 1] = pack2(p10, p11);
        }
        l0 = l0 * f0 + ps0, l1 = l1 * f1 + ps1;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[nt][0] *= f0, acc[nt][1] *= f0, acc[nt][2] *= f1, acc[nt][3] *= f1;
        // O += P V: 4 k-steps (16 keys) x 4 n-tiles (8 dims); V fragments through ldmatrix.trans (rows = keys)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t b0, b1;
                const uint32_t addr = smem_u32(sV + (16 * kk + (lane & 15)) * AT_PITCH + 8 * nt);
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
                mma16816(acc[nt], pa[kk], b0, b1);
            }
        }
    }
    // row sums live in the four threads of a quad
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        if (r0 < Nt) *reinterpret_cast<uint32_t*>(o + (tok0 + r0) * opitch + head * AT_HD + 8 * nt + 2 * t) = pack2(acc[nt][0] * i0, acc[nt][1] * i0);
        if (r1 < Nt) *reinterpret_cast<uint32_t*>(o + (tok0 + r1) * opitch + head * AT_HD + 8 * nt + 2 * t) = pack2(acc[nt][2] * i1, acc[nt][3] * i1);
    }
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_area_attention(const dbl_tensor* qkv, const dbl_tensor* o, const dbl_tensor* v_out, int heads, int area, dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(qkv) && valid_tensor(o) && valid_tensor(v_out), DBL_E_ARG, "dbl_area_attention: null/invalid tensor");
    DBL_REQUIRE(vec8_ok(qkv) && vec8_ok(o) && vec8_ok(v_out), DBL_E_ARG, "dbl_area_attention: bf16 NHWC views, 16-byte aligned, pitch %% 8 == 0");
    DBL_REQUIRE(heads > 0 && area > 0 && qkv->c == 3 * heads * AT_HD, DBL_E_UNSUPPORTED, "dbl_area_attention: head dim must be %d (channels %d, heads %d)", AT_HD,
                qkv->c, heads);
    DBL_REQUIRE(o->c == heads * AT_HD && v_out->c == o->c && o->n == qkv->n && o->h == qkv->h && o->w == qkv->w && v_out->n == qkv->n && v_out->h == qkv->h &&
                    v_out->w == qkv->w,
                DBL_E_ARG, "dbl_area_attention: output shape mismatch");
    const int N = qkv->h * qkv->w;
    DBL_REQUIRE(N % area == 0, DBL_E_ARG, "dbl_area_attention: %d tokens do not split into %d areas", N, area);
    DBL_REQUIRE(qkv->data != o->data && qkv->data != v_out->data, DBL_E_ARG, "dbl_area_attention: outputs must not alias qkv");
    const int Nt = N / area;
    DBL_REQUIRE((long long)qkv->n * area <= 65535 && heads <= 65535, DBL_E_UNSUPPORTED, "dbl_area_attention: grid too large");
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)AT_HD);
    launch_k(area_attention_kernel, dim3(cdiv(Nt, AT_QB), heads, qkv->n * area), 128, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(qkv->data),
             reinterpret_cast<__nv_bfloat16*>(o->data), reinterpret_cast<__nv_bfloat16*>(v_out->data), N, Nt, heads, area, qkv->pitch, o->pitch, v_out->pitch,
             scale_log2e);
    count_launch();
    return check_launch("dbl_area_attention");
}
