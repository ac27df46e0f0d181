// Fused Detect tail: the last 1x1 convolutions of both head branches + DFL decode + confidence filter in ONE kernel.
//
// Reference: Detect.forward's final nn.Conv2d(c2, 64, 1) / nn.Conv2d(c3, nc, 1) (nn/modules/head.py:86-100,108-118),
// Detect._inference + DFL + make_anchors + dist2bbox (head.py:143-181, block.py:80-83, utils/tal.py:333-357) and stage 1
// of non_max_suppression (utils/ops.py:234-279).
//
// Unfused, the head writes 64+nc fp32 logits per anchor (310 MB per 64-image batch at nc=80) that the decode kernel reads
// straight back; here they never leave registers.  A CTA takes 128 anchors of one level of one image: the two branch
// features (bf16, 64 and c3 channels) are copied to shared memory once (cp.async, planar layout), each warp runs the two
// small GEMMs for 16 anchors on mma.sync (fp32 accumulate, fp32 bias).  The packed-weight channel permutation gives every
// thread of a quad one whole DFL side (16 contiguous bins) and 2*NB contiguous classes, so softmax-expectation, sigmoid
// and arg-max are register-local; the quad exchanges four distances / four partial maxima by shuffle.
// Optional outputs: raw logits (fp32 NHWC, what DetectionModel.forward returns as `feats`), y (the reference's
// [B, 4+nc, A] tensor) and the candidate lists of the NMS.  The arithmetic after the GEMMs is the one of detect_post.cu.
#include <cstring>

#include "mma_util.cuh"
#include "post_util.cuh"

namespace dbl {

struct HfLevel {
    const __nv_bfloat16* f2;  // [B,h,w,pitch2]: cv2 branch feature (64 channels)
    const __nv_bfloat16* f3;  // [B,h,w,pitch3]: cv3 branch feature (C3 channels)
    const uint2* w2;          // packed (64 x 64, 1 tap)
    const uint2* w3;          // packed (NCP x C3, 1 tap)
    const float* b2;
    const float* b3;
    float* logits;            // optional [B,h,w,lpitch]
    int pitch2, pitch3, h, w, start, blocks;  // start: first anchor index; blocks: CTAs per image for this level
    float stride;
};
struct HfParams {
    HfLevel lv[3];
    int nl, batch, nc, A, lpitch, blocks_per_img;
    float* y;  // optional
    float conf;
    int multi_label, ncls, want_cand;
    const int32_t* classes;
    dbl_candidates cd;
};

template <int C3, int NCP>
struct HfCfg {
    static constexpr int C2 = 64, MPX = 128;
    static constexpr int G2 = C2 / 8, G3 = C3 / 8;
    static constexpr int PS = (MPX + 1) * 16;  // plane stride = 16 mod 128 bytes: the 8 planes of one pixel (one quarter-warp of cp.async) hit 8 different bank groups
    static constexpr int NB2 = 8, NB3 = NCP / 8;
    static constexpr int KS2 = C2 / 16, KS3 = C3 / 16;
    static constexpr int W2B = KS2 * NB2 * 256, W3B = KS3 * NB3 * 256;
    static constexpr int A2B = G2 * PS, A3B = G3 * PS;
    static constexpr int SMEM = A2B + A3B + W2B + W3B + (64 + NCP) * 4 + 128;
    static_assert(SMEM <= 227 * 1024, "head tile does not fit shared memory");
};

template <int C3, int NCP>
__global__ void __launch_bounds__(256) head_decode_kernel(const HfParams p) {
    using K = HfCfg<C3, NCP>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t sA2 = smem_u32(smem);
    const uint32_t sA3 = sA2 + K::A2B;
    const uint32_t sW2 = sA3 + K::A3B;
    const uint32_t sW3 = sW2 + K::W2B;
    float* sB2 = reinterpret_cast<float*>(smem + K::A2B + K::A3B + K::W2B + K::W3B);
    float* sB3 = sB2 + 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    pdl_launch_dependents();

    // which (image, level, pixel block) is this CTA?
    const int b = blockIdx.x / p.blocks_per_img;
    int r = blockIdx.x - b * p.blocks_per_img, lvl = 0;
    while (lvl + 1 < p.nl && r >= p.lv[lvl].blocks) r -= p.lv[lvl].blocks, ++lvl;
    const HfLevel& L = p.lv[lvl];
    const int hw = L.h * L.w, px0 = r * K::MPX;

    for (int i = tid; i < K::W2B / 16; i += 256) cp_async16(sW2 + i * 16, reinterpret_cast<const uint8_t*>(L.w2) + i * 16, true);
    for (int i = tid; i < K::W3B / 16; i += 256) cp_async16(sW3 + i * 16, reinterpret_cast<const uint8_t*>(L.w3) + i * 16, true);
    for (int i = tid; i < 64; i += 256) sB2[i] = L.b2 ? L.b2[i] : 0.f;
    for (int i = tid; i < NCP; i += 256) sB3[i] = (L.b3 && i < p.nc) ? L.b3[i] : 0.f;
    pdl_wait();  // the branch features are the previous kernels' outputs
    {
        const __nv_bfloat16* f2 = L.f2 + ((size_t)b * hw + px0) * L.pitch2;
        const __nv_bfloat16* f3 = L.f3 + ((size_t)b * hw + px0) * L.pitch3;
        for (int i = tid; i < K::MPX * K::G2; i += 256) {
            const int gi = i % K::G2, px = i / K::G2;
            const bool ok = px0 + px < hw;
            cp_async16(sA2 + gi * K::PS + px * 16, ok ? f2 + (size_t)px * L.pitch2 + gi * 8 : L.f2, ok);
        }
        for (int i = tid; i < K::MPX * K::G3; i += 256) {
            const int gi = i % K::G3, px = i / K::G3;
            const bool ok = px0 + px < hw;
            cp_async16(sA3 + gi * K::PS + px * 16, ok ? f3 + (size_t)px * L.pitch3 + gi * 8 : L.f3, ok);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const uint32_t row16 = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 16);
    const uint32_t a_lane = row16 + (lane >= 16 ? K::PS : 0) + (uint32_t)warp * 256;  // this warp's 16 anchors

    // ---------------- class branch first: C3 -> nc logits; thread t owns classes [t*2*NB3, (t+1)*2*NB3) ----------------
    float acc3[K::NB3][4];
#pragma unroll
    for (int j = 0; j < K::NB3; ++j) {
        const float2 bb = lds_f2(sB3 + t * 2 * K::NB3 + 2 * j);
        acc3[j][0] = acc3[j][2] = bb.x, acc3[j][1] = acc3[j][3] = bb.y;
    }
#pragma unroll
    for (int ks = 0; ks < K::KS3; ++ks) {
        uint32_t a[4];
        ldsm_x4(sA3 + a_lane + (uint32_t)(2 * ks) * K::PS, a);
#pragma unroll
        for (int j = 0; j < K::NB3; ++j) {
            const uint2 bw = lds64(sW3 + (uint32_t)((ks * K::NB3 + j) * 256) + lane * 8);
            mma16816(acc3[j], a, bw.x, bw.y);
        }
    }
    // Cheap pre-filter: sigmoid is monotonic, so max_c sigmoid(l_c) = sigmoid(max_c l_c) up to an ulp of expf; a warp none of
    // whose 16 anchors can reach the threshold (with a 1e-6 safety margin for that ulp) needs neither the 2*NB3 exact
    // sigmoids per thread nor the box branch at all.  The exact per-class arithmetic below decides for everybody else.
    const bool want_all = p.y != nullptr || L.logits != nullptr;
    bool maybe = false;
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
        float m = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < K::NB3; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (t * 2 * K::NB3 + 2 * j + e < p.nc) m = fmaxf(m, acc3[j][2 * rh + e]);
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        maybe = maybe || (px0 + warp * 16 + g + 8 * rh < hw && sigmoid_exact(m) > p.conf - 1e-6f);
    }
    if (!want_all && !__any_sync(0xffffffffu, maybe)) return;

    // ---------------- box branch: 64 -> 64 logits = 4 sides x 16 DFL bins; thread t of a quad owns side t ----------------
    float acc2[K::NB2][4];
#pragma unroll
    for (int j = 0; j < K::NB2; ++j) {
        const float2 bb = lds_f2(sB2 + t * 16 + 2 * j);
        acc2[j][0] = acc2[j][2] = bb.x, acc2[j][1] = acc2[j][3] = bb.y;
    }
#pragma unroll
    for (int ks = 0; ks < K::KS2; ++ks) {
        uint32_t a[4];
        ldsm_x4(sA2 + a_lane + (uint32_t)(2 * ks) * K::PS, a);
#pragma unroll
        for (int j = 0; j < K::NB2; ++j) {
            const uint2 bw = lds64(sW2 + (uint32_t)((ks * K::NB2 + j) * 256) + lane * 8);
            mma16816(acc2[j], a, bw.x, bw.y);
        }
    }

    // ---------------- per anchor: decode + filter (+ optional logits / y) ----------------
#pragma unroll 1
    for (int rh = 0; rh < 2; ++rh) {
        const int cell = px0 + warp * 16 + g + 8 * rh;
        const bool live = cell < hw;
        const int a = L.start + cell;
        float v[REG_MAX];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[2 * j] = rh ? acc2[j][2] : acc2[j][0], v[2 * j + 1] = rh ? acc2[j][3] : acc2[j][1];
        float lg3[2 * K::NB3];
#pragma unroll
        for (int j = 0; j < K::NB3; ++j) lg3[2 * j] = rh ? acc3[j][2] : acc3[j][0], lg3[2 * j + 1] = rh ? acc3[j][3] : acc3[j][1];
        if (L.logits && live) {
            float* lp = L.logits + ((size_t)b * hw + cell) * p.lpitch;
#pragma unroll
            for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(lp + t * 16)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
            for (int i = 0; i < 2 * K::NB3; ++i) {
                const int c = t * 2 * K::NB3 + i;
                if (c < p.nc) lp[64 + c] = lg3[i];
            }
        }
        // scores of this thread's classes, best class of the anchor (lowest index among equal maxima, like the reference)
        // Only classes that can reach the anchor's best score need the exact sigmoid when just (best, class) is wanted: for a
        // maximum logit m <= 8 two logits more than ~2e-4 apart never round to the same float sigmoid, so a 1e-3 window around m
        // contains every class that could win or tie (ties keep the lowest index, like the reference); saturated scores (m > 8),
        // multi_label and the y output take every class.
        float mq = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < 2 * K::NB3; ++i)
            if (t * 2 * K::NB3 + i < p.nc) mq = fmaxf(mq, lg3[i]);
        mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, 1));
        mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, 2));
        const float lo = (p.y == nullptr && !p.multi_label && mq <= 8.0f) ? mq - 1e-3f : -CUDART_INF_F;
        float sc[2 * K::NB3];
        float best = -CUDART_INF_F;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < 2 * K::NB3; ++i) {
            const int c = t * 2 * K::NB3 + i;
            float s = -CUDART_INF_F;
            if (c < p.nc && lg3[i] >= lo) s = sigmoid_exact(lg3[i]);
            sc[i] = s;
            if (s > best) best = s, bi = c;
        }
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {  // quad reduction; classes are ordered by t, so ties keep the lower t
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        const bool any = live && best > p.conf;
        const bool need_box = (p.y != nullptr && live) || any;
        // the quad's four DFL sides (every lane computes its own side; uniform control flow keeps the shuffles legal)
        const float dside = __any_sync(0xffffffffu, need_box) ? dfl_side_regs(v) : 0.f;
        const float dl = __shfl_sync(0xffffffffu, dside, (lane & ~3) | 0), dt = __shfl_sync(0xffffffffu, dside, (lane & ~3) | 1);
        const float dr = __shfl_sync(0xffffffffu, dside, (lane & ~3) | 2), db = __shfl_sync(0xffffffffu, dside, (lane & ~3) | 3);
        const float ax = (float)(cell % L.w) + 0.5f, ay = (float)(cell / L.w) + 0.5f;
        const float4 xywh = dist_to_xywh(dl, dt, dr, db, ax, ay, L.stride);
        if (p.y && live) {
            float* yo = p.y + (size_t)b * (4 + p.nc) * p.A + a;
            if (t == 0) yo[0] = xywh.x, yo[(size_t)p.A] = xywh.y, yo[(size_t)2 * p.A] = xywh.z, yo[(size_t)3 * p.A] = xywh.w;
#pragma unroll
            for (int i = 0; i < 2 * K::NB3; ++i) {
                const int c = t * 2 * K::NB3 + i;
                if (c < p.nc) yo[(size_t)(4 + c) * p.A] = sc[i];
            }
        }
        if (p.want_cand) {
            const float4 box = xywh_to_xyxy(xywh);
            if (!p.multi_label) {
                push_candidate(any && t == 0 && class_allowed(bi, p.classes, p.ncls), p.cd, b, box, best, bi, a);
            } else {
#pragma unroll 1
                for (int i = 0; i < 2 * K::NB3; ++i) {
                    const int c = t * 2 * K::NB3 + i;
                    float s = sc[0];
#pragma unroll
                    for (int q = 1; q < 2 * K::NB3; ++q) s = (q == i) ? sc[q] : s;
                    push_candidate(any && c < p.nc && s > p.conf && class_allowed(c, p.classes, p.ncls), p.cd, b, box, s, c, a * p.nc + c);
                }
            }
        }
    }
}

template <int C3, int NCP>
static int launch_head(const HfParams& p, cudaStream_t st) {
    using K = HfCfg<C3, NCP>;
    auto kern = head_decode_kernel<C3, NCP>;
    static DevOnce once;  // one per template instantiation, one bit per device
    if (once.need()) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("dbl_head_decode: cannot opt in to %d bytes of shared memory: %s", (int)K::SMEM, cudaGetErrorString(e));
            return DBL_E_CUDA;
        }
        once.mark();
    }
    launch_k(kern, p.batch * p.blocks_per_img, 256, (size_t)K::SMEM, st, p);
    count_launch();
    return check_launch("dbl_head_decode");
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_head_decode_supported(int c2, int c3, int nc) {
    if (c2 != 64) return 0;
    return ((c3 == 80 && nc == 80) || (c3 == 64 && nc >= 1 && nc <= 8) || (c3 == 128 && nc >= 1 && nc <= 8) || (c3 == 128 && nc == 80)) ? 1 : 0;
}

extern "C" int dbl_head_decode(const dbl_head_fused* hd, float* y, float conf_thres, int multi_label, const int32_t* classes,
                               int num_classes_kept, const dbl_candidates* cand, dbl_stream_t stream) {
    DBL_REQUIRE(hd && hd->num_levels >= 1 && hd->num_levels <= 3 && hd->batch >= 1 && hd->nc >= 1, DBL_E_ARG, "dbl_head_decode: invalid head");
    DBL_REQUIRE(dbl_head_decode_supported(hd->c2, hd->c3, hd->nc), DBL_E_UNSUPPORTED, "dbl_head_decode: no instantiation for c2 %d c3 %d nc %d",
                hd->c2, hd->c3, hd->nc);
    HfParams p;
    memset(&p, 0, sizeof(p));
    int a = 0, blocks = 0;
    for (int i = 0; i < hd->num_levels; ++i) {
        const dbl_head_fused_level& s = hd->level[i];
        DBL_REQUIRE(s.f2 && s.f3 && s.w2_packed && s.w3_packed && s.h > 0 && s.w > 0, DBL_E_ARG, "dbl_head_decode: level %d incomplete", i);
        DBL_REQUIRE(s.pitch2 % 8 == 0 && s.pitch3 % 8 == 0 && ((reinterpret_cast<uintptr_t>(s.f2) | reinterpret_cast<uintptr_t>(s.f3)) & 15) == 0,
                    DBL_E_ARG, "dbl_head_decode: features must be 16-byte aligned bf16 NHWC views");
        HfLevel& L = p.lv[i];
        L.f2 = reinterpret_cast<const __nv_bfloat16*>(s.f2), L.f3 = reinterpret_cast<const __nv_bfloat16*>(s.f3);
        L.w2 = reinterpret_cast<const uint2*>(s.w2_packed), L.w3 = reinterpret_cast<const uint2*>(s.w3_packed);
        L.b2 = s.b2, L.b3 = s.b3, L.logits = s.logits;
        L.pitch2 = s.pitch2, L.pitch3 = s.pitch3, L.h = s.h, L.w = s.w, L.stride = s.stride;
        L.start = a, L.blocks = cdiv(s.h * s.w, 128);
        a += s.h * s.w, blocks += L.blocks;
    }
    p.nl = hd->num_levels, p.batch = hd->batch, p.nc = hd->nc, p.A = a, p.lpitch = hd->logit_pitch, p.blocks_per_img = blocks;
    for (int i = 0; i < hd->num_levels; ++i)
        if (p.lv[i].logits) DBL_REQUIRE(hd->logit_pitch >= 64 + hd->nc && hd->logit_pitch % 4 == 0, DBL_E_ARG, "dbl_head_decode: logit pitch %d", hd->logit_pitch);
    p.y = y, p.conf = conf_thres, p.multi_label = (multi_label && hd->nc > 1) ? 1 : 0, p.classes = classes, p.ncls = num_classes_kept;
    cudaStream_t st = as_stream(stream);
    if (cand) {
        DBL_REQUIRE(cand->box && cand->score && cand->cls && cand->pos && cand->count && cand->batch == hd->batch, DBL_E_ARG,
                    "dbl_head_decode: invalid candidate buffers");
        DBL_REQUIRE(cand->cap >= (p.multi_label ? a * hd->nc : a), DBL_E_ARG, "dbl_head_decode: candidate capacity %d too small", cand->cap);
        p.cd = *cand, p.want_cand = 1;
        const cudaError_t e = cudaMemsetAsync(cand->count, 0, sizeof(int32_t) * cand->batch, st);
        DBL_REQUIRE(e == cudaSuccess, DBL_E_CUDA, "dbl_head_decode: memset: %s", cudaGetErrorString(e));
    }
    DBL_REQUIRE(p.batch * (long long)blocks < (1ll << 31), DBL_E_UNSUPPORTED, "dbl_head_decode: grid too large");
    if (hd->c3 == 80 && hd->nc == 80) return launch_head<80, 80>(p, st);
    if (hd->c3 == 64) return launch_head<64, 8>(p, st);
    if (hd->c3 == 128 && hd->nc <= 8) return launch_head<128, 8>(p, st);
    return launch_head<128, 80>(p, st);
}
