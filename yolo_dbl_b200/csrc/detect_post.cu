// Detect post-process: DFL decode, confidence filter with compaction, stable sort + class-offset greedy NMS.
// Arithmetic that decides WHICH boxes survive (xywh->xyxy, class offset, IoU) is written with explicit round-to-nearest
// intrinsics so the compiler can never contract a multiply-add: the keep set must equal the reference's
// torchvision.ops.nms result bit for bit.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "post_util.cuh"

namespace dbl {

struct LevelGeom {
    const float* data[3];
    int h[3], w[3];
    float stride[3];
    int start[3];  // first anchor index of each level
    int nl, nc, pitch, A;
};

static bool make_geom(const dbl_head_levels* lv, LevelGeom* g) {
    if (!lv || lv->num_levels < 1 || lv->num_levels > 3 || lv->nc < 1 || lv->batch < 1 || lv->pitch < 64 + lv->nc) return false;
    int a = 0;
    for (int i = 0; i < 3; ++i) {
        const bool on = i < lv->num_levels;
        if (on && (!lv->data[i] || lv->h[i] < 1 || lv->w[i] < 1)) return false;
        g->data[i] = on ? lv->data[i] : nullptr;
        g->h[i] = on ? lv->h[i] : 0;
        g->w[i] = on ? lv->w[i] : 0;
        g->stride[i] = on ? lv->stride[i] : 0.f;
        g->start[i] = a;
        a += g->h[i] * g->w[i];
    }
    g->nl = lv->num_levels, g->nc = lv->nc, g->pitch = lv->pitch, g->A = a;
    return true;
}

__device__ __forceinline__ void locate_anchor(const LevelGeom& g, int a, int& lvl, int& cell) {
    lvl = (g.nl > 2 && a >= g.start[2]) ? 2 : ((g.nl > 1 && a >= g.start[1]) ? 1 : 0);
    cell = a - g.start[lvl];
}

// ---- decode to the reference's y layout ---------------------------------------------------------------------------
// grid (ceil(A/256), B); thread per anchor.
__global__ void __launch_bounds__(256) decode_kernel(LevelGeom g, float* __restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (a >= g.A) return;
    int lvl, cell;
    locate_anchor(g, a, lvl, cell);
    const float* row = g.data[lvl] + ((size_t)b * g.h[lvl] * g.w[lvl] + cell) * g.pitch;
    const float ax = (float)(cell % g.w[lvl]) + 0.5f, ay = (float)(cell / g.w[lvl]) + 0.5f;
    const float4 v = dist_to_xywh(dfl_side(row), dfl_side(row + 16), dfl_side(row + 32), dfl_side(row + 48), ax, ay, g.stride[lvl]);
    float* yo = y + (size_t)b * (4 + g.nc) * g.A + a;
    yo[0] = v.x;
    yo[(size_t)g.A] = v.y;
    yo[(size_t)2 * g.A] = v.z;
    yo[(size_t)3 * g.A] = v.w;
    for (int c = 0; c < g.nc; ++c) yo[(size_t)(4 + c) * g.A] = sigmoid_exact(row[64 + c]);
}

// ---- fused decode + filter ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_filter_kernel(LevelGeom g, float conf, int multi_label, const int32_t* __restrict__ classes,
                                                            int ncls, dbl_candidates cd) {
    pdl_launch_dependents();
    pdl_wait();
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const bool live = a < g.A;
    int lvl = 0, cell = 0;
    if (live) locate_anchor(g, a, lvl, cell);
    const float* row = live ? g.data[lvl] + ((size_t)b * g.h[lvl] * g.w[lvl] + cell) * g.pitch : nullptr;
    // Cheap pre-filter (same as head_fused.cu): sigmoid is monotonic, so max_c sigmoid(l_c) = sigmoid(max_c l_c) up to an ulp of expf;
    // an anchor whose largest logit cannot reach the threshold (1e-6 safety margin for that ulp) needs none of the nc exact
    // sigmoids (an expf and an IEEE division each).  The exact per-class arithmetic decides for everybody else.
    float best = -CUDART_INF_F;
    int bi = 0;
    if (live) {
        float m = -CUDART_INF_F;
        {   // 16-byte loads, eight independent ones in flight per thread (the scalar walk along the row kept ~one 32-byte sector per
            // thread in flight: 2 TB/s)
            const float4* r4 = reinterpret_cast<const float4*>(row + 64);
            const int n4 = g.nc >> 2;
#pragma unroll 8
            for (int c4 = 0; c4 < n4; ++c4) {
                const float4 v = __ldg(r4 + c4);
                m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            }
            for (int c = n4 * 4; c < g.nc; ++c) m = fmaxf(m, row[64 + c]);
        }
        if (sigmoid_exact(m) > conf - 1e-6f) {
            for (int c = 0; c < g.nc; ++c) {
                const float s = sigmoid_exact(row[64 + c]);
                if (s > best) best = s, bi = c;
            }
        }
    }
    const bool any = live && best > conf;
    float4 box = make_float4(0, 0, 0, 0);
    if (any) {
        const float ax = (float)(cell % g.w[lvl]) + 0.5f, ay = (float)(cell / g.w[lvl]) + 0.5f;
        box = xywh_to_xyxy(dist_to_xywh(dfl_side(row), dfl_side(row + 16), dfl_side(row + 32), dfl_side(row + 48), ax, ay, g.stride[lvl]));
    }
    if (!multi_label) {
        push_candidate(any && class_allowed(bi, classes, ncls), cd, b, box, best, bi, a);
    } else {
        for (int c = 0; c < g.nc; ++c) {
            const float s = any ? sigmoid_exact(row[64 + c]) : 0.f;
            push_candidate(any && s > conf && class_allowed(c, classes, ncls), cd, b, box, s, c, a * g.nc + c);
        }
    }
}

// ---- filter from a materialised prediction (the reference's y) ----------------------------------------------------
__global__ void __launch_bounds__(256) filter_kernel(const float* __restrict__ pred, int nc, int A, long long sb, long long sc, long long sa,
                                                     float conf, int multi_label, const int32_t* __restrict__ classes, int ncls,
                                                     dbl_candidates cd) {
    pdl_launch_dependents();
    pdl_wait();
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const bool live = a < A;
    const float* p = pred + (size_t)b * sb + (size_t)(live ? a : 0) * sa;
    float best = -CUDART_INF_F;
    int bi = 0;
    if (live) {
        for (int c = 0; c < nc; ++c) {
            const float s = p[(size_t)(4 + c) * sc];
            if (s > best) best = s, bi = c;
        }
    }
    const bool any = live && best > conf;
    float4 box = make_float4(0, 0, 0, 0);
    if (any) box = xywh_to_xyxy(make_float4(p[0], p[(size_t)sc], p[(size_t)2 * sc], p[(size_t)3 * sc]));
    if (!multi_label) {
        push_candidate(any && class_allowed(bi, classes, ncls), cd, b, box, best, bi, a);
    } else {
        for (int c = 0; c < nc; ++c) {
            const float s = any ? p[(size_t)(4 + c) * sc] : 0.f;
            push_candidate(any && s > conf && class_allowed(c, classes, ncls), cd, b, box, s, c, a * nc + c);
        }
    }
}

// IoU(p, q) > thr, evaluated exactly like torchvision's CPU kernel (strict fp32, no FMA contraction).  Pairs that do
// not intersect are by far the most common; for them inter == 0, so ovr is 0 (or NaN when both boxes are empty) and the
// comparison is false for every thr >= 0 -- the IEEE division is skipped without changing any result.
__device__ __forceinline__ bool iou_exceeds(const float4& pb, float pa, const float4& q, float qa, float thr) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(pb.z, q.z), fmaxf(pb.x, q.x)));
    if (!(w > 0.f)) return false;
    const float h = fmaxf(0.f, __fsub_rn(fminf(pb.w, q.w), fmaxf(pb.y, q.y)));
    if (!(h > 0.f)) return false;
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(pa, qa), inter)) > thr;
}

// ---- sort + NMS ---------------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 1024;
constexpr int NMS_FLAG_CAP = 40960;  // candidates per image the suppression flags can hold (smem bytes)
constexpr int NMS_SMEM_SORT = 8192;  // slow path: lists up to this many keys (12 bytes each) are sorted in shared memory
constexpr int NMS_FAST_N = 1024;     // images with at most this many candidates are handled entirely in shared memory
constexpr int NMS_SMEM = NMS_FAST_N * (8 + 16 + 4 + 32 * 4);  // keys | boxes | slots | suppression bit-matrix (> NMS_FLAG_CAP)

__host__ __device__ inline size_t next_pow2(size_t v) {
    size_t p = 32;
    while (p < v) p <<= 1;
    return p;
}

// Scratch per image (in uint64 units): keys[P] | payload u32[P] (P/2 u64) | boxes float4[P] (2P u64)
// Epilogue of the NMS: one output row.  With a scale record (pad_x, pad_y, gain, w0, h0) the kept box leaves in the ORIGINAL image's
// coordinates: scale_boxes + clip_boxes (utils/ops.py:92-127, 319-338) -- subtract the letterbox padding, divide by the gain (an
// IEEE division like the reference's fp32 tensor / scalar), clamp to the image -- so no separate pass touches the rows again.
__device__ __forceinline__ void emit_row(float* __restrict__ ro, float4 bx, float score, int cls, const float* __restrict__ sc) {
    if (sc) {
        const float px = sc[0], py = sc[1], g = sc[2], w0 = sc[3], h0 = sc[4];
        bx.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, px), g), 0.f), w0);
        bx.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, py), g), 0.f), h0);
        bx.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, px), g), 0.f), w0);
        bx.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, py), g), 0.f), h0);
    }
    ro[0] = bx.x, ro[1] = bx.y, ro[2] = bx.z, ro[3] = bx.w, ro[4] = score, ro[5] = (float)cls;
}

__global__ void __launch_bounds__(NMS_THREADS) sort_nms_kernel(dbl_candidates cd, float iou_thr, float cls_scale, int max_nms, int max_det,
                                                               uint64_t* __restrict__ scratch, size_t per_image, const float* __restrict__ scale,
                                                               float* __restrict__ out_rows, int32_t* __restrict__ out_pos,
                                                               int32_t* __restrict__ out_count) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char nms_smem[];
    unsigned char* dead = nms_smem;  // slow path: one flag per candidate (NMS_FLAG_CAP bytes), overlays the fast-path area
    __shared__ int s_pivot;
    __shared__ int s_keep[1024];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(cd.count[b], cd.cap);
    const float* sc = scale ? scale + 5 * (size_t)b : nullptr;
    if (n <= NMS_FAST_N) {
        // ---------------- fast path (the usual case): everything stays in shared memory --------------------------------
        // stable sort -> class-offset boxes -> full suppression bit-matrix in parallel -> one warp walks it greedily.
        uint64_t* fkeys = reinterpret_cast<uint64_t*>(nms_smem);                       // [1024]
        float4* fbox = reinterpret_cast<float4*>(nms_smem + NMS_FAST_N * 8);           // [1024]
        uint32_t* fslot = reinterpret_cast<uint32_t*>(nms_smem + NMS_FAST_N * 24);     // [1024]
        uint32_t* fmask = reinterpret_cast<uint32_t*>(nms_smem + NMS_FAST_N * 28);     // [1024][32]
        const size_t cb = (size_t)b * cd.cap;
        int P = 32;
        while (P < n) P <<= 1;
        for (int i = tid; i < P; i += NMS_THREADS) {
            if (i < n) {
                fkeys[i] = ((uint64_t)(~__float_as_uint(cd.score[cb + i])) << 32) | (uint32_t)cd.pos[cb + i];
                fslot[i] = (uint32_t)i;
            } else {
                fkeys[i] = ~0ull;
                fslot[i] = 0xffffffffu;
            }
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < P / 2; t += NMS_THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const uint64_t ki = fkeys[i], kl = fkeys[l];
                    if ((ki > kl) == ((i & k) == 0)) {
                        fkeys[i] = kl, fkeys[l] = ki;
                        const uint32_t si = fslot[i];
                        fslot[i] = fslot[l], fslot[l] = si;
                    }
                }
                __syncthreads();
            }
        }
        const int m = min(n, max_nms);
        for (int i = tid; i < m; i += NMS_THREADS) {
            const size_t sidx = cb + fslot[i];
            const float4 bx = reinterpret_cast<const float4*>(cd.box)[sidx];
            const float off = __fmul_rn((float)cd.cls[sidx], cls_scale);
            fbox[i] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
        }
        __syncthreads();
        const int W = (m + 31) >> 5;
        for (int item = tid; item < m * W; item += NMS_THREADS) {
            const int i = item / W, wq = item - i * W;
            uint32_t word = 0;
            if (32 * wq + 31 > i) {
                const float4 pb = fbox[i];
                const float pa = __fmul_rn(__fsub_rn(pb.z, pb.x), __fsub_rn(pb.w, pb.y));
                for (int bit = 0; bit < 32; ++bit) {
                    const int j = 32 * wq + bit;
                    if (j > i && j < m) {
                        const float4 q = fbox[j];
                        const float qa = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
                        if (iou_exceeds(pb, pa, q, qa, iou_thr)) word |= 1u << bit;
                    }
                }
            }
            fmask[i * 32 + wq] = word;
        }
        __syncthreads();
        int kept = 0;
        if (tid < 32) {  // lane l owns the "removed" bits of candidates [32l, 32l+32)
            uint32_t removed = 0;
            for (int i = 0; i < m && kept < max_det; ++i) {
                const uint32_t r = __shfl_sync(0xffffffffu, removed, i >> 5);
                if (!((r >> (i & 31)) & 1u)) {
                    if (tid == 0) s_keep[kept] = i;
                    ++kept;
                    if (tid < W) removed |= fmask[i * 32 + tid];
                }
            }
            if (tid == 0) s_pivot = kept;
        }
        __syncthreads();
        kept = s_pivot;
        for (int k = tid; k < max_det; k += NMS_THREADS) {
            float* ro = out_rows + ((size_t)b * max_det + k) * 6;
            if (k < kept) {
                const size_t sidx = cb + fslot[s_keep[k]];
                emit_row(ro, reinterpret_cast<const float4*>(cd.box)[sidx], cd.score[sidx], cd.cls[sidx], sc);
                out_pos[(size_t)b * max_det + k] = cd.pos[sidx];
            } else {
                ro[0] = ro[1] = ro[2] = ro[3] = ro[4] = ro[5] = 0.f;
                out_pos[(size_t)b * max_det + k] = -1;
            }
        }
        if (tid == 0) out_count[b] = kept;
        return;
    }
    const size_t P = next_pow2((size_t)(n > 0 ? n : 1));
    uint64_t* keys = scratch + (size_t)b * per_image;
    const size_t Pcap = next_pow2((size_t)cd.cap);
    uint32_t* slot_of = reinterpret_cast<uint32_t*>(keys + Pcap);
    float4* sbox = reinterpret_cast<float4*>(keys + Pcap + Pcap / 2);
    const size_t cbase = (size_t)b * cd.cap;

    // 1. keys: descending score, ascending position (scores are positive floats: their bit patterns order like the values)
    for (size_t i = tid; i < P; i += NMS_THREADS) {
        if (i < (size_t)n) {
            const uint32_t sbits = __float_as_uint(cd.score[cbase + i]);
            keys[i] = ((uint64_t)(~sbits) << 32) | (uint32_t)cd.pos[cbase + i];
            slot_of[i] = (uint32_t)i;
        } else {
            keys[i] = ~0ull;
            slot_of[i] = 0xffffffffu;
        }
    }
    __syncthreads();
    // 2. bitonic sort (keys are unique, so the result equals a stable sort by score).  Up to NMS_SMEM_SORT keys are
    //    sorted in shared memory (the flags that later live there are not needed yet) and written back; larger lists
    //    sort in place in global memory.
    if (P <= (size_t)NMS_SMEM_SORT) {
        uint64_t* skeys = reinterpret_cast<uint64_t*>(nms_smem);
        uint32_t* sslot = reinterpret_cast<uint32_t*>(nms_smem + NMS_SMEM_SORT * 8);
        const int Pi = (int)P;
        for (int i = tid; i < Pi; i += NMS_THREADS) skeys[i] = keys[i], sslot[i] = slot_of[i];
        __syncthreads();
        for (int k = 2; k <= Pi; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < Pi / 2; t += NMS_THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const uint64_t ki = skeys[i], kl = skeys[l];
                    if ((ki > kl) == ((i & k) == 0)) {
                        skeys[i] = kl, skeys[l] = ki;
                        const uint32_t si = sslot[i];
                        sslot[i] = sslot[l], sslot[l] = si;
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < Pi; i += NMS_THREADS) slot_of[i] = sslot[i];
        __syncthreads();
    } else {
        for (size_t k = 2; k <= P; k <<= 1) {
            for (size_t j = k >> 1; j > 0; j >>= 1) {
                for (size_t t = tid; t < P / 2; t += NMS_THREADS) {
                    const size_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const size_t l = i | j;
                    const uint64_t ki = keys[i], kl = keys[l];
                    const bool up = (i & k) == 0;
                    if ((ki > kl) == up) {
                        keys[i] = kl, keys[l] = ki;
                        const uint32_t si = slot_of[i];
                        slot_of[i] = slot_of[l], slot_of[l] = si;
                    }
                }
                __syncthreads();
            }
        }
    }
    // 3. class-offset boxes in sorted order (ops.py:289-295), truncated to max_nms (ops.py:285-286)
    int m = min(n, max_nms);
    m = min(m, NMS_FLAG_CAP);
    for (int i = tid; i < m; i += NMS_THREADS) {
        const size_t s = cbase + slot_of[i];
        const float4 bx = reinterpret_cast<const float4*>(cd.box)[s];
        const float off = __fmul_rn((float)cd.cls[s], cls_scale);
        sbox[i] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
        dead[i] = 0;
    }
    __shared__ float4 s_pbox[32];
    __shared__ unsigned s_kmask;
    __syncthreads();
    // 4. greedy suppression, 32 candidates per round.  "Pull" form (max_det <= 1024, the kept boxes fit shared memory): warp w
    //    tests chunk member w against every box kept so far (lanes stride over the kept list) and against the later members
    //    of its chunk (one IoU per lane -> one ballot = its row of the 32x32 chunk matrix); warp 0 then resolves the chunk in
    //    order.  Every round costs ~kept/32 + 1 IoUs per thread and two barriers, independent of how long the tail is; the
    //    older "push" form below (kept boxes applied to the whole tail each round) remains for max_det > 1024.
    int kept = 0;
    if (max_det <= 1024) {
        float4* kbox = reinterpret_cast<float4*>(nms_smem);                     // [1024] kept boxes, class-offset
        uint32_t* cmask = reinterpret_cast<uint32_t*>(nms_smem + 1024 * 16);    // [32] rows of the chunk matrix
        uint32_t* calive = cmask + 32;                                          // chunk members no kept box suppresses
        __shared__ int s_nkept;
        if (tid == 0) s_nkept = 0, *calive = 0;
        __syncthreads();
        const int w = tid >> 5, lane = tid & 31;
        for (int cursor = 0; cursor < m; cursor += 32) {
            const int kept0 = s_nkept;  // block-uniform: written before the barrier that ended the previous round
            if (kept0 >= max_det) break;
            const int c = cursor + w, j = cursor + lane;
            const bool valid = c < m;
            const float4 bx = valid ? sbox[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ar = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
            bool hit = false;
            for (int k = lane; k < kept0 && !hit; k += 32) {
                const float4 kb = kbox[k];
                hit = iou_exceeds(kb, __fmul_rn(__fsub_rn(kb.z, kb.x), __fsub_rn(kb.w, kb.y)), bx, ar, iou_thr);
            }
            const bool sup = __any_sync(0xffffffffu, hit);
            const float4 q = j < m ? sbox[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float qa = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
            const unsigned row = __ballot_sync(0xffffffffu, valid && j < m && lane > w && iou_exceeds(bx, ar, q, qa, iou_thr));
            if (lane == 0) {
                cmask[w] = row;
                if (valid && !sup) atomicOr(calive, 1u << w);
            }
            __syncthreads();
            if (w == 0) {  // here q is the box of chunk member `lane`
                const unsigned alive = *calive, myrow = cmask[lane];
                unsigned removed = 0;
                int k2 = kept0;
                for (int l = 0; l < 32 && k2 < max_det; ++l) {
                    const unsigned rl = __shfl_sync(0xffffffffu, myrow, l);
                    if (((alive >> l) & 1u) && !((removed >> l) & 1u)) {  // uniform across the warp
                        removed |= rl;
                        const float kx = __shfl_sync(0xffffffffu, q.x, l), ky = __shfl_sync(0xffffffffu, q.y, l);
                        const float kz = __shfl_sync(0xffffffffu, q.z, l), kw = __shfl_sync(0xffffffffu, q.w, l);
                        if (lane == 0) s_keep[k2] = cursor + l, kbox[k2] = make_float4(kx, ky, kz, kw);
                        ++k2;
                    }
                }
                if (lane == 0) s_nkept = k2, *calive = 0;
            }
            __syncthreads();
        }
        kept = s_nkept;
    }
    // push form:
    int cursor = 0;
    while (max_det > 1024 && kept < max_det && cursor < m) {
        if (tid < 32) {
            const int c = cursor + tid;
            const bool alive = c < m && !dead[c];
            const float4 bx = c < m ? sbox[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ar = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
            unsigned sup = 0;  // later chunk members this lane's box would suppress
            for (int j = 0; j < 32; ++j) {
                const float qx = __shfl_sync(0xffffffffu, bx.x, j), qy = __shfl_sync(0xffffffffu, bx.y, j);
                const float qz = __shfl_sync(0xffffffffu, bx.z, j), qw = __shfl_sync(0xffffffffu, bx.w, j);
                const float qa = __shfl_sync(0xffffffffu, ar, j);
                if (j > tid && iou_exceeds(bx, ar, make_float4(qx, qy, qz, qw), qa, iou_thr)) sup |= 1u << j;
            }
            unsigned live = __ballot_sync(0xffffffffu, alive), kmask = 0;
            int k2 = kept;
            for (int l = 0; l < 32 && k2 < max_det; ++l) {
                const unsigned sl = __shfl_sync(0xffffffffu, sup, l);
                if ((live >> l) & 1u) {
                    kmask |= 1u << l;
                    live &= ~sl;
                    if (tid == 0) s_keep[k2] = cursor + l;
                    ++k2;
                }
            }
            s_pbox[tid] = bx;
            if (tid == 0) s_kmask = kmask;
        }
        __syncthreads();
        const unsigned kmask = s_kmask;
        kept += __popc(kmask);
        if (kept < max_det && kmask) {
            for (int j = cursor + 32 + tid; j < m; j += NMS_THREADS) {
                if (dead[j]) continue;
                const float4 q = sbox[j];
                const float qa = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
                unsigned km = kmask;
                while (km) {
                    const int l = __ffs(km) - 1;
                    km &= km - 1;
                    const float4 pb = s_pbox[l];
                    const float pa = __fmul_rn(__fsub_rn(pb.z, pb.x), __fsub_rn(pb.w, pb.y));
                    if (iou_exceeds(pb, pa, q, qa, iou_thr)) {
                        dead[j] = 1;
                        break;
                    }
                }
            }
        }
        cursor += 32;
        __syncthreads();
    }
    // 5. emit rows in keep order: x1,y1,x2,y2 (no class offset), conf, cls
    for (int k = tid; k < max_det; k += NMS_THREADS) {
        float* ro = out_rows + ((size_t)b * max_det + k) * 6;
        if (k < kept) {
            const size_t s = cbase + slot_of[s_keep[k]];
            emit_row(ro, reinterpret_cast<const float4*>(cd.box)[s], cd.score[s], cd.cls[s], sc);
            out_pos[(size_t)b * max_det + k] = cd.pos[s];
        } else {
            ro[0] = ro[1] = ro[2] = ro[3] = ro[4] = ro[5] = 0.f;
            out_pos[(size_t)b * max_det + k] = -1;
        }
    }
    if (tid == 0) out_count[b] = kept;
}

// ---- validator path: a-priori labels (utils/ops.py:255-261) -------------------------------------------------------------
// The reference appends, per image, one row per ground-truth label AFTER the rows that passed the confidence filter: box =
// xywh2xyxy(label box), score 1.0 on the label's class.  Here: one extra candidate per label with the position key of that row
// (anchors + k, or (anchors + k) * nc + cls with multi_label), so the stable sort orders it exactly like the reference does.
// grid (ceil(max_labels/128), B)
__global__ void __launch_bounds__(128) append_labels_kernel(const float* __restrict__ labels, const int32_t* __restrict__ offs, int anchors,
                                                            int nc, int multi_label, float conf, const int32_t* __restrict__ classes, int ncls,
                                                            dbl_candidates cd) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = offs[b + 1] - offs[b];
    const bool live = k < n;
    float4 box = make_float4(0, 0, 0, 0);
    int c = 0;
    if (live) {
        const float* lb = labels + 5 * (size_t)(offs[b] + k);
        c = (int)lb[0];  // lb[:, 0].long()
        box = xywh_to_xyxy(make_float4(lb[1], lb[2], lb[3], lb[4]));
    }
    const bool ok = live && c >= 0 && c < nc && 1.0f > conf && class_allowed(c, classes, ncls);
    push_candidate(ok, cd, b, box, 1.0f, c, multi_label ? (anchors + k) * nc + c : anchors + k);
}

// ---- validator path: DetectionValidator._process_batch = box_iou + match_predictions ----------------------------------------
// box_iou (utils/metrics.py:52-71): inter / ((area_a + area_b - inter) + eps), strict fp32; match_predictions
// (engine/validator.py:222-260, use_scipy=False) for every IoU threshold t: keep the pairs (label l, detection d) of equal class with
// iou >= thr_t, order them by descending IoU, keep the first pair of every detection, then (in detection order) the first pair of
// every label.  Equivalent closed form: best(d) = arg max_l iou(l, d) over the admissible labels; label l is won by the SMALLEST d
// with best(d) == l; correct[d, t] = (d won best(d)).  Pairs of exactly equal IoU (the reference's unstable argsort leaves their order
// unspecified) resolve to the larger label index.  One CTA per image.
constexpr int MATCH_MAX_THR = 16;
__global__ void __launch_bounds__(256) match_predictions_kernel(const float* __restrict__ rows, const int32_t* __restrict__ counts, int max_det,
                                                                const float* __restrict__ gt_box, const float* __restrict__ gt_cls,
                                                                const int32_t* __restrict__ gt_offs, const float* __restrict__ iouv, int T,
                                                                uint8_t* __restrict__ correct, int32_t* __restrict__ winner_ws, int max_labels) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nd = min(counts[b], max_det), g0 = gt_offs[b], nl = min(gt_offs[b + 1] - g0, max_labels);
    int32_t* winner = winner_ws + (size_t)b * max_labels * T;  // [T][max_labels]
    for (int i = tid; i < nl * T; i += blockDim.x) winner[(i / nl) * max_labels + (i % nl)] = 0x7fffffff;
    for (int i = tid; i < max_det * T; i += blockDim.x) correct[(size_t)b * max_det * T + i] = 0;
    __syncthreads();
    float thr[MATCH_MAX_THR];
#pragma unroll
    for (int t = 0; t < MATCH_MAX_THR; ++t) thr[t] = t < T ? iouv[t] : CUDART_INF_F;
    for (int d0 = 0; d0 < nd; d0 += blockDim.x) {  // block-uniform trip count: the barriers below are reached by every thread
        const int d = d0 + tid;
        int best[MATCH_MAX_THR];
        float biou[MATCH_MAX_THR];
#pragma unroll
        for (int t = 0; t < MATCH_MAX_THR; ++t) best[t] = -1, biou[t] = 0.f;
        if (d < nd) {
            const float* r = rows + ((size_t)b * max_det + d) * 6;
            const float bx1 = r[0], by1 = r[1], bx2 = r[2], by2 = r[3], bc = r[5];
            const float area_b = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
            for (int l = 0; l < nl; ++l) {
                const float4 a = reinterpret_cast<const float4*>(gt_box)[g0 + l];
                const float w = fmaxf(__fsub_rn(fminf(a.z, bx2), fmaxf(a.x, bx1)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(a.w, by2), fmaxf(a.y, by1)), 0.f);
                const float inter = __fmul_rn(w, h);
                const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
                float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-7f));
                iou = gt_cls[g0 + l] == bc ? iou : __fmul_rn(iou, 0.f);  // iou * correct_class
#pragma unroll
                for (int t = 0; t < MATCH_MAX_THR; ++t)
                    if (iou >= thr[t] && iou >= biou[t]) best[t] = l, biou[t] = iou;
            }
#pragma unroll
            for (int t = 0; t < MATCH_MAX_THR; ++t)
                if (t < T && best[t] >= 0) atomicMin(winner + t * max_labels + best[t], d);
        }
        __syncthreads();  // all detections of this image vote before anyone reads: process chunks in order, smaller d first
        // a later chunk can never take a label from an earlier one (atomicMin keeps the smallest d), so reading now is final
        // for every d of this chunk only if no EARLIER d is still to come -- chunks are processed in ascending order.
        if (d < nd) {
#pragma unroll
            for (int t = 0; t < MATCH_MAX_THR; ++t)
                if (t < T && best[t] >= 0 && winner[t * max_labels + best[t]] == d) correct[((size_t)b * max_det + d) * T + t] = 1;
        }
        __syncthreads();
    }
}

static bool valid_cand(const dbl_candidates* c) {
    return c && c->box && c->score && c->cls && c->pos && c->count && c->cap > 0 && c->batch > 0 &&
           (reinterpret_cast<uintptr_t>(c->box) & 15) == 0;
}

}  // namespace dbl

using namespace dbl;

extern "C" int dbl_detect_decode(const dbl_head_levels* lv, float* y, dbl_stream_t stream) {
    LevelGeom g;
    DBL_REQUIRE(make_geom(lv, &g) && y, DBL_E_ARG, "dbl_detect_decode: invalid head description");
    DBL_REQUIRE(g.pitch % 4 == 0, DBL_E_ARG, "dbl_detect_decode: pitch must be a multiple of 4 floats");
    launch_k(decode_kernel, dim3(cdiv(g.A, 256), lv->batch), 256, 0, as_stream(stream), g, y);
    count_launch();
    return check_launch("dbl_detect_decode");
}

extern "C" int dbl_decode_filter(const dbl_head_levels* lv, float conf_thres, int multi_label, const int32_t* classes,
                                 int num_classes_kept, const dbl_candidates* cand, dbl_stream_t stream) {
    LevelGeom g;
    DBL_REQUIRE(make_geom(lv, &g) && valid_cand(cand), DBL_E_ARG, "dbl_decode_filter: invalid argument");
    DBL_REQUIRE(g.pitch % 4 == 0 && cand->batch == lv->batch, DBL_E_ARG, "dbl_decode_filter: pitch/batch mismatch");
    const int ml = (multi_label && g.nc > 1) ? 1 : 0;
    DBL_REQUIRE(cand->cap >= (ml ? g.A * g.nc : g.A), DBL_E_ARG, "dbl_decode_filter: candidate capacity %d too small", cand->cap);
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(cand->count, 0, sizeof(int32_t) * cand->batch, st);
    DBL_REQUIRE(e == cudaSuccess, DBL_E_CUDA, "dbl_decode_filter: memset: %s", cudaGetErrorString(e));
    // (a variant that staged 128 rows per CTA through shared memory with coalesced loads was measured SLOWER: 167 us vs 86 us for
    // 64 x 8400 x 144 logits -- two CTAs per SM and a load -> barrier -> compute sequence per CTA; the thread-per-anchor form with
    // 16-byte loads keeps far more requests in flight)
    launch_k(decode_filter_kernel, dim3(cdiv(g.A, 256), lv->batch), 256, 0, st, g, conf_thres, ml, classes, num_classes_kept, *cand);
    count_launch();
    return check_launch("dbl_decode_filter");
}

extern "C" int dbl_nms_filter(const float* pred, int batch, int nc, int anchors, int64_t stride_b, int64_t stride_c, int64_t stride_a,
                              float conf_thres, int multi_label, const int32_t* classes, int num_classes_kept,
                              const dbl_candidates* cand, dbl_stream_t stream) {
    DBL_REQUIRE(pred && batch > 0 && nc > 0 && anchors > 0 && valid_cand(cand) && cand->batch == batch, DBL_E_ARG,
                "dbl_nms_filter: invalid argument");
    const int ml = (multi_label && nc > 1) ? 1 : 0;
    DBL_REQUIRE(cand->cap >= (ml ? anchors * nc : anchors), DBL_E_ARG, "dbl_nms_filter: candidate capacity %d too small", cand->cap);
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(cand->count, 0, sizeof(int32_t) * batch, st);
    DBL_REQUIRE(e == cudaSuccess, DBL_E_CUDA, "dbl_nms_filter: memset: %s", cudaGetErrorString(e));
    launch_k(filter_kernel, dim3(cdiv(anchors, 256), batch), 256, 0, st, pred, nc, anchors, stride_b, stride_c, stride_a, conf_thres, ml, classes,
                                                                   num_classes_kept, *cand);
    count_launch();
    return check_launch("dbl_nms_filter");
}

extern "C" size_t dbl_nms_sort_capacity(int cap) {
    if (cap <= 0) return 0;
    const size_t P = next_pow2((size_t)cap);
    return P + P / 2 + 2 * P;
}

extern "C" int dbl_nms(const dbl_candidates* cand, float iou_thres, int agnostic, float max_wh, int max_nms, int max_det,
                       uint64_t* sort_keys, float* out_rows, int32_t* out_pos, int32_t* out_count, dbl_stream_t stream) {
    return dbl_nms_scaled(cand, iou_thres, agnostic, max_wh, max_nms, max_det, sort_keys, nullptr, out_rows, out_pos, out_count, stream);
}

extern "C" int dbl_nms_scaled(const dbl_candidates* cand, float iou_thres, int agnostic, float max_wh, int max_nms, int max_det,
                              uint64_t* sort_keys, const dbl_box_scale* scale, float* out_rows, int32_t* out_pos, int32_t* out_count,
                              dbl_stream_t stream) {
    DBL_REQUIRE(valid_cand(cand) && sort_keys && out_rows && out_pos && out_count, DBL_E_ARG, "dbl_nms: invalid argument");
    DBL_REQUIRE(max_det >= 1 && max_det <= 1024, DBL_E_UNSUPPORTED, "dbl_nms: max_det %d not in [1,1024]", max_det);
    DBL_REQUIRE(max_nms >= 1, DBL_E_ARG, "dbl_nms: max_nms");
    DBL_REQUIRE(std::min(cand->cap, max_nms) <= NMS_FLAG_CAP, DBL_E_UNSUPPORTED,
                "dbl_nms: min(cap,max_nms)=%d exceeds the %d candidates one image may hold", std::min(cand->cap, max_nms), NMS_FLAG_CAP);
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(sort_keys) & 15) == 0, DBL_E_ARG, "dbl_nms: scratch must be 16-byte aligned");
    static DevOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(sort_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_SMEM);
        DBL_REQUIRE(e == cudaSuccess, DBL_E_CUDA, "dbl_nms: shared memory opt-in failed: %s", cudaGetErrorString(e));
        once.mark();
    }
    launch_k(sort_nms_kernel, cand->batch, NMS_THREADS, NMS_SMEM, as_stream(stream), *cand, iou_thres, agnostic ? 0.f : max_wh, max_nms, max_det,
             sort_keys, dbl_nms_sort_capacity(cand->cap), reinterpret_cast<const float*>(scale), out_rows, out_pos, out_count);
    count_launch();
    return check_launch("dbl_nms");
}

extern "C" int dbl_nms_append_labels(const float* labels, const int32_t* label_offsets, int max_labels_per_image, int anchors, int nc,
                                     int multi_label, float conf_thres, const int32_t* classes, int num_classes_kept,
                                     const dbl_candidates* cand, dbl_stream_t stream) {
    DBL_REQUIRE(labels && label_offsets && valid_cand(cand) && anchors > 0 && nc > 0 && max_labels_per_image >= 0, DBL_E_ARG,
                "dbl_nms_append_labels: invalid argument");
    if (max_labels_per_image == 0) return DBL_OK;
    const int ml = (multi_label && nc > 1) ? 1 : 0;
    DBL_REQUIRE((long long)(anchors + max_labels_per_image) * (ml ? nc : 1) < (1ll << 31), DBL_E_UNSUPPORTED,
                "dbl_nms_append_labels: position keys exceed 31 bits");
    DBL_REQUIRE(cand->batch <= 65535, DBL_E_UNSUPPORTED, "dbl_nms_append_labels: batch too large");
    launch_k(append_labels_kernel, dim3(cdiv(max_labels_per_image, 128), cand->batch), 128, 0, as_stream(stream), labels, label_offsets, anchors, nc,
             ml, conf_thres, classes, num_classes_kept, *cand);
    count_launch();
    return check_launch("dbl_nms_append_labels");
}

extern "C" size_t dbl_match_workspace_bytes(int batch, int max_labels_per_image, int num_thresholds) {
    if (batch <= 0 || max_labels_per_image <= 0 || num_thresholds <= 0) return 0;
    return sizeof(int32_t) * (size_t)batch * max_labels_per_image * num_thresholds;
}

extern "C" int dbl_match_predictions(const float* rows, const int32_t* counts, int batch, int max_det, const float* gt_boxes,
                                     const float* gt_cls, const int32_t* gt_offsets, int max_labels_per_image, const float* iou_thresholds,
                                     int num_thresholds, uint8_t* correct, void* ws, size_t ws_bytes, dbl_stream_t stream) {
    DBL_REQUIRE(rows && counts && gt_offsets && iou_thresholds && correct && batch > 0 && max_det > 0, DBL_E_ARG,
                "dbl_match_predictions: invalid argument");
    DBL_REQUIRE(num_thresholds >= 1 && num_thresholds <= MATCH_MAX_THR, DBL_E_UNSUPPORTED, "dbl_match_predictions: 1..%d thresholds", MATCH_MAX_THR);
    DBL_REQUIRE(max_labels_per_image >= 0 && (max_labels_per_image == 0 || (gt_boxes && gt_cls)), DBL_E_ARG, "dbl_match_predictions: labels");
    DBL_REQUIRE((reinterpret_cast<uintptr_t>(gt_boxes) & 15) == 0, DBL_E_ARG, "dbl_match_predictions: gt_boxes must be 16-byte aligned");
    const int ml = max_labels_per_image > 0 ? max_labels_per_image : 1;
    DBL_REQUIRE(ws && ws_bytes >= dbl_match_workspace_bytes(batch, ml, num_thresholds), DBL_E_WORKSPACE, "dbl_match_predictions: workspace too small");
    launch_k(match_predictions_kernel, batch, 256, 0, as_stream(stream), rows, counts, max_det, gt_boxes, gt_cls, gt_offsets, iou_thresholds,
             num_thresholds, correct, reinterpret_cast<int32_t*>(ws), ml);
    count_launch();
    return check_launch("dbl_match_predictions");
}
