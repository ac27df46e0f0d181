// Arithmetic shared by the Detect post-process kernels (detect_post.cu) and the fused head (head_fused.cu).  Everything that
// decides WHICH boxes survive is written with explicit round-to-nearest intrinsics (no multiply-add contraction).
#pragma once
#include <math_constants.h>

#include "dbl_common.cuh"

namespace dbl {

constexpr int REG_MAX = 16;

// ---- shared arithmetic -------------------------------------------------------------------------------------------
// One DFL side: expectation of softmax over 16 bins (block.py:80-83).
__device__ __forceinline__ float dfl_side(const float* __restrict__ p) {
    float v[REG_MAX];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
        v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
    }
    float m = v[0];
#pragma unroll
    for (int j = 1; j < REG_MAX; ++j) m = fmaxf(m, v[j]);
    float s = 0.f, e = 0.f;
#pragma unroll
    for (int j = 0; j < REG_MAX; ++j) {
        const float t = expf(v[j] - m);
        s += t;
        e = fmaf((float)j, t, e);
    }
    return e / s;
}

// dist2bbox(xywh=True) * stride (tal.py:348-357, head.py:178)
__device__ __forceinline__ float4 dist_to_xywh(float l, float t, float r, float b, float ax, float ay, float stride) {
    const float x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b;
    return make_float4((x1 + x2) * 0.5f * stride, (y1 + y2) * 0.5f * stride, (x2 - x1) * stride, (y2 - y1) * stride);
}

// xywh2xyxy (ops.py:416-433): half = wh/2, xy -+ half, each step rounded once.
__device__ __forceinline__ float4 xywh_to_xyxy(float4 v) {
    const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
    return make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
}

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ bool class_allowed(int c, const int32_t* __restrict__ classes, int ncls) {
    if (!classes) return true;
    for (int i = 0; i < ncls; ++i)
        if (classes[i] == c) return true;
    return false;
}

// Append one candidate; one atomic per warp (all lanes of a warp belong to the same image).
__device__ __forceinline__ void push_candidate(bool pred, const dbl_candidates& cd, int b, float4 box, float score, int cls, int pos) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cd.count + b, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        const int slot = base + __popc(mask & ((1u << lane) - 1));
        if (slot < cd.cap) {
            const size_t o = (size_t)b * cd.cap + slot;
            reinterpret_cast<float4*>(cd.box)[o] = box;
            cd.score[o] = score;
            cd.cls[o] = cls;
            cd.pos[o] = pos;
        }
    }
}


// DFL expectation of one side from 16 logits already in registers: the SAME operation order as dfl_side()
__device__ __forceinline__ float dfl_side_regs(const float (&v)[REG_MAX]) {
    float m = v[0];
#pragma unroll
    for (int j = 1; j < REG_MAX; ++j) m = fmaxf(m, v[j]);
    float s = 0.f, e = 0.f;
#pragma unroll
    for (int j = 0; j < REG_MAX; ++j) {
        const float t = expf(v[j] - m);
        s += t;
        e = fmaf((float)j, t, e);
    }
    return e / s;
}

}  // namespace dbl
