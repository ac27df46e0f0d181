/* CPU oracle for greedy NMS -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates the published algorithm of torchvision.ops.nms (CPU kernel), the third-party routine the reference
 * calls at ultralytics/utils/ops.py:296.  torchvision is NOT vendored under /root/reference (pinned
 * torchvision==0.23.0 in requirements.txt:2; this container has 0.26.0).  Parity is pinned by
 * tests/test_oracle_golden.py against outputs of the reference's own non_max_suppression run in the authoring
 * container (tests/golden/nms_*.npz) and against torchvision.ops.nms itself on random/tied inputs.
 *
 * Algorithm: areas = (x2-x1)*(y2-y1); visit boxes in STABLE descending score order; an unsuppressed box i is kept and
 * suppresses every later box j with inter/(area_i+area_j-inter) > thr, inter = max(0,xx2-xx1)*max(0,yy2-yy1).
 * All arithmetic is strict IEEE fp32 (compile with -ffp-contract=off: no FMA contraction).
 *
 * `order` = indices sorted by descending score (stable), computed by the caller.
 * Returns the number of kept boxes; keep[] receives original indices in score order.
 */
#include <stdint.h>
#include <stdlib.h>

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

int64_t dbl_oracle_nms(const float* boxes, const int64_t* order, int64_t n, float thr, int64_t* keep) {
    unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    float* area = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        volatile float w = b[2] - b[0], h = b[3] - b[1];
        area[i] = w * h;
    }
    int64_t kept = 0;
    for (int64_t oi = 0; oi < n; ++oi) {
        int64_t i = order[oi];
        if (dead[i]) continue;
        keep[kept++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (int64_t oj = oi + 1; oj < n; ++oj) {
            int64_t j = order[oj];
            if (dead[j]) continue;
            const float* b = boxes + 4 * j;
            float xx1 = fmaxf_(ix1, b[0]), yy1 = fmaxf_(iy1, b[1]);
            float xx2 = fminf_(ix2, b[2]), yy2 = fminf_(iy2, b[3]);
            volatile float w = fmaxf_(0.0f, xx2 - xx1);
            volatile float h = fmaxf_(0.0f, yy2 - yy1);
            volatile float inter = w * h;
            volatile float uni = ia + area[j];
            uni = uni - inter;
            volatile float ovr = inter / uni;
            if (ovr > thr) dead[j] = 1;
        }
    }
    free(dead);
    free(area);
    return kept;
}
