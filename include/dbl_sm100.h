/* libdbl_sm100.so -- C ABI of the B200-native YOLO-DBL inference path.
 *
 * The reference (player4771/YOLO-DBL, a fork of ultralytics 8.3.63) has no C boundary: its hot path is Python that
 * calls torch / torchvision library ops.  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md).  Each entry point names the reference code it replaces (paths relative to
 * models/YOLO/ultralytics/ of the reference).
 *
 * Conventions (seam B4)
 *   - plain pointers and sizes only; no torch / C++ types; every function returns 0 on success or a DBL_E_* code and
 *     never throws, aborts or calls cudaDeviceSynchronize.  dbl_last_error() returns a per-thread message.
 *   - the CALLER owns every buffer including scratch; the library never allocates or frees device memory.
 *   - every launch goes to the stream that is passed in, so a whole forward pass is CUDA-graph capturable.
 *   - there is NO CPU fallback: without an sm_100 device every compute call fails with DBL_E_CUDA.
 *
 * Activation layout: NHWC, channels innermost.  A dbl_tensor may be a channel slice of a wider buffer
 * (`pitch` = elements per pixel of the underlying buffer), which is how concat/chunk cost nothing.
 */
#ifndef DBL_SM100_H
#define DBL_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dbl_stream_t; /* a cudaStream_t */

enum { DBL_OK = 0, DBL_E_ARG = 1, DBL_E_UNSUPPORTED = 2, DBL_E_CUDA = 3, DBL_E_WORKSPACE = 4 };
enum { DBL_BF16 = 0, DBL_F32 = 1 };
enum { DBL_ACT_NONE = 0, DBL_ACT_SILU = 1, DBL_ACT_GELU = 2 };
enum { DBL_EPI_NONE = 0, DBL_EPI_ADD = 1 /* y = act(conv)+r */, DBL_EPI_MUL = 2 /* y = act(conv)*r */ };
enum {
    DBL_IMPL_AUTO = 0,
    DBL_IMPL_DIRECT = 1,      /* CUDA-core direct convolution */
    DBL_IMPL_TCGEN05 = 2,     /* tensor cores: halo-reuse kernel for 3x3, TMA kernel for 1x1 */
    DBL_IMPL_TCGEN05_TMA = 3, /* tensor cores, first-generation kernel (one TMA box per filter tap) */
    DBL_IMPL_TCGEN05_MULTI = 4 /* tensor cores, 3x3 stride 1: halo reuse + weights streamed once per multi-accumulator super-tile */
};

typedef struct dbl_tensor {
    void* data;    /* device pointer to element (n=0,h=0,w=0,c=0) of this view */
    int32_t n, h, w, c;
    int32_t pitch; /* elements between consecutive pixels (>= c) */
    int32_t dtype; /* DBL_BF16 or DBL_F32 */
} dbl_tensor;

const char* dbl_last_error(void);
int dbl_version(void);
/* 1 if the current CUDA device can run this library (compute capability 10.x), else 0. */
int dbl_device_ok(void);
/* Programmatic dependent launch (on by default): kernels are launched so that the prologue of kernel i+1 overlaps the
 * tail of kernel i on the same stream / graph branch.  Returns the previous setting. */
int dbl_set_pdl(int enabled);
/* number of kernels this library has launched from the calling process (bench.py reports it as gpu_launches) */
int64_t dbl_launch_count(void);

/* ---- dense convolutions ------------------------------------------------------------------------------------------
 * Conv.forward_fuse + fuse_conv_and_bn (nn/modules/conv.py:39-63, utils/torch_utils.py:238-265), the pointwise half of
 * DSConv (conv.py:91-108), the biased nn.Conv2d 1x1s of LSKblock / DySample / Detect, Bottleneck's residual add
 * (nn/modules/block.py:344-357) and LSKblock's final x*attn (nn/modules_attention/LSKA.py:52).
 *   x: bf16 NHWC (n,h,w,cin)      w: bf16 [cout][k][k][cin] (BatchNorm already folded)      bias: f32 [cout] or NULL
 *   r: bf16 NHWC like y, or NULL  y: bf16 or f32 NHWC (n,ho,wo,cout), ho = (h+2*pad-dil*(k-1)-1)/stride+1
 * impl = DBL_IMPL_TCGEN05 requires dil == 1, k in {1,3}, stride in {1,2} (k==1: stride 1), cin % 16 == 0 (k==3 also
 * accepts cin == 8). */
int dbl_conv2d(const dbl_tensor* x, const void* w, const float* bias, const dbl_tensor* r, const dbl_tensor* y,
               int k, int stride, int pad, int dil, int act, int epi, int impl, dbl_stream_t stream);

/* Which kernel dbl_conv2d runs for this problem: 0 CUDA-core direct, 1 tcgen05 + TMA (1x1 GEMM / one box per tap), 3 tcgen05 halo reuse
 * with resident weights, 4 tcgen05 halo reuse with multi-accumulator super-tiles, 5 tcgen05 thin stride-2 (pixel-pair planes), 6 tcgen05
 * stride-1 3x3 whose whole reduction depth is one shared-memory tile; -1 for invalid tensors. */
int dbl_conv2d_which(const dbl_tensor* x, const dbl_tensor* y, int k, int stride, int pad, int dil, int impl);

/* Layer 0 of the graph (cfg/models/v13/yolov13_DBL.yaml:9, Conv [32,3,1]) fused with the input cast of
 * BasePredictor.preprocess (engine/predictor.py:116-134): x is the user's f32 NCHW image batch (n,3,h,w) in 0..1,
 * w f32 [cout][3][3][3], y bf16 NHWC, SiLU applied. */
int dbl_stem_conv(const float* x_nchw, int n, int h, int w_, const float* w, const float* bias, const dbl_tensor* y,
                  int stride, dbl_stream_t stream);
/* Same for the image sources of the reference: x is uint8 NCHW 0..255 and the `/ 255` of BasePredictor.preprocess
 * (engine/predictor.py:131-133) happens in the loader of the kernel, so only one byte per sample crosses PCIe. */
int dbl_stem_conv_u8(const uint8_t* x_nchw, int n, int h, int w_, const float* w, const float* bias, const dbl_tensor* y,
                     int stride, dbl_stream_t stream);

/* ---- image-source pre-process -----------------------------------------------------------
 * LetterBox.__call__ (data/augment.py:1558-1600, auto=False, centre, pad value 114) + BGR->RGB + HWC->CHW of
 * BasePredictor.preprocess (engine/predictor.py:116-134) for a batch of uint8 HWC BGR images of different sizes in one launch; the
 * bilinear resize repeats cv2.resize(INTER_LINEAR)'s 8-bit fixed-point arithmetic.  `table_dev` is a DEVICE array of `batch`
 * descriptors (src = device pointer to the image, pitch = bytes per source row, new_h/new_w = resized size, top/left = where it
 * goes inside the H x W output).  dst: uint8 NCHW (batch,3,H,W), RGB -- exactly what dbl_stem_pair(x_is_u8 = 1) consumes. */
typedef struct dbl_lb_image {
    const void* src;
    int32_t h, w, pitch, new_h, new_w, top, left, reserved;
} dbl_lb_image;
int dbl_letterbox_u8(const dbl_lb_image* table_dev, int batch, uint8_t* dst_nchw, int H, int W, int pad_value, dbl_stream_t stream);

/* ---- thin high-resolution layers: warp-level MMA kernels (csrc/conv_mma.cu) --------------------------------------
 * Layers whose arithmetic intensity is far below the B200 ridge (3..64 channels at 640^2..80^2) are HBM-bound; these
 * entry points fuse neighbouring reference modules so every activation crosses HBM once.
 * Weights are consumed in MMA B-fragment order; dbl_mma_pack_weights produces that layout ONCE from the dbl_conv2d
 * layout (bf16 [cout][taps][cin], BatchNorm folded) into caller-owned memory of dbl_mma_packed_bytes() bytes.
 * cin must be 8 or a multiple of 16, cout a multiple of 8. */
size_t dbl_mma_packed_bytes(int cout, int cin, int taps);
int dbl_mma_pack_weights(const void* w, int cout, int cin, int taps, void* packed, dbl_stream_t stream);
/* Bottleneck.forward (nn/modules/block.py:344-357) with e=0.5, k=(3,3): y = x + SiLU(cv2(SiLU(cv1(x)))) in one
 * kernel; cv1: C -> C/2, cv2: C/2 -> C, C in {16,32,64}; the hidden tensor stays in shared memory.  shortcut = 0 drops
 * the residual.  x and y must not alias. */
int dbl_bottleneck3x3(const dbl_tensor* x, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                      int shortcut, const dbl_tensor* y, dbl_stream_t stream);
/* The same Bottleneck on tcgen05 (csrc/bottleneck_tc.cu) for C in {16, 32}: a GEMM row is a group of 64/C consecutive pixels (one 128-byte
 * row of the NHWC tensor), the filters are band matrices over the group's window, the hidden tensor stays in shared memory.  x and y must
 * be dense (pitch == C) with W % (64/C) == 0 -- dbl_bottleneck_tc_supported(c, w, pitch) -- and must not alias.  The band matrices and
 * biases are produced ONCE by dbl_bottleneck_tc_pack from the dbl_conv2d layouts (w1: bf16 [C/2][3][3][C], w2: bf16 [C][3][3][C/2], BN
 * folded) into caller-owned, 16-byte-aligned memory of dbl_bottleneck_tc_packed_bytes(c) bytes. */
int dbl_bottleneck_tc_supported(int c, int w, int pitch);
size_t dbl_bottleneck_tc_packed_bytes(int c);
int dbl_bottleneck_tc_pack(const void* w1, const float* b1, const void* w2, const float* b2, int c, void* packed, dbl_stream_t stream);
int dbl_bottleneck3x3_tc(const dbl_tensor* x, const void* packed, int shortcut, const dbl_tensor* y, dbl_stream_t stream);
/* One thin 3x3 Conv (pad 1, stride 1|2), act in {NONE, SILU}, optional residual; (cin,cout,stride) from the set the
 * kernel is instantiated for (else DBL_E_UNSUPPORTED): the stride-2 down-samplers 8->16, 16->32, 32->64 and the
 * stride-1 pairs 16<->8, 32<->16, 64<->32. */
int dbl_conv3x3_mma(const dbl_tensor* x, const void* w_packed, const float* bias, const dbl_tensor* r, const dbl_tensor* y,
                    int stride, int act, int epi, dbl_stream_t stream);
/* Layers 0 and 1 of the graph fused (cfg/models/v13/yolov13_DBL.yaml:9-10: Conv [c0,3,1] then Conv [c1,3,2]) together
 * with the input cast / scale of BasePredictor.preprocess (engine/predictor.py:116-134).  x: NCHW (n,3,h,w), f32 in 0..1
 * (x_is_u8 = 0) or uint8 0..255 (x_is_u8 = 1; the /255 is folded into layer 0's weights).  w0: f32 [c0][3][3][3], w1:
 * packed (c1, c0, 9 taps).  y: bf16 NHWC (n, h/2, w/2, c1).  (c0,c1) in {(8,16), (16,32)}.  h and w even; x aligned to 8 bytes
 * (f32) / 2 bytes (u8): the kernel loads pixel pairs (DBL_E_ARG otherwise). */
int dbl_stem_pair(const void* x_nchw, int x_is_u8, int n, int h, int w, const float* w0, const float* b0, int c0,
                  const void* w1_packed, const float* b1, const dbl_tensor* y, dbl_stream_t stream);

/* The same stem pair on tcgen05 (csrc/stem_tc.cu) for (c0, c1) = (8, 16): 8-pixel groups as GEMM rows, band-matrix filters, layer 0's output
 * in shared memory.  y must be dense (pitch == 16), h even, w % 16 == 0, x 16-byte aligned -- dbl_stem_tc_supported(c0, c1, h, w, ypitch).  The band matrices are
 * produced ONCE by dbl_stem_tc_pack (w0: f32 [8][3][3][3] = [c0][ci][ky][kx], w1: bf16 [16][3][3][8] in the dbl_conv2d layout, BN folded;
 * x_is_u8 folds the /255 into layer 0) into caller-owned, 16-byte-aligned memory of dbl_stem_tc_packed_bytes() bytes; a blob packed for
 * u8 input must be used with x_is_u8 = 1. */
int dbl_stem_tc_supported(int c0, int c1, int h, int w, int ypitch);
size_t dbl_stem_tc_packed_bytes(void);
int dbl_stem_tc_pack(const float* w0, const float* b0, const void* w1, const float* b1, int x_is_u8, void* packed, dbl_stream_t stream);
int dbl_stem_pair_tc(const void* x_nchw, int x_is_u8, int n, int h, int w, const void* packed, const dbl_tensor* y, dbl_stream_t stream);

/* ---- fused depthwise-separable convolution (csrc/dsconv_mma.cu) -------------------------------------------------
 * DSConv.forward (nn/modules/conv.py:91-108): y = SiLU(pw(dw(x)) + b_pw) [+|* r], depthwise k x k (pad (k-1)/2, stride 1|2) and
 * pointwise 1x1 in ONE kernel; with dw_act = 1 the depthwise stage is DWConv + BN + SiLU (bias b_dw, Detect's class branch,
 * nn/modules/head.py:93-100).  w_dw: f32 [k*k][cin]; w_pw_packed: dbl_mma_pack_weights of bf16 [cout][1][cin] (BN folded).
 * Only the (k, stride, cin, cout) combinations dbl_dsconv_supported() reports are instantiated.  x and y must not alias
 * (r may alias y). */
int dbl_dsconv_supported(int k, int stride, int cin, int cout);
int dbl_dsconv(const dbl_tensor* x, const float* w_dw, const float* b_dw, int dw_act, const void* w_pw_packed, const float* b_pw,
               const dbl_tensor* r, const dbl_tensor* y, int k, int stride, int epi, dbl_stream_t stream);

/* The same operation on the Blackwell data path (csrc/dsconv_tc.cu), stride 1: a persistent warp-specialised kernel -- TMA halo
 * tiles (zero fill = padding), depthwise on the fp32 pipes writing the UMMA-swizzled A tile in place, pointwise on tcgen05 with the
 * accumulator in TMEM, TMA-store epilogue -- whose phases overlap across consecutive tiles.  Differences to dbl_dsconv:
 * w_pw is the PLAIN bf16 [cout][cin] matrix (the dbl_conv2d layout, BN folded), not the mma.sync fragment order; dw_act and a residual
 * are not combined; the stride (1 or 2) is implied by the shape of y.  (k, stride, cin, cout) from dbl_dsconv_tc_supported(); x and y
 * must not alias (r may alias y). */
int dbl_dsconv_tc_supported(int k, int stride, int cin, int cout);
int dbl_dsconv_tc(const dbl_tensor* x, const float* w_dw, const float* b_dw, int dw_act, const void* w_pw, const float* b_pw,
                  const dbl_tensor* r, const dbl_tensor* y, int k, int epi, dbl_stream_t stream);
/* Developer aid (tools/ds_trace.py, tools/t5_trace.py): with a non-NULL DEVICE buffer of the returned size, CTA 0 of every later dbl_dsconv_tc
 * launch (and of the stride-1 single-tile 3x3 kernel behind dbl_conv2d)
 * records %globaltimer stamps [role: producer, mma, depthwise, epilogue][first 16 tiles][4 events]; NULL switches it off. */
int dbl_dsconv_tc_debug_trace(void* device_buffer);

/* ---- depthwise convolution ---------------------------------------------------------------------------------------
 * DSConv.dw (conv.py:96-99), DWConv (conv.py:128-133), LSKblock.conv0 / conv_spatial (LSKA.py:31-32).
 *   w: f32 [k*k][c]   bias: f32 [c] or NULL   pad = dil*(k-1)/2 */
int dbl_dwconv2d(const dbl_tensor* x, const float* w, const float* bias, const dbl_tensor* y, int k, int stride, int dil,
                 int act, dbl_stream_t stream);

/* ---- bandwidth-bound glue ---------------------------------------------------------------------------------------- */
/* nn.AvgPool2d(2) of FuseModule / DownsampleConv (block.py:1825, 1919) */
int dbl_avgpool2(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t stream);
/* nn.Upsample(scale_factor=2, mode='nearest') of FuseModule (block.py:1826) */
int dbl_upsample2_nearest(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t stream);
/* FullPAD_Tunnel.forward (block.py:1950-1956): y = a + gate*b.  y2 (optional) receives a second copy, which is how a
 * tunnel output lands in its consumer's concat buffer as well as in its own. */
int dbl_gated_add(const dbl_tensor* a, const dbl_tensor* b, float gate, const dbl_tensor* y, const dbl_tensor* y2,
                  dbl_stream_t stream);
/* A2C2f's learnable residual (nn/modules/block.py:1403-1404): y = a + gamma[c] * b, gamma f32 [c] (16-byte aligned); y may alias a or b */
int dbl_channel_scale_add(const dbl_tensor* a, const dbl_tensor* b, const float* gamma, const dbl_tensor* y, dbl_stream_t stream);
/* copy a view (Concat of a tensor that could not be produced in place, conv.py:349-359) */
int dbl_copy(const dbl_tensor* x, const dbl_tensor* y, dbl_stream_t stream);
/* DySample.forward_lp + sample (nn/modules_upsample/DySample.py:49-68), scale 2.  off: f32 NHWC (n,h,w,2*groups*4) =
 * raw output of the `offset` 1x1 conv incl. bias; the kernel applies *0.25 + init_pos, the pixel shuffle and the
 * bilinear border-clamped gather (grid_sample align_corners=False).  y: (n,2h,2w,c). */
int dbl_dysample(const dbl_tensor* x, const dbl_tensor* off, const dbl_tensor* y, int groups, dbl_stream_t stream);
/* LSKblock (LSKA.py:43-50): attn = [attn1|attn2] bf16 (n,h,w,2*half); agg f32 (n,h,w,2) <- channel mean / max */
int dbl_lsk_squeeze(const dbl_tensor* attn, const dbl_tensor* agg, dbl_stream_t stream);
/* sig = sigmoid(conv7x7(agg; w_sq f32 [2][2][7][7], b_sq[2])); y = attn1*sig0 + attn2*sig1  (n,h,w,half) bf16 */
int dbl_lsk_gate(const dbl_tensor* attn, const dbl_tensor* agg, const float* w_sq, const float* b_sq, const dbl_tensor* y,
                 dbl_stream_t stream);

/* ---- area attention (stock yolov13.yaml) ----------------------------------------------------
 * The core of AAttn.forward (nn/modules/block.py:1236-1262) between the qkv convolution and the pe / proj convolutions.
 *   qkv: bf16 NHWC (n,h,w,3*heads*32) as the 1x1 qkv Conv writes it (per head [q | k | v], 32 channels each); tokens = pixels in
 *   row-major order, split into `area` contiguous bands that attend within themselves (h*w %% area == 0).
 *   o (n,h,w,heads*32) = softmax(q k^T / sqrt(32)) v per head; v_out (n,h,w,heads*32) = the v part (input of the pe depthwise conv). */
int dbl_area_attention(const dbl_tensor* qkv, const dbl_tensor* o, const dbl_tensor* v_out, int heads, int area, dbl_stream_t stream);

/* ---- adaptive hypergraph (AdaHGComputation / AdaHGConv / AdaHyperedgeGen, block.py:1582-1752) -------------------
 * x: bf16 tokens (n, h, w, d) == (B, N=h*w, D).  All weights f32, row-major [out][in] like nn.Linear.
 * Scratch `ws` must hold dbl_hypergraph_workspace_bytes(B, N, D, E) bytes.  y = node_proj(A·edge_proj(AᵀX)) + X. */
typedef struct dbl_hypergraph_weights {
    const float* prototype_base; /* [E][D] */
    const float* context_w;      /* [E*D][2D] */
    const float* context_b;      /* [E*D] */
    const float* pre_head_w;     /* [D][D] */
    const float* pre_head_b;     /* [D] */
    const float* edge_w;         /* [D][D] */
    const float* edge_b;         /* [D] */
    const float* node_w;         /* [D][D] */
    const float* node_b;         /* [D] */
    int32_t num_edges;           /* E */
    int32_t num_heads;
} dbl_hypergraph_weights;
size_t dbl_hypergraph_workspace_bytes(int batch, int tokens, int dim, int edges);
int dbl_hypergraph(const dbl_tensor* x, const dbl_hypergraph_weights* wts, const dbl_tensor* y, void* ws, size_t ws_bytes,
                   dbl_stream_t stream);

/* ---- Detect post-process -----------------------------------------------------------------------------------------
 * Head logits live in one f32 NHWC tensor per level: channels [0,64) = cv2 (4 sides x 16 DFL bins), [64,64+nc) = cv3. */
typedef struct dbl_head_levels {
    const float* data[3];
    int32_t h[3], w[3];
    float stride[3];
    int32_t num_levels, batch, nc, pitch; /* pitch = elements per pixel (>= 64+nc) */
} dbl_head_levels;

/* Detect._inference (nn/modules/head.py:143-181) = DFL (block.py:80-83) + make_anchors (utils/tal.py:333-345) +
 * dist2bbox (tal.py:348-357) + sigmoid.  y: f32 [B][4+nc][A] exactly like the reference's first output. */
int dbl_detect_decode(const dbl_head_levels* lv, float* y, dbl_stream_t stream);

/* ---- fused Detect tail (csrc/head_fused.cu) ---------------------------------------------------------------------
 * The last 1x1 convolutions of both head branches (nn/modules/head.py:86-100: nn.Conv2d(c2, 64, 1) and nn.Conv2d(c3, nc, 1))
 * + dbl_detect_decode + dbl_decode_filter in one kernel: the 64+nc fp32 logits per anchor never reach HBM.
 *   f2 / f3: bf16 NHWC branch features (c2 = 64 and c3 channels); w2_packed / w3_packed: dbl_mma_pack_weights layout of the
 *   [64][1][c2] and [ceil8(nc)][1][c3] weights (rows >= nc zero); b2 / b3: f32 biases.
 * Outputs, each optional: level[i].logits (f32 NHWC, `logit_pitch` floats per pixel: what forward() returns as feats),
 * y (f32 [B][4+nc][A], the reference's first output) and `cand` (filled exactly like dbl_decode_filter does). */
typedef struct dbl_head_fused_level {
    const void* f2;
    const void* f3;
    const void* w2_packed;
    const void* w3_packed;
    const float* b2;
    const float* b3;
    float* logits;
    int32_t pitch2, pitch3, h, w;
    float stride;
} dbl_head_fused_level;
typedef struct dbl_head_fused {
    dbl_head_fused_level level[3];
    int32_t num_levels, batch, nc, c2, c3, logit_pitch;
} dbl_head_fused;

/* Candidate lists: per image up to `cap` entries, unordered; pos = anchor*npos_per_anchor + k is the position the
 * entry has in the reference's filtered list and is the tie-break key of the stable sort. */
typedef struct dbl_candidates {
    float* box;      /* [B][cap][4] xyxy, NOT class-offset */
    float* score;    /* [B][cap] */
    int32_t* cls;    /* [B][cap] */
    int32_t* pos;    /* [B][cap] */
    int32_t* count;  /* [B]  (zeroed by the filter call) */
    int32_t cap, batch;
} dbl_candidates;

/* non_max_suppression stage 1 (utils/ops.py:234-279): xc = amax > conf, xywh2xyxy (ops.py:416-433), best class (or
 * multi_label expansion), optional class filter.  pred: f32, element (b,ch,a) at b*stride_b + ch*stride_c + a*stride_a
 * (the reference's y is stride_c = A, stride_a = 1).  classes: device int32 list or NULL. */
int dbl_nms_filter(const float* pred, int batch, int nc, int anchors, int64_t stride_b, int64_t stride_c, int64_t stride_a,
                   float conf_thres, int multi_label, const int32_t* classes, int num_classes_kept,
                   const dbl_candidates* cand, dbl_stream_t stream);
/* decode + filter fused: same candidates as dbl_detect_decode -> dbl_nms_filter without materialising y. */
int dbl_decode_filter(const dbl_head_levels* lv, float conf_thres, int multi_label, const int32_t* classes,
                      int num_classes_kept, const dbl_candidates* cand, dbl_stream_t stream);
/* 1 when dbl_head_decode has an instantiation for these branch widths / class count */
int dbl_head_decode_supported(int c2, int c3, int nc);
int dbl_head_decode(const dbl_head_fused* hd, float* y, float conf_thres, int multi_label, const int32_t* classes,
                    int num_classes_kept, const dbl_candidates* cand, dbl_stream_t stream);
/* non_max_suppression stage 2 (ops.py:281-297) + torchvision.ops.nms: stable descending sort, max_nms truncation,
 * boxes + cls*max_wh, greedy IoU suppression in strict fp32, max_det cut.
 *   out_rows f32 [B][max_det][6] (x1,y1,x2,y2,conf,cls), out_pos int32 [B][max_det], out_count int32 [B]
 *   sort_keys: scratch uint64 [B][dbl_nms_sort_capacity(cap)] */
size_t dbl_nms_sort_capacity(int cap);
int dbl_nms(const dbl_candidates* cand, float iou_thres, int agnostic, float max_wh, int max_nms, int max_det,
            uint64_t* sort_keys, float* out_rows, int32_t* out_pos, int32_t* out_count, dbl_stream_t stream);


/* Same, with DetectionPredictor.postprocess's tail fused into the epilogue: scale_boxes + clip_boxes
 * (utils/ops.py:92-127, 319-338; called at models/yolo/detect/predict.py:38).  scale: DEVICE array of `batch` records, or NULL
 * (= dbl_nms).  A kept box (x1,y1,x2,y2) leaves as clamp((x - pad_x) / gain, 0, w0), clamp((y - pad_y) / gain, 0, h0): the
 * letterbox padding and gain of that image (pad = round((W - w0*gain)/2 - 0.1) etc., computed by the caller like the reference
 * does in Python) and the original image size.  IEEE fp32 subtract / divide / clamp, bit-identical to the reference on CPU. */
typedef struct dbl_box_scale {
    float pad_x, pad_y, gain, w0, h0;
} dbl_box_scale;
int dbl_nms_scaled(const dbl_candidates* cand, float iou_thres, int agnostic, float max_wh, int max_nms, int max_det,
                   uint64_t* sort_keys, const dbl_box_scale* scale, float* out_rows, int32_t* out_pos, int32_t* out_count,
                   dbl_stream_t stream);

/* ---- validator path ----------------------------------------------------------------------------
 * A-priori labels of non_max_suppression (utils/ops.py:255-261; DetectionValidator.postprocess passes labels=self.lb,
 * models/yolo/detect/val.py:92-102): after a filter call (dbl_nms_filter / dbl_decode_filter / dbl_head_decode) has filled
 * `cand`, append one candidate per ground-truth label: box = xywh2xyxy(label), score 1.0, class = label class, ordered after every
 * prediction row like the reference's torch.cat.  labels: DEVICE f32 [L][5] = (cls, x, y, w, h) in network-input pixels, all
 * images concatenated; label_offsets: DEVICE int32 [batch+1].  anchors / nc / multi_label / conf / classes as in the filter
 * call.  cand->cap must leave room for the labels. */
int dbl_nms_append_labels(const float* labels, const int32_t* label_offsets, int max_labels_per_image, int anchors, int nc,
                          int multi_label, float conf_thres, const int32_t* classes, int num_classes_kept,
                          const dbl_candidates* cand, dbl_stream_t stream);

/* DetectionValidator._process_batch (models/yolo/detect/val.py:213-227) = box_iou (utils/metrics.py:52-71) + match_predictions
 * (engine/validator.py:222-260, use_scipy=False) for a whole batch in one launch.
 *   rows f32 [batch][max_det][6] + counts int32 [batch]: the NMS output (dbl_nms / dbl_nms_scaled);
 *   gt_boxes f32 [G][4] xyxy (16-byte aligned), gt_cls f32 [G], gt_offsets int32 [batch+1]: ground truth, images concatenated;
 *   iou_thresholds f32 [T], T <= 16 (the validator's iouv = linspace(0.5, 0.95, 10));
 *   correct uint8 [batch][max_det][T]: 1 where detection d is a true positive at threshold t (rows >= counts[b] are 0).
 * ws: scratch of dbl_match_workspace_bytes(batch, max_labels_per_image, T) bytes.  Pairs with exactly equal IoU resolve to
 * the larger label index (the reference's argsort leaves that order unspecified). */
size_t dbl_match_workspace_bytes(int batch, int max_labels_per_image, int num_thresholds);
int dbl_match_predictions(const float* rows, const int32_t* counts, int batch, int max_det, const float* gt_boxes,
                          const float* gt_cls, const int32_t* gt_offsets, int max_labels_per_image, const float* iou_thresholds,
                          int num_thresholds, uint8_t* correct, void* ws, size_t ws_bytes, dbl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DBL_SM100_H */
