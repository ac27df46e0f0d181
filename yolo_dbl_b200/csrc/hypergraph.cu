// Adaptive hypergraph computation (AdaHGComputation -> AdaHGConv -> AdaHyperedgeGen, reference
// nn/modules/block.py:1582-1752) on NHWC tokens.  In NHWC a feature map IS the (B, N, D) token matrix, so the
// reference's flatten/transpose pairs vanish.  Algebra used (exact in real arithmetic):
//   logits[n,e] = mean_h (Xproj_h[n] . proto_h[e]) / sqrt(hd) = (X[n] . P[e] + c[e]),
//       P[e,i] = s * sum_o proto[e,o] Wp[o,i],  c[e] = s * sum_o bp[o] proto[e,o],  s = 1/(heads*sqrt(hd))
//   A = softmax over the N tokens;  He = A^T X;  He2 = GELU(He We^T + be)
//   y[n] = GELU(sum_e A[n,e] Q[e] + bn) + X[n],   Q = He2 Wn^T
// so the per-token work is two E-term dot products and no D x D GEMM ever touches the token matrix.
// Six small launches: context partials -> prototypes -> logits (+softmax partials) -> edge partials -> edge finalise
// -> nodes.
#include "dbl_common.cuh"

namespace dbl {

constexpr int HG_MAX_E = 16;
constexpr int HG_CTX_SPLIT = 32;   // token splits for the mean/max context
constexpr int HG_TOK_BLOCK = 128;  // tokens per block in the per-token kernels
constexpr int HG_EDGE_SPLIT = 32;  // token slices per image in the A^T X kernel: short, latency-bound loops -> more CTAs and more loads in flight

struct HgLayout {  // float offsets into the caller's workspace
    size_t ctx_part, P, cvec, logits, stat_part, stats, Q, He, total;
    int nblk;
};
static HgLayout hg_layout(int B, int N, int D, int E) {
    HgLayout L;
    L.nblk = cdiv(N, HG_TOK_BLOCK);
    size_t o = 0;
    L.ctx_part = o, o += (size_t)B * HG_CTX_SPLIT * 2 * D;
    L.P = o, o += (size_t)B * E * D;
    L.cvec = o, o += (size_t)B * HG_MAX_E;
    L.logits = o, o += (size_t)B * N * E;
    L.stat_part = o, o += (size_t)B * L.nblk * E * 2;
    L.stats = o, o += (size_t)B * E * 2;
    L.Q = o, o += (size_t)B * E * D;
    L.He = o, o += (size_t)B * HG_EDGE_SPLIT * E * D;  // per-slice partials, reduced in a fixed order (bitwise reproducible)
    L.total = o;
    return L;
}

// K1: partial sum / max over a token range.  grid (HG_CTX_SPLIT, B), block D threads.
__global__ void hg_context_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part, int N, int D, int pitch) {
    pdl_launch_dependents();
    pdl_wait();
    const int d = threadIdx.x, s = blockIdx.x, b = blockIdx.y;
    const int per = (N + HG_CTX_SPLIT - 1) / HG_CTX_SPLIT;
    const int n0 = s * per, n1 = min(N, n0 + per);
    float sum = 0.f, mx = -INFINITY;
    for (int n = n0; n < n1; ++n) {
        const float v = __bfloat162float(x[((size_t)b * N + n) * pitch + d]);
        sum += v;
        mx = fmaxf(mx, v);
    }
    float* p = part + ((size_t)b * HG_CTX_SPLIT + s) * 2 * D;
    p[d] = sum;
    p[D + d] = mx;
}

// K2: prototypes and their folded forms.  grid (B), block 1024 (one warp per output row: 32 warps keep 32 rows of loads in flight).
__global__ void __launch_bounds__(1024) hg_proto_kernel(const float* __restrict__ part, dbl_hypergraph_weights wt, float* __restrict__ P,
                                                       float* __restrict__ cvec, int N, int D) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // ctx[2D] | proto[E*D]
    float* ctx = sm;
    float* proto = sm + 2 * D;
    const int b = blockIdx.x, E = wt.num_edges;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float sum = 0.f, mx = -INFINITY;
        for (int s = 0; s < HG_CTX_SPLIT; ++s) {
            const float* p = part + ((size_t)b * HG_CTX_SPLIT + s) * 2 * D;
            sum += p[d];
            mx = fmaxf(mx, p[D + d]);
        }
        ctx[d] = sum / (float)N;
        ctx[D + d] = mx;
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31, nwarp = blockDim.x / 32;
    for (int row = warp; row < E * D; row += nwarp) {  // context_net: Linear(2D -> E*D)
        const float* wr = wt.context_w + (size_t)row * 2 * D;
        float acc = 0.f;
        for (int k = lane; k < 2 * D; k += 32) acc = fmaf(__ldg(wr + k), ctx[k], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) proto[row] = acc + wt.context_b[row] + wt.prototype_base[row];
    }
    __syncthreads();
    const int hd = D / wt.num_heads;
    const float scale = 1.0f / ((float)wt.num_heads * sqrtf((float)hd));
    for (int idx = threadIdx.x; idx < E * D; idx += blockDim.x) {  // P[e,i] = s * sum_o proto[e,o] * Wp[o,i]
        const int e = idx / D, i = idx % D;
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < D; ++o) acc = fmaf(proto[e * D + o], __ldg(wt.pre_head_w + (size_t)o * D + i), acc);  // 8 loads in flight
        P[((size_t)b * E + e) * D + i] = acc * scale;
    }
    for (int e = warp; e < E; e += nwarp) {
        float acc = 0.f;
        for (int o = lane; o < D; o += 32) acc = fmaf(proto[e * D + o], __ldg(wt.pre_head_b + o), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) cvec[(size_t)b * HG_MAX_E + e] = acc * scale;
    }
}

// K3: logits[n,e] and per-block online-softmax partials (max, sum exp).  grid (nblk, B), block HG_TOK_BLOCK.
__global__ void __launch_bounds__(HG_TOK_BLOCK) hg_logits_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ P,
                                                                 const float* __restrict__ cvec, float* __restrict__ logits,
                                                                 float* __restrict__ stat_part, int N, int D, int E, int pitch) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // P[E*D] | red_max[E*NW] | red_sum[E*NW]
    float* sP = sm;
    float* red = sm + E * D;
    const int b = blockIdx.y, n = blockIdx.x * HG_TOK_BLOCK + threadIdx.x;
    for (int i = threadIdx.x; i < E * D; i += blockDim.x) sP[i] = P[(size_t)b * E * D + i];
    __syncthreads();
    float lg[HG_MAX_E];
#pragma unroll
    for (int e = 0; e < HG_MAX_E; ++e) lg[e] = (e < E) ? cvec[(size_t)b * HG_MAX_E + e] : 0.f;
    const bool live = n < N;
    if (live) {
        const __nv_bfloat16* xr = x + ((size_t)b * N + n) * pitch;
        for (int d0 = 0; d0 < D; d0 += 8) {
            float xv[8];
            unpack8(ldg8(xr + d0), xv);
#pragma unroll
            for (int e = 0; e < HG_MAX_E; ++e) {
                if (e >= E) break;
                const float* pr = sP + e * D + d0;
#pragma unroll
                for (int j = 0; j < 8; ++j) lg[e] = fmaf(xv[j], pr[j], lg[e]);
            }
        }
        float* lo = logits + ((size_t)b * N + n) * E;
        for (int e = 0; e < E; ++e) lo[e] = lg[e];
    }
    // block partials: (max, sum exp(l - max)) per hyperedge
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    constexpr int NW = HG_TOK_BLOCK / 32;
    float* red_max = red;            // [E][NW]
    float* red_sum = red + E * NW;   // [E][NW]
    for (int e = 0; e < E; ++e) {
        float m = live ? lg[e] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red_max[e * NW + warp] = m;
    }
    __syncthreads();
    for (int e = 0; e < E; ++e) {
        float m = red_max[e * NW];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, red_max[e * NW + w]);
        float s = live ? __expf(lg[e] - m) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red_sum[e * NW + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < E) {
        const int e = threadIdx.x;
        float m = red_max[e * NW], tot = 0.f;
        for (int w = 1; w < NW; ++w) m = fmaxf(m, red_max[e * NW + w]);
        for (int w = 0; w < NW; ++w) tot += red_sum[e * NW + w];
        float* sp = stat_part + (((size_t)b * gridDim.x + blockIdx.x) * E + e) * 2;
        sp[0] = m;
        sp[1] = tot;
    }
}

// K4a: softmax normalisation in place + partial He = A^T X over a token slice.  grid (HG_EDGE_SPLIT, B), block 256.
__global__ void __launch_bounds__(256) hg_edges_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ logits,
                                                               const float* __restrict__ stat_part, float* __restrict__ He, int N, int D,
                                                               int E, int nblk, int pitch) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // mstat[2*HG_MAX_E] | red[R*E*D]
    const int b = blockIdx.y;
    float* mstat = sm;
    float* red = sm + 2 * HG_MAX_E;
    if (threadIdx.x < E) {
        const int e = threadIdx.x;
        float m = -INFINITY;
        for (int k = 0; k < nblk; ++k) m = fmaxf(m, stat_part[(((size_t)b * nblk + k) * E + e) * 2]);
        float s = 0.f;
        for (int k = 0; k < nblk; ++k) {
            const float* sp = stat_part + (((size_t)b * nblk + k) * E + e) * 2;
            s += sp[1] * __expf(sp[0] - m);
        }
        mstat[e] = m;
        mstat[HG_MAX_E + e] = 1.0f / s;
    }
    __syncthreads();
    const int per = (N + HG_EDGE_SPLIT - 1) / HG_EDGE_SPLIT;
    const int n0 = blockIdx.x * per, n1 = min(N, n0 + per);
    float* lg = logits + (size_t)b * N * E;
    for (int i = n0 * E + threadIdx.x; i < n1 * E; i += blockDim.x) {  // A[n,e] in place
        const int e = i % E;
        lg[i] = __expf(lg[i] - mstat[e]) * mstat[HG_MAX_E + e];
    }
    __syncthreads();
    const int R = blockDim.x / D;
    const int d = threadIdx.x % D, r = threadIdx.x / D;
    float acc[HG_MAX_E];
#pragma unroll
    for (int e = 0; e < HG_MAX_E; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int n = n0 + r; n < n1; n += R) {
        const float xv = __bfloat162float(x[((size_t)b * N + n) * pitch + d]);
        const float* a = lg + (size_t)n * E;
#pragma unroll
        for (int e = 0; e < HG_MAX_E; ++e) {
            if (e >= E) break;
            acc[e] = fmaf(a[e], xv, acc[e]);
        }
    }
    for (int e = 0; e < E; ++e) red[((size_t)r * E + e) * D + d] = acc[e];
    __syncthreads();
    for (int i = threadIdx.x; i < E * D; i += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < R; ++rr) s += red[(size_t)rr * E * D + i];
        He[((size_t)b * HG_EDGE_SPLIT + blockIdx.x) * E * D + i] = s;  // no atomics: the final kernel adds the slices in order
    }
}

// K4b: He2 = GELU(He We^T + be), Q = He2 Wn^T.  grid (B), block 1024.
__global__ void __launch_bounds__(1024) hg_edges_final_kernel(const float* __restrict__ He, dbl_hypergraph_weights wt, float* __restrict__ Q,
                                                             int D) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // he[E*D] | he2[E*D]
    const int b = blockIdx.x, E = wt.num_edges;
    float* he = sm;
    float* he2 = sm + E * D;
    for (int i = threadIdx.x; i < E * D; i += blockDim.x) {
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < HG_EDGE_SPLIT; ++k) s += He[((size_t)b * HG_EDGE_SPLIT + k) * E * D + i];
        he[i] = s;
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31, nwarp = blockDim.x / 32;
    for (int row = warp; row < E * D; row += nwarp) {  // He2[e,j] = GELU(sum_d He[e,d] We[j,d] + be[j])
        const int e = row / D, j = row % D;
        float s = 0.f;
        for (int k = lane; k < D; k += 32) s = fmaf(he[e * D + k], __ldg(wt.edge_w + (size_t)j * D + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) he2[row] = gelu_f(s + wt.edge_b[j]);
    }
    __syncthreads();
    for (int row = warp; row < E * D; row += nwarp) {  // Q[e,j] = sum_d He2[e,d] Wn[j,d]
        const int e = row / D, j = row % D;
        float s = 0.f;
        for (int k = lane; k < D; k += 32) s = fmaf(he2[e * D + k], __ldg(wt.node_w + (size_t)j * D + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) Q[(size_t)b * E * D + row] = s;
    }
}

// K5: y[n] = GELU(sum_e A[n,e] Q[e] + bn) + X[n].  grid (nblk, B), block HG_TOK_BLOCK, thread per token.
__global__ void __launch_bounds__(HG_TOK_BLOCK) hg_nodes_kernel(const __nv_bfloat16* x, const float* __restrict__ A,
                                                                const float* __restrict__ Q, const float* __restrict__ node_b,
                                                                __nv_bfloat16* y, int N, int D, int E, int xpitch,
                                                                int ypitch) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];  // Q[E*D] | bn[D]
    float* sQ = sm;
    float* sb = sm + E * D;
    const int b = blockIdx.y, n = blockIdx.x * HG_TOK_BLOCK + threadIdx.x;
    for (int i = threadIdx.x; i < E * D; i += blockDim.x) sQ[i] = Q[(size_t)b * E * D + i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) sb[i] = node_b[i];
    __syncthreads();
    if (n >= N) return;
    float a[HG_MAX_E];
#pragma unroll
    for (int e = 0; e < HG_MAX_E; ++e) a[e] = (e < E) ? A[((size_t)b * N + n) * E + e] : 0.f;
    const __nv_bfloat16* xr = x + ((size_t)b * N + n) * xpitch;
    __nv_bfloat16* yr = y + ((size_t)b * N + n) * ypitch;
    for (int d0 = 0; d0 < D; d0 += 8) {
        float xv[8], o[8];
        unpack8(ld8(xr + d0), xv);  // x may alias y (in-place use inside C3AH): coherent load
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = sb[d0 + j];
#pragma unroll
        for (int e = 0; e < HG_MAX_E; ++e) {
            if (e >= E) break;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(a[e], sQ[e * D + d0 + j], o[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = gelu_f(o[j]) + xv[j];
        stg8(yr + d0, pack8(o));
    }
}

}  // namespace dbl

using namespace dbl;

extern "C" size_t dbl_hypergraph_workspace_bytes(int batch, int tokens, int dim, int edges) {
    if (batch <= 0 || tokens <= 0 || dim <= 0 || edges <= 0) return 0;
    return hg_layout(batch, tokens, dim, edges).total * sizeof(float);
}

extern "C" int dbl_hypergraph(const dbl_tensor* x, const dbl_hypergraph_weights* wt, const dbl_tensor* y, void* ws, size_t ws_bytes,
                              dbl_stream_t stream) {
    DBL_REQUIRE(valid_tensor(x) && valid_tensor(y) && wt && ws, DBL_E_ARG, "dbl_hypergraph: null/invalid argument");
    DBL_REQUIRE(vec8_ok(x) && vec8_ok(y) && same_shape(x, y), DBL_E_ARG, "dbl_hypergraph: needs bf16 views, c %% 8 == 0, same shape");
    const int B = x->n, N = x->h * x->w, D = x->c, E = wt->num_edges;
    DBL_REQUIRE(E >= 1 && E <= HG_MAX_E, DBL_E_UNSUPPORTED, "dbl_hypergraph: num_edges %d > %d", E, HG_MAX_E);
    DBL_REQUIRE(D <= 256 && 256 % D == 0, DBL_E_UNSUPPORTED, "dbl_hypergraph: dim %d must divide 256", D);
    DBL_REQUIRE(wt->num_heads > 0 && D % wt->num_heads == 0, DBL_E_ARG, "dbl_hypergraph: heads");
    const HgLayout L = hg_layout(B, N, D, E);
    DBL_REQUIRE(ws_bytes >= L.total * sizeof(float), DBL_E_WORKSPACE, "dbl_hypergraph: workspace %zu < %zu bytes", ws_bytes,
                L.total * sizeof(float));
    float* f = reinterpret_cast<float*>(ws);
    cudaStream_t st = as_stream(stream);
    auto* xp = reinterpret_cast<const __nv_bfloat16*>(x->data);
    auto* yp = reinterpret_cast<__nv_bfloat16*>(y->data);

    launch_k(hg_context_kernel, dim3(HG_CTX_SPLIT, B), D, 0, st, xp, f + L.ctx_part, N, D, x->pitch);
    launch_k(hg_proto_kernel, B, 1024, (2 * D + E * D) * sizeof(float), st, f + L.ctx_part, *wt, f + L.P, f + L.cvec, N, D);
    const size_t sm3 = ((size_t)E * D + 2 * E * (HG_TOK_BLOCK / 32)) * sizeof(float);
    launch_k(hg_logits_kernel, dim3(L.nblk, B), HG_TOK_BLOCK, sm3, st, xp, f + L.P, f + L.cvec, f + L.logits, f + L.stat_part, N, D, E,
                                                                 x->pitch);
    const int R = 256 / D;
    const size_t sm4 = (2 * HG_MAX_E + (size_t)R * E * D) * sizeof(float);
    launch_k(hg_edges_partial_kernel, dim3(HG_EDGE_SPLIT, B), 256, sm4, st, xp, f + L.logits, f + L.stat_part, f + L.He, N, D, E, L.nblk, x->pitch);
    launch_k(hg_edges_final_kernel, B, 1024, (size_t)2 * E * D * sizeof(float), st, f + L.He, *wt, f + L.Q, D);
    launch_k(hg_nodes_kernel, dim3(L.nblk, B), HG_TOK_BLOCK, ((size_t)E * D + D) * sizeof(float), st, 
        xp, f + L.logits, f + L.Q, wt->node_b, yp, N, D, E, x->pitch, y->pitch);
    count_launch(6);
    return check_launch("dbl_hypergraph");
}
