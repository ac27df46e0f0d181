// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu mufu.cu   (run on the GPU box: gpurun -- tools/ubench/mufu)
// Micro-benchmark: MUFU.TANH f32 vs packed f16x2 throughput, and the accuracy of SiLU built on each.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ float tanh_f32(float x) { float t; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ unsigned tanh_h2(unsigned x) { unsigned t; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(x)); return t; }
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned h[4];
    for (int i = 0; i < 4; ++i) h[i] = 0x3c003800u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = tanh_f32(a[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = tanh_h2(h[i]);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    for (int i = 0; i < 4; ++i) s += (float)h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ float silu_a(float v) { const float h = 0.5f * v; return fmaf(h, tanh_f32(h), h); }
__device__ __forceinline__ void silu_b(float v0, float v1, float& o0, float& o1) {
    const float h0 = 0.5f * v0, h1 = 0.5f * v1;
    const __half2 hh = __floats2half2_rn(h0, h1);
    const unsigned t = tanh_h2(*reinterpret_cast<const unsigned*>(&hh));
    const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&t));
    o0 = fmaf(h0, tf.x, h0), o1 = fmaf(h1, tf.y, h1);
}
__global__ void acc(float* err) {
    // x in [-20, 20]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float x = -20.f + 40.f * i / (1 << 20);
    const double ref = (double)x / (1.0 + exp(-(double)x));
    float b0, b1;
    silu_b(x, x, b0, b1);
    const float a0 = silu_a(x);
    const double ea = fabs(a0 - ref), eb = fabs(b0 - ref);
    const double den = fmax(fabs(ref), 1e-3);
    atomicMax((int*)&err[0], __float_as_int((float)ea));
    atomicMax((int*)&err[1], __float_as_int((float)eb));
    atomicMax((int*)&err[2], __float_as_int((float)(ea / den)));
    atomicMax((int*)&err[3], __float_as_int((float)(eb / den)));
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double n = 148.0 * 8 * 256 * iters * 8;  // tanh results
            printf("mode %d: %.3f ms  %.2f Ttanh/s (%.1f results/clk/SM at 1.965 GHz)\n", mode, ms, n / ms / 1e9, n / ms / 1e3 / 148 / 1.965e6);
        }
    float* err;
    cudaMalloc(&err, 16);
    cudaMemset(err, 0, 16);
    acc<<<(1 << 20) / 256, 256>>>(err);
    float h[4];
    cudaMemcpy(h, err, 16, cudaMemcpyDeviceToHost);
    printf("silu max abs err: f32-tanh %.3e  f16x2-tanh %.3e ; max rel err (|ref|>1e-3): %.3e  %.3e\n", h[0], h[1], h[2], h[3]);
    return 0;
}
