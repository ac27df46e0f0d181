// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_warps ffma2_warps.cu
// FFMA2 issue rate as a function of resident warps per SM (one CTA per SM) and of the number of independent accumulators per thread:
// how many depthwise warps does a warp-specialised CTA need before the fp32 pipes are full?
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(float* out, const float* in, int iters) {
    unsigned long long A[NACC], X[4], W[7];
    for (int i = 0; i < NACC; ++i) A[i] = 0ull;
    for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(X[i]) : "f"(in[threadIdx.x + i]), "f"(in[threadIdx.x + i + 1]));
    for (int i = 0; i < 7; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(W[i]) : "f"(in[64 + threadIdx.x + i]), "f"(in[65 + threadIdx.x + i]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 28; ++r)
#pragma unroll
            for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(X[(i + r) & 3]), "l"(W[r % 7]));
    }
    float s = 0;
    for (int i = 0; i < NACC; ++i) {
        float a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(A[i]));
        s += a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(float* out, float* in, int warps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int iters = 4000;
    k<NACC><<<148, warps * 32>>>(out, in, 10);
    cudaEventRecord(e0);
    k<NACC><<<148, warps * 32>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double f2 = 148.0 * warps * 32 * (double)iters * 28 * NACC;
    printf("acc %2d warps/SM %2d: %.3f ms  %.1f FFMA2-lanes/clk/SM = %.1f FMA/clk/SM, %.2f clk per warp-FFMA2 per SMSP\n", NACC, warps, ms, f2 / ms / 1e3 / 148 / 1.965e6,
           2 * f2 / ms / 1e3 / 148 / 1.965e6, 1.0 / (f2 / 32 / 4 / ms / 1e3 / 148 / 1.965e6));
}
int main() {
    float *in, *out;
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&out, 148 * 1024 * 4);
    for (int w : {4, 8, 10, 12, 16, 20, 24, 32}) run<10>(out, in, w);
    for (int w : {4, 8, 10, 16}) run<20>(out, in, w);
    for (int w : {4, 8, 10, 16}) run<4>(out, in, w);
    return 0;
}
