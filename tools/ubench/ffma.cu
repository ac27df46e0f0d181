// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma ffma.cu   (run on the GPU box: gpurun -- tools/ubench/ffma)
// Micro-benchmark: 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, const float* in, int iters) {
    float a[16], b[4], c[4];
    for (int i = 0; i < 16; ++i) a[i] = in[threadIdx.x + i];
    for (int i = 0; i < 4; ++i) b[i] = in[64 + threadIdx.x + i], c[i] = in[128 + threadIdx.x + i];
    unsigned long long A2[8], B2[2], C2[2];
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(A2[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    for (int i = 0; i < 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(B2[i]) : "f"(b[2 * i]), "f"(b[2 * i + 1]));
    for (int i = 0; i < 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(C2[i]) : "f"(c[2 * i]), "f"(c[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(b[(i + r) & 3], c[r], a[i]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A2[i]) : "l"(B2[(i + r) & 1]), "l"(C2[r & 1]));
        }
    }
    if (MODE == 1)
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(A2[i]));
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *in, *out;
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, in, iters); else k<1><<<148 * 8, 256>>>(out, in, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * iters * 64;
            printf("mode %d: %.3f ms  %.2f TFMA/s (%.1f FMA/clk/SM at 1.965 GHz)\n", mode, ms, fma / ms / 1e9, fma / ms / 1e3 / 148 / 1.965e6);
        }
    return 0;
}
