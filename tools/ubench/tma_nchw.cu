// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_nchw tma_nchw.cu -lcuda
// Does a 4-D TMA box over a raw NCHW f32 / u8 image batch load (negative start = zero fill)?  Prints the sum of the loaded box against the host's.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename T>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int count, uint32_t bytes, double* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                     : "memory");
    }
    uint32_t done = 0;
    int spins = 0;
    while (!done && ++spins < (1 << 22))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (threadIdx.x == 0) {
        double s = 0;
        const T* v = reinterpret_cast<const T*>(smem);
        for (int i = 0; i < count; ++i) s += (double)v[i];
        out[0] = done ? s : -1.0;
    }
}
template <typename T>
void run(int w, int h, int n, int bw, int bh, int x0, int y0) {
    std::vector<T> host((size_t)w * h * 3 * n);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (T)(i % 7 + 1);
    T* d;
    cudaMalloc(&d, host.size() * sizeof(T));
    cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
    double* out;
    cudaMalloc(&out, 8);
    CUtensorMap map;
    cuuint64_t gd[4] = {(cuuint64_t)w, (cuuint64_t)h, 3, (cuuint64_t)n};
    cuuint64_t gs[3] = {(cuuint64_t)w * sizeof(T), (cuuint64_t)w * h * sizeof(T), (cuuint64_t)3 * w * h * sizeof(T)};
    cuuint32_t bx[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 3, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    double ref = 0;
    for (int pl = 0; pl < 3; ++pl)
        for (int yy = 0; yy < bh; ++yy)
            for (int xx = 0; xx < bw; ++xx) {
                const int gx = x0 + xx, gy = y0 + yy;
                if (gx >= 0 && gx < w && gy >= 0 && gy < h) ref += (double)host[((size_t)(1 * 3 + pl) * h + gy) * w + gx];
            }
    const int count = bw * bh * 3;
    cudaFuncSetAttribute(k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<T><<<1, 32, 200 * 1024>>>(map, x0, y0, 0, 1, count, (uint32_t)(count * sizeof(T)), out);
    double got = 0;
    cudaError_t e = cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost);
    printf("%s w %d h %d box %dx%d at (%d,%d): encode %d, kernel %s, sum %.0f (host %.0f)\n", sizeof(T) == 4 ? "f32" : "u8 ", w, h, bw, bh, x0, y0, (int)r, cudaGetErrorString(e), got, ref);
    if (e != cudaSuccess) exit(1);
}
int main() {
    cuInit(0);
    cudaFree(0);
    run<uint8_t>(96, 64, 2, 96, 19, 0, 0);
    run<uint8_t>(96, 64, 2, 128, 19, -16, -2);   // innermost start must be a multiple of 16 bytes (-10 faults: "illegal instruction")
    run<float>(96, 64, 2, 116, 19, -12, -2);
    run<float>(640, 640, 2, 100, 23, 628, 630);
    run<uint8_t>(640, 640, 2, 112, 23, 624, 630);
    run<float>(96, 64, 2, 112, 19, -10, -2);
    return 0;
}
