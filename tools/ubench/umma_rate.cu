// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
// Issue-to-completion rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16) with BOTH operands in shared memory, as a function
// of N and of the row width / swizzle of the A tile: what does a narrow-N implicit-GEMM convolution pay per MMA?
//   mode 0: every MMA reads the same A tile        mode 1: the A start walks through a 64 KB window in steps of one row (conv taps)
//   mode 4: mode 1 + the B tile changes with every MMA (eight 1 KB blocks)      mode 5: mode 4 + the A start also steps through the 32-byte pieces of its row
//   mode 3: the A start walks in steps of 40 rows (8-row aligned)
//   accumulators: consecutive MMAs accumulate into 1 / 2 / 4 different TMEM tiles in turn (dependent chain vs interleaved)
//   mode 2: no-swizzle A, the two 8-byte... (LBO) core matrices `lbo` bytes apart (two filter taps of an 8-channel tensor in one k-step)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t row_bytes) {  // swizzled K-major
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) | ((uint64_t)1 << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int rowb, int mode, int iters, int lbo, int nacc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
    if (warp == 0) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        const uint64_t bd = desc(b0, 32);
        // eight descriptors precomputed (the issue loop must not be bound by its own address arithmetic)
        uint64_t ad[8], bdv[8];
        for (int j = 0; j < 8; ++j) bdv[j] = desc(b0 + (uint32_t)((mode == 4 || mode == 5) ? j * 1024 : 0), 32);
        for (int j = 0; j < 8; ++j) {
            const uint32_t row = mode == 0 ? 0u : (mode == 3 ? (uint32_t)(j * 8 * 5) : (uint32_t)(j * 37 + 1));
            ad[j] = mode == 2 ? desc_noswz(a0 + row * 16, lbo, 128) : desc(a0 + row * rowb + (mode == 5 ? (uint32_t)((j & 3) * 32) % rowb : 0u), rowb);
        }
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((j % nacc) * n)),
                    "l"(ad[j]), "l"(bdv[j]), "r"(idesc), "r"(1u)
                    : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (blockIdx.x == 0) out[0] = t1 - t0, out[1] = (long long)(g1 - g0);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 64);
    const int smem = 162 * 1024 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 40000;
    const int ns[] = {16, 32, 64, 128, 256};
    for (int nacc : {1, 2, 4})
    for (int mode : {1, 4, 5})
        for (int rowb : {64, 128})
            for (int n : ns) {
                if (nacc * n > 256) continue;
                k<<<148, 128, smem>>>(out, n, rowb, mode, iters, 0, nacc);
                long long c2[2] = {0, 0};
                cudaMemcpy(c2, out, 16, cudaMemcpyDeviceToHost);
                const long long c = c2[0];
                printf("accumulators %d  mode %d  A rows %3d B  N %3d: %6.1f clk per MMA = %5.1f ns (SM clock %.0f MHz)  (%s)\n", nacc, mode, rowb, n, (double)c / iters,
                       (double)c2[1] / iters, 1e3 * c / (double)c2[1], cudaGetErrorString(cudaGetLastError()));
            }
    for (int lbo : {16, 48, 1040})
        for (int n : {16, 32}) {
            k<<<148, 128, smem>>>(out, n, 16, 2, iters, lbo, 2);
            long long c = 0;
            cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            printf("mode 2  no swizzle, LBO %4d  N %3d: %6.1f clk per MMA  (%s)\n", lbo, n, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
