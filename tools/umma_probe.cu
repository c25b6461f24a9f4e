// umma_probe.cu -- which shared-memory word does a tf32 tcgen05.mma read for B[n][k] under a given descriptor?
// A = selector (A[m][k] = 1 iff k == m % 8, K-major SWIZZLE_32B, known-good), B region filled with its own word index,
// so D[m][n] = index of the word read as B[n][k = m % 8].  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(float *out, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t b_mn, uint32_t N)
{
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *gen = raw + (base - smem_u32(raw));
    float *A = reinterpret_cast<float *>(gen);                  // 4 KB: 128 rows x 32 B
    float *B = reinterpret_cast<float *>(gen + 4096);           // 8 KB
    const uint32_t bar = base + 4096 + 8192, slot = bar + 8;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 1024; i += 128) A[i] = 0.f;
    for (int i = t; i < 2048; i += 128) B[i] = (float)i;
    __syncthreads();
    {   // A[m][k] = (k == m % 8): row m at m * 32, 16-byte chunk swizzled by address bit 7
        const int m = t, k = m & 7;
        const int chunk = (k >> 2) ^ ((m >> 2) & 1);
        A[m * 8 + chunk * 4 + (k & 3)] = 1.f;
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(gen + (slot - base));
    if (t == 0) {
        const uint64_t adesc = (uint64_t)((base & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
        const uint64_t bdesc = (uint64_t)(((base + 4096) & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                               ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (b_mn << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     :: "r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
            if (clock64() - t0 > (1ll << 31)) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (uint32_t c0 = 0; c0 < N; c0 += 8) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[t * N + c0 + i] = __uint_as_float(u[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

int main()
{
    const uint32_t N = 64;
    float *d;
    cudaMalloc(&d, 128 * N * 4);
    float *h = (float *)malloc(128 * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    struct Cfg { uint32_t lbo, sbo, layout, mn; } cfgs[] = {
        {1024, 512, 1, 1}, {512, 1024, 1, 1}, {256, 2048, 1, 1}, {256, 2048, 6, 1}, {1024, 512, 2, 1}, {256, 2048, 0, 1}, {128, 256, 0, 1}, {256, 128, 0, 1},
        {16, 256, 6, 0},
    };
    for (auto c : cfgs) {
        cudaMemset(d, 0, 128 * N * 4);
        probe<<<1, 128, 16384>>>(d, c.lbo, c.sbo, c.layout, c.mn, N);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== lbo %u sbo %u layout %u b_mn %u : %s\n", c.lbo, c.sbo, c.layout, c.mn, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, 128 * N * 4, cudaMemcpyDeviceToHost);
        for (int k = 0; k < 8; ++k) {
            printf(" k=%d:", k);
            for (uint32_t n = 0; n < N; ++n) printf(" %4d", (int)h[k * N + n]);
            printf("\n");
        }
    }
    return 0;
}
