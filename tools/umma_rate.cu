// umma_rate.cu -- issue rate of tf32 tcgen05.mma (M = 128, K = 8) from one thread: clocks per UMMA for N = 64..256, B K-major
// (SWIZZLE_32B) or MN-major (SWIZZLE_128B_BASE32B), A from shared memory or from TMEM, one commit per `group` UMMAs, with or
// without waiting for each group (latency vs throughput).  Operand contents do not matter (zeros).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) rate(long long *out, uint32_t N, uint32_t b_mn, uint32_t a_tmem, uint32_t groups, uint32_t group,
                                               uint32_t wait_each, uint32_t n_acc, uint32_t bf16)
{
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *gen = raw + (base - smem_u32(raw));
    const uint32_t bar = base + 4096 + 8192, slot = bar + 8;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 3072; i += 128) reinterpret_cast<float *>(gen)[i] = 0.f;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(gen + (slot - base));
    if (t == 0) {
        const uint64_t adesc = (uint64_t)((base & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
        const uint64_t bdesc_k = (uint64_t)(((base + 4096) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
        const uint64_t bdesc_mn = (uint64_t)(((base + 4096) & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
        const uint64_t bdesc = b_mn ? bdesc_mn : bdesc_k;
        const uint32_t fmt = bf16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (b_mn << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (uint32_t g = 0; g < groups; ++g) {
            for (uint32_t i = 0; i < group; ++i) {
                const uint32_t d = tmem + (n_acc == 2 ? 256u * (g & 1) : n_acc == 3 ? 256u * ((g * group + i) & 1) : n_acc == 4 ? 128u * ((g * group + i) & 3) : 0u);
                if (bf16)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                 :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                else if (a_tmem)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                                 :: "r"(d), "r"(tmem + 480u), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                                 :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
            }
            if (wait_each || g + 1 == groups) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
                phase ^= 1;
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int main()
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    struct Cfg { uint32_t N, b_mn, a_tmem, groups, group, wait_each, n_acc; const char *what; uint32_t bf16; } cfgs[] = {
        {256, 0, 0, 2000, 3, 0, 3, "N=256 B K-major, A smem, accumulator alternating per UMMA (2 x 256 columns)"},
        {128, 0, 0, 2000, 4, 0, 4, "N=128 B K-major, A smem, 4 accumulators round robin"},
        {256, 0, 0, 2000, 3, 0, 1, "bf16 (K = 16) N=256 B K-major, A smem, back to back", 1},
        {128, 0, 0, 2000, 3, 0, 1, "bf16 (K = 16) N=128 B K-major, A smem, back to back", 1},
        {256, 0, 0, 2000, 3, 0, 3, "bf16 (K = 16) N=256, accumulator alternating per UMMA", 1},
        {256, 0, 0, 2000, 3, 0, 1, "N=256 B K-major, A smem, back to back"},
        {256, 1, 0, 2000, 3, 0, 1, "N=256 B MN-major, A smem, back to back"},
        {256, 1, 1, 2000, 3, 0, 1, "N=256 B MN-major, A tmem, back to back"},
        {256, 0, 1, 2000, 3, 0, 1, "N=256 B K-major, A tmem, back to back"},
        {128, 1, 1, 2000, 3, 0, 1, "N=128 B MN-major, A tmem, back to back"},
        {128, 0, 0, 2000, 3, 0, 1, "N=128 B K-major, A smem, back to back"},
        {64, 0, 0, 2000, 3, 0, 1, "N=64  B K-major, A smem, back to back"},
        {256, 1, 1, 2000, 3, 0, 2, "N=256 B MN-major, A tmem, alternating accumulators per group of 3"},
        {256, 1, 1, 2000, 3, 1, 1, "N=256 B MN-major, A tmem, commit + wait after every 3 (latency)"},
        {256, 1, 1, 2000, 1, 1, 1, "N=256 B MN-major, A tmem, commit + wait after every 1 (latency)"},
        {256, 0, 0, 2000, 1, 1, 1, "N=256 B K-major, A smem, commit + wait after every 1 (latency)"},
    };
    for (auto c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) rate<<<148, 128, 16384>>>(d, c.N, c.b_mn, c.a_tmem, c.groups, c.group, c.wait_each, c.n_acc, c.bf16);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("%-75s: %s  %.1f clocks per UMMA\n", c.what, cudaGetErrorString(e), s / 148 / (c.groups * c.group));
    }
    return 0;
}
