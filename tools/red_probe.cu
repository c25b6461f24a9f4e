// red_probe.cu -- what bounds fp32 global reductions (RED.ADD.F32) on B200: elements, sectors or instructions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu && tools/red_probe
// Each variant adds N floats into a 100 MB accumulator (larger than nothing: it stays in L2 / HBM as it likes).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red1(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float *p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float *p, float a, float b, float c, float d) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory"); }

// mode 0: scalar, lane -> consecutive floats (coalesced)        1 element / lane-op
// mode 1: v2 coalesced   mode 2: v4 coalesced
// mode 3: scalar, each lane its own 32-byte sector (stride 8)
// mode 4: scalar, pseudo-random address within +-64 KB
// mode 5: scalar coalesced but unaligned (+1 float)
// mode 6: plain st.global coalesced (reference: no atomics)
__global__ void k(float *acc, size_t n_elem, int mode, size_t span)
{
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    if (mode == 0) for (size_t i = t; i < n_elem; i += stride) red1(acc + i % span, 1.f);
    if (mode == 5) for (size_t i = t; i < n_elem; i += stride) red1(acc + (i + 1) % span, 1.f);
    if (mode == 1) for (size_t i = t; i < n_elem / 2; i += stride) red2(acc + (2 * i) % span, 1.f, 2.f);
    if (mode == 2) for (size_t i = t; i < n_elem / 4; i += stride) red4(acc + (4 * i) % span, 1.f, 2.f, 3.f, 4.f);
    if (mode == 3) for (size_t i = t; i < n_elem; i += stride) red1(acc + (8 * i) % span, 1.f);
    if (mode == 4) for (size_t i = t; i < n_elem; i += stride) {
        unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15;
        red1(acc + ((i & ~(size_t)1023) + (h & 16383)) % span, 1.f);
    }
    if (mode == 6) for (size_t i = t; i < n_elem; i += stride) acc[i % span] = 1.f;
    // modes 7..9: G = 2, 4, 8 consecutive lanes share one sector, every group of G lanes sits in another row (4 KB apart)
    if (mode >= 7 && mode <= 9) {
        const int G = mode == 7 ? 2 : (mode == 8 ? 4 : 8);
        const unsigned lane = threadIdx.x & 31;
        for (size_t i = t; i < n_elem; i += stride) {
            const size_t warp_base = (i - lane) * 8;                       // spread warps over the accumulator
            red1(acc + (warp_base + (lane / G) * 1024 + (lane % G)) % span, 1.f);
        }
    }
    // mode 10: like mode 8 (4 lanes per sector) but the 4 lanes are NOT sector aligned (straddle two sectors)
    if (mode == 10) {
        const unsigned lane = threadIdx.x & 31;
        for (size_t i = t; i < n_elem; i += stride) {
            const size_t warp_base = (i - lane) * 8;
            red1(acc + (warp_base + (lane / 4) * 1024 + (lane % 4) + 6) % span, 1.f);
        }
    }
}

int main()
{
    const size_t span = 32u << 20;            // 32 M floats = 128 MB (power of two: the modulo is a mask)
    const size_t n = 100u << 20;              // 100 M element-adds per launch
    float *acc; cudaMalloc(&acc, span * 4 + 64); cudaMemset(acc, 0, span * 4 + 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"scalar coalesced", "v2 coalesced", "v4 coalesced", "scalar 1 lane/sector", "scalar random",
                           "scalar coalesced unaligned", "plain store coalesced", "2 lanes/sector, 16 rows", "4 lanes/sector, 8 rows",
                           "8 lanes/sector, 4 rows", "4 lanes straddling 2 sectors"};
    for (int mode = 0; mode < 11; ++mode) {
        k<<<148 * 8, 256>>>(acc, n, mode, span);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) k<<<148 * 8, 256>>>(acc, n, mode, span);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-28s %8.1f us  %7.1f G elem/s  %6.2f elem/clk/SM\n", names[mode], ms * 1e3, n / ms / 1e6, n / ms / 1e6 / 148 / 1.92);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
