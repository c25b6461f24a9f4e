// ffma_probe.cu -- FP32 FMA issue-rate probe for sm_100a (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma_probe tools/ffma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long *>(&d))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return d;
}

template <int V>
__global__ void __launch_bounds__(256, (V >= 4) ? 1 : 2) k(float *sink, int iters, float a, float b, long long *cyc)
{
    long long t0 = clock64();
    float s = 0.f;
    if (V == 0) {            // scalar 8x8 outer product
        float acc[8][8], u[8], v[8];
        for (int i = 0; i < 8; ++i) { u[i] = a + (threadIdx.x + i) * 1e-7f; v[i] = b + (threadIdx.x * 8 + i) * 1e-7f;
            for (int j = 0; j < 8; ++j) acc[i][j] = (float)(i - j); }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(u[i], v[j], acc[i][j]);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    } else if (V == 1) {     // scalar, one register-file operand
        float acc[64]; float x = a + threadIdx.x * 1e-7f, y = b;
        for (int i = 0; i < 64; ++i) acc[i] = (float)i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i] = __fmaf_rn(acc[i], x, y);
        for (int i = 0; i < 64; ++i) s += acc[i];
    } else if (V == 2) {     // packed f32x2 outer product: 8 u-pairs(dup) x 4 v-pairs -> 32 pair accumulators (64 FMAs)
        float2 acc[8][4], u[8], v[4];
        for (int i = 0; i < 8; ++i) { float t = a + (threadIdx.x + i) * 1e-7f; u[i] = make_float2(t, t); }
        for (int j = 0; j < 4; ++j) v[j] = make_float2(b + j * 1e-7f, b + (j + threadIdx.x) * 1e-7f);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = make_float2((float)i, (float)j);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma2(u[i], v[j], acc[i][j]);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    } else if (V == 3) {     // packed, one register-file operand
        float2 acc[32]; float2 x = make_float2(a + threadIdx.x * 1e-7f, a), y = make_float2(b, b);
        for (int i = 0; i < 32; ++i) acc[i] = make_float2((float)i, 1.f);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fma2(acc[i], x, y);
        for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
    } else if (V == 4) {     // correlation-like sliding window, scalar: acc[21][8] += a[k]*w[i+k]
        float acc[21][8], aa[8], w[28];
        for (int i = 0; i < 8; ++i) aa[i] = a + (threadIdx.x + i) * 1e-7f;
        for (int i = 0; i < 28; ++i) w[i] = b + (threadIdx.x * 3 + i) * 1e-7f;
        for (int i = 0; i < 21; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 21; ++i)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) acc[i][kk] = __fmaf_rn(aa[kk], w[i + kk], acc[i][kk]);
        for (int i = 0; i < 21; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    } else if (V == 5) {     // correlation-like sliding window, packed over i (pairs start where i+k is even)
        float2 accp[8][10]; float accs[8]; float2 ad[8]; float2 w2[14];
        for (int i = 0; i < 8; ++i) { float t = a + (threadIdx.x + i) * 1e-7f; ad[i] = make_float2(t, t); accs[i] = 0.f; }
        for (int i = 0; i < 14; ++i) w2[i] = make_float2(b + (threadIdx.x * 3 + i) * 1e-7f, b + i * 2e-7f);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 10; ++j) accp[i][j] = make_float2(0.f, 0.f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                // even k: pairs (i=2p,2p+1) use w2[(2p+k)/2], single i=20 uses w[20+k] = w2[(20+k)/2].x
                // odd  k: single i=0 uses w[k] = w2[k/2].y, pairs (i=2p+1,2p+2) use w2[(2p+1+k)/2]
#pragma unroll
                for (int p = 0; p < 10; ++p) {
                    const int j = (kk & 1) ? (2 * p + 1 + kk) / 2 : (2 * p + kk) / 2;
                    accp[kk][p] = fma2(ad[kk], w2[j], accp[kk][p]);
                }
                accs[kk] = (kk & 1) ? __fmaf_rn(ad[kk].x, w2[kk / 2].y, accs[kk]) : __fmaf_rn(ad[kk].x, w2[(20 + kk) / 2].x, accs[kk]);
            }
        }
        for (int i = 0; i < 8; ++i) { s += accs[i]; for (int j = 0; j < 10; ++j) s += accp[i][j].x + accp[i][j].y; }
    }
    long long t1 = clock64();
    if (s == 123.456f) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
void run(const char *name, double fma_per_thread_iter, int iters)
{
    float *sink; long long *cyc, hc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    const int per_sm = (V >= 4) ? 1 : 2;
    const int grid = 148 * per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<grid, 256>>>(sink, iters, 0.999f, 0.001f, cyc);
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(sink, iters, 0.999f, 0.001f, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    double fma = fma_per_thread_iter * iters * 256.0 * grid;
    // per SM: 2 CTAs x 256 threads
    double fma_per_clk_sm = fma_per_thread_iter * iters * 256.0 * per_sm / (double)hc;
    printf("%-28s %8.3f ms  %6.2f TFLOP/s  %6.1f FMA/clk/SM  (%lld cycles, %.2f GHz)  %s\n", name, ms, 2 * fma / ms / 1e9,
           fma_per_clk_sm, hc, hc / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int iters = 20000;
    run<0>("scalar 8x8 outer", 64, iters);
    run<1>("scalar acc=fma(acc,x,y)", 64, iters);
    run<2>("f32x2 8x4 outer", 64, iters);
    run<3>("f32x2 acc=fma(acc,x,y)", 64, iters);
    run<4>("scalar corr window 21x8", 168, iters / 2);
    run<5>("f32x2 corr window", 168, iters / 2);
    return 0;
}
