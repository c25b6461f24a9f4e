// tma_probe.cu -- standalone probe of TMA tiled loads on sm_100a (development aid, not shipped).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe <W> <H> <C> <N> <boxW> <boxH> <boxC> <x0> <y0> <style>
// style 0: thread 0 in a divergent branch issues; 1: warp 0 + elect.sync
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int nfloats, int x0, int y0, int style, uint32_t bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = smem_u32(&bar), dst = smem_u32(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    bool issuer = false;
    if (style == 1) {
        if (threadIdx.x < 32) {
            uint32_t pred;
            asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
            issuer = pred != 0;
        }
    } else {
        issuer = threadIdx.x == 0;
    }
    if (issuer) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     :: "r"(dst), "l"(&tm), "r"(x0), "r"(y0), "r"(0), "r"(0), "r"(b) : "memory");
    }
    asm volatile("{ .reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" :: "r"(b), "r"(0) : "memory");
    const float *s = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char **argv)
{
    if (argc < 11) { printf("usage\n"); return 2; }
    int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), bc = atoi(argv[7]), x0 = atoi(argv[8]), y0 = atoi(argv[9]), style = atoi(argv[10]);
    size_t n = (size_t)W * H * C * N;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003) + 1.0f;
    float *d, *o;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int nf = bw * bh * bc;
    cudaMalloc(&o, nf * 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((Fn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("[%s] encode rc=%d ", argv[11 < argc ? 11 : 0], (int)r);
    if (r) { printf("\n"); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<1, 128, nf * 4 + 128, 0>>>(tm, o, nf, x0, y0, style, (uint32_t)nf * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s ", cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    std::vector<float> ho(nf);
    cudaMemcpy(ho.data(), o, nf * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < bc; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
        int gx = x0 + x, gy = y0 + y;
        float exp = (gx < 0 || gx >= W || gy < 0 || gy >= H || c >= C) ? 0.f : h[((size_t)c * H + gy) * W + gx];
        if (ho[(c * bh + y) * bw + x] != exp) ++bad;
    }
    printf("mismatches=%d of %d\n", bad, nf);
    return bad != 0;
}
