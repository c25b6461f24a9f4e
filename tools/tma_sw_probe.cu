// tma_sw_probe.cu -- where does a 5-D TMA box with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and a 32-byte inner dimension land in
// shared memory?  Global tensor = P8 planes [cb=4][Y=8][X=8][8] with value cb*1000 + Y*100 + X*10 + c%8, map as in
// csrc/corr_tc_bwd.cu: dims (c%8, cb%4, X, Y, group).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_sw_probe tma_sw_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int c1, int c2, int c3, int bytes)
{
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    float *dst = reinterpret_cast<float *>(raw + (base - smem_u32(raw)));
    __shared__ __align__(8) uint64_t bar_mem;
    const uint32_t bar = smem_u32(&bar_mem);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) dst[i] = -1.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     :: "r"(base), "l"(&tm), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(0), "r"(bar) : "memory");
    }
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = dst[i];
}

int main(int argc, char **argv)
{
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", (void **)&enc, 12000, cudaEnableDefault, &q);
    const int CB = 4, PH = 8, PW = 8;
    float *h = (float *)malloc(CB * PH * PW * 8 * 4);
    for (int cb = 0; cb < CB; ++cb) for (int Y = 0; Y < PH; ++Y) for (int X = 0; X < PW; ++X) for (int c = 0; c < 8; ++c)
        h[((cb * PH + Y) * PW + X) * 8 + c] = cb * 1000 + Y * 100 + X * 10 + c;
    float *d, *o;
    cudaMalloc(&d, CB * PH * PW * 8 * 4);
    cudaMalloc(&o, 512 * 4);
    cudaMemcpy(d, h, CB * PH * PW * 8 * 4, cudaMemcpyHostToDevice);
    const cuuint64_t row = 32, line = PW * row, img = line * PH;
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int order = variant / 10, swz = variant % 10;
    // order 0: (c%8, cb%4, X, Y, group)   order 1: (c%8, X, Y, cb, 1) natural   order 2: (c%8, X, cb%4, Y, group)
    const cuuint64_t dims0[5] = {8, 4, PW, PH, CB / 4}, str0[4] = {img, row, line, 4 * img};
    const cuuint64_t dims1[5] = {8, PW, PH, CB, 1}, str1[4] = {row, line, img, CB * img};
    const cuuint64_t dims2[5] = {8, PW, 4, PH, CB / 4}, str2[4] = {row, img, line, 4 * img};
    const cuuint32_t box0[5] = {8, 4, 4, 2, 1}, box1[5] = {8, 4, 2, 4, 1}, box2[5] = {8, 4, 4, 2, 1};
    // order 3: the same memory seen as [Y'=32][X=8][32 floats] rows of 128 bytes: dims (32, X, Y', 1, 1), value = linear index
    const cuuint64_t dims3[5] = {32, 8, 8, 1, 1}, str3[4] = {128, 1024, 8192, 8192};
    const cuuint32_t box3[5] = {32, 4, 2, 1, 1};
    if (order == 3) {
        for (int i = 0; i < CB * PH * PW * 8; ++i) h[i] = (float)i;
        cudaMemcpy(d, h, CB * PH * PW * 8 * 4, cudaMemcpyHostToDevice);
    }
    const cuuint64_t *dims = order == 0 ? dims0 : order == 1 ? dims1 : order == 2 ? dims2 : dims3;
    const cuuint64_t *strides = order == 0 ? str0 : order == 1 ? str1 : order == 2 ? str2 : str3;
    const cuuint32_t *box = order == 0 ? box0 : order == 1 ? box1 : order == 2 ? box2 : box3;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    float ho[512];
    const CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== order %d swizzle enum %d: encode %d\n", order, (int)sw, (int)r);
    if (r) return 0;
    probe<<<1, 128, 4096>>>(tm, o, order == 0 ? 0 : order == 3 ? 0 : 2, order == 0 ? 2 : order == 1 ? 1 : 0, order == 2 ? 1 : order == 0 ? 1 : 0, 1024);
    cudaError_t e = cudaDeviceSynchronize();
    printf("   run: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    for (int r32 = 0; r32 < 34; ++r32) {       // 32-byte rows
        printf("  +%4d:", r32 * 32);
        for (int i = 0; i < 8; ++i) printf(" %5d", (int)ho[r32 * 8 + i]);
        printf("\n");
    }
    return 0;
}
