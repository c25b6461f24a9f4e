// tma.cuh -- TMA / mbarrier / packed-FMA primitives shared by the Correlation kernels (sm_100a).
#pragma once
#include <cuda.h>   // CUtensorMap + enums only; the encoder is resolved at run time (no libcuda link)

#include "common.cuh"

namespace flowops {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        :: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// packed FP32 FMA (Blackwell FFMA2): d = a * b + c on both halves, round-to-nearest each
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<unsigned long long *>(&d))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return d;
}

// acc / nelems (reference correlation_cuda_kernel.cu:143) without the call-based divide sequence, which
// would spill the 168 live accumulators: q = t * (1/n), one residual correction.  Exact for a power-of-
// two channel count (FlowNetC: 256) and correctly rounded otherwise up to rare last-bit cases.
__device__ __forceinline__ float div_nelems(float t, float n, float inv_n)
{
    const float q = t * inv_n;
    return __fmaf_rn(__fmaf_rn(-q, n, t), inv_n, q);
}

// ---- host side: cuTensorMapEncodeTiled through the runtime's driver entry point -------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}


// 4-D fp32 tensor map, no swizzle/interleave, out-of-bounds elements read as zero.
// dims/strides are given innermost first; strides (bytes) for dimensions 1..3.
static inline int encode_map4(CUtensorMap *tm, const void *base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                              const cuuint32_t box[4], const char *who)
{
    EncodeTiledFn enc = get_encoder();
    FLOWOPS_REQUIRE(enc, FLOWOPS_EUNSUPPORTED, "%s: cuTensorMapEncodeTiled is not available from the driver", who);
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLOWOPS_REQUIRE(r == CUDA_SUCCESS, FLOWOPS_EINVAL, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
    return 0;
}

}  // namespace flowops
