// tc.cuh -- tcgen05 / TMEM / 5-D TMA primitives shared by the tensor-core Correlation kernels (sm_100a).
#pragma once
#include "tma.cuh"

namespace flowops {
namespace tc {

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol error traps (and surfaces as a launch failure) instead of hanging the GPU
__device__ __forceinline__ long long mbar_wait_b(uint32_t bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return 0;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > (1ll << 32)) __trap();
    return clock64() - t0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        :: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v)
{
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v)
{
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits 0-13,
// leading byte offset (unused for swizzled K-major, 1) in 16-29, stride byte offset = 256 B (one 8-row core matrix
// of 32-byte rows) >> 4 in 32-45, descriptor version 1 (Blackwell) in 46-47, layout type SWIZZLE_32B = 6 in 61-63.
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32 (1 << 4), A and B = TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

}  // namespace tc

// 5-D fp32 tensor map with the 32-byte swizzle (32-byte innermost rows), out-of-bounds elements read as zero.
static inline int encode_map5_sw32(CUtensorMap *tm, const void *base, const cuuint64_t dims[5], const cuuint64_t strides[4],
                            const cuuint32_t box[5], const char *who, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_32B)
{
    EncodeTiledFn enc = get_encoder();
    FLOWOPS_REQUIRE(enc, FLOWOPS_EUNSUPPORTED, "%s: cuTensorMapEncodeTiled is not available from the driver", who);
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLOWOPS_REQUIRE(r == CUDA_SUCCESS, FLOWOPS_EINVAL, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
    return 0;
}

}  // namespace flowops
