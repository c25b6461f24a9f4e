// io16.cuh -- 16-bit storage (fp16 / bf16) helpers for the "16-bit storage, fp32 math" operator variants
// (reference: flownet2_pytorch/main.py:59 "pseudo-fp16 mode"; models.py:22-28 fp16_resample2d, FlowNetC.py:86-87,
// channelnorm_kernel.cu's at::Half instantiation, models/base_model.py:123-127).
//
// Elements travel as raw 16-bit patterns; Io16<T> converts.  float -> T is round-to-nearest-even with the
// hardware conversion instructions -- the same ones ATen's `.half()` / `.bfloat16()` and c10::Half's
// float constructor compile to on the device, NaN and overflow behaviour included.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace flowops {

template <typename T> struct Io16;
template <> struct Io16<__half> {
    static __device__ __forceinline__ float to_float(unsigned short b) { return __half2float(__ushort_as_half(b)); }
    static __device__ __forceinline__ unsigned short from_float(float f) { return __half_as_ushort(__float2half_rn(f)); }
    // two floats -> packed pair (lo in bits 0..15): one F2FP on the ALU pipe instead of two F2F on the XU pipe
    static __device__ __forceinline__ unsigned pack2(float lo, float hi)
    {
        unsigned r;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};
template <> struct Io16<__nv_bfloat16> {
    static __device__ __forceinline__ float to_float(unsigned short b) { return __uint_as_float((unsigned)b << 16); }
    static __device__ __forceinline__ unsigned short from_float(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
    static __device__ __forceinline__ unsigned pack2(float lo, float hi)
    {
        unsigned r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};
// identity for fp32 storage, so that shared code can be written once
template <> struct Io16<float> {};

// the value a float takes after a round trip through storage type T (T = float: unchanged)
template <typename T> __device__ __forceinline__ float round_io(float f) { return Io16<T>::to_float(Io16<T>::from_float(f)); }
template <> __device__ __forceinline__ float round_io<float>(float f) { return f; }

// ---- streaming (read-once / write-once) accesses, L1-bypassing ----
__device__ __forceinline__ unsigned short ldg_stream_u16(const unsigned short *p)
{
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream_u16(unsigned short *p, unsigned short v)
{
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" :: "l"(p), "h"(v) : "memory");
}
__device__ __forceinline__ void stg_stream_u4(uint4 *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// eight consecutive 16-bit elements <-> eight floats
template <typename T> __device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8])
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[2 * k] = Io16<T>::to_float((unsigned short)(w[k] & 0xffffu));
        f[2 * k + 1] = Io16<T>::to_float((unsigned short)(w[k] >> 16));
    }
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = Io16<T>::pack2(f[2 * k], f[2 * k + 1]);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Two packed elements squared IN the storage type: (float)a * (float)a rounded to T.  The float product of two 16-bit
// values is exact (at most 22 significant bits), so the packed multiply's single rounding gives the same bits as
// widening, multiplying and narrowing -- subnormals, overflow to inf and NaN included -- in one instruction per pair.
template <typename T> __device__ __forceinline__ unsigned square2_io(unsigned packed);
template <> __device__ __forceinline__ unsigned square2_io<__half>(unsigned packed)
{
    unsigned r;
    asm("mul.rn.f16x2 %0, %1, %1;" : "=r"(r) : "r"(packed));
    return r;
}
template <> __device__ __forceinline__ unsigned square2_io<__nv_bfloat16>(unsigned packed)
{
    unsigned r;
    asm("mul.rn.bf16x2 %0, %1, %1;" : "=r"(r) : "r"(packed));
    return r;
}

static inline bool dtype16_ok(int dtype) { return dtype == FLOWOPS_DTYPE_F16 || dtype == FLOWOPS_DTYPE_BF16; }

}  // namespace flowops
