// common.cuh -- shared helpers for libflowops (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flowops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libflowops is written for sm_100a (B200) only"
#endif

namespace flowops {

constexpr int kNumSMs = 148;  // B200; used for grid sizing only (grids are still correct elsewhere)

void set_error(const char *fmt, ...);
int check_launch(const char *what);

#define FLOWOPS_REQUIRE(cond, code, ...)      \
    do {                                      \
        if (!(cond)) {                        \
            flowops::set_error(__VA_ARGS__);  \
            return (code);                    \
        }                                     \
    } while (0)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- streaming 128-bit global accesses (read-once / write-once data: keep it out of L1) ----
__device__ __forceinline__ float4 ldg_stream4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// fire-and-forget fp32 reduction (RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

}  // namespace flowops
