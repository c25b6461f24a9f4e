// cnorm.cu -- ChannelNorm forward / backward for sm_100a.
//
// Replaces kernel_channelnorm_update_output / kernel_channelnorm_backward_input1
// (reference channelnorm_package/channelnorm_kernel.cu:19-96).  Both are pure streaming kernels:
// every byte is touched once, so they are written for the HBM roofline -- one thread owns four
// consecutive pixels, every access is a 128-bit L1-bypassing load/store, and the grid is a whole
// number of waves of 148 SMs.
//
// Algorithmic bytes per pixel: fwd 4*(C+1), bwd 4*(2C+2).
//
// Arithmetic is kept bit-compatible with the reference: the forward is an FFMA chain in channel
// order followed by an IEEE sqrt; the backward forms gy*x in fp32 and divides by (y + 1e-9) in
// fp64.  The fp64 divide is done once per pixel as a reciprocal and each channel's quotient is
// Newton-corrected with two DFMAs, which reproduces the correctly rounded fp64 quotient before
// the final rounding to fp32.
#include "common.cuh"

namespace flowops {

template <int CT>  // CT > 0: compile-time channel count, 0: runtime
__global__ void __launch_bounds__(256) cnorm_fwd_v4(const float *__restrict__ x, float *__restrict__ y,
                                                    int C, unsigned hw4, size_t total4)
{
    const int c_n = CT > 0 ? CT : C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw4, p = i - b * hw4;
        const float *src = x + (b * c_n * hw4 + p) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int c = 0; c < c_n; ++c) {
            const float4 v = ldg_stream4(src + (size_t)c * hw4 * 4);
            acc.x = __fmaf_rn(v.x, v.x, acc.x);
            acc.y = __fmaf_rn(v.y, v.y, acc.y);
            acc.z = __fmaf_rn(v.z, v.z, acc.z);
            acc.w = __fmaf_rn(v.w, v.w, acc.w);
        }
        stg_stream4(y + i * 4, make_float4(__fsqrt_rn(acc.x), __fsqrt_rn(acc.y),
                                           __fsqrt_rn(acc.z), __fsqrt_rn(acc.w)));
    }
}

// scalar path for H*W not a multiple of 4 (or unaligned base pointers)
__global__ void __launch_bounds__(256) cnorm_fwd_s(const float *__restrict__ x, float *__restrict__ y,
                                                   int C, size_t hw, size_t total)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = ldg_stream(x + (b * C + c) * hw + p);
            acc = __fmaf_rn(v, v, acc);
        }
        y[i] = __fsqrt_rn(acc);
    }
}

// (float)((double)(g*x) / d) with r = 1/d precomputed: q = p*r, then one residual correction.
__device__ __forceinline__ float div_by_recip(float prod, double d, double r)
{
    const double p = (double)prod;
    double q = p * r;
    const double rem = __fma_rn(-q, d, p);
    q = __fma_rn(rem, r, q);
    return (float)q;
}

template <int CT>
__global__ void __launch_bounds__(256) cnorm_bwd_v4(const float *__restrict__ x, const float *__restrict__ y,
                                                    const float *__restrict__ gy, float *__restrict__ gx,
                                                    int C, unsigned hw4, size_t total4)
{
    const int c_n = CT > 0 ? CT : C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw4, p = i - b * hw4;
        const float4 yo = ldg_stream4(y + i * 4);
        const float4 g = ldg_stream4(gy + i * 4);
        const double d0 = (double)yo.x + 1e-9, d1 = (double)yo.y + 1e-9;
        const double d2 = (double)yo.z + 1e-9, d3 = (double)yo.w + 1e-9;
        const double r0 = 1.0 / d0, r1 = 1.0 / d1, r2 = 1.0 / d2, r3 = 1.0 / d3;
        const size_t base = (b * c_n * hw4 + p) * 4;
#pragma unroll 4
        for (int c = 0; c < c_n; ++c) {
            const size_t off = base + (size_t)c * hw4 * 4;
            const float4 v = ldg_stream4(x + off);
            float4 o;
            o.x = div_by_recip(__fmul_rn(g.x, v.x), d0, r0);
            o.y = div_by_recip(__fmul_rn(g.y, v.y), d1, r1);
            o.z = div_by_recip(__fmul_rn(g.z, v.z), d2, r2);
            o.w = div_by_recip(__fmul_rn(g.w, v.w), d3, r3);
            stg_stream4(gx + off, o);
        }
    }
}

__global__ void __launch_bounds__(256) cnorm_bwd_s(const float *__restrict__ x, const float *__restrict__ y,
                                                   const float *__restrict__ gy, float *__restrict__ gx,
                                                   int C, size_t hw, size_t total)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const double d = (double)y[i] + 1e-9;
        const double r = 1.0 / d;
        const float g = gy[i];
        for (int c = 0; c < C; ++c) {
            const size_t off = (b * C + c) * hw + p;
            gx[off] = div_by_recip(__fmul_rn(g, x[off]), d, r);
        }
    }
}

static inline unsigned grid_for(size_t work_items, int block, int waves_cap = 16)
{
    size_t blocks = (work_items + block - 1) / block;
    const size_t cap = (size_t)kNumSMs * 8 * waves_cap;   // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_cnorm_fwd(const float *x, float *y, int B, int C, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(x && y, FLOWOPS_EINVAL, "cnorm_fwd: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "cnorm_fwd: bad shape %dx%dx%dx%d", B, C, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)H * W;
    if (hw % 4 == 0 && aligned16(x) && aligned16(y)) {
        const size_t total4 = (size_t)B * (hw / 4);
        const unsigned grid = grid_for(total4, 256);
        if (C == 3) cnorm_fwd_v4<3><<<grid, 256, 0, st>>>(x, y, C, (unsigned)(hw / 4), total4);
        else if (C == 2) cnorm_fwd_v4<2><<<grid, 256, 0, st>>>(x, y, C, (unsigned)(hw / 4), total4);
        else cnorm_fwd_v4<0><<<grid, 256, 0, st>>>(x, y, C, (unsigned)(hw / 4), total4);
    } else {
        const size_t total = (size_t)B * hw;
        cnorm_fwd_s<<<grid_for(total, 256), 256, 0, st>>>(x, y, C, hw, total);
    }
    return check_launch("cnorm_fwd");
}

extern "C" int flowops_cnorm_bwd(const float *x, const float *y, const float *gy, float *gx,
                                 int B, int C, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(x && y && gy && gx, FLOWOPS_EINVAL, "cnorm_bwd: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "cnorm_bwd: bad shape %dx%dx%dx%d", B, C, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)H * W;
    if (hw % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(gy) && aligned16(gx)) {
        const size_t total4 = (size_t)B * (hw / 4);
        const unsigned grid = grid_for(total4, 256);
        if (C == 3) cnorm_bwd_v4<3><<<grid, 256, 0, st>>>(x, y, gy, gx, C, (unsigned)(hw / 4), total4);
        else if (C == 2) cnorm_bwd_v4<2><<<grid, 256, 0, st>>>(x, y, gy, gx, C, (unsigned)(hw / 4), total4);
        else cnorm_bwd_v4<0><<<grid, 256, 0, st>>>(x, y, gy, gx, C, (unsigned)(hw / 4), total4);
    } else {
        const size_t total = (size_t)B * hw;
        cnorm_bwd_s<<<grid_for(total, 256), 256, 0, st>>>(x, y, gy, gx, C, hw, total);
    }
    return check_launch("cnorm_bwd");
}
