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
#include "io16.cuh"

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
// Valid on the guarded fast path only (bwd_fast_ok below: finite product, 0 < d < inf); there the quotient has the
// sign of the product, which is merged in with one LOP3 because the correction computes (+0) + (-0) = +0 for a
// product of -0 where the true quotient is -0.
__device__ __forceinline__ float div_by_recip(float prod, double d, double r)
{
    const double p = (double)prod;
    double q = p * r;
    const double rem = __fma_rn(-q, d, p);
    q = __fma_rn(rem, r, q);
    return __uint_as_float((__float_as_uint((float)q) & 0x7fffffffu) | (__float_as_uint(prod) & 0x80000000u));
}
// r ~ 1/d for the fast path: the hardware's ~20-bit seed (MUFU.RCP64H) and two Newton steps leave a relative error
// e far below 2^-27, which is all the quotient needs -- after the residual correction in div_by_recip the error of q is
// e^2 + 2^-53.  The correctly rounded `1.0 / d` costs twice the DFMAs plus a guarded slow path per pixel, and this
// kernel is bound by instruction issue, not by HBM.  d is a normal double in [1e-9, 3.5e38] on the fast path.
__device__ __forceinline__ double recip_fast(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = __fma_rn(r, __fma_rn(-d, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-d, r, 1.0), r);
    return r;
}
// Per-pixel guard of the fast path.  y finite (and not NaN) means no square overflowed, so |x_c| <= y < 2^64 for every
// channel, and with |gy| <= 2^63 no product gy * x_c overflows; d = y + 1e-9 is then finite and positive.  Pixels that
// fail it (an overflowed norm -- common with fp16 storage --, inf / NaN anywhere, absurd gradients) take the literal
// fp64 divide, so that finite / inf is 0 and inf / finite is inf exactly as in the reference instead of the NaN
// the residual correction would make of 0 * inf.
__device__ __forceinline__ bool bwd_fast_ok(float g, float y) { return fabsf(g) <= 9.2233720e18f && y < __int_as_float(0x7f800000); }
__device__ __noinline__ float div_literal(float prod, double d) { return (float)((double)prod / d); }

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
        const double r0 = recip_fast(d0), r1 = recip_fast(d1), r2 = recip_fast(d2), r3 = recip_fast(d3);
        const size_t base = (b * c_n * hw4 + p) * 4;
        // compile-time channel counts: every load of the pixel group is in flight before the guard's branch (the kernel
        // is bound by memory latency per thread: with the x loads behind the branch it runs 9 % slower)
        float4 xv[CT > 0 ? CT : 1];
        if (CT > 0) {
#pragma unroll
            for (int c = 0; c < CT; ++c) xv[c] = ldg_stream4(x + base + (size_t)c * hw4 * 4);
        }
        if (__builtin_expect(bwd_fast_ok(g.x, yo.x) && bwd_fast_ok(g.y, yo.y) && bwd_fast_ok(g.z, yo.z) && bwd_fast_ok(g.w, yo.w), 1)) {
#pragma unroll 4
            for (int c = 0; c < c_n; ++c) {
                const size_t off = base + (size_t)c * hw4 * 4;
                const float4 v = CT > 0 ? xv[CT > 0 ? c : 0] : ldg_stream4(x + off);
                float4 o;
                o.x = div_by_recip(__fmul_rn(g.x, v.x), d0, r0);
                o.y = div_by_recip(__fmul_rn(g.y, v.y), d1, r1);
                o.z = div_by_recip(__fmul_rn(g.z, v.z), d2, r2);
                o.w = div_by_recip(__fmul_rn(g.w, v.w), d3, r3);
                stg_stream4(gx + off, o);
            }
        } else {
#pragma unroll
            for (int c = 0; c < c_n; ++c) {
                const size_t off = base + (size_t)c * hw4 * 4;
                const float4 v = CT > 0 ? xv[CT > 0 ? c : 0] : ldg_stream4(x + off);
                stg_stream4(gx + off, make_float4(div_literal(__fmul_rn(g.x, v.x), d0), div_literal(__fmul_rn(g.y, v.y), d1),
                                                  div_literal(__fmul_rn(g.z, v.z), d2), div_literal(__fmul_rn(g.w, v.w), d3)));
            }
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
        const double r = recip_fast(d);
        const float g = gy[i];
        const bool fast = bwd_fast_ok(g, y[i]);
        for (int c = 0; c < C; ++c) {
            const size_t off = (b * C + c) * hw + p;
            const float prod = __fmul_rn(g, x[off]);
            gx[off] = fast ? div_by_recip(prod, d, r) : div_literal(prod, d);
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


// ---------------------------------------------------------------------------------------------
// 16-bit storage (fp16 / bf16): the reference kernels instantiated for at::Half
// (channelnorm_kernel.cu:111,152 dispatch AT_DISPATCH_FLOATING_TYPES_AND_HALF).  Their arithmetic is NOT
// "fp32 kernel with casts": the square is formed in the storage type (`val * val` on two at::Half values is a
// float product rounded back to half, :55), the sum and the sqrt are fp32, the result is rounded once more.
// bf16 follows the same rules with bf16 rounding (the reference has no bf16 dispatch).
// One thread owns eight consecutive pixels: 128-bit loads / stores of 16-bit elements.
// Algorithmic bytes per pixel: fwd 2*(C+1), bwd 2*(2C+2).
// ---------------------------------------------------------------------------------------------
template <typename T, int CT>
__global__ void __launch_bounds__(256) cnorm16_fwd_v8(const uint4 *__restrict__ x, uint4 *__restrict__ y,
                                                      int C, unsigned hw8, size_t total8)
{
    const int c_n = CT > 0 ? CT : C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw8, p = i - b * hw8;
        const uint4 *src = x + b * c_n * hw8 + p;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int c = 0; c < c_n; ++c) {
            const uint4 raw = ldg_stream_u4(src + (size_t)c * hw8);
            float sq[8];       // val * val in the storage type (:55), one packed multiply per two elements
            unpack8<T>(make_uint4(square2_io<T>(raw.x), square2_io<T>(raw.y), square2_io<T>(raw.z), square2_io<T>(raw.w)), sq);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(acc[k], sq[k]);                               // :56
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = __fsqrt_rn(acc[k]);
        stg_stream_u4(y + i, pack8<T>(acc));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cnorm16_fwd_s(const unsigned short *__restrict__ x, unsigned short *__restrict__ y,
                                                     int C, size_t hw, size_t total)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = Io16<T>::to_float(ldg_stream_u16(x + (b * C + c) * hw + p));
            acc = __fadd_rn(acc, round_io<T>(__fmul_rn(v, v)));
        }
        y[i] = Io16<T>::from_float(__fsqrt_rn(acc));
    }
}

// channelnorm_kernel.cu:93-94: float(gO) * float(x) is an fp32 product, the divide by (float(out) + 1e-9) is fp64,
// the quotient is rounded to fp32 (`val`) and then to the storage type.
template <typename T, int CT>
__global__ void __launch_bounds__(256) cnorm16_bwd_v8(const uint4 *__restrict__ x, const uint4 *__restrict__ y,
                                                      const uint4 *__restrict__ gy, uint4 *__restrict__ gx,
                                                      int C, unsigned hw8, size_t total8)
{
    const int c_n = CT > 0 ? CT : C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw8, p = i - b * hw8;
        float yo[8], g[8];
        unpack8<T>(ldg_stream_u4(y + i), yo);
        unpack8<T>(ldg_stream_u4(gy + i), g);
        double d[8], r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { d[k] = (double)yo[k] + 1e-9; r[k] = recip_fast(d[k]); }
        const size_t base = b * c_n * hw8 + p;
        uint4 xv[CT > 0 ? CT : 1];                 // as in cnorm_bwd_v4: loads in flight before the guard's branch
        if (CT > 0) {
#pragma unroll
            for (int c = 0; c < CT; ++c) xv[c] = ldg_stream_u4(x + base + (size_t)c * hw8);
        }
        bool fast = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) fast = fast && bwd_fast_ok(g[k], yo[k]);
        if (__builtin_expect(fast, 1)) {
#pragma unroll(CT > 0 ? CT : 2)
            for (int c = 0; c < c_n; ++c) {
                const size_t off = base + (size_t)c * hw8;
                float v[8], o[8];
                unpack8<T>(CT > 0 ? xv[CT > 0 ? c : 0] : ldg_stream_u4(x + off), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = div_by_recip(__fmul_rn(g[k], v[k]), d[k], r[k]);
                stg_stream_u4(gx + off, pack8<T>(o));
            }
        } else {
#pragma unroll
            for (int c = 0; c < c_n; ++c) {
                const size_t off = base + (size_t)c * hw8;
                float v[8], o[8];
                unpack8<T>(CT > 0 ? xv[CT > 0 ? c : 0] : ldg_stream_u4(x + off), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = div_literal(__fmul_rn(g[k], v[k]), d[k]);
                stg_stream_u4(gx + off, pack8<T>(o));
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cnorm16_bwd_s(const unsigned short *__restrict__ x, const unsigned short *__restrict__ y,
                                                     const unsigned short *__restrict__ gy, unsigned short *__restrict__ gx,
                                                     int C, size_t hw, size_t total)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const float yf = Io16<T>::to_float(y[i]);
        const double d = (double)yf + 1e-9;
        const double r = recip_fast(d);
        const float g = Io16<T>::to_float(gy[i]);
        const bool fast = bwd_fast_ok(g, yf);
        for (int c = 0; c < C; ++c) {
            const size_t off = (b * C + c) * hw + p;
            const float prod = __fmul_rn(g, Io16<T>::to_float(x[off]));
            gx[off] = Io16<T>::from_float(fast ? div_by_recip(prod, d, r) : div_literal(prod, d));
        }
    }
}

template <typename T>
static int cnorm16_fwd_launch(const void *x, void *y, int B, int C, int H, int W, cudaStream_t st)
{
    const size_t hw = (size_t)H * W;
    if (hw % 8 == 0 && aligned16(x) && aligned16(y)) {
        const size_t total8 = (size_t)B * (hw / 8);
        const unsigned grid = grid_for(total8, 256);
        const uint4 *xs = static_cast<const uint4 *>(x);
        uint4 *ys = static_cast<uint4 *>(y);
        if (C == 3) cnorm16_fwd_v8<T, 3><<<grid, 256, 0, st>>>(xs, ys, C, (unsigned)(hw / 8), total8);
        else if (C == 2) cnorm16_fwd_v8<T, 2><<<grid, 256, 0, st>>>(xs, ys, C, (unsigned)(hw / 8), total8);
        else cnorm16_fwd_v8<T, 0><<<grid, 256, 0, st>>>(xs, ys, C, (unsigned)(hw / 8), total8);
    } else {
        const size_t total = (size_t)B * hw;
        cnorm16_fwd_s<T><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const unsigned short *>(x),
                                                               static_cast<unsigned short *>(y), C, hw, total);
    }
    return check_launch("cnorm_fwd_16");
}

template <typename T>
static int cnorm16_bwd_launch(const void *x, const void *y, const void *gy, void *gx, int B, int C, int H, int W,
                              cudaStream_t st)
{
    const size_t hw = (size_t)H * W;
    if (hw % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(gy) && aligned16(gx)) {
        const size_t total8 = (size_t)B * (hw / 8);
        const unsigned grid = grid_for(total8, 256);
        const uint4 *xs = static_cast<const uint4 *>(x), *ys = static_cast<const uint4 *>(y), *gs = static_cast<const uint4 *>(gy);
        uint4 *os = static_cast<uint4 *>(gx);
        if (C == 3) cnorm16_bwd_v8<T, 3><<<grid, 256, 0, st>>>(xs, ys, gs, os, C, (unsigned)(hw / 8), total8);
        else if (C == 2) cnorm16_bwd_v8<T, 2><<<grid, 256, 0, st>>>(xs, ys, gs, os, C, (unsigned)(hw / 8), total8);
        else cnorm16_bwd_v8<T, 0><<<grid, 256, 0, st>>>(xs, ys, gs, os, C, (unsigned)(hw / 8), total8);
    } else {
        const size_t total = (size_t)B * hw;
        cnorm16_bwd_s<T><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const unsigned short *>(x),
                                                               static_cast<const unsigned short *>(y),
                                                               static_cast<const unsigned short *>(gy),
                                                               static_cast<unsigned short *>(gx), C, hw, total);
    }
    return check_launch("cnorm_bwd_16");
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

extern "C" int flowops_cnorm_fwd_16(const void *x, void *y, int B, int C, int H, int W, int dtype, void *stream)
{
    FLOWOPS_REQUIRE(x && y, FLOWOPS_EINVAL, "cnorm_fwd_16: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "cnorm_fwd_16: bad shape %dx%dx%dx%d", B, C, H, W);
    FLOWOPS_REQUIRE(dtype16_ok(dtype), FLOWOPS_EINVAL, "cnorm_fwd_16: dtype must be FLOWOPS_DTYPE_F16 or FLOWOPS_DTYPE_BF16, got %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == FLOWOPS_DTYPE_F16 ? cnorm16_fwd_launch<__half>(x, y, B, C, H, W, st)
                                      : cnorm16_fwd_launch<__nv_bfloat16>(x, y, B, C, H, W, st);
}

extern "C" int flowops_cnorm_bwd_16(const void *x, const void *y, const void *gy, void *gx,
                                    int B, int C, int H, int W, int dtype, void *stream)
{
    FLOWOPS_REQUIRE(x && y && gy && gx, FLOWOPS_EINVAL, "cnorm_bwd_16: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "cnorm_bwd_16: bad shape %dx%dx%dx%d", B, C, H, W);
    FLOWOPS_REQUIRE(dtype16_ok(dtype), FLOWOPS_EINVAL, "cnorm_bwd_16: dtype must be FLOWOPS_DTYPE_F16 or FLOWOPS_DTYPE_BF16, got %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == FLOWOPS_DTYPE_F16 ? cnorm16_bwd_launch<__half>(x, y, gy, gx, B, C, H, W, st)
                                      : cnorm16_bwd_launch<__nv_bfloat16>(x, y, gy, gx, B, C, H, W, st);
}
