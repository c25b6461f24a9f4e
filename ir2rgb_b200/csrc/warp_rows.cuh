// warp_rows.cuh -- the row-pipelined forward gather shared by flowops_warp_fwd (warp.cu) and the fused
// FlowNet2 glue kernels (fused.cu).
//
// Why it looks like this (ncu, profiles/ncu_warp_fwd_smooth_r01.txt): a thread-per-pixel forward warp is
// neither HBM- nor L1-bound but LATENCY-bound (long_scoreboard 9.6 of 16 stall cycles): every pixel pays a
// flow load and then a dependent gather, one after the other, and the bit-exact fp64 weights of the
// reference keep the conversion (XU) pipe 63 % busy, with ~240 instructions per pixel, most of them 64-bit
// address arithmetic.  Here a thread walks `rows` consecutive rows of one column:
//   * the flow of row y+1 is in flight while row y is gathered and blended, so the two dependent latencies
//     overlap (a deeper pipeline with double-buffered gathers was measured: no better, 24 more registers);
//   * consecutive rows share their corner rows in L1 (sector hit rate 57 % -> 72 %);
//   * the batch index is block-uniform and corners are 32-bit offsets from one opaque base: one IMAD.WIDE per
//     corner row, the right-hand neighbour is an immediate +4 (134 instructions per pixel, 37 of them the
//     reference's fp64 arithmetic);
//   * floor / int<->float conversions run on the FP32/ALU pipes with exact magic-number arithmetic, which
//     leaves the XU pipe only the 20 F2F conversions per pixel that the reference's mixed precision forces.
// Measured on B200, config 3, smooth flow: 110 -> 81 us (RESAMPLE2D), 114 -> 80 us (GRIDSAMPLE), bit-identical.
#pragma once
#include "io16.cuh"
#include "warp.cuh"

namespace flowops {

constexpr float kMagic15 = 12582912.f;   // 1.5 * 2^23: ulp is 1 on [2^23, 2^24)

// Tuning macro (variant builds: python -m ir2rgb_b200.build --out ... -DFLOWOPS_TUNE_WARP2D=4): shape of the patch of
// pixels one warp of the forward row-walking kernel covers: 0 = 32 x 1 (a row segment), 2 = 16 x 2, 4 = 8 x 4.
#ifndef FLOWOPS_TUNE_WARP2D
#define FLOWOPS_TUNE_WARP2D 0
#endif
#if FLOWOPS_TUNE_WARP2D
constexpr int kRowStep = FLOWOPS_TUNE_WARP2D, kWarpCols = 32 / FLOWOPS_TUNE_WARP2D;
#endif

// v is an integer-valued float in [0, 2^22)
__device__ __forceinline__ int small_float_as_int(float v)
{
    return __float_as_int(__fadd_rn(v, kMagic15)) - 0x4B400000;
}
// i in [0, 2^23)
__device__ __forceinline__ float small_int_as_float(int i)
{
    return __fsub_rn(__int_as_float(0x4B000000 | i), 8388608.f);
}

struct WarpArgs {
    const float *img;  size_t img_bs;    // gather source [B,C,H,W] view, batch stride in floats
    const float *flow;                   // [B,2,H,W] contiguous, pixels
    const float *ref;  size_t ref_bs;    // fused epilogues: the frame the warp is compared with
    float *out;        size_t out_bs;    // warped frame (may be null for DIFF_NORM)
    float *aux;        size_t aux_bs;    // DIFF_NORM: norm plane; CONF: mask plane
    int B, C, H, W, rows;
    float wm1, hm1;                      // (float)(W-1), (float)(H-1)
    const float *lin_x, *lin_y;          // GRIDSAMPLE tables
    float invx, invy, thresh;
    int c_dst;                           // CONCAT: channels per pixel of the channels-last destination
    float inv_div_flow;                  // CONCAT: the flow is stored times 1/div_flow (models.py:112; torch divides a CUDA
                                         // tensor by a Python scalar as a multiply by the fp32 reciprocal)
    const float *flow_lo;                // UP4: [B,2,H/4,W/4] network-unit flow; the kernel forms
    float flow_mul;                      //      upsample_bilinear_x4(flow_lo * flow_mul) itself (models.py:106,118)
};

// nn.Upsample(scale_factor=4, mode='bilinear') (align_corners=False) of (lo * mul) at full-resolution pixel (x, y):
// ATen upsample_bilinear2d -- source index 0.25 * (dst + 0.5) - 0.5 clamped at 0, neighbour +1 unless at the last
// row / column, value = h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11).  Index and weights are exact in
// fp32 (multiples of 1/8); the blend differs from ATen's compiled kernel at most in FMA contraction (<= 2 ulp).
struct Up4Col { unsigned x1; unsigned x1p; float w0, w1; };
__device__ __forceinline__ Up4Col up4_col(int x, int Wl)
{
    Up4Col c;
    float sx = __fmaf_rn(0.25f, (float)x + 0.5f, -0.5f);
    sx = sx < 0.f ? 0.f : sx;
    const int x1 = (int)sx;
    c.x1 = (unsigned)x1; c.x1p = x1 < Wl - 1 ? 1u : 0u;
    c.w1 = sx - (float)x1; c.w0 = 1.f - c.w1;
    return c;
}
__device__ __forceinline__ void up4_flow(const float *__restrict__ lo, unsigned hwl, int Wl, int Hl, const Up4Col &c, int y,
                                         float mul, float &dx, float &dy)
{
    float sy = __fmaf_rn(0.25f, (float)y + 0.5f, -0.5f);
    sy = sy < 0.f ? 0.f : sy;
    const int y1 = (int)sy;
    const unsigned y1p = y1 < Hl - 1 ? (unsigned)Wl : 0u;
    const float h1 = sy - (float)y1, h0 = 1.f - h1;
    const float *p = lo + (unsigned)y1 * (unsigned)Wl + c.x1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float v00 = __fmul_rn(__ldg(p), mul), v01 = __fmul_rn(__ldg(p + c.x1p), mul);
        const float v10 = __fmul_rn(__ldg(p + y1p), mul), v11 = __fmul_rn(__ldg(p + y1p + c.x1p), mul);
        const float top = __fmaf_rn(c.w1, v01, __fmul_rn(c.w0, v00)), bot = __fmaf_rn(c.w1, v11, __fmul_rn(c.w0, v10));
        const float v = __fmaf_rn(h1, bot, __fmul_rn(h0, top));
        if (k == 0) dx = v; else dy = v;
        p += hwl;
    }
}

// EPI_CONCAT (C = 3 only): the whole `concat1` / `concat2` tensor of models.py:112-114,124-126 --
// (frame 0, frame 1, warped frame 1, flow / div_flow, |frame 0 - warped|) -- written channels-last with the pixel
// padded to c_dst channels (zeros), i.e. the layout and channel count the next FlowNetS's first convolution wants
enum { EPI_STORE = 0, EPI_DIFF_NORM = 1, EPI_CONF = 2, EPI_CONCAT = 3 };

// per-pixel state between the pipeline stages.  Corner addressing is kept as two 32-bit row offsets
// (top-left, bottom-left) plus "the right-hand column is one to the right" -- one 64-bit address per
// corner ROW and an immediate +4 for the right neighbour, instead of four 64-bit addresses per channel.
template <int MODE> struct PixPrep;
template <> struct PixPrep<FLOWOPS_WARP_RESAMPLE2D> {
    unsigned o_t, o_b;      // yT*W + xL, yB*W + xL
    bool ex;                // xR == xL + 1 (false when both clamp to the same border column)
    R2dWeights w;
};
// Tolerance mode of RESAMPLE2D (internal; flowops_warp_set_impl bit 1): the same coordinates and corners, but the
// bilinear weights and the blend stay in fp32 -- the reference's three fp64 weight products are an accident of a
// `1.` literal (resample2d_kernel.cu:55-58), they cost 20 F2F conversions per pixel on the quarter-rate XU pipe, and the
// contract for floating point is max-relative 1e-5, which an fp32 blend meets by two orders of magnitude.
constexpr int kWarpResample2dF32 = 2;
template <> struct PixPrep<kWarpResample2dF32> {
    unsigned o_t, o_b;
    bool ex;
    float w_tl, w_tr, w_bl, w_br;
};
template <> struct PixPrep<FLOWOPS_WARP_GRIDSAMPLE> {
    unsigned o_t, o_b;      // iy_nw*W + ix_nw, and the row below (same row when it is out of bounds: never read)
    bool in_e, in_s;
    float nw, ne, sw, se;
};

// resample2d_kernel.cu:40-51 with the conversions moved off the XU pipe (same values bit for bit)
__device__ __forceinline__ void pix_prep(PixPrep<FLOWOPS_WARP_RESAMPLE2D> &q, const WarpArgs &a, float xfl, float yfl,
                                         int /*x*/, int /*y*/, float dx, float dy)
{
    const float xf = __fadd_rn(xfl, dx), yf = __fadd_rn(yfl, dy);
    // floor on the FP32 pipe: round to nearest integer with the 1.5*2^23 trick, step down if that rounded up.
    // Exact for |v| < 2^22; anything larger (or inf / NaN) takes the FRND path, one rare branch for both axes.
    float fx = __fsub_rn(__fadd_rn(xf, kMagic15), kMagic15); fx = fx > xf ? __fsub_rn(fx, 1.f) : fx;
    float fy = __fsub_rn(__fadd_rn(yf, kMagic15), kMagic15); fy = fy > yf ? __fsub_rn(fy, 1.f) : fy;
    if (__builtin_expect(!(fmaxf(fabsf(xf), fabsf(yf)) < 4194304.f), 0)) { fx = floorf(xf); fy = floorf(yf); }
    // max(min((int)f, n-1), 0) == (int)clamp(f, 0, n-1) for every f, NaN and inf included;
    // the right / bottom neighbour is a different pixel iff the floor lies in [0, n-2]
    const unsigned xL = (unsigned)small_float_as_int(fminf(fmaxf(fx, 0.f), a.wm1));
    const unsigned yT = (unsigned)small_float_as_int(fminf(fmaxf(fy, 0.f), a.hm1));
    q.ex = fx >= 0.f && fx < a.wm1;
    q.o_t = yT * (unsigned)a.W + xL;
    q.o_b = (fy >= 0.f && fy < a.hm1) ? q.o_t + (unsigned)a.W : q.o_t;
    const float alpha = __fsub_rn(xf, fx), beta = __fsub_rn(yf, fy);
    const double wa = 1. - (double)alpha, wb = 1. - (double)beta;
    q.w.w_tl = wa * wb; q.w.w_tr = (double)alpha * wb; q.w.w_bl = wa * (double)beta;
    q.w.w_br = __fmul_rn(alpha, beta);
}

__device__ __forceinline__ void pix_prep(PixPrep<kWarpResample2dF32> &q, const WarpArgs &a, float xfl, float yfl,
                                         int /*x*/, int /*y*/, float dx, float dy)
{
    const float xf = __fadd_rn(xfl, dx), yf = __fadd_rn(yfl, dy);
    float fx = __fsub_rn(__fadd_rn(xf, kMagic15), kMagic15); fx = fx > xf ? __fsub_rn(fx, 1.f) : fx;
    float fy = __fsub_rn(__fadd_rn(yf, kMagic15), kMagic15); fy = fy > yf ? __fsub_rn(fy, 1.f) : fy;
    if (__builtin_expect(!(fmaxf(fabsf(xf), fabsf(yf)) < 4194304.f), 0)) { fx = floorf(xf); fy = floorf(yf); }
    const unsigned xL = (unsigned)small_float_as_int(fminf(fmaxf(fx, 0.f), a.wm1));
    const unsigned yT = (unsigned)small_float_as_int(fminf(fmaxf(fy, 0.f), a.hm1));
    q.ex = fx >= 0.f && fx < a.wm1;
    q.o_t = yT * (unsigned)a.W + xL;
    q.o_b = (fy >= 0.f && fy < a.hm1) ? q.o_t + (unsigned)a.W : q.o_t;
    const float alpha = __fsub_rn(xf, fx), beta = __fsub_rn(yf, fy);
    const float wa = 1.f - alpha, wb = 1.f - beta;
    q.w_tl = wa * wb; q.w_tr = alpha * wb; q.w_bl = wa * beta; q.w_br = alpha * beta;
}

// models/networks.py:97-98 + ATen grid_sampler (align_corners=False, border); see gs_coords / gs_weights.
// T is the dtype the reference evaluates the grid in (the flow's: networks.py:96-98, base_model.py:131-134): for
// 16-bit flows the normalised flow and its sum with the (equally rounded) linspace table are rounded to T;
// T = float leaves both as they are.
template <typename T = float>
__device__ __forceinline__ void pix_prep(PixPrep<FLOWOPS_WARP_GRIDSAMPLE> &q, const WarpArgs &a, float /*xfl*/, float /*yfl*/,
                                         int x, int y, float dx, float dy)
{
    const float gx = round_io<T>(__fadd_rn(__ldg(a.lin_x + x), round_io<T>(__fmul_rn(dx, a.invx))));
    const float gy = round_io<T>(__fadd_rn(__ldg(a.lin_y + y), round_io<T>(__fmul_rn(dy, a.invy))));
    float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), a.wm1 + 1.f, -1.f), 0.5f);
    float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), a.hm1 + 1.f, -1.f), 0.5f);
    ix = fminf(a.wm1, fmaxf(ix, 0.f));
    iy = fminf(a.hm1, fmaxf(iy, 0.f));
    // the clipped coordinate is in [0, n-1]: the magic-number floor needs no slow path
    float fx = __fsub_rn(__fadd_rn(ix, kMagic15), kMagic15); fx = fx > ix ? __fsub_rn(fx, 1.f) : fx;
    float fy = __fsub_rn(__fadd_rn(iy, kMagic15), kMagic15); fy = fy > iy ? __fsub_rn(fy, 1.f) : fy;
    const float ex = __fadd_rn(fx, 1.f), ey = __fadd_rn(fy, 1.f);        // (float)ix_se, (float)iy_se
    q.nw = __fmul_rn(__fsub_rn(ex, ix), __fsub_rn(ey, iy));
    q.ne = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(ey, iy));
    q.sw = __fmul_rn(__fsub_rn(ex, ix), __fsub_rn(iy, fy));
    q.se = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
    q.in_e = ex <= a.wm1; q.in_s = ey <= a.hm1;
    q.o_t = (unsigned)small_float_as_int(fy) * (unsigned)a.W + (unsigned)small_float_as_int(fx);
    q.o_b = q.in_s ? q.o_t + (unsigned)a.W : q.o_t;
}

template <int CT> struct PixVals { float v[CT][4]; float r[CT]; float s[CT]; };    // corners, reference frame, source frame

template <int CT, bool NEED_REF, bool NEED_SELF>
__device__ __forceinline__ void pix_gather(PixVals<CT> &g, const PixPrep<FLOWOPS_WARP_RESAMPLE2D> &q,
                                           const float *__restrict__ src, const float *__restrict__ ref,
                                           const float *__restrict__ self, unsigned hw)
{
    const float *pt = src + q.o_t, *pb = src + q.o_b;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        // the right-hand loads must not wait for the left-hand values: a clamped column is patched in at blend time
        g.v[c][0] = __ldg(pt); g.v[c][2] = __ldg(pb);
        g.v[c][1] = q.ex ? __ldg(pt + 1) : 0.f;
        g.v[c][3] = q.ex ? __ldg(pb + 1) : 0.f;
        if (NEED_REF) { g.r[c] = ldg_stream(ref); ref += hw; }
        if (NEED_SELF) { g.s[c] = __ldg(self); self += hw; }
        pt += hw; pb += hw;
    }
}
template <int CT, bool NEED_REF, bool NEED_SELF>
__device__ __forceinline__ void pix_gather(PixVals<CT> &g, const PixPrep<kWarpResample2dF32> &q,
                                           const float *__restrict__ src, const float *__restrict__ ref,
                                           const float *__restrict__ self, unsigned hw)
{
    const float *pt = src + q.o_t, *pb = src + q.o_b;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        g.v[c][0] = __ldg(pt); g.v[c][2] = __ldg(pb);
        g.v[c][1] = q.ex ? __ldg(pt + 1) : 0.f;
        g.v[c][3] = q.ex ? __ldg(pb + 1) : 0.f;
        if (NEED_REF) { g.r[c] = ldg_stream(ref); ref += hw; }
        if (NEED_SELF) { g.s[c] = __ldg(self); self += hw; }
        pt += hw; pb += hw;
    }
}
template <int CT, bool NEED_REF, bool NEED_SELF>
__device__ __forceinline__ void pix_gather(PixVals<CT> &g, const PixPrep<FLOWOPS_WARP_GRIDSAMPLE> &q,
                                           const float *__restrict__ src, const float *__restrict__ ref,
                                           const float *__restrict__ /*self*/, unsigned hw)
{
    const float *pt = src + q.o_t, *pb = src + q.o_b;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        g.v[c][0] = __ldg(pt);
        g.v[c][1] = q.in_e ? __ldg(pt + 1) : 0.f;
        g.v[c][2] = q.in_s ? __ldg(pb) : 0.f;
        g.v[c][3] = (q.in_e && q.in_s) ? __ldg(pb + 1) : 0.f;
        if (NEED_REF) { g.r[c] = ldg_stream(ref); ref += hw; }
        pt += hw; pb += hw;
    }
}

__device__ __forceinline__ float pix_blend(const PixPrep<FLOWOPS_WARP_RESAMPLE2D> &q, const float (&v)[4])
{
    return r2d_blend(q.w, v[0], q.ex ? v[1] : v[0], v[2], q.ex ? v[3] : v[2]);
}
__device__ __forceinline__ float pix_blend(const PixPrep<kWarpResample2dF32> &q, const float (&v)[4])
{
    float val = q.w_tl * v[0];
    val = __fmaf_rn(q.w_tr, q.ex ? v[1] : v[0], val);
    val = __fmaf_rn(q.w_bl, v[2], val);
    return __fmaf_rn(q.w_br, q.ex ? v[3] : v[2], val);
}
// out_acc += val * weight in nw, ne, sw, se order, skipping out-of-bounds corners (ATen grid_sampler_2d_kernel)
__device__ __forceinline__ float pix_blend(const PixPrep<FLOWOPS_WARP_GRIDSAMPLE> &q, const float (&v)[4])
{
    float acc = __fmaf_rn(v[0], q.nw, 0.f);
    if (q.in_e) acc = __fmaf_rn(v[1], q.ne, acc);
    if (q.in_s) acc = __fmaf_rn(v[2], q.sw, acc);
    if (q.in_e && q.in_s) acc = __fmaf_rn(v[3], q.se, acc);
    return acc;
}

// `out` / `aux` point at this pixel in channel 0 of this batch item
template <int MODE, int CT, int EPI, bool WRITE_WARPED>
__device__ __forceinline__ void pix_finish(const WarpArgs &a, const PixPrep<MODE> &q, const PixVals<CT> &g,
                                           float *__restrict__ out, float *__restrict__ aux, unsigned hw, float dx, float dy)
{
    float acc = 0.f;
    float warped[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const float val = pix_blend(q, g.v[c]);
        warped[c] = val;
        if (WRITE_WARPED) { stg_stream(out, val); out += hw; }
        if (EPI == EPI_DIFF_NORM || EPI == EPI_CONCAT) {
            const float d = __fsub_rn(g.r[c], val);                   // img0 - warped (models.py:110)
            acc = __fmaf_rn(d, d, acc);                               // channelnorm_kernel.cu:55-56
        } else if (EPI == EPI_CONF) {
            const float d = __fsub_rn(g.r[c], val);
            // torch.sum(t*t, dim=1): products are rounded before they are added (flownet.py:56-57)
            acc = c == 0 ? __fmul_rn(d, d) : __fadd_rn(acc, __fmul_rn(d, d));
        }
    }
    if (EPI == EPI_DIFF_NORM) stg_stream(aux, __fsqrt_rn(acc));
    if (EPI == EPI_CONF) stg_stream(aux, acc < a.thresh ? 1.f : 0.f);
    if (EPI == EPI_CONCAT) {
        // written for 3-channel frames (the launcher also instantiates CT = 1, 2, which are never launched)
        float4 *dst = reinterpret_cast<float4 *>(aux);            // c_dst is a multiple of 4: 16-byte aligned pixels
        dst[0] = make_float4(g.r[0], g.r[1 % CT], g.r[2 % CT], g.s[0]);
        dst[1] = make_float4(g.s[1 % CT], g.s[2 % CT], warped[0], warped[1 % CT]);
        dst[2] = make_float4(warped[2 % CT], __fmul_rn(dx, a.inv_div_flow), __fmul_rn(dy, a.inv_div_flow), __fsqrt_rn(acc));
        for (int q4 = 3; q4 < a.c_dst / 4; ++q4) dst[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// grid: (W / blockDim.x, H / (blockDim.y * rows), B)  -- the batch index is block-uniform, so every base
// pointer below lives in uniform registers and per-thread addressing is 32-bit offsets
template <int MODE, int CT, int EPI, bool WRITE_WARPED, bool UP4 = false>
__global__ void __launch_bounds__(256, 5) warp_rows_kernel(const __grid_constant__ WarpArgs a)
{
    constexpr bool NEED_REF = EPI != EPI_STORE;
    constexpr bool NEED_SELF = EPI == EPI_CONCAT;
#if FLOWOPS_TUNE_WARP2D
    // a warp covers a 2-D patch (kWarpCols columns x kRowStep rows) instead of 32 x 1: its 32 gather targets then span
    // fewer image rows (fewer L1 lines per request) at the price of 32 B .. 64 B row segments for the coalesced accesses
    constexpr int RS = kRowStep;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, wrp = tid >> 5, lane = tid & 31;
    const int x = blockIdx.x * (8 * kWarpCols) + wrp * kWarpCols + (lane % kWarpCols);
    const int yb = blockIdx.y * (RS * a.rows);
    const int y0 = yb + lane / kWarpCols;
    if (x >= a.W || y0 >= a.H) return;
    const int y1 = min(yb + RS * a.rows, a.H);
#else
    constexpr int RS = 1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * a.rows;
    if (x >= a.W || y0 >= a.H) return;
    const int y1 = min(y0 + a.rows, a.H);
#endif
    const unsigned hw = (unsigned)a.H * a.W, W = (unsigned)a.W;
    const size_t b = blockIdx.z;
    const float *src = a.img + b * a.img_bs;
    asm("" : "+l"(src));     // keep the batch base as one opaque 64-bit value (no b*stride folded into every gather)
    const unsigned p0 = (unsigned)y0 * W + (unsigned)x;
    const float *fl = UP4 ? nullptr : a.flow + b * 2 * hw + p0;       // walks down the column
    const float *ref = NEED_REF ? a.ref + b * a.ref_bs + p0 : src;    // unused when !NEED_REF
    float *out = WRITE_WARPED ? a.out + b * a.out_bs + p0 : nullptr;
    const unsigned aux_px = EPI == EPI_CONCAT ? (unsigned)a.c_dst : 1u;      // floats per destination pixel
    float *aux = EPI != EPI_STORE ? a.aux + b * a.aux_bs + (size_t)p0 * aux_px : nullptr;
    const float *self = src + p0;                                            // CONCAT: frame 1 at this pixel
    const float xfl = small_int_as_float(x);
    float yfl = small_int_as_float(y0);
    // UP4: the flow is formed from the quarter-resolution field (L1 / L2 resident) instead of read at full resolution
    const int Wl = a.W >> 2, Hl = a.H >> 2;
    const unsigned hwl = (unsigned)Wl * Hl;
    const float *lo = UP4 ? a.flow_lo + b * 2 * hwl : nullptr;
    Up4Col ucol{};
    if (UP4) ucol = up4_col(x, Wl);
    float dx, dy;
    if (UP4) up4_flow(lo, hwl, Wl, Hl, ucol, y0, a.flow_mul, dx, dy);
    else { dx = ldg_stream(fl); dy = ldg_stream(fl + hw); }
    for (int y = y0; y < y1; y += RS) {
        float ndx = 0.f, ndy = 0.f;
        if (y + RS < y1) {                                                                // next row's flow
            if (UP4) up4_flow(lo, hwl, Wl, Hl, ucol, y + RS, a.flow_mul, ndx, ndy);
            else { ndx = ldg_stream(fl + RS * W); ndy = ldg_stream(fl + RS * W + hw); }
        }
        PixPrep<MODE> cur;
        PixVals<CT> vcur;
        pix_prep(cur, a, xfl, yfl, x, y, dx, dy);
        pix_gather<CT, NEED_REF, NEED_SELF>(vcur, cur, src, ref, self, hw);
        pix_finish<MODE, CT, EPI, WRITE_WARPED>(a, cur, vcur, out, aux, hw, dx, dy);
        dx = ndx; dy = ndy;
        fl += RS * W; ref += RS * W; self += RS * W; out += RS * W; aux += (size_t)(RS * W) * aux_px; yfl = __fadd_rn(yfl, (float)RS);
    }
}

// ---------------------------------------------------------------------------------------------
// The forward store with more loads in flight per thread (plain warps only: EPI_STORE, full-resolution flow).
// tools/warp_flow_sweep.py showed that the row-walking kernel above needs 70 us at config 3 even for a ZERO flow (0.59 of
// the HBM roofline, perfectly coalesced gathers): its time is set by the dependent chain flow -> gather -> blend of one row
// at a time, not by the access pattern of the flow.  Here a thread owns R rows of its column, loads the flow of all R rows
// at once (2R independent loads), and keeps the corner gathers of DEPTH rows in flight (12 * DEPTH loads for 3 channels)
// while it blends and stores the oldest one.  Fully unrolled: every buffer index is a compile-time constant.  Same
// pix_prep / pix_gather / pix_finish as above, so the results are bit-identical to the row-walking kernel in every mode.
// ---------------------------------------------------------------------------------------------
template <int MODE, int CT, int R, int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB) warp_rows_mlp_kernel(const __grid_constant__ WarpArgs a)
{
    static_assert(DEPTH >= 2 && DEPTH <= R, "DEPTH rows of gathers in flight out of R rows per thread");
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * R;
    if (x >= a.W || y0 >= a.H) return;
    const unsigned hw = (unsigned)a.H * a.W, W = (unsigned)a.W;
    const size_t b = blockIdx.z;
    const float *src = a.img + b * a.img_bs;
    asm("" : "+l"(src));
    const unsigned p0 = (unsigned)y0 * W + (unsigned)x;
    const float *fl = a.flow + b * 2 * hw + p0;
    float *out = a.out + b * a.out_bs + p0;
    const float xfl = small_int_as_float(x);
    float dx[R], dy[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned o = (y0 + r < a.H ? (unsigned)r : 0u) * W;          // rows below the image redo row y0 (never stored)
        dx[r] = ldg_stream(fl + o); dy[r] = ldg_stream(fl + o + hw);
    }
    PixPrep<MODE> q[DEPTH];
    PixVals<CT> v[DEPTH];
#pragma unroll
    for (int r = 0; r < DEPTH - 1; ++r) {
        const int yr = y0 + r < a.H ? y0 + r : y0;
        pix_prep(q[r], a, xfl, small_int_as_float(yr), x, yr, dx[r], dy[r]);
        pix_gather<CT, false, false>(v[r], q[r], src, src, src, hw);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        constexpr int D1 = DEPTH - 1;
        if (r + D1 < R) {
            const int yr = y0 + r + D1 < a.H ? y0 + r + D1 : y0;
            pix_prep(q[(r + D1) % DEPTH], a, xfl, small_int_as_float(yr), x, yr, dx[r + D1], dy[r + D1]);
            pix_gather<CT, false, false>(v[(r + D1) % DEPTH], q[(r + D1) % DEPTH], src, src, src, hw);
        }
        if (y0 + r < a.H)
            pix_finish<MODE, CT, EPI_STORE, true>(a, q[r % DEPTH], v[r % DEPTH], out + (unsigned)r * W, nullptr, hw, dx[r], dy[r]);
    }
}

template <int MODE, int CT, int R, int DEPTH, int MINB>
static inline void launch_warp_rows_mlp(WarpArgs a, cudaStream_t st)
{
    const int B = a.B;
    const size_t hw = (size_t)a.H * a.W;
    a.rows = R;
    for (int b0 = 0; b0 < B; b0 += 65535) {            // gridDim.z limit
        WarpArgs c = a;
        c.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + (size_t)b0 * a.img_bs;
        c.flow = a.flow + (size_t)b0 * 2 * hw;
        c.out = a.out + (size_t)b0 * a.out_bs;
        const int bx = c.W >= 256 ? 256 : (c.W > 64 ? 128 : (c.W > 32 ? 64 : 32));
        const dim3 block(bx, 256 / bx, 1);
        const int strip = (int)block.y * R;
        const dim3 grid((c.W + bx - 1) / bx, (c.H + strip - 1) / strip, c.B);
        warp_rows_mlp_kernel<MODE, CT, R, DEPTH, MINB><<<grid, block, 0, st>>>(c);
    }
}

// 256-thread blocks shaped to the image width: whole warps lie along x (coalescing), the block's rows
// are `rows` apart so that each thread walks its own strip
static inline void warp_rows_shape(int B, int H, int W, int rows, dim3 &grid, dim3 &block)
{
    const int bx = W >= 256 ? 256 : (W > 64 ? 128 : (W > 32 ? 64 : 32));
    block = dim3(bx, 256 / bx, 1);
    const int strip = (int)block.y * rows;
    grid = dim3((W + bx - 1) / bx, (H + strip - 1) / strip, B);     // callers split B > 65535
}

// rows per thread: long strips amortise the pipeline fill, but the grid should still be several waves
static inline int warp_rows_pick(int B, int H, int W, int rows = 4)
{
    const long long px = (long long)B * H * W;
    while (rows > 1 && px / (256LL * rows) < 4LL * kNumSMs * 5) rows >>= 1;
    return rows;
}

// C must be 1..3 (compile-time channel counts; the register pipeline is sized by them)
template <int MODE, int EPI, bool WRITE_WARPED, bool UP4 = false>
static inline void launch_warp_rows(WarpArgs a, cudaStream_t st)
{
    const int B = a.B;
    const size_t hw = (size_t)a.H * a.W;
    for (int b0 = 0; b0 < B; b0 += 65535) {            // gridDim.z limit
        WarpArgs c = a;
        c.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + (size_t)b0 * a.img_bs;
        if (a.flow) c.flow = a.flow + (size_t)b0 * 2 * hw;
        if (a.flow_lo) c.flow_lo = a.flow_lo + (size_t)b0 * 2 * (size_t)(a.H >> 2) * (a.W >> 2);
        if (a.ref) c.ref = a.ref + (size_t)b0 * a.ref_bs;
        if (a.out) c.out = a.out + (size_t)b0 * a.out_bs;
        if (a.aux) c.aux = a.aux + (size_t)b0 * a.aux_bs;
        dim3 grid, block;
        warp_rows_shape(c.B, c.H, c.W, c.rows, grid, block);
#if FLOWOPS_TUNE_WARP2D
        block = dim3(256, 1, 1);
        grid = dim3((c.W + 8 * kWarpCols - 1) / (8 * kWarpCols), (c.H + kRowStep * c.rows - 1) / (kRowStep * c.rows), c.B);
#endif
        if (a.C == 3) warp_rows_kernel<MODE, 3, EPI, WRITE_WARPED, UP4><<<grid, block, 0, st>>>(c);
        else if (a.C == 2) warp_rows_kernel<MODE, 2, EPI, WRITE_WARPED, UP4><<<grid, block, 0, st>>>(c);
        else warp_rows_kernel<MODE, 1, EPI, WRITE_WARPED, UP4><<<grid, block, 0, st>>>(c);
    }
}

}  // namespace flowops
