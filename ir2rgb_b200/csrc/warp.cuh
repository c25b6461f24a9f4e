// warp.cuh -- device helpers shared by warp.cu and fused.cu (Resample2d coordinate and weight rules).
#pragma once
#include "common.cuh"

namespace flowops {

// process-wide switches of the warp kernels (flowops_warp_set_impl; defined in warp.cu)
int warp_impl_flags();

struct Corners {
    int o_tl, o_tr, o_bl, o_br;   // offsets inside one H*W plane
};

// ---------------------------------------------------------------------------------------------
// coordinate conventions
// ---------------------------------------------------------------------------------------------

// resample2d_kernel.cu:40-51
__device__ __forceinline__ void r2d_coords(int x, int y, float dx, float dy, int H, int W,
                                           float &xf, float &yf, Corners &k)
{
    xf = __fadd_rn((float)x, dx);
    yf = __fadd_rn((float)y, dy);
    const float fx = floorf(xf), fy = floorf(yf);
    const int xL = max(min((int)fx, W - 1), 0);
    const int xR = max(min((int)(__fadd_rn(fx, 1.f)), W - 1), 0);
    const int yT = max(min((int)fy, H - 1), 0);
    const int yB = max(min((int)(__fadd_rn(fy, 1.f)), H - 1), 0);
    k.o_tl = yT * W + xL; k.o_tr = yT * W + xR; k.o_bl = yB * W + xL; k.o_br = yB * W + xR;
}


// resample2d_kernel.cu:45-46,55-58: bilinear weights with the reference's mixed precision -- the first
// three products are fp64 (literal `1.`), the fourth is fp32.
struct R2dWeights {
    double w_tl, w_tr, w_bl;
    float w_br;
};
__device__ __forceinline__ R2dWeights r2d_weights(float xf, float yf)
{
    const float alpha = __fsub_rn(xf, floorf(xf));
    const float beta = __fsub_rn(yf, floorf(yf));
    const double wa = 1. - (double)alpha, wb = 1. - (double)beta;
    R2dWeights w;
    w.w_tl = wa * wb; w.w_tr = (double)alpha * wb; w.w_bl = wa * (double)beta;
    w.w_br = __fmul_rn(alpha, beta);
    return w;
}
// fp32 adds in TL, TR, BL, BR order (val starts at 0.0f as in the reference)
__device__ __forceinline__ float r2d_blend(const R2dWeights &w, float tl, float tr, float bl, float br)
{
    float val = 0.0f;
    val = __fadd_rn(val, (float)(w.w_tl * (double)tl));
    val = __fadd_rn(val, (float)(w.w_tr * (double)tr));
    val = __fadd_rn(val, (float)(w.w_bl * (double)bl));
    return __fmaf_rn(w.w_br, br, val);
}

// models/networks.py:97-98 + ATen grid_sampler_compute_source_index (align_corners=False, border)
struct GsCoord {
    float ix, iy;        // clipped source coordinates
    int ix_nw, iy_nw;    // floor
    float gmx, gmy;      // d(ix)/d(flow_x), d(iy)/d(flow_y) incl. the clip mask (backward only)
};
__device__ __forceinline__ GsCoord gs_coords(int x, int y, float dx, float dy, int H, int W,
                                             const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                             float invx, float invy)
{
    GsCoord g;
    const float gx = __fadd_rn(__ldg(lin_x + x), __fmul_rn(dx, invx));
    const float gy = __fadd_rn(__ldg(lin_y + y), __fmul_rn(dy, invy));
    // ((coord + 1) * size - 1) / 2 ; nvcc contracts the multiply-subtract in ATen's build
    float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), (float)W, -1.f), 0.5f);
    float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), (float)H, -1.f), 0.5f);
    // clip_coordinates_set_grad: gradient is zero at and beyond both borders
    g.gmx = (ix <= 0.f || ix >= (float)(W - 1)) ? 0.f : 1.f;
    g.gmy = (iy <= 0.f || iy >= (float)(H - 1)) ? 0.f : 1.f;
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    g.ix = ix; g.iy = iy;
    g.ix_nw = (int)floorf(ix);
    g.iy_nw = (int)floorf(iy);
    return g;
}

// bilinear weights / corner validity of ATen's grid_sampler_2d_kernel (nw, ne, sw, se as differences)
struct GsWeights {
    float nw, ne, sw, se;
    int o_nw;
    bool in_e, in_s;
};
__device__ __forceinline__ GsWeights gs_weights(const GsCoord &g, int H, int W)
{
    GsWeights w;
    const int ix_se = g.ix_nw + 1, iy_se = g.iy_nw + 1;
    w.nw = __fmul_rn(__fsub_rn((float)ix_se, g.ix), __fsub_rn((float)iy_se, g.iy));
    w.ne = __fmul_rn(__fsub_rn(g.ix, (float)g.ix_nw), __fsub_rn((float)iy_se, g.iy));
    w.sw = __fmul_rn(__fsub_rn((float)ix_se, g.ix), __fsub_rn(g.iy, (float)g.iy_nw));
    w.se = __fmul_rn(__fsub_rn(g.ix, (float)g.ix_nw), __fsub_rn(g.iy, (float)g.iy_nw));
    // after the border clip (ix_nw, iy_nw) is always inside; the +1 neighbours may be one past the edge
    w.in_e = ix_se < W; w.in_s = iy_se < H;
    w.o_nw = g.iy_nw * W + g.ix_nw;
    return w;
}
// out_acc += val * weight in nw, ne, sw, se order (FFMA chain), skipping out-of-bounds corners
__device__ __forceinline__ float gs_gather(const GsWeights &w, const float *__restrict__ plane, int W)
{
    const float *pl = plane + w.o_nw;
    float acc = __fmaf_rn(__ldg(pl), w.nw, 0.f);
    if (w.in_e) acc = __fmaf_rn(__ldg(pl + 1), w.ne, acc);
    if (w.in_s) acc = __fmaf_rn(__ldg(pl + W), w.sw, acc);
    if (w.in_e && w.in_s) acc = __fmaf_rn(__ldg(pl + W + 1), w.se, acc);
    return acc;
}

// fp32 constants of the models/networks.py:97 normalisation: flow / ((W-1)/2) runs on CUDA as a
// multiply by the fp32 reciprocal of the fp32 scalar.
static inline void gs_scales(int H, int W, float &invx, float &invy, float &mulx, float &muly)
{
    const float sx = (float)((W - 1.0) / 2.0), sy = (float)((H - 1.0) / 2.0);
    invx = 1.0f / sx; invy = 1.0f / sy;
    // d(ix)/d(flow_x) = (W/2) * invx  (unnormalize gradient times the reciprocal above)
    mulx = ((float)W * 0.5f) * invx;
    muly = ((float)H * 0.5f) * invy;
}

}  // namespace flowops
