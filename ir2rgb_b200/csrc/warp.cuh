// warp.cuh -- device helpers shared by warp.cu and fused.cu (Resample2d coordinate and weight rules).
#pragma once
#include "common.cuh"

namespace flowops {

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

}  // namespace flowops
