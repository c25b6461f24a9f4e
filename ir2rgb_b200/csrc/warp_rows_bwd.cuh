// warp_rows_bwd.cuh -- row-walking backward of the flow warp (both coordinate conventions).
//
// What bounds the image gradient (tools/red_probe.cu, B200): RED.ADD.F32 costs one L2 sector operation per
// 32-byte sector a warp instruction touches (~180 G sector-ops/s chip-wide), not one per lane -- eight lanes
// landing in one sector cost the same as one lane.  So the kernel is organised to issue as few, as dense,
// reductions as possible:
//   * lanes are horizontally adjacent pixels: lane i's right-hand corners usually are lane i+1's left-hand
//     corners and are handed over with a shuffle (as before);
//   * a thread walks `rows` rows of its column and keeps the bottom-row contribution of row y in registers:
//     when row y+1's top row lands on the same pixel (smooth flow: half to all of the time) the two are
//     added in registers and leave as ONE reduction -- 1 to 1.5 RED per element instead of 2;
//   * contributions that are exactly zero are not issued (the accumulator starts at +0 and can never become
//     -0, so x + 0 == x bit for bit; NaN is not zero and still propagates): integer-valued flows -- static
//     background -- issue a single RED per element;
//   * the flow gradient is the forward's gather, with the same prefetch of the next row's flow and the same
//     32-bit corner addressing (warp_rows.cuh).
#pragma once
#include <stdlib.h>

#include "warp_rows.cuh"

namespace flowops {

struct WarpBwdArgs {
    const float *img, *flow, *gout;      // [B,C,H,W], [B,2,H,W], [B,C,H,W] contiguous
    float *gimg, *gflow;                 // nullable by template flag
    int B, C, H, W, rows;
    float wm1, hm1;
    const float *lin_x, *lin_y;
    float invx, invy, mulx, muly;
};

__device__ __forceinline__ void red_add_nz(float *p, float v)
{
    if (!(v == 0.f)) red_add(p, v);
}

// Per-pixel state of the backward: corner offsets (top-left and the row below; the right-hand column is +1 when `ex`),
// image-gradient weights, and what the flow gradient of either coordinate convention needs.
struct BwdPix {
    unsigned o_t, o_b, xL, yT;
    bool ex, ey;
    float w_tl, w_tr, w_bl, w_br;                    // image-gradient weights
    float gam_x, gam_y;                              // RESAMPLE2D flow-gradient weights
    float ax, ay, bx, by;                            // GRIDSAMPLE: (ix_se-ix),(iy_se-iy),(ix-ix_nw),(iy-iy_nw)
    float gmx, gmy;
};

template <int MODE>
__device__ __forceinline__ void bwd_pix_setup(BwdPix &q, const WarpBwdArgs &a, float xfl, float yfl, float lin_xv, int y, float dx, float dy)
{
    const unsigned W = (unsigned)a.W;
    q.gam_x = q.gam_y = q.ax = q.ay = q.bx = q.by = q.gmx = q.gmy = 0.f;
    if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
        const float xf = __fadd_rn(xfl, dx), yf = __fadd_rn(yfl, dy);
        float fx = __fsub_rn(__fadd_rn(xf, kMagic15), kMagic15); fx = fx > xf ? __fsub_rn(fx, 1.f) : fx;
        float fy = __fsub_rn(__fadd_rn(yf, kMagic15), kMagic15); fy = fy > yf ? __fsub_rn(fy, 1.f) : fy;
        float tx = (xf < 0.f && fx != xf) ? __fadd_rn(fx, 1.f) : fx;      // (float)(int)xf: truncation
        float ty = (yf < 0.f && fy != yf) ? __fadd_rn(fy, 1.f) : fy;
        if (__builtin_expect(!(fmaxf(fabsf(xf), fabsf(yf)) < 4194304.f), 0)) {
            fx = floorf(xf); fy = floorf(yf); tx = (float)(int)xf; ty = (float)(int)yf;
        }
        const unsigned xL = q.xL = (unsigned)small_float_as_int(fminf(fmaxf(fx, 0.f), a.wm1));
        const unsigned yT = q.yT = (unsigned)small_float_as_int(fminf(fmaxf(fy, 0.f), a.hm1));
        q.ex = fx >= 0.f && fx < a.wm1;
        q.ey = fy >= 0.f && fy < a.hm1;
        q.o_t = yT * W + xL;
        // resample2d_kernel.cu:97-98: truncation, not floor, in the image-gradient kernel
        const float alpha = __fsub_rn(xf, tx), beta = __fsub_rn(yf, ty);
        q.w_tl = (1 - alpha) * (1 - beta); q.w_tr = alpha * (1 - beta);
        q.w_bl = (1 - alpha) * beta;       q.w_br = alpha * beta;
        q.gam_x = 1 - __fsub_rn(xf, fx);     // :160 (used for d/d dy)
        q.gam_y = 1 - __fsub_rn(yf, fy);     // :172 (used for d/d dx)
    } else {
        const float gx = __fadd_rn(lin_xv, __fmul_rn(dx, a.invx));
        const float gy = __fadd_rn(__ldg(a.lin_y + y), __fmul_rn(dy, a.invy));
        float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), a.wm1 + 1.f, -1.f), 0.5f);
        float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), a.hm1 + 1.f, -1.f), 0.5f);
        // clip_coordinates_set_grad: gradient is zero at and beyond both borders
        q.gmx = (ix <= 0.f || ix >= a.wm1) ? 0.f : 1.f;
        q.gmy = (iy <= 0.f || iy >= a.hm1) ? 0.f : 1.f;
        ix = fminf(a.wm1, fmaxf(ix, 0.f));
        iy = fminf(a.hm1, fmaxf(iy, 0.f));
        float fx = __fsub_rn(__fadd_rn(ix, kMagic15), kMagic15); fx = fx > ix ? __fsub_rn(fx, 1.f) : fx;
        float fy = __fsub_rn(__fadd_rn(iy, kMagic15), kMagic15); fy = fy > iy ? __fsub_rn(fy, 1.f) : fy;
        q.ex = __fadd_rn(fx, 1.f) <= a.wm1;           // weight of a clamped neighbour is 0
        q.ey = __fadd_rn(fy, 1.f) <= a.hm1;
        q.xL = (unsigned)small_float_as_int(fx); q.yT = (unsigned)small_float_as_int(fy);
        q.o_t = q.yT * W + q.xL;
        q.ax = __fadd_rn(fx, 1.f) - ix; q.ay = __fadd_rn(fy, 1.f) - iy;
        q.bx = ix - fx;                 q.by = iy - fy;
        q.w_tl = q.ax * q.ay; q.w_tr = q.bx * q.ay; q.w_bl = q.ax * q.by; q.w_br = q.bx * q.by;
    }
    q.o_b = q.ey ? q.o_t + W : q.o_t;
}

// rows y0 .. y1-1 of column xr of batch item b, for one whole warp (lane = position inside a 32-pixel row segment)
template <int MODE, int CT, bool NEED_IMG, bool NEED_FLOW>
__device__ __forceinline__ void warp_rows_bwd_body(const WarpBwdArgs &a, int xr, int y0, int y1, int lane, size_t b)
{
    const unsigned full = 0xffffffffu;
    const bool valid_x = xr < a.W;                    // out-of-image lanes stay alive for the shuffles
    const int x = valid_x ? xr : a.W - 1;
    const unsigned hw = (unsigned)a.H * a.W, W = (unsigned)a.W;
    const float *src = a.img + b * CT * hw;
    asm("" : "+l"(src));
    float *gi = NEED_IMG ? a.gimg + b * CT * hw : nullptr;
    asm("" : "+l"(gi));
    const unsigned p0 = (unsigned)y0 * W + (unsigned)x;
    const float *fl = a.flow + b * 2 * hw + p0;
    const float *go = a.gout + b * CT * hw + p0;
    float *gf = NEED_FLOW ? a.gflow + b * 2 * hw + p0 : nullptr;
    const float xfl = small_int_as_float(x);
    float yfl = small_int_as_float(y0);
    const float lin_xv = MODE == FLOWOPS_WARP_GRIDSAMPLE ? __ldg(a.lin_x + x) : 0.f;

    float pend[CT];                   // bottom-row contribution of the previous row, waiting for a partner
    unsigned pend_off = 0;
    bool pend_valid = false;
#pragma unroll
    for (int c = 0; c < CT; ++c) pend[c] = 0.f;

    float dx = ldg_stream(fl), dy = ldg_stream(fl + hw);
    for (int y = y0; y < y1; ++y) {
        float ndx = 0.f, ndy = 0.f;
        if (y + 1 < y1) { ndx = ldg_stream(fl + W); ndy = ldg_stream(fl + W + hw); }
        float g[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) g[c] = valid_x ? ldg_stream(go + (size_t)c * hw) : 0.f;

        BwdPix q;
        bwd_pix_setup<MODE>(q, a, xfl, yfl, lin_xv, y, dx, dy);
        const unsigned o_t = q.o_t, o_b = q.o_b;
        const bool ex = q.ex, ey = q.ey;
        const float w_tl = q.w_tl, w_tr = q.w_tr, w_bl = q.w_bl, w_br = q.w_br;
        const float gam_x = q.gam_x, gam_y = q.gam_y, ax = q.ax, ay = q.ay, bx = q.bx, by = q.by, gmx = q.gmx, gmy = q.gmy;

        float gfx = 0.f, gfy = 0.f;
        if (NEED_FLOW) {
            const float *pt = src + o_t, *pb = src + o_b;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float tl = __ldg(pt), bl = __ldg(pb);
                const float trv = ex ? __ldg(pt + 1) : 0.f, brv = ex ? __ldg(pb + 1) : 0.f;
                const float tr = ex ? trv : tl, br = ex ? brv : bl;
                pt += hw; pb += hw;
                if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
                    // resample2d_kernel.cu:159-184, same operation order
                    gfy = __fmaf_rn(gam_x * g[c], bl, gfy);
                    gfy = __fmaf_rn(-(gam_x * g[c]), tl, gfy);
                    gfy = __fmaf_rn((1 - gam_x) * g[c], br, gfy);
                    gfy = __fmaf_rn(-((1 - gam_x) * g[c]), tr, gfy);
                    gfx = __fmaf_rn(gam_y * g[c], tr, gfx);
                    gfx = __fmaf_rn(-(gam_y * g[c]), tl, gfx);
                    gfx = __fmaf_rn((1 - gam_y) * g[c], br, gfx);
                    gfx = __fmaf_rn(-((1 - gam_y) * g[c]), bl, gfx);
                } else {
                    // ATen grid_sampler_2d_backward_kernel, bilinear branch: gix -= nw * (iy_se - iy) * gOut, ... -- the same
                    // eight terms, with the per-pixel products hoisted and each term one FMA (rounding differs from ATen's
                    // mul-mul-sub by < 1 ulp per term; the backward is held to 1e-4, not to bit equality)
                    const float ayg = ay * g[c], axg = ax * g[c], byg = by * g[c], bxg = bx * g[c];
                    gfx = __fmaf_rn(-tl, ayg, gfx); gfy = __fmaf_rn(-tl, axg, gfy);
                    gfx = __fmaf_rn(tr, ayg, gfx);  gfy = __fmaf_rn(-tr, bxg, gfy);
                    gfx = __fmaf_rn(-bl, byg, gfx); gfy = __fmaf_rn(bl, axg, gfy);
                    gfx = __fmaf_rn(br, byg, gfx);  gfy = __fmaf_rn(br, bxg, gfy);
                }
            }
            if (valid_x) {
                if (MODE == FLOWOPS_WARP_GRIDSAMPLE) { gfx *= gmx * a.mulx; gfy *= gmy * a.muly; }
                stg_stream(gf, gfx);
                stg_stream(gf + hw, gfy);
            }
        }

        if (NEED_IMG) {
            // ---- horizontal hand-over plan (addresses are the same for every channel) ----
            const unsigned a_tl = o_t, a_tr = o_t + (ex ? 1u : 0u), a_bl = o_b, a_br = o_b + (ex ? 1u : 0u);
            // invalid lanes carry keys that match nothing
            const unsigned k_tl = valid_x ? a_tl : 0xffffffffu - 4u * lane, k_tr = valid_x ? a_tr : 0xfffffffeu - 4u * lane;
            const unsigned k_bl = valid_x ? a_bl : 0xfffffffdu - 4u * lane, k_br = valid_x ? a_br : 0xfffffffcu - 4u * lane;
            const unsigned left_tr = __shfl_up_sync(full, k_tr, 1), left_br = __shfl_up_sync(full, k_br, 1);
            const unsigned right_tl = __shfl_down_sync(full, k_tl, 1), right_bl = __shfl_down_sync(full, k_bl, 1);
            // only the regular pattern is merged (both rows shift together); anything else goes out unmerged
            const bool take = valid_x && lane > 0 && left_tr == k_tl && left_br == k_bl;
            const bool give = valid_x && lane < 31 && right_tl == k_tr && right_bl == k_br;
            const bool same = pend_valid && pend_off == o_t;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                float v_tl = w_tl * g[c], v_tr = w_tr * g[c], v_bl = w_bl * g[c], v_br = w_br * g[c];
                if (!ex) { v_tl += v_tr; v_bl += v_br; v_tr = 0.f; v_br = 0.f; }      // same column twice
                if (!ey) { v_tl += v_bl; v_tr += v_br; v_bl = 0.f; v_br = 0.f; }      // same row twice
                const float in_tr = __shfl_up_sync(full, v_tr, 1), in_br = __shfl_up_sync(full, v_br, 1);
                if (take) { v_tl += in_tr; v_bl += in_br; }
                if (valid_x) {
                    float *pl = gi + (size_t)c * hw;
                    if (ex && !give) {
                        red_add_nz(pl + a_tr, v_tr);
                        if (ey) red_add_nz(pl + a_br, v_br);
                    }
                    // vertical pairing: last row's bottom contribution and this row's top contribution
                    if (same) v_tl += pend[c];
                    else if (pend_valid) red_add_nz(pl + pend_off, pend[c]);
                    red_add_nz(pl + a_tl, v_tl);
                    pend[c] = v_bl;
                }
            }
            pend_off = o_b;
            pend_valid = ey && valid_x;
        }
        dx = ndx; dy = ndy;
        fl += W; go += W; gf += W; yfl = __fadd_rn(yfl, 1.f);
    }
    if (NEED_IMG && pend_valid) {
#pragma unroll
        for (int c = 0; c < CT; ++c) red_add_nz(gi + (size_t)c * hw + pend_off, pend[c]);
    }
}

// Tuning macro (variant builds, tools/ab_ops.py): resident CTAs per SM the register allocation of the backward aims at
#ifndef FLOWOPS_TUNE_BWD_MINB
#define FLOWOPS_TUNE_BWD_MINB 3
#endif

template <int MODE, int CT, bool NEED_IMG, bool NEED_FLOW>
__global__ void __launch_bounds__(256, FLOWOPS_TUNE_BWD_MINB) warp_rows_bwd_kernel(const __grid_constant__ WarpBwdArgs a)
{
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * a.rows;
    if (y0 >= a.H) return;                            // uniform per warp (a warp never spans two rows)
    warp_rows_bwd_body<MODE, CT, NEED_IMG, NEED_FLOW>(a, blockIdx.x * blockDim.x + threadIdx.x, y0, min(y0 + a.rows, a.H),
                                                      threadIdx.x & 31, blockIdx.z);
}

template <int MODE, bool NEED_IMG, bool NEED_FLOW>
static inline void launch_warp_rows_bwd(WarpBwdArgs a, cudaStream_t st)
{
    const int B = a.B;
    const size_t hw = (size_t)a.H * a.W, chw = (size_t)a.C * hw;
    for (int b0 = 0; b0 < B; b0 += 65535) {            // gridDim.z limit
        WarpBwdArgs c = a;
        c.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + b0 * chw; c.flow = a.flow + (size_t)b0 * 2 * hw; c.gout = a.gout + b0 * chw;
        if (a.gimg) c.gimg = a.gimg + b0 * chw;
        if (a.gflow) c.gflow = a.gflow + (size_t)b0 * 2 * hw;
        dim3 grid, block;
        warp_rows_shape(c.B, c.H, c.W, c.rows, grid, block);
        if (const char *e = getenv("FLOWOPS_TUNE_BWD_GEOM")) {          // "bx,by,rows": A/B timing of the CTA footprint
            int bx = 0, by = 0, rows = 0;
            if (sscanf(e, "%d,%d,%d", &bx, &by, &rows) == 3 && bx >= 32 && bx % 32 == 0 && by >= 1 && bx * by <= 256 && rows >= 1) {
                c.rows = rows;
                block = dim3(bx, by, 1);
                grid = dim3((c.W + bx - 1) / bx, (c.H + by * rows - 1) / (by * rows), c.B);
            }
        }
        if (a.C == 3) warp_rows_bwd_kernel<MODE, 3, NEED_IMG, NEED_FLOW><<<grid, block, 0, st>>>(c);
        else if (a.C == 2) warp_rows_bwd_kernel<MODE, 2, NEED_IMG, NEED_FLOW><<<grid, block, 0, st>>>(c);
        else warp_rows_bwd_kernel<MODE, 1, NEED_IMG, NEED_FLOW><<<grid, block, 0, st>>>(c);
    }
}

}  // namespace flowops
