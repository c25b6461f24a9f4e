// warp_fx_bwd.cuh -- backward of the flow warp with the image gradient accumulated per CTA in shared memory, in FIXED POINT.
//
// What bounds the image gradient of warp_rows_bwd.cuh is the L2 reduction unit: one sector operation per 32-byte sector a
// warp's RED touches, 14 x more of them than the gradient tensor has sectors on a smooth flow (a warp's 32 targets spread
// over ~10 image rows).  Accumulating a tile's contributions on chip needs atomics, and a float atomicAdd on shared memory
// is a compare-and-swap loop on sm_100 (ATOMS.CAST.SPIN); 32-bit INTEGER adds are native (ATOMS.ADD).  So:
//   * a CTA owns a 32 x 32 tile of source pixels; pass 1 loads its flows and gradients, finds the bounding box of the
//     bilinear targets and G = max |gO| of the tile;
//   * if the box fits a 56 x 56 window (any flow whose Jacobian stays below ~0.35 px/px -- real optical flow), every
//     contribution w * gO is scaled by the power of two 2^27 / G', G' >= G (exact; three bits of headroom because the
//     reference's truncation-based weights reach 4 at negative coordinates), rounded to an integer q (the one rounding:
//     <= 2^-28 of the tile's largest gradient) and added as q >> 15 and q & 0x7fff to two int32 cells with native
//     shared-memory atomics.  8192 contributions at most per tile: neither cell can overflow, and integer sums do not depend
//     on their order;
//   * the window is flushed with coalesced, zero-skipping red.global.add.f32 -- one sector operation per sector instead of 14;
//   * a tile whose box does not fit, or whose gradients are not finite (Inf / NaN must propagate as such), runs the
//     row-walking direct-reduction code of warp_rows_bwd.cuh unchanged (same function, same geometry).
// The flow gradient is the same gather as in warp_rows_bwd.cuh.
#pragma once
#include "warp_rows_bwd.cuh"

namespace flowops {
namespace fx {

constexpr int TX = 32, TY = 32, RPT = TY / 8;       // tile of source pixels; rows per thread (8 warps of 32 lanes)
constexpr int WX = 56, WY = 56;                     // window of target pixels
constexpr int CELLS = WX * WY;

template <int CT> constexpr int smem_bytes() { return 2 * CT * CELLS * (int)sizeof(int); }

template <int MODE, int CT, bool NEED_FLOW>
__global__ void __launch_bounds__(256, 2) warp_fx_bwd_kernel(const __grid_constant__ WarpBwdArgs a)
{
    extern __shared__ __align__(16) int win[];          // [hi | lo][CT][WY][WX]
    __shared__ unsigned s_red[3][8];
    __shared__ unsigned s_dec[4];                       // origin x, origin y, bits of G, fits?
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int xr = blockIdx.x * TX + lane;
    const bool valid_x = xr < a.W;
    const int x = valid_x ? xr : a.W - 1;
    const int y0 = blockIdx.y * TY + wrp * RPT;
    const int y1 = min(y0 + RPT, a.H);
    const unsigned hw = (unsigned)a.H * a.W, W = (unsigned)a.W;
    const size_t b = blockIdx.z;

    // ---- pass 1: flows and gradients of my pixels, bounding box of the targets, largest gradient ----
    BwdPix px[RPT];
    float g[RPT][CT];
    unsigned mnx = 0xffffffffu, mny = 0xffffffffu, mxx = 0u, mxy = 0u, gbits = 0u;
    bool bad_w = false;                                  // non-finite weights (NaN / Inf flow): must reach the output as such
    {
        const unsigned p0 = (unsigned)min(y0, a.H - 1) * W + (unsigned)x;
        const float *fl = a.flow + b * 2 * hw + p0;
        const float *go = a.gout + b * CT * hw + p0;
        const float xfl = small_int_as_float(x);
        const float lin_xv = MODE == FLOWOPS_WARP_GRIDSAMPLE ? __ldg(a.lin_x + x) : 0.f;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int y = y0 + k;
            const bool ok = valid_x && y < y1;
            float dx = 0.f, dy = 0.f;
            if (y < y1) { dx = ldg_stream(fl + k * W); dy = ldg_stream(fl + k * W + hw); }
#pragma unroll
            for (int c = 0; c < CT; ++c) g[k][c] = ok ? ldg_stream(go + k * W + (size_t)c * hw) : 0.f;
            bwd_pix_setup<MODE>(px[k], a, xfl, small_int_as_float(min(y, a.H - 1)), lin_xv, min(y, a.H - 1), dx, dy);
            if (ok) {
                mnx = min(mnx, px[k].xL); mny = min(mny, px[k].yT);
                mxx = max(mxx, px[k].xL + (px[k].ex ? 1u : 0u)); mxy = max(mxy, px[k].yT + (px[k].ey ? 1u : 0u));
#pragma unroll
                for (int c = 0; c < CT; ++c) gbits = max(gbits, __float_as_uint(fabsf(g[k][c])));     // NaN bits compare above Inf
                bad_w = bad_w || !(fabsf(px[k].w_tl) + fabsf(px[k].w_tr) + fabsf(px[k].w_bl) + fabsf(px[k].w_br) < 64.f);
            }
        }
    }
    // zero the window while the loads are in flight
    for (int i = tid; i < 2 * CT * CELLS / 4; i += 256) reinterpret_cast<int4 *>(win)[i] = make_int4(0, 0, 0, 0);
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    gbits = __reduce_max_sync(0xffffffffu, gbits);
    if (lane == 0) { s_red[0][wrp] = mnx; s_red[1][wrp] = mny; s_red[2][wrp] = gbits; }
    __syncthreads();
    if (tid < 32) {
        unsigned v0 = tid < 8 ? s_red[0][tid] : 0xffffffffu, v1 = tid < 8 ? s_red[1][tid] : 0xffffffffu, v2 = tid < 8 ? s_red[2][tid] : 0u;
        v0 = __reduce_min_sync(0xffffffffu, v0); v1 = __reduce_min_sync(0xffffffffu, v1); v2 = __reduce_max_sync(0xffffffffu, v2);
        if (tid == 0) { s_dec[0] = v0; s_dec[1] = v1; s_dec[2] = v2; }
    }
    __syncthreads();
    const unsigned ox = s_dec[0], oy = s_dec[1], G = s_dec[2];
    // does every target of the tile lie inside the window placed at (ox, oy)?  (all warps vote through shared memory)
    const bool mine_fits = !__any_sync(0xffffffffu, bad_w) && (mnx == 0xffffffffu || (mxx - ox < (unsigned)WX && mxy - oy < (unsigned)WY));
    if (lane == 0) s_red[0][wrp] = mine_fits ? 1u : 0u;
    __syncthreads();
    bool fits = true;
#pragma unroll
    for (int w = 0; w < 8; ++w) fits = fits && s_red[0][w] != 0u;
    const bool finite = G < 0x7f800000u && (G == 0u || G >= (30u << 23));      // Inf / NaN, or too small to scale: direct path

    if (!(fits && finite)) {
        // incoherent flow (or Inf / NaN gradients): the direct-reduction kernel's code on this tile
        if (y0 < a.H) warp_rows_bwd_body<MODE, CT, true, NEED_FLOW>(a, xr, y0, y1, lane, b);
        return;
    }

    // ---- pass 2: flow gradient (gather) and fixed-point accumulation of the image gradient ----
    // scale = 2^27 / G' with G' = 2^(e - 126) the power of two above G (e = G's exponent field >= 30): field 127 + 27 - (e - 126)
    const int se = 280 - (int)(G >> 23);
    const float scale = __int_as_float(se << 23);
    const float unit = __int_as_float((254 - se) << 23);           // 1 / scale, exact (a subnormal unit only when G ~ 1e38)
    const float *src = a.img + b * CT * hw;
    float *gf = NEED_FLOW ? a.gflow + b * 2 * hw + (unsigned)min(y0, a.H - 1) * W + (unsigned)x : nullptr;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const bool ok = valid_x && y0 + k < y1;
        const BwdPix &q = px[k];
        if (NEED_FLOW) {
            float gfx = 0.f, gfy = 0.f;
            const float *pt = src + q.o_t, *pb = src + q.o_b;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float tl = __ldg(pt), bl = __ldg(pb);
                const float trv = q.ex ? __ldg(pt + 1) : 0.f, brv = q.ex ? __ldg(pb + 1) : 0.f;
                const float tr = q.ex ? trv : tl, br = q.ex ? brv : bl;
                pt += hw; pb += hw;
                const float gc = g[k][c];
                if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
                    // resample2d_kernel.cu:159-184, same operation order (as warp_rows_bwd.cuh)
                    gfy = __fmaf_rn(q.gam_x * gc, bl, gfy);
                    gfy = __fmaf_rn(-(q.gam_x * gc), tl, gfy);
                    gfy = __fmaf_rn((1 - q.gam_x) * gc, br, gfy);
                    gfy = __fmaf_rn(-((1 - q.gam_x) * gc), tr, gfy);
                    gfx = __fmaf_rn(q.gam_y * gc, tr, gfx);
                    gfx = __fmaf_rn(-(q.gam_y * gc), tl, gfx);
                    gfx = __fmaf_rn((1 - q.gam_y) * gc, br, gfx);
                    gfx = __fmaf_rn(-((1 - q.gam_y) * gc), bl, gfx);
                } else {
                    const float ayg = q.ay * gc, axg = q.ax * gc, byg = q.by * gc, bxg = q.bx * gc;
                    gfx = __fmaf_rn(-tl, ayg, gfx); gfy = __fmaf_rn(-tl, axg, gfy);
                    gfx = __fmaf_rn(tr, ayg, gfx);  gfy = __fmaf_rn(-tr, bxg, gfy);
                    gfx = __fmaf_rn(-bl, byg, gfx); gfy = __fmaf_rn(bl, axg, gfy);
                    gfx = __fmaf_rn(br, byg, gfx);  gfy = __fmaf_rn(br, bxg, gfy);
                }
            }
            if (ok) {
                if (MODE == FLOWOPS_WARP_GRIDSAMPLE) { gfx *= q.gmx * a.mulx; gfy *= q.gmy * a.muly; }
                stg_stream(gf + k * W, gfx);
                stg_stream(gf + k * W + hw, gfy);
            }
        }
        if (ok && G != 0u) {
            const int cell = (int)(q.yT - oy) * WX + (int)(q.xL - ox);
            const int dxr = q.ex ? 1 : 0, dyr = q.ey ? WX : 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float gc = g[k][c];
                float v_tl = q.w_tl * gc, v_tr = q.w_tr * gc, v_bl = q.w_bl * gc, v_br = q.w_br * gc;
                if (!q.ex) { v_tl += v_tr; v_bl += v_br; v_tr = 0.f; v_br = 0.f; }      // same column twice
                if (!q.ey) { v_tl += v_bl; v_tr += v_br; v_bl = 0.f; v_br = 0.f; }      // same row twice
                int *hi = win + c * CELLS + cell, *lo = hi + CT * CELLS;
                const int q0 = __float2int_rn(v_tl * scale), q1 = __float2int_rn(v_tr * scale);
                const int q2 = __float2int_rn(v_bl * scale), q3 = __float2int_rn(v_br * scale);
                if (q0) { atomicAdd(hi, q0 >> 15); atomicAdd(lo, q0 & 0x7fff); }
                if (q1) { atomicAdd(hi + dxr, q1 >> 15); atomicAdd(lo + dxr, q1 & 0x7fff); }
                if (q2) { atomicAdd(hi + dyr, q2 >> 15); atomicAdd(lo + dyr, q2 & 0x7fff); }
                if (q3) { atomicAdd(hi + dyr + dxr, q3 >> 15); atomicAdd(lo + dyr + dxr, q3 & 0x7fff); }
            }
        }
    }
    __syncthreads();

    // ---- flush: cell -> float -> one reduction per non-zero cell, a warp per window row segment ----
    if (G == 0u) return;
    float *gi = a.gimg + b * CT * hw;
    for (int i = tid; i < CT * CELLS; i += 256) {
        const int c = i / CELLS, cell = i - c * CELLS;
        const int h = win[i], l = win[CT * CELLS + i];
        if ((h | l) != 0) {
            const int wy = cell / WX, wx = cell - wy * WX;
            const long long v = ((long long)h << 15) + (long long)l;
            red_add(gi + (size_t)c * hw + (oy + (unsigned)wy) * W + ox + (unsigned)wx, __fmul_rn(__ll2float_rn(v), unit));
        }
    }
}

}  // namespace fx

template <int MODE, bool NEED_FLOW>
static inline int launch_warp_fx_bwd(WarpBwdArgs a, cudaStream_t st)
{
    const int B = a.B;
    const size_t hw = (size_t)a.H * a.W, chw = (size_t)a.C * hw;
    for (int b0 = 0; b0 < B; b0 += 65535) {            // gridDim.z limit
        WarpBwdArgs c = a;
        c.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + b0 * chw; c.flow = a.flow + (size_t)b0 * 2 * hw; c.gout = a.gout + b0 * chw;
        c.gimg = a.gimg + b0 * chw;
        if (a.gflow) c.gflow = a.gflow + (size_t)b0 * 2 * hw;
        const dim3 grid((a.W + fx::TX - 1) / fx::TX, (a.H + fx::TY - 1) / fx::TY, c.B);
#define FLOWOPS_FX_LAUNCH(CT)                                                                                                         \
        {                                                                                                                             \
            const cudaError_t e = cudaFuncSetAttribute(fx::warp_fx_bwd_kernel<MODE, CT, NEED_FLOW>,                                   \
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, fx::smem_bytes<CT>());            \
            if (e != cudaSuccess) { set_error("warp_bwd: cannot reserve %d bytes of shared memory: %s", fx::smem_bytes<CT>(), cudaGetErrorString(e)); return (int)e; } \
            fx::warp_fx_bwd_kernel<MODE, CT, NEED_FLOW><<<grid, 256, fx::smem_bytes<CT>(), st>>>(c);                                  \
        }
        if (a.C == 3) FLOWOPS_FX_LAUNCH(3)
        else if (a.C == 2) FLOWOPS_FX_LAUNCH(2)
        else FLOWOPS_FX_LAUNCH(1)
#undef FLOWOPS_FX_LAUNCH
    }
    return 0;
}

}  // namespace flowops
