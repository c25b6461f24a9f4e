// warp.cu -- bilinear flow warp, forward and backward, for sm_100a.
//
// One kernel family serves both warps on the hot path:
//   mode RESAMPLE2D : FlowNet2's Resample2d (reference resample2d_package/resample2d_kernel.cu:16-190)
//   mode GRIDSAMPLE : vid2vid's `resample` (reference models/networks.py:15-28,89-100 ==
//                     models/base_model.py:123-136), i.e. get_grid + flow normalisation +
//                     F.grid_sample(bilinear, border, align_corners=False), without ever
//                     materialising the normalised grid.
//
// Design (HBM-bound op; algorithmic bytes per pixel: fwd 4*(2C+2), bwd 4*(3C+4) + gimg memset):
//   * one thread per pixel handles ALL channels: the flow is read once (coalesced, L1-bypassing),
//     the bilinear weights are computed once, the 4*C corner gathers go through L1 (neighbouring
//     lanes hit the same 128-byte lines), the C stores are coalesced and L1-bypassing;
//   * the reference runs one thread per output ELEMENT and re-reads the flow / recomputes the
//     weights C times;
//   * backward: the flow gradient is a gather (deterministic); the image gradient is a scatter whose
//     atomics are aggregated inside the warp before they are issued -- lane i's right-hand corners
//     usually coincide with lane i+1's left-hand corners, and clamped corners coincide within a
//     lane -- so a smooth flow costs ~2 RED.ADD per element instead of the reference's 4.
//
// Arithmetic: RESAMPLE2D forward reproduces the reference's mixed precision exactly (three weight
// products in fp64, the fourth in fp32, fp32 adds in TL,TR,BL,BR order) and is bit-identical to it.
// GRIDSAMPLE reproduces ATen's fp32 op chain (GridSampler.cuh: unnormalize -> clip -> floor ->
// weights as differences -> nw,ne,sw,se FFMA chain).
#include <stdlib.h>

#include "warp_win_bwd.cuh"
#include "warp_fx_bwd.cuh"

namespace flowops {

// process-wide switches of the warp kernels (flowops_warp_set_impl): bit 0 = image gradient by owned accumulation in
// per-warp shared-memory windows (warp_win_bwd.cuh; default OFF: measured on B200 it cuts the L2 reduction sector
// operations 4x but only wins on incoherent flows), bit 1 = forward blend with fp32 weights instead of the
// reference's accidental fp64 weight products (tolerance mode, see flowops.h; default off), bit 2 = fixed-point shared-memory
// image gradient (warp_fx_bwd.cuh; default off), bit 3 = GRIDSAMPLE forward on the row-walking kernel instead of
// warp_rows_mlp_kernel (A/B timing; default off)
static int g_warp_impl = -1;
int warp_impl_flags()
{
    if (g_warp_impl < 0) {
        const char *e = getenv("FLOWOPS_WARP_IMPL");
        g_warp_impl = e ? atoi(e) : 0;
    }
    return g_warp_impl;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int MODE, int CT>
__global__ void __launch_bounds__(256) warp_fwd_kernel(const float *__restrict__ img, const float *__restrict__ flow,
                                                       float *__restrict__ out, int B, int C, int H, int W,
                                                       const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                                       float invx, float invy)
{
    const int c_n = CT > 0 ? CT : C;
    // 2-D thread blocks over (x, y), batch in grid.z: no integer divisions, 32-bit in-plane offsets
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const unsigned hw = (unsigned)H * W;
    const unsigned p = (unsigned)y * W + x;
    for (int b = blockIdx.z; b < B; b += gridDim.z) {
        const float dx = ldg_stream(flow + ((size_t)b * 2) * hw + p);
        const float dy = ldg_stream(flow + ((size_t)b * 2 + 1) * hw + p);
        const float *src = img + (size_t)b * c_n * hw;
        float *dst = out + (size_t)b * c_n * hw + p;

        if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
            float xf, yf; Corners k;
            r2d_coords(x, y, dx, dy, H, W, xf, yf, k);
            const R2dWeights w = r2d_weights(xf, yf);
#pragma unroll
            for (int c = 0; c < c_n; ++c) {
                const float *pl = src + (size_t)c * hw;
                stg_stream(dst + (size_t)c * hw,
                           r2d_blend(w, __ldg(pl + k.o_tl), __ldg(pl + k.o_tr), __ldg(pl + k.o_bl), __ldg(pl + k.o_br)));
            }
        } else {
            const GsCoord g = gs_coords(x, y, dx, dy, H, W, lin_x, lin_y, invx, invy);
            const GsWeights w = gs_weights(g, H, W);
#pragma unroll
            for (int c = 0; c < c_n; ++c) stg_stream(dst + (size_t)c * hw, gs_gather(w, src + (size_t)c * hw, W));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------

// Warp-aggregated scatter of the four corner contributions of one channel.
// Lanes hold horizontally adjacent pixels.  `m` holds the per-pixel merge plan computed once
// (addresses are the same for every channel).
struct MergePlan {
    bool take_tr_from_left;   // lane-1's TR lands on my TL  -> I add it to my TL
    bool take_br_from_left;   // lane-1's BR lands on my BL
    bool give_tr_to_right;    // my TR is taken by lane+1    -> I do not issue it
    bool give_br_to_right;
    bool fold_x;              // xL == xR : TL/TR (and BL/BR) are the same address
    bool fold_y;              // yT == yB
};

__device__ __forceinline__ MergePlan make_plan(const Corners &k, bool valid, size_t plane_key)
{
    // keys compare plane (batch) and in-plane offset; invalid lanes never match
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long base = valid ? (long long)plane_key : -(long long)(lane + 2);
    const long long k_tl = valid ? base + k.o_tl : -1 - lane * 4;
    const long long k_tr = valid ? base + k.o_tr : -2 - lane * 4;
    const long long k_bl = valid ? base + k.o_bl : -3 - lane * 4;
    const long long k_br = valid ? base + k.o_br : -4 - lane * 4;
    const long long left_tr = __shfl_up_sync(full, k_tr, 1);
    const long long left_br = __shfl_up_sync(full, k_br, 1);
    const long long right_tl = __shfl_down_sync(full, k_tl, 1);
    const long long right_bl = __shfl_down_sync(full, k_bl, 1);
    MergePlan m;
    m.fold_x = (k.o_tl == k.o_tr);
    m.fold_y = (k.o_tl == k.o_bl);
    // only merge the regular pattern (both rows shift together); anything else goes out unmerged
    const bool l_ok = valid && lane > 0 && left_tr == k_tl && left_br == k_bl;
    const bool r_ok = valid && lane < 31 && right_tl == k_tr && right_bl == k_br;
    m.take_tr_from_left = l_ok; m.take_br_from_left = l_ok;
    m.give_tr_to_right = r_ok;  m.give_br_to_right = r_ok;
    return m;
}

__device__ __forceinline__ void scatter4(float *__restrict__ plane, const Corners &k, const MergePlan &m,
                                         bool valid, float v_tl, float v_tr, float v_bl, float v_br)
{
    const unsigned full = 0xffffffffu;
    if (m.fold_x) { v_tl += v_tr; v_bl += v_br; v_tr = 0.f; v_br = 0.f; }
    if (m.fold_y) { v_tl += v_bl; v_tr += v_br; v_bl = 0.f; v_br = 0.f; }
    const float in_tr = __shfl_up_sync(full, v_tr, 1);
    const float in_br = __shfl_up_sync(full, v_br, 1);
    if (m.take_tr_from_left) v_tl += in_tr;
    if (m.take_br_from_left) v_bl += in_br;
    if (!valid) return;
    red_add(plane + k.o_tl, v_tl);
    if (!m.fold_y) red_add(plane + k.o_bl, v_bl);
    if (!m.fold_x && !m.give_tr_to_right) {
        red_add(plane + k.o_tr, v_tr);
        if (!m.fold_y) red_add(plane + k.o_br, v_br);
    }
}

template <int MODE, int CT, bool NEED_IMG, bool NEED_FLOW>
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float *__restrict__ img, const float *__restrict__ flow,
                                                       const float *__restrict__ gout, float *__restrict__ gimg,
                                                       float *__restrict__ gflow, int B, int C, int H, int W,
                                                       const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                                       float invx, float invy, float mulx, float muly)
{
    const int c_n = CT > 0 ? CT : C;
    const size_t hw = (size_t)H * W;
    const size_t total = (size_t)B * hw;
    const size_t rounded = (total + 31) & ~(size_t)31;   // whole warps stay in the loop (shuffles)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rounded;
         i += (size_t)gridDim.x * blockDim.x) {
        const bool valid = i < total;
        const size_t ii = valid ? i : total - 1;
        const size_t b = ii / hw;
        const int p = (int)(ii - b * hw);
        const int y = p / W, x = p - y * W;
        const float dx = ldg_stream(flow + (b * 2) * hw + p);
        const float dy = ldg_stream(flow + (b * 2 + 1) * hw + p);
        const float *src = img + b * c_n * hw;
        const float *go = gout + b * c_n * hw + p;
        float *gi = NEED_IMG ? gimg + b * c_n * hw : nullptr;

        Corners k;
        float w_tl, w_tr, w_bl, w_br;      // image-gradient weights
        float gam_x = 0.f, gam_y = 0.f;    // RESAMPLE2D flow-gradient weights
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;   // GRIDSAMPLE: (ix_se-ix),(iy_se-iy),(ix-ix_nw),(iy-iy_nw)
        float gmx = 0.f, gmy = 0.f;
        if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
            float xf, yf;
            r2d_coords(x, y, dx, dy, H, W, xf, yf, k);
            // resample2d_kernel.cu:97-98: truncation, not floor, in the image-gradient kernel
            const float alpha = __fsub_rn(xf, (float)(int)xf);
            const float beta = __fsub_rn(yf, (float)(int)yf);
            w_tl = (1 - alpha) * (1 - beta); w_tr = alpha * (1 - beta);
            w_bl = (1 - alpha) * beta;       w_br = alpha * beta;
            gam_x = 1 - __fsub_rn(xf, floorf(xf));     // :160 (used for d/d dy)
            gam_y = 1 - __fsub_rn(yf, floorf(yf));     // :172 (used for d/d dx)
        } else {
            const GsCoord g = gs_coords(x, y, dx, dy, H, W, lin_x, lin_y, invx, invy);
            const int xe = min(g.ix_nw + 1, W - 1), ys = min(g.iy_nw + 1, H - 1);   // weight is 0 when clamped
            k.o_tl = g.iy_nw * W + g.ix_nw; k.o_tr = g.iy_nw * W + xe;
            k.o_bl = ys * W + g.ix_nw;      k.o_br = ys * W + xe;
            ax = (float)(g.ix_nw + 1) - g.ix; ay = (float)(g.iy_nw + 1) - g.iy;
            bx = g.ix - (float)g.ix_nw;       by = g.iy - (float)g.iy_nw;
            w_tl = ax * ay; w_tr = bx * ay; w_bl = ax * by; w_br = bx * by;
            gmx = g.gmx; gmy = g.gmy;
        }

        MergePlan m;
        if (NEED_IMG) m = make_plan(k, valid, b * (size_t)c_n * hw);

        float gfx = 0.f, gfy = 0.f;
#pragma unroll
        for (int c = 0; c < c_n; ++c) {
            const float g = valid ? ldg_stream(go + (size_t)c * hw) : 0.f;
            if (NEED_FLOW) {
                const float *pl = src + (size_t)c * hw;
                const float tl = __ldg(pl + k.o_tl), tr = __ldg(pl + k.o_tr);
                const float bl = __ldg(pl + k.o_bl), br = __ldg(pl + k.o_br);
                if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
                    // resample2d_kernel.cu:159-184, same operation order
                    gfy = __fmaf_rn(gam_x * g, bl, gfy);
                    gfy = __fmaf_rn(-(gam_x * g), tl, gfy);
                    gfy = __fmaf_rn((1 - gam_x) * g, br, gfy);
                    gfy = __fmaf_rn(-((1 - gam_x) * g), tr, gfy);
                    gfx = __fmaf_rn(gam_y * g, tr, gfx);
                    gfx = __fmaf_rn(-(gam_y * g), tl, gfx);
                    gfx = __fmaf_rn((1 - gam_y) * g, br, gfx);
                    gfx = __fmaf_rn(-((1 - gam_y) * g), bl, gfx);
                } else {
                    // ATen grid_sampler_2d_backward_kernel, bilinear branch
                    gfx -= tl * ay * g; gfy -= tl * ax * g;
                    gfx += tr * ay * g; gfy -= tr * bx * g;
                    gfx -= bl * by * g; gfy += bl * ax * g;
                    gfx += br * by * g; gfy += br * bx * g;
                }
            }
            if (NEED_IMG)
                scatter4(gi + (size_t)c * hw, k, m, valid, w_tl * g, w_tr * g, w_bl * g, w_br * g);
        }
        if (NEED_FLOW && valid) {
            if (MODE == FLOWOPS_WARP_GRIDSAMPLE) { gfx *= gmx * mulx; gfy *= gmy * muly; }
            stg_stream(gflow + (b * 2) * hw + p, gfx);
            stg_stream(gflow + (b * 2 + 1) * hw + p, gfy);
        }
    }
}

static inline unsigned warp_grid(size_t total)
{
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

// 256-thread blocks shaped to the image width: whole warps lie along x (coalescing, and the backward's
// lane-neighbour merging), rows fill the rest of the block
static inline void warp_launch_shape(int B, int H, int W, dim3 &grid, dim3 &block)
{
    const int bx = W >= 256 ? 256 : (W > 64 ? 128 : (W > 32 ? 64 : 32));
    block = dim3(bx, 256 / bx, 1);
    grid = dim3((W + bx - 1) / bx, (H + block.y - 1) / block.y, B < 65535 ? B : 65535);
}

template <int MODE>
static int launch_fwd(const float *img, const float *flow, float *out, int B, int C, int H, int W,
                      const float *lx, const float *ly, float invx, float invy, cudaStream_t st)
{
    dim3 grid, block;
    if (C <= 3) {
        // row-pipelined kernel (warp_rows.cuh)
        WarpArgs a{};
        a.img = img; a.img_bs = (size_t)C * H * W; a.flow = flow; a.out = out; a.out_bs = (size_t)C * H * W;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W);
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        a.lin_x = lx; a.lin_y = ly; a.invx = invx; a.invy = invy;
#ifdef FLOWOPS_TUNE_WARP_MLP
        // variant build (python -m ir2rgb_b200.build --out ... -DFLOWOPS_TUNE_WARP_MLP): flag bits 3..5 pick rows per thread /
        // rows of gathers in flight / resident CTAs of warp_rows_mlp_kernel in every mode (tools/warp_mlp_probe.py)
        const int variant = (warp_impl_flags() >> 3) & 7;
        if (variant && C == 3 && a.rows == 4) {
            switch (variant) {
            case 1: launch_warp_rows_mlp<MODE, 3, 4, 2, 4>(a, st); break;
            case 2: launch_warp_rows_mlp<MODE, 3, 4, 2, 3>(a, st); break;
            case 3: launch_warp_rows_mlp<MODE, 3, 4, 4, 2>(a, st); break;
            case 4: launch_warp_rows_mlp<MODE, 3, 8, 2, 4>(a, st); break;
            case 5: launch_warp_rows_mlp<MODE, 3, 8, 3, 3>(a, st); break;
            case 6: launch_warp_rows_mlp<MODE, 3, 4, 4, 3>(a, st); break;
            default: launch_warp_rows_mlp<MODE, 3, 4, 2, 5>(a, st); break;
            }
            return check_launch("warp_fwd");
        }
#else
        if constexpr (MODE == FLOWOPS_WARP_GRIDSAMPLE) if (a.rows == 4 && !(warp_impl_flags() & 8)) {
            // frames large enough for 4-row strips, grid_sample arithmetic: two rows of corner gathers in flight per thread
            // (4 rows per thread, 4 CTAs per SM).  Bit-identical to the row-walking kernel; measured at config 3
            // (profiles/warp_fwd_mlp_probe_r02.json): 76.5 -> 72.4 us on the smooth flow, 166.5 -> 148.1 us on the bilinear-
            // upsampled noise flow, 101.6 -> 94.0 us on per-pixel noise, equal on a zero flow.  The same variant LOSES in
            // both Resample2d modes (nearest 78.8 -> 88.6 us), which therefore keep the row-walking kernel; flag bit 3 of
            // flowops_warp_set_impl switches it off (A/B timing).
            if (C == 3) launch_warp_rows_mlp<MODE, 3, 4, 2, 4>(a, st);
            else if (C == 2) launch_warp_rows_mlp<MODE, 2, 4, 2, 4>(a, st);
            else launch_warp_rows_mlp<MODE, 1, 4, 2, 4>(a, st);
            return check_launch("warp_fwd");
        }
#endif
        launch_warp_rows<MODE, EPI_STORE, true>(a, st);
        return check_launch("warp_fwd");
    }
    warp_launch_shape(B, H, W, grid, block);      // any other channel count: channel loop at run time
    warp_fwd_kernel<MODE, 0><<<grid, block, 0, st>>>(img, flow, out, B, C, H, W, lx, ly, invx, invy);
    return check_launch("warp_fwd");
}

template <int MODE, bool NI, bool NF>
static void launch_bwd_c(const float *img, const float *flow, const float *gout, float *gimg, float *gflow,
                         int B, int C, int H, int W, const float *lx, const float *ly,
                         float invx, float invy, float mulx, float muly, cudaStream_t st)
{
    const unsigned grid = warp_grid((size_t)B * H * W);
    // only channel counts above 3 get here (run-time channel loop); 1..3 use warp_rows_bwd.cuh
    warp_bwd_kernel<MODE, 0, NI, NF><<<grid, 256, 0, st>>>(img, flow, gout, gimg, gflow, B, C, H, W, lx, ly, invx, invy, mulx, muly);
}

template <int MODE>
static int launch_bwd(const float *img, const float *flow, const float *gout, float *gimg, float *gflow,
                      int B, int C, int H, int W, const float *lx, const float *ly,
                      float invx, float invy, float mulx, float muly, cudaStream_t st)
{
    if (gimg) {
        cudaError_t e = cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)B * C * H * W, st);
        if (e != cudaSuccess) { set_error("warp_bwd: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if (gimg && C <= 3 && (W % 4) == 0 && aligned16(gimg) && (warp_impl_flags() & 1)) {
        // owned accumulation in per-warp shared-memory windows (warp_win_bwd.cuh)
        WarpBwdArgs a{};
        a.img = img; a.flow = flow; a.gout = gout; a.gimg = gimg; a.gflow = gflow;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = 32;
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        a.lin_x = lx; a.lin_y = ly; a.invx = invx; a.invy = invy; a.mulx = mulx; a.muly = muly;
        const int rc = gflow ? launch_warp_win_bwd<MODE, true>(a, st) : launch_warp_win_bwd<MODE, false>(a, st);
        if (rc) return rc;
        return check_launch("warp_bwd");
    }
    if (gimg && C <= 3 && (warp_impl_flags() & 4)) {
        // image gradient accumulated per tile in shared memory in fixed point; tiles whose targets do not fit the window run
        // the row-walking code below (warp_fx_bwd.cuh).  Optional (flowops_warp_set_impl bit 2), default OFF: exact and
        // parity-green, 20 % faster than the direct reductions on per-pixel-random flows, but 310 us flat where the direct
        // kernel needs 190 - 270 us on coherent flows (instruction-bound at 2 CTAs per SM).
        WarpBwdArgs a{};
        a.img = img; a.flow = flow; a.gout = gout; a.gimg = gimg; a.gflow = gflow;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = fx::RPT;
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        a.lin_x = lx; a.lin_y = ly; a.invx = invx; a.invy = invy; a.mulx = mulx; a.muly = muly;
        const int rc = gflow ? launch_warp_fx_bwd<MODE, true>(a, st) : launch_warp_fx_bwd<MODE, false>(a, st);
        if (rc) return rc;
        return check_launch("warp_bwd");
    }
    if (C <= 3) {
        // row-walking kernel (warp_rows_bwd.cuh)
        WarpBwdArgs a{};
        a.img = img; a.flow = flow; a.gout = gout; a.gimg = gimg; a.gflow = gflow;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W, 8);
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        a.lin_x = lx; a.lin_y = ly; a.invx = invx; a.invy = invy; a.mulx = mulx; a.muly = muly;
        if (gimg && gflow) launch_warp_rows_bwd<MODE, true, true>(a, st);
        else if (gimg) launch_warp_rows_bwd<MODE, true, false>(a, st);
        else launch_warp_rows_bwd<MODE, false, true>(a, st);
        return check_launch("warp_bwd");
    }
    if (gimg && gflow) launch_bwd_c<MODE, true, true>(img, flow, gout, gimg, gflow, B, C, H, W, lx, ly, invx, invy, mulx, muly, st);
    else if (gimg) launch_bwd_c<MODE, true, false>(img, flow, gout, gimg, gflow, B, C, H, W, lx, ly, invx, invy, mulx, muly, st);
    else launch_bwd_c<MODE, false, true>(img, flow, gout, gimg, gflow, B, C, H, W, lx, ly, invx, invy, mulx, muly, st);
    return check_launch("warp_bwd");
}

}  // namespace flowops

using namespace flowops;

static int warp_check(const char *who, const void *img, const void *flow, int B, int C, int H, int W, int mode,
                      const float *lin_x, const float *lin_y)
{
    FLOWOPS_REQUIRE(img && flow, FLOWOPS_EINVAL, "%s: null pointer", who);
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "%s: bad shape %dx%dx%dx%d", who, B, C, H, W);
    FLOWOPS_REQUIRE((size_t)C * H * W < (1ull << 31), FLOWOPS_EUNSUPPORTED, "%s: C*H*W exceeds int32 indexing", who);
    FLOWOPS_REQUIRE(H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED, "%s: H, W above 2^22 are not supported", who);
    FLOWOPS_REQUIRE(mode == FLOWOPS_WARP_RESAMPLE2D || mode == FLOWOPS_WARP_GRIDSAMPLE, FLOWOPS_EINVAL, "%s: unknown mode %d", who, mode);
    if (mode == FLOWOPS_WARP_GRIDSAMPLE) {
        FLOWOPS_REQUIRE(lin_x && lin_y, FLOWOPS_EINVAL, "%s: GRIDSAMPLE mode needs the linspace tables", who);
        FLOWOPS_REQUIRE(H > 1 && W > 1, FLOWOPS_EUNSUPPORTED, "%s: GRIDSAMPLE mode needs H, W > 1", who);
    }
    return 0;
}

extern "C" int flowops_warp_fwd(const float *img, const float *flow, float *out, int B, int C, int H, int W,
                                int mode, const float *lin_x, const float *lin_y, void *stream)
{
    int rc = warp_check("warp_fwd", img, flow, B, C, H, W, mode, lin_x, lin_y);
    if (rc) return rc;
    FLOWOPS_REQUIRE(out, FLOWOPS_EINVAL, "warp_fwd: null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == FLOWOPS_WARP_RESAMPLE2D) {
        if (C <= 3 && (warp_impl_flags() & 2))         // tolerance mode: fp32 weights (row-walking kernel only)
            return launch_fwd<kWarpResample2dF32>(img, flow, out, B, C, H, W, nullptr, nullptr, 0.f, 0.f, st);
        return launch_fwd<FLOWOPS_WARP_RESAMPLE2D>(img, flow, out, B, C, H, W, nullptr, nullptr, 0.f, 0.f, st);
    }
    float invx, invy, mulx, muly;
    gs_scales(H, W, invx, invy, mulx, muly);
    return launch_fwd<FLOWOPS_WARP_GRIDSAMPLE>(img, flow, out, B, C, H, W, lin_x, lin_y, invx, invy, st);
}

extern "C" int flowops_warp_bwd(const float *img, const float *flow, const float *gout, float *gimg, float *gflow,
                                int B, int C, int H, int W, int mode,
                                const float *lin_x, const float *lin_y, void *stream)
{
    int rc = warp_check("warp_bwd", img, flow, B, C, H, W, mode, lin_x, lin_y);
    if (rc) return rc;
    FLOWOPS_REQUIRE(gout, FLOWOPS_EINVAL, "warp_bwd: null grad_output");
    FLOWOPS_REQUIRE(gimg || gflow, FLOWOPS_EINVAL, "warp_bwd: both gradient outputs are null");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == FLOWOPS_WARP_RESAMPLE2D)
        return launch_bwd<FLOWOPS_WARP_RESAMPLE2D>(img, flow, gout, gimg, gflow, B, C, H, W, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, st);
    float invx, invy, mulx, muly;
    gs_scales(H, W, invx, invy, mulx, muly);
    return launch_bwd<FLOWOPS_WARP_GRIDSAMPLE>(img, flow, gout, gimg, gflow, B, C, H, W, lin_x, lin_y, invx, invy, mulx, muly, st);
}

extern "C" int flowops_warp_set_impl(int flags)
{
    flowops::g_warp_impl = flags;
    return 0;
}

extern "C" int flowops_warp_get_impl(void) { return flowops::warp_impl_flags(); }
