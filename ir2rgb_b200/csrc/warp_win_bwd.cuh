// warp_win_bwd.cuh -- backward of the flow warp with OWNED accumulation of the image gradient.
//
// What bounded the previous kernel (warp_rows_bwd.cuh; profiles/ncu_warp_bwd_rows_smooth_r01.txt): the image gradient is
// a scatter, and RED.ADD.F32 costs one L2 sector operation per 32-byte sector a warp instruction touches.  A warp's 32
// targets spread over ~10 image rows, so 3.1 M reduction requests touched 43 M sectors -- 14x the 3.1 M sectors the
// gradient tensor has -- and the kernel ran at the L2 reduction rate (0.25 of HBM bandwidth).
//
// Here every warp owns a private window of the gradient image in shared memory and is the only writer of it:
//   * a warp walks a segment of S rows of a 32-column strip; its window (WH = 16 rows x WW = 64 columns x C channels)
//     is anchored around the target of the warp's middle lane and ROLLS with that lane's target row (circular in y);
//   * contributions are added with plain ld.shared / add / st.shared -- no atomics: within one warp instruction the
//     targets are made distinct first (horizontal hand-over of the right-hand corners by shuffle as before).  When the
//     lanes' target columns increase strictly -- every flow that does not fold over inside the warp -- each lane owns a
//     column and two phases suffice; otherwise coinciding corners are found with __match_any_sync (slow: it iterates
//     over the distinct keys, which is why it is kept off the common path), summed by shuffles, and written in four phases;
//   * the window row that rolls out is flushed with dense `red.global.add.v4.f32` (one sector operation per sector,
//     all-zero quads skipped) and zeroed; targets outside the window fall back to scalar RED, so any flow is handled
//     correctly and only coherent flows are fast.
// fp32 `atomicAdd` on shared memory is a CAS loop on sm_100 (ATOMS.CAST.SPIN) -- a CTA-wide window with shared atomics
// was measured slower than the direct reductions; ownership is what makes the window pay.
//
// The flow gradient is the gather of warp_rows_bwd.cuh, unchanged.  Reference: resample2d_kernel.cu:68-190; autograd of
// models/networks.py:93-100 for mode GRIDSAMPLE.
#pragma once
#include "warp_rows_bwd.cuh"

namespace flowops {

constexpr int kWinW = 64, kWinH = 16;
constexpr int kWinWarps = 4;

__device__ __forceinline__ void red_add4(float *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct WinState {
    float *win;        // this warp's window: [CT][kWinH][kWinW]
    int ox, ylo0;      // image column of window column 0 (multiple of 4); image row of slot 0 at segment start
    int ylo;           // image row of the oldest window row (rolls)
};

// flush (and zero) the window row holding image row `gy`
template <int CT>
__device__ __forceinline__ void win_flush_row(const WinState &s, float *__restrict__ gi, unsigned hw, int H, int W, int gy, int lane)
{
    const int slot = (unsigned)(gy - s.ylo0) % kWinH;
    constexpr int quads = CT * (kWinW / 4);
#pragma unroll
    for (int j0 = 0; j0 < quads; j0 += 32) {
        const int j = j0 + lane;
        if (j < quads) {
            const int c = j / (kWinW / 4), xq = j - c * (kWinW / 4);
            float4 *cell = reinterpret_cast<float4 *>(s.win + (c * kWinH + slot) * kWinW + 4 * xq);
            const float4 v = *cell;
            if (!(v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f)) {
                // cells outside the image are never written (targets are clamped image pixels), so a non-zero quad is inside
                red_add4(gi + (size_t)c * hw + (size_t)gy * W + (s.ox + 4 * xq), v);
                *cell = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// one phase: lanes with `act` add v[c] at image offset `off` = ty * W + tx (distinct among the active lanes)
template <int CT>
__device__ __forceinline__ void win_add(const WinState &s, float *__restrict__ gi, unsigned hw, int W, bool act, int ty, int tx,
                                        const float (&v)[CT])
{
    if (!act) return;
    const unsigned wy = (unsigned)(ty - s.ylo), wx = (unsigned)(tx - s.ox);
    if (wy < (unsigned)kWinH && wx < (unsigned)kWinW) {
        const int slot = (unsigned)(ty - s.ylo0) % kWinH;
        float *p = s.win + slot * kWinW + wx;
#pragma unroll
        for (int c = 0; c < CT; ++c) p[c * kWinH * kWinW] += v[c];
    } else {
        float *p = gi + (size_t)ty * W + tx;
#pragma unroll
        for (int c = 0; c < CT; ++c) red_add_nz(p + (size_t)c * hw, v[c]);
    }
}

// sum v over the lanes that share this lane's target (peers: bit mask incl. this lane); every peer ends with the total
template <int CT>
__device__ __forceinline__ void win_dedupe(unsigned peers, int lane, float (&v)[CT])
{
    unsigned rem = peers & ~(1u << lane);
    float own[CT];                                   // the shuffles must hand out the ORIGINAL values, not partial sums
#pragma unroll
    for (int c = 0; c < CT; ++c) own[c] = v[c];
    while (__any_sync(0xffffffffu, rem != 0u)) {
        const int src = rem ? __ffs(rem) - 1 : lane;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float o = __shfl_sync(0xffffffffu, own[c], src);
            if (rem) v[c] += o;
        }
        rem &= rem - 1;
    }
}

// grid: (ceil(W / 32), ceil(n_segments / kWinWarps), B); block: (32, kWinWarps); dynamic shared memory:
// kWinWarps * CT * kWinH * kWinW floats.  Requires W % 4 == 0 and a 16-byte aligned gimg (v4 reductions).
template <int MODE, int CT, bool NEED_FLOW>
__global__ void __launch_bounds__(32 * kWinWarps, 4) warp_win_bwd_kernel(const __grid_constant__ WarpBwdArgs a)
{
    extern __shared__ __align__(16) float win_smem[];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const int S = a.rows;                                   // rows per segment
    const int xr = blockIdx.x * 32 + lane;
    const bool valid_x = xr < a.W;
    const int x = valid_x ? xr : a.W - 1;
    const int y0 = (blockIdx.y * kWinWarps + threadIdx.y) * S;
    WinState s;
    s.win = win_smem + threadIdx.y * (CT * kWinH * kWinW);
    // the window starts clean and every flush leaves its row clean
    for (int i = lane; i < CT * kWinH * kWinW / 4; i += 32) reinterpret_cast<float4 *>(s.win)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y0 >= a.H) return;                                  // uniform per warp
    const int y1 = min(y0 + S, a.H);
    const unsigned hw = (unsigned)a.H * a.W, W = (unsigned)a.W;
    const size_t b = blockIdx.z;
    const float *src = a.img + b * CT * hw;
    asm("" : "+l"(src));
    float *gi = a.gimg + b * CT * hw;
    asm("" : "+l"(gi));
    const unsigned p0 = (unsigned)y0 * W + (unsigned)x;
    const float *fl = a.flow + b * 2 * hw + p0;
    const float *go = a.gout + b * CT * hw + p0;
    float *gf = NEED_FLOW ? a.gflow + b * 2 * hw + p0 : nullptr;
    const float xfl = small_int_as_float(x);
    float yfl = small_int_as_float(y0);
    const float lin_xv = MODE == FLOWOPS_WARP_GRIDSAMPLE ? __ldg(a.lin_x + x) : 0.f;
    bool anchored = false;
    __syncwarp();

    float dx = ldg_stream(fl), dy = ldg_stream(fl + hw);
    for (int y = y0; y < y1; ++y) {
        float ndx = 0.f, ndy = 0.f;
        if (y + 1 < y1) { ndx = ldg_stream(fl + W); ndy = ldg_stream(fl + W + hw); }
        float g[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) g[c] = valid_x ? ldg_stream(go + (size_t)c * hw) : 0.f;

        // ---- coordinates and weights: identical to warp_rows_bwd_kernel ----
        unsigned o_t;
        int xL, yT;
        bool ex, ey;
        float w_tl, w_tr, w_bl, w_br;
        float gam_x = 0.f, gam_y = 0.f;
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
        float gmx = 0.f, gmy = 0.f;
        if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
            const float xf = __fadd_rn(xfl, dx), yf = __fadd_rn(yfl, dy);
            float fx = __fsub_rn(__fadd_rn(xf, kMagic15), kMagic15); fx = fx > xf ? __fsub_rn(fx, 1.f) : fx;
            float fy = __fsub_rn(__fadd_rn(yf, kMagic15), kMagic15); fy = fy > yf ? __fsub_rn(fy, 1.f) : fy;
            float tx = (xf < 0.f && fx != xf) ? __fadd_rn(fx, 1.f) : fx;      // (float)(int)xf: truncation
            float ty = (yf < 0.f && fy != yf) ? __fadd_rn(fy, 1.f) : fy;
            if (__builtin_expect(!(fmaxf(fabsf(xf), fabsf(yf)) < 4194304.f), 0)) {
                fx = floorf(xf); fy = floorf(yf); tx = (float)(int)xf; ty = (float)(int)yf;
            }
            xL = small_float_as_int(fminf(fmaxf(fx, 0.f), a.wm1));
            yT = small_float_as_int(fminf(fmaxf(fy, 0.f), a.hm1));
            ex = fx >= 0.f && fx < a.wm1;
            ey = fy >= 0.f && fy < a.hm1;
            const float alpha = __fsub_rn(xf, tx), beta = __fsub_rn(yf, ty);   // resample2d_kernel.cu:97-98
            w_tl = (1 - alpha) * (1 - beta); w_tr = alpha * (1 - beta);
            w_bl = (1 - alpha) * beta;       w_br = alpha * beta;
            gam_x = 1 - __fsub_rn(xf, fx);
            gam_y = 1 - __fsub_rn(yf, fy);
        } else {
            const float gx = __fadd_rn(lin_xv, __fmul_rn(dx, a.invx));
            const float gy = __fadd_rn(__ldg(a.lin_y + y), __fmul_rn(dy, a.invy));
            float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), a.wm1 + 1.f, -1.f), 0.5f);
            float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), a.hm1 + 1.f, -1.f), 0.5f);
            gmx = (ix <= 0.f || ix >= a.wm1) ? 0.f : 1.f;
            gmy = (iy <= 0.f || iy >= a.hm1) ? 0.f : 1.f;
            ix = fminf(a.wm1, fmaxf(ix, 0.f));
            iy = fminf(a.hm1, fmaxf(iy, 0.f));
            float fx = __fsub_rn(__fadd_rn(ix, kMagic15), kMagic15); fx = fx > ix ? __fsub_rn(fx, 1.f) : fx;
            float fy = __fsub_rn(__fadd_rn(iy, kMagic15), kMagic15); fy = fy > iy ? __fsub_rn(fy, 1.f) : fy;
            ex = __fadd_rn(fx, 1.f) <= a.wm1;
            ey = __fadd_rn(fy, 1.f) <= a.hm1;
            xL = small_float_as_int(fx); yT = small_float_as_int(fy);
            ax = __fadd_rn(fx, 1.f) - ix; ay = __fadd_rn(fy, 1.f) - iy;
            bx = ix - fx;                 by = iy - fy;
            w_tl = ax * ay; w_tr = bx * ay; w_bl = ax * by; w_br = bx * by;
        }
        o_t = (unsigned)yT * W + (unsigned)xL;
        const unsigned o_b = ey ? o_t + W : o_t;

        if (NEED_FLOW) {
            float gfx = 0.f, gfy = 0.f;
            const float *pt = src + o_t, *pb = src + o_b;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float tl = __ldg(pt), bl = __ldg(pb);
                const float trv = ex ? __ldg(pt + 1) : 0.f, brv = ex ? __ldg(pb + 1) : 0.f;
                const float tr = ex ? trv : tl, br = ex ? brv : bl;
                pt += hw; pb += hw;
                if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
                    gfy = __fmaf_rn(gam_x * g[c], bl, gfy);
                    gfy = __fmaf_rn(-(gam_x * g[c]), tl, gfy);
                    gfy = __fmaf_rn((1 - gam_x) * g[c], br, gfy);
                    gfy = __fmaf_rn(-((1 - gam_x) * g[c]), tr, gfy);
                    gfx = __fmaf_rn(gam_y * g[c], tr, gfx);
                    gfx = __fmaf_rn(-(gam_y * g[c]), tl, gfx);
                    gfx = __fmaf_rn((1 - gam_y) * g[c], br, gfx);
                    gfx = __fmaf_rn(-((1 - gam_y) * g[c]), bl, gfx);
                } else {
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

        // ---- window anchor: around the target of the warp's middle lane; rows then follow that lane's target row ----
        const int yT_mid = __shfl_sync(full, yT, 16);
        if (!anchored) {
            const int xL_mid = __shfl_sync(full, xL, 16);
            s.ox = (xL_mid - kWinW / 2) & ~3;                     // floor to a multiple of 4 (two's complement)
            s.ylo0 = s.ylo = yT_mid - (kWinH / 2 - 1);
            anchored = true;
        }

        // ---- image gradient: four corner values per channel, hand-over, dedupe, owned accumulation ----
        {
            const unsigned k_tl = valid_x ? o_t : 0xffffffffu - 4u * lane, k_tr = valid_x ? o_t + (ex ? 1u : 0u) : 0xfffffffeu - 4u * lane;
            const unsigned k_bl = valid_x ? o_b : 0xfffffffdu - 4u * lane, k_br = valid_x ? o_b + (ex ? 1u : 0u) : 0xfffffffcu - 4u * lane;
            const unsigned left_tr = __shfl_up_sync(full, k_tr, 1), left_br = __shfl_up_sync(full, k_br, 1);
            const unsigned right_tl = __shfl_down_sync(full, k_tl, 1), right_bl = __shfl_down_sync(full, k_bl, 1);
            const bool take = valid_x && lane > 0 && left_tr == k_tl && left_br == k_bl;
            const bool give = valid_x && lane < 31 && right_tl == k_tr && right_bl == k_br;
            float v_l[CT], v_r[CT], v_lb[CT], v_rb[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                float v_tl = w_tl * g[c], v_tr = w_tr * g[c], v_bl = w_bl * g[c], v_br = w_br * g[c];
                if (!ex) { v_tl += v_tr; v_bl += v_br; v_tr = 0.f; v_br = 0.f; }
                if (!ey) { v_tl += v_bl; v_tr += v_br; v_bl = 0.f; v_br = 0.f; }
                const float in_tr = __shfl_up_sync(full, v_tr, 1), in_br = __shfl_up_sync(full, v_br, 1);
                if (take) { v_tl += in_tr; v_bl += in_br; }
                v_l[c] = v_tl; v_r[c] = v_tr; v_lb[c] = v_bl; v_rb[c] = v_br;
            }
            const bool wr = valid_x && ex && !give;                // this lane writes its own right-hand corners
            const int yB = yT + 1;
            // Fast path: the lanes' left-hand columns are strictly increasing (any flow that does not fold over or hit
            // the clamp inside this warp).  Then every lane writes its own column, so the left-hand corners of all
            // lanes (top and bottom rows) are distinct and go out together; the right-hand corners that were not handed
            // over (column + 1, again distinct) follow after a __syncwarp.
            const int xl_left = __shfl_up_sync(full, xL, 1);
            const bool fast = __all_sync(full, !valid_x || lane == 0 || xL > xl_left);
            if (fast) {
                win_add<CT>(s, gi, hw, a.W, valid_x, yT, xL, v_l);
                win_add<CT>(s, gi, hw, a.W, valid_x && ey, yB, xL, v_lb);
                __syncwarp();
                if (__any_sync(full, wr)) {
                    win_add<CT>(s, gi, hw, a.W, wr, yT, xL + 1, v_r);
                    win_add<CT>(s, gi, hw, a.W, wr && ey, yB, xL + 1, v_rb);
                    __syncwarp();
                }
            } else {
                // general case: lanes whose corner coincides are found with match.any (slow: it iterates over the distinct
                // keys), summed by shuffles, and the lowest lane writes; four phases
                const unsigned peers = __match_any_sync(full, k_tl);
                const bool ey_uniform = __all_sync(full, ey || !valid_x) || __all_sync(full, !ey || !valid_x);
                unsigned peers_b = peers;
                // mixed ey (image bottom border inside the warp): group only the lanes that do write a bottom row
                if (!ey_uniform) peers_b = __match_any_sync(full, (valid_x && ey) ? o_b : 0xfffffffbu - 4u * lane);
                const unsigned wr_mask = __ballot_sync(full, wr);
                win_dedupe<CT>(peers, lane, v_l);
                win_dedupe<CT>(peers & wr_mask, lane, v_r);       // lanes outside wr_mask keep their (unused) values
                win_dedupe<CT>(peers_b, lane, v_lb);
                win_dedupe<CT>(peers_b & wr_mask, lane, v_rb);
                const bool lead = (peers & ((1u << lane) - 1u)) == 0u;
                const bool lead_r = ((peers & wr_mask) & ((1u << lane) - 1u)) == 0u;
                const bool lead_b = (peers_b & ((1u << lane) - 1u)) == 0u;
                const bool lead_rb = ((peers_b & wr_mask) & ((1u << lane) - 1u)) == 0u;
                win_add<CT>(s, gi, hw, a.W, valid_x && lead, yT, xL, v_l);
                __syncwarp();
                win_add<CT>(s, gi, hw, a.W, wr && lead_r, yT, xL + 1, v_r);
                __syncwarp();
                win_add<CT>(s, gi, hw, a.W, valid_x && ey && lead_b, yB, xL, v_lb);
                __syncwarp();
                win_add<CT>(s, gi, hw, a.W, wr && ey && lead_rb, yB, xL + 1, v_rb);
                __syncwarp();
            }
        }
        // roll: rows the middle lane's target has moved past leave the window (at most a few per source row)
        {
            const int want = yT_mid + 1 - (kWinH / 2 - 1);       // where the window should start for the next source row
            while (s.ylo < want) {
                win_flush_row<CT>(s, gi, hw, a.H, a.W, s.ylo, lane);
                s.ylo += 1;
            }
            __syncwarp();
        }

        dx = ndx; dy = ndy;
        fl += W; go += W; gf += W; yfl = __fadd_rn(yfl, 1.f);
    }
    // drain what the window still holds
    for (int r = 0; r < kWinH; ++r) win_flush_row<CT>(s, gi, hw, a.H, a.W, s.ylo + r, lane);
}

template <int MODE, bool NEED_FLOW>
static inline int launch_warp_win_bwd(WarpBwdArgs a, cudaStream_t st)
{
    const int B = a.B;
    const size_t hw = (size_t)a.H * a.W, chw = (size_t)a.C * hw;
    const int xb = (a.W + 31) / 32;
    // rows per segment: long segments amortise the final drain of the window, but the grid should fill the GPU
    int S = 32;
    while (S > 8 && (long long)B * xb * ((a.H + S - 1) / S) < 2LL * kNumSMs * 4 * kWinWarps) S >>= 1;
    a.rows = S;
    const int segs = (a.H + S - 1) / S;
    const size_t smem = sizeof(float) * kWinWarps * a.C * kWinH * kWinW;
    const void *fn = a.C == 3 ? (const void *)warp_win_bwd_kernel<MODE, 3, NEED_FLOW>
                   : a.C == 2 ? (const void *)warp_win_bwd_kernel<MODE, 2, NEED_FLOW>
                              : (const void *)warp_win_bwd_kernel<MODE, 1, NEED_FLOW>;
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("warp_bwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    for (int b0 = 0; b0 < B; b0 += 65535) {
        WarpBwdArgs c = a;
        c.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + b0 * chw; c.flow = a.flow + (size_t)b0 * 2 * hw; c.gout = a.gout + b0 * chw;
        c.gimg = a.gimg + b0 * chw;
        if (a.gflow) c.gflow = a.gflow + (size_t)b0 * 2 * hw;
        const dim3 grid(xb, (segs + kWinWarps - 1) / kWinWarps, c.B), block(32, kWinWarps, 1);
        if (a.C == 3) warp_win_bwd_kernel<MODE, 3, NEED_FLOW><<<grid, block, smem, st>>>(c);
        else if (a.C == 2) warp_win_bwd_kernel<MODE, 2, NEED_FLOW><<<grid, block, smem, st>>>(c);
        else warp_win_bwd_kernel<MODE, 1, NEED_FLOW><<<grid, block, smem, st>>>(c);
    }
    return 0;
}

}  // namespace flowops
