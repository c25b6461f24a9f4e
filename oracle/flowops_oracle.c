/*
 * flowops_oracle.c -- CPU restatement of the reference's flow hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it, and only as the checker.  The shipped operators (ir2rgb_b200/) never import it and have
 * no CPU fallback.
 *
 * Each function is a literal, thread-for-thread transliteration of one reference CUDA kernel
 * (paths relative to /root/reference/models/flownet2_pytorch/networks/), including the order of
 * the floating-point operations, the places where the reference promotes to double, and the fused
 * multiply-adds nvcc emits for `a += b * c` (default -fmad=true).  Build with -ffp-contract=off so
 * that the only contractions are the explicit fmaf() calls below.
 *
 * Parity pin: the reference ships no tests or golden vectors.  The pins are
 * (1) tests/golden/*.npz -- outputs of the reference's own CUDA extensions, rebuilt for sm_100 by
 *     oracle/build_ref.py and executed on a B200 (generator: tests/golden/make_golden_gpu.py), and
 * (2) for networks.resample, outputs of the reference's own Python code path (get_grid +
 *     F.grid_sample) imported from /root/reference (generator: tests/golden/make_golden_cpu.py).
 * tests/test_oracle.py checks this file against both.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX4(b, c, y, x, C, H, W) ((((size_t)(b) * (C) + (c)) * (H) + (y)) * (size_t)(W) + (x))

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
/* float -> int as the GPU does it (cvt.rzi.s32.f32): saturating, NaN -> 0.  In C the conversion of an
 * out-of-range value is undefined (x86 yields INT_MIN), so a flow of 3e38 would clamp to column 0 here and
 * to column W-1 in the reference kernel. */
static inline int f2i_cuda(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* ------------------------------------------------------------------------------------------ */
/* ChannelNorm                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* channelnorm_package/channelnorm_kernel.cu:19-60 (kernel_channelnorm_update_output).
 * One output per (b, y, x): fp32 accumulator, `result += val * val` in channel order (FFMA),
 * IEEE sqrt.  norm_deg is ignored by the reference kernel. */
void oracle_cnorm_fwd(const float *x, float *y, int B, int C, int H, int W)
{
    const size_t hw = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)B * (long long)hw; ++i) {
        const size_t b = (size_t)i / hw, p = (size_t)i % hw;
        float result = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float val = x[(b * C + c) * hw + p];
            result = fmaf(val, val, result);                    /* :55-56 */
        }
        y[i] = sqrtf(result);                                   /* :58-59 */
    }
}

/* channelnorm_kernel.cu:64-96 (kernel_channelnorm_backward_input1).
 * gO * x is an fp32 product; the divide is fp64 because of the 1e-9 literal (:93). */
void oracle_cnorm_bwd(const float *x, const float *y, const float *gy, float *gx,
                      int B, int C, int H, int W)
{
    const size_t hw = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)B * C * (long long)hw; ++i) {
        const size_t b = (size_t)i / (hw * C), p = (size_t)i % hw;
        const float prod = gy[b * hw + p] * x[i];
        gx[i] = (float)((double)prod / ((double)y[b * hw + p] + 1e-9));
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Resample2d (kernel_size == 1 only; larger kernels read out of bounds in the reference)     */
/* ------------------------------------------------------------------------------------------ */

/* resample2d_package/resample2d_kernel.cu:16-64 (kernel_resample2d_update_output<float>).
 * The first three weight products are fp64 (literal `1.`), the fourth is fp32 (:55-58). */
void oracle_resample2d_fwd(const float *img, const float *flow, float *out,
                           int B, int C, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float dx = flow[IDX4(b, 0, y, x, 2, H, W)];
                const float dy = flow[IDX4(b, 1, y, x, 2, H, W)];
                const float xf = (float)x + dx;                            /* :43 */
                const float yf = (float)y + dy;
                const float alpha = xf - floorf(xf);                       /* :45 */
                const float beta = yf - floorf(yf);
                const int xL = imax(imin(f2i_cuda(floorf(xf)), W - 1), 0);      /* :48-51 */
                const int xR = imax(imin(f2i_cuda(floorf(xf) + 1), W - 1), 0);
                const int yT = imax(imin(f2i_cuda(floorf(yf)), H - 1), 0);
                const int yB = imax(imin(f2i_cuda(floorf(yf) + 1), H - 1), 0);
                for (int c = 0; c < C; ++c) {
                    const float tl = img[IDX4(b, c, yT, xL, C, H, W)];
                    const float tr = img[IDX4(b, c, yT, xR, C, H, W)];
                    const float bl = img[IDX4(b, c, yB, xL, C, H, W)];
                    const float br = img[IDX4(b, c, yB, xR, C, H, W)];
                    float val = 0.0f;
                    val += (float)((1. - alpha) * (1. - beta) * tl);       /* :55 */
                    val += (float)((alpha) * (1. - beta) * tr);            /* :56 */
                    val += (float)((1. - alpha) * (beta) * bl);            /* :57 */
                    val = fmaf(alpha * beta, br, val);                     /* :58, all-fp32 FFMA */
                    out[IDX4(b, c, y, x, C, H, W)] = val;
                }
            }
}

/* resample2d_kernel.cu:68-117 (kernel_resample2d_backward_input1<float>): atomic scatter of the
 * four fp32 weights.  alpha/beta use truncation here (:97-98).  The CPU restatement accumulates
 * in element-index order (the GPU order is non-deterministic). gimg must be zero on entry
 * (resample2d.py:29). */
void oracle_resample2d_bwd_img(const float *flow, const float *gout, float *gimg,
                               int B, int C, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const float dx = flow[IDX4(b, 0, y, x, 2, H, W)];
                    const float dy = flow[IDX4(b, 1, y, x, 2, H, W)];
                    const float xf = (float)x + dx;
                    const float yf = (float)y + dy;
                    const float alpha = xf - (float)f2i_cuda(xf);               /* :97 */
                    const float beta = yf - (float)f2i_cuda(yf);                /* :98 */
                    const int xL = imax(imin(f2i_cuda(floorf(xf)), W - 1), 0);  /* :103-106 */
                    const int xR = imax(imin(f2i_cuda(floorf(xf) + 1), W - 1), 0);
                    const int yT = imax(imin(f2i_cuda(floorf(yf)), H - 1), 0);
                    const int yB = imax(imin(f2i_cuda(floorf(yf) + 1), H - 1), 0);
                    const float g = gout[IDX4(b, c, y, x, C, H, W)];
                    gimg[IDX4(b, c, yT, xL, C, H, W)] += (1 - alpha) * (1 - beta) * g;   /* :110 */
                    gimg[IDX4(b, c, yT, xR, C, H, W)] += (alpha) * (1 - beta) * g;       /* :111 */
                    gimg[IDX4(b, c, yB, xL, C, H, W)] += (1 - alpha) * (beta) * g;       /* :112 */
                    gimg[IDX4(b, c, yB, xR, C, H, W)] += (alpha) * (beta) * g;           /* :113 */
                }
}

/* resample2d_kernel.cu:120-190 (kernel_resample2d_backward_input2<float>): flow gradient, one
 * thread per (b, c in {0,1}, y, x); `output +=/-= gamma * gO * I` contracts to FMUL + FFMA. */
void oracle_resample2d_bwd_flow(const float *img, const float *flow, const float *gout,
                                float *gflow, int B, int C, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < 2; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const float dx = flow[IDX4(b, 0, y, x, 2, H, W)];
                    const float dy = flow[IDX4(b, 1, y, x, 2, H, W)];
                    const float xf = (float)x + dx;
                    const float yf = (float)y + dy;
                    const int xL = imax(imin(f2i_cuda(floorf(xf)), W - 1), 0);  /* :154-157 */
                    const int xR = imax(imin(f2i_cuda(floorf(xf) + 1), W - 1), 0);
                    const int yT = imax(imin(f2i_cuda(floorf(yf)), H - 1), 0);
                    const int yB = imax(imin(f2i_cuda(floorf(yf) + 1), H - 1), 0);
                    float output = 0.0f;
                    if (c % 2) {                                           /* :159-170, d/d(dy) */
                        const float gamma = 1 - (xf - floorf(xf));
                        for (int ch = 0; ch < C; ++ch) {
                            const float g = gout[IDX4(b, ch, y, x, C, H, W)];
                            output = fmaf(gamma * g, img[IDX4(b, ch, yB, xL, C, H, W)], output);
                            output = fmaf(-(gamma * g), img[IDX4(b, ch, yT, xL, C, H, W)], output);
                            output = fmaf((1 - gamma) * g, img[IDX4(b, ch, yB, xR, C, H, W)], output);
                            output = fmaf(-((1 - gamma) * g), img[IDX4(b, ch, yT, xR, C, H, W)], output);
                        }
                    } else {                                               /* :171-184, d/d(dx) */
                        const float gamma = 1 - (yf - floorf(yf));
                        for (int ch = 0; ch < C; ++ch) {
                            const float g = gout[IDX4(b, ch, y, x, C, H, W)];
                            output = fmaf(gamma * g, img[IDX4(b, ch, yT, xR, C, H, W)], output);
                            output = fmaf(-(gamma * g), img[IDX4(b, ch, yT, xL, C, H, W)], output);
                            output = fmaf((1 - gamma) * g, img[IDX4(b, ch, yB, xR, C, H, W)], output);
                            output = fmaf(-((1 - gamma) * g), img[IDX4(b, ch, yB, xL, C, H, W)], output);
                        }
                    }
                    gflow[IDX4(b, c, y, x, 2, H, W)] = output;
                }
}

/* ------------------------------------------------------------------------------------------ */
/* Correlation                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* correlation_package/correlation_cuda.cc:19-34: output shape. */
void oracle_corr_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                       int *oC, int *oH, int *oW)
{
    const int kr = (k - 1) / 2, br = kr + md;
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int D = (md / s2) * 2 + 1;
    *oC = D * D;
    *oH = (int)ceilf((float)(pH - 2 * br) / (float)s1);
    *oW = (int)ceilf((float)(pW - 2 * br) / (float)s1);
}

/* Element (yy, xx, ch) of the zero-padded NHWC scratch the reference builds with
 * channels_first (correlation_cuda_kernel.cu:47-70) after fill_(0) (correlation_cuda.cc:40-41).
 * Reads outside the padded array (possible in the reference only for kernel_size > 1, where it
 * is an out-of-bounds read) are defined as 0 here. */
static inline float padded(const float *in, int n, int C, int H, int W, int pad,
                           int yy, int xx, int ch)
{
    const int y = yy - pad, x = xx - pad;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    return in[IDX4(n, ch, y, x, C, H, W)];
}

/* The reference's shfl_down tree (correlation_cuda_kernel.cu:17-21) as seen by lane 0.
 * __shfl_down_sync with an out-of-range source lane returns the caller's own value. */
static float warp_reduce_lane0(float v[32])
{
    float t[32];
    for (int off = 16; off > 0; off /= 2) {
        for (int l = 0; l < 32; ++l) t[l] = v[l] + (l + off < 32 ? v[l + off] : v[l]);
        memcpy(v, t, sizeof(t));
    }
    return v[0];
}

/* correlation_cuda_kernel.cu:74-147 (correlation_forward<float>), block = one warp per output
 * pixel (:400-401): lane l accumulates channels l, l+32, ... with FFMA (:124), then the shuffle
 * tree, then a true fp32 divide by nelems = k*k*C (:143).  corr_multiply is ignored. */
void oracle_corr_fwd(const float *in1, const float *in2, float *out,
                     int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    int oC, oH, oW;
    oracle_corr_shape(H, W, pad, k, md, s1, s2, &oC, &oH, &oW);
    const int kr = (k - 1) / 2, dr = md / s2, D = 2 * dr + 1;
    const int nelems = k * k * C;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
        for (int oy = 0; oy < oH; ++oy)
            for (int ox = 0; ox < oW; ++ox) {
                const int y1 = oy * s1 + md, x1 = ox * s1 + md;            /* :90-91 */
                for (int tj = -dr; tj <= dr; ++tj)
                    for (int ti = -dr; ti <= dr; ++ti) {
                        const int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
                        float acc[32];
                        for (int l = 0; l < 32; ++l) acc[l] = 0.0f;
                        for (int j = -kr; j <= kr; ++j)
                            for (int i = -kr; i <= kr; ++i)
                                for (int l = 0; l < 32; ++l)
                                    for (int ch = l; ch < C; ch += 32)
                                        acc[l] = fmaf(padded(in1, n, C, H, W, pad, y1 + j, x1 + i, ch),
                                                      padded(in2, n, C, H, W, pad, y2 + j, x2 + i, ch),
                                                      acc[l]);
                        const float total = warp_reduce_lane0(acc);
                        const int tc = (tj + dr) * D + (ti + dr);          /* :139-140 */
                        out[IDX4(n, tc, oy, ox, oC, oH, oW)] = total / nelems;
                    }
            }
}

/* correlation_cuda_kernel.cu:151-241 (backward_input1) and :244-334 (backward_input2): block of
 * 32 threads per (y, x, c); thread t handles output channels t, t+32, ...; partial sums in shared
 * memory (`prod_sum[t] += gO * val`, FFMA), thread 0 adds the 32 partials in order and divides by
 * nelems (fp32).  Integer divisions by stride1 are C truncating divisions, as in the kernel. */
void oracle_corr_bwd(const float *in1, const float *in2, const float *gout,
                     float *gin1, float *gin2,
                     int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    int oC, oH, oW;
    oracle_corr_shape(H, W, pad, k, md, s1, s2, &oC, &oH, &oW);
    const int kr = (k - 1) / 2, dr = md / s2, D = 2 * dr + 1;
    const float nelems = (float)(k * k * C);
    memset(gin1, 0, sizeof(float) * (size_t)B * C * H * W);                /* correlation_cuda.cc:114-115 */
    memset(gin2, 0, sizeof(float) * (size_t)B * C * H * W);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c)
            for (int by = 0; by < H; ++by)
                for (int bx = 0; bx < W; ++bx) {
                    const int y = by * s1 + pad, x = bx * s1 + pad;        /* :164-165 */
                    /* ---- grad wrt input1 ---- */
                    {
                        int xmin = (x - kr - md) / s1, ymin = (y - kr - md) / s1;   /* :173-177 */
                        int xmax = (x + kr - md) / s1, ymax = (y + kr - md) / s1;
                        if (!(xmax < 0 || ymax < 0 || xmin >= oW || ymin >= oH) &&
                            !(xmin > xmax || ymin > ymax)) {
                            xmin = imax(0, xmin); xmax = imin(oW - 1, xmax);
                            ymin = imax(0, ymin); ymax = imin(oH - 1, ymax);
                            float prod_sum[32];
                            for (int t = 0; t < 32; ++t) prod_sum[t] = 0.0f;
                            for (int t = 0; t < 32; ++t)
                                for (int tc = t; tc < oC; tc += 32) {
                                    const int i2 = (tc % D - dr) * s2, j2 = (tc / D - dr) * s2;
                                    const float val2 = padded(in2, n, C, H, W, pad, y + j2, x + i2, c);
                                    for (int j = ymin; j <= ymax; ++j)
                                        for (int i = xmin; i <= xmax; ++i)
                                            prod_sum[t] = fmaf(gout[IDX4(n, tc, j, i, oC, oH, oW)], val2, prod_sum[t]);
                                }
                            float reduce_sum = 0.0f;
                            for (int t = 0; t < 32; ++t) reduce_sum += prod_sum[t];
                            if (y - pad < H && x - pad < W)
                                gin1[IDX4(n, c, y - pad, x - pad, C, H, W)] = reduce_sum / nelems;
                        }
                    }
                    /* ---- grad wrt input2 ---- */
                    {
                        float prod_sum[32];
                        for (int t = 0; t < 32; ++t) prod_sum[t] = 0.0f;
                        for (int t = 0; t < 32; ++t)
                            for (int tc = t; tc < oC; tc += 32) {
                                const int i2 = (tc % D - dr) * s2, j2 = (tc / D - dr) * s2;
                                int xmin = (x - kr - md - i2) / s1, ymin = (y - kr - md - j2) / s1;   /* :289-293 */
                                int xmax = (x + kr - md - i2) / s1, ymax = (y + kr - md - j2) / s1;
                                if (xmax < 0 || ymax < 0 || xmin >= oW || ymin >= oH) continue;
                                if (xmin > xmax || ymin > ymax) continue;
                                xmin = imax(0, xmin); xmax = imin(oW - 1, xmax);
                                ymin = imax(0, ymin); ymax = imin(oH - 1, ymax);
                                const float val1 = padded(in1, n, C, H, W, pad, y - j2, x - i2, c);
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i)
                                        prod_sum[t] = fmaf(gout[IDX4(n, tc, j, i, oC, oH, oW)], val1, prod_sum[t]);
                            }
                        float reduce_sum = 0.0f;
                        for (int t = 0; t < 32; ++t) reduce_sum += prod_sum[t];
                        if (y - pad < H && x - pad < W)
                            gin2[IDX4(n, c, y - pad, x - pad, C, H, W)] = reduce_sum / nelems;
                    }
                }
}

/* ------------------------------------------------------------------------------------------ */
/* networks.resample (models/networks.py:15-28, 89-100; models/base_model.py:123-136)         */
/* ------------------------------------------------------------------------------------------ */

/* The arithmetic lives in third-party PyTorch (pinned torch~=1.6.0 in requirements.txt:1; the
 * installed 2.11 has the same semantics for bilinear / border / align_corners=False):
 *   ATen/native/cuda/GridSampler.cuh  grid_sampler_unnormalize, clip_coordinates
 *   ATen/native/cuda/GridSampler.cu   grid_sampler_2d_kernel (bilinear branch)
 * restated here op for op in fp32:
 *   g   = lin[x] + flow * inv            (networks.py:97-98; CUDA `div` by a Python scalar is a
 *                                         multiply by the fp32 reciprocal, `inv_mode` = 1; the
 *                                         CPU kernel really divides, `inv_mode` = 0)
 *   ix  = ((g + 1) * W - 1) / 2          (unnormalize, align_corners=False; nvcc contracts the
 *                                         multiply-subtract into one FFMA, `fma_mode` = 1)
 *   ix  = min(W - 1, max(ix, 0))         (border clip)
 *   nw/ne/sw/se weights as differences, out = sum in nw, ne, sw, se order (FFMA chain).
 * lin_x / lin_y are torch.linspace(-1, 1, W|H) tables supplied by the caller (get_grid,
 * networks.py:15-28). */
void oracle_gridwarp_fwd(const float *img, const float *flow, float *out,
                         const float *lin_x, const float *lin_y,
                         int B, int C, int H, int W, int inv_mode, int fma_mode)
{
    const float sx = (float)((W - 1.0) / 2.0), sy = (float)((H - 1.0) / 2.0);
    const float invx = 1.0f / sx, invy = 1.0f / sy;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float dx = flow[IDX4(b, 0, y, x, 2, H, W)];
                const float dy = flow[IDX4(b, 1, y, x, 2, H, W)];
                const float gx = lin_x[x] + (inv_mode ? dx * invx : dx / sx);
                const float gy = lin_y[y] + (inv_mode ? dy * invy : dy / sy);
                float ix = fma_mode ? fmaf(gx + 1.f, (float)W, -1.f) / 2 : ((gx + 1.f) * W - 1) / 2;
                float iy = fma_mode ? fmaf(gy + 1.f, (float)H, -1.f) / 2 : ((gy + 1.f) * H - 1) / 2;
                ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
                iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
                const float fx = floorf(ix), fy = floorf(iy);
                const int ix_nw = (int)fx, iy_nw = (int)fy;
                const int ix_ne = ix_nw + 1, iy_ne = iy_nw;
                const int ix_sw = ix_nw, iy_sw = iy_nw + 1;
                const int ix_se = ix_nw + 1, iy_se = iy_nw + 1;
                const float nw = ((float)ix_se - ix) * ((float)iy_se - iy);
                const float ne = (ix - (float)ix_sw) * ((float)iy_sw - iy);
                const float sw = ((float)ix_ne - ix) * (iy - (float)iy_ne);
                const float se = (ix - (float)ix_nw) * (iy - (float)iy_nw);
                for (int c = 0; c < C; ++c) {
                    float acc = 0.f;
                    if (iy_nw >= 0 && iy_nw < H && ix_nw >= 0 && ix_nw < W)
                        acc = fmaf(img[IDX4(b, c, iy_nw, ix_nw, C, H, W)], nw, acc);
                    if (iy_ne >= 0 && iy_ne < H && ix_ne >= 0 && ix_ne < W)
                        acc = fmaf(img[IDX4(b, c, iy_ne, ix_ne, C, H, W)], ne, acc);
                    if (iy_sw >= 0 && iy_sw < H && ix_sw >= 0 && ix_sw < W)
                        acc = fmaf(img[IDX4(b, c, iy_sw, ix_sw, C, H, W)], sw, acc);
                    if (iy_se >= 0 && iy_se < H && ix_se >= 0 && ix_se < W)
                        acc = fmaf(img[IDX4(b, c, iy_se, ix_se, C, H, W)], se, acc);
                    out[IDX4(b, c, y, x, C, H, W)] = acc;
                }
            }
}

/* ------------------------------------------------------------------------------------------ */
/* 16-bit storage (fp16 / bf16): the reference's fp16 mode                                    */
/* ------------------------------------------------------------------------------------------ */
/* Tensors cross this interface as float arrays whose values are representable in the storage type;
 * `dtype` is 1 (IEEE binary16) or 2 (bfloat16), as FLOWOPS_DTYPE_F16 / _BF16 in include/flowops.h.
 * round16() is float -> storage type -> float with round-to-nearest-even, i.e. what c10::Half's /
 * c10::BFloat16's float constructor and `.half()` / `.bfloat16()` do (cvt.rn.f16.f32 /
 * cvt.rn.bf16.f32 on the device); written out on the bit pattern so that it does not depend on the
 * compiler's _Float16 support.  tests/test_oracle.py pins it against numpy.float16 and torch.bfloat16. */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static float round_half(float f)
{
    const uint32_t u = f2u(f), sign = u & 0x80000000u, a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return u2f(sign | 0x7fc00000u);          /* NaN */
    if (a >= 0x477ff000u) return u2f(sign | 0x7f800000u);         /* >= 65520 rounds to inf (inf stays inf) */
    if (a < 0x33000001u) return u2f(sign);                        /* <= 2^-25 rounds to zero (ties to even) */
    if (a >= 0x38800000u) {                                       /* normal half: drop 13 mantissa bits */
        const uint32_t r = a + 0xfffu + ((a >> 13) & 1u);          /* nearest, ties to even; a carry moves the exponent */
        return u2f(sign | (r & ~0x1fffu));
    }
    /* subnormal half: a multiple of 2^-24.  |f| = m * 2^(e-150) with the implicit one in m, so the number of
     * quanta is m >> (126 - e), 14 <= shift <= 24 here */
    const uint32_t m = (a & 0x007fffffu) | 0x00800000u, sh = 126u - (a >> 23);
    const uint32_t q = (m + ((1u << (sh - 1)) - 1u) + ((m >> sh) & 1u)) >> sh;
    return (sign ? -1.0f : 1.0f) * (float)q * 5.9604644775390625e-08f;          /* exact: q <= 1024 */
}

static float round_bf16(float f)
{
    const uint32_t u = f2u(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return u2f((u & 0x80000000u) | 0x7fc00000u);     /* NaN */
    const uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);                                  /* may carry into inf: correct */
    return u2f(r & 0xffff0000u);
}

static inline float round16(float f, int dtype) { return dtype == 1 ? round_half(f) : round_bf16(f); }

void oracle_round16(const float *in, float *out, long long n, int dtype)
{
    for (long long i = 0; i < n; ++i) out[i] = round16(in[i], dtype);
}

/* kernel_channelnorm_update_output<at::Half> (channelnorm_kernel.cu:19-60): `val * val` multiplies two
 * at::Half values -- a float product rounded back to half (c10/util/Half-inl.h operator*) -- before
 * static_cast<float> and the fp32 `result +=` (:55-56); sqrt in fp32 (:58), rounded to half on store (:59). */
void oracle_cnorm_fwd_16(const float *x, float *y, int B, int C, int H, int W, int dtype)
{
    const size_t hw = (size_t)H * W;
    for (long long i = 0; i < (long long)B * (long long)hw; ++i) {
        const size_t b = (size_t)i / hw, p = (size_t)i % hw;
        float result = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float val = x[(b * C + c) * hw + p];
            result = result + round16(val * val, dtype);
        }
        y[i] = round16(sqrtf(result), dtype);
    }
}

/* kernel_channelnorm_backward_input1<at::Half> (channelnorm_kernel.cu:64-96): fp32 product of the widened
 * operands, fp64 divide (the 1e-9 literal), `val` is a float, the store rounds it to half (:93-94). */
void oracle_cnorm_bwd_16(const float *x, const float *y, const float *gy, float *gx,
                         int B, int C, int H, int W, int dtype)
{
    const size_t hw = (size_t)H * W;
    for (long long i = 0; i < (long long)B * C * (long long)hw; ++i) {
        const size_t b = (size_t)i / (hw * C), p = (size_t)i % hw;
        const float prod = gy[b * hw + p] * x[i];
        const float val = (float)((double)prod / ((double)y[b * hw + p] + 1e-9));
        gx[i] = round16(val, dtype);
    }
}

/* Model.resample with opt['fp16'] on 16-bit tensors (models/base_model.py:123-136):
 *   grid  = get_grid(..., dtype=flow.dtype)            -> lin_x / lin_y arrive rounded to the storage type
 *   nflow = flow / ((w-1)/2)                           -> rounded to the storage type (CUDA: multiply by the fp32
 *                                                         reciprocal, inv_mode 1; CPU: true divide, inv_mode 0)
 *   g     = grid + nflow                               -> rounded to the storage type
 *   out   = grid_sample(image.float(), g.float(), bilinear, border).half()      -> fp32 chain as above, rounded */
void oracle_gridwarp_fwd_16(const float *img, const float *flow, float *out,
                            const float *lin_x, const float *lin_y,
                            int B, int C, int H, int W, int inv_mode, int fma_mode, int dtype)
{
    const float sx = (float)((W - 1.0) / 2.0), sy = (float)((H - 1.0) / 2.0);
    const float invx = 1.0f / sx, invy = 1.0f / sy;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const float dx = flow[IDX4(b, 0, y, x, 2, H, W)];
                const float dy = flow[IDX4(b, 1, y, x, 2, H, W)];
                const float gx = round16(lin_x[x] + round16(inv_mode ? dx * invx : dx / sx, dtype), dtype);
                const float gy = round16(lin_y[y] + round16(inv_mode ? dy * invy : dy / sy, dtype), dtype);
                float ix = fma_mode ? fmaf(gx + 1.f, (float)W, -1.f) / 2 : ((gx + 1.f) * W - 1) / 2;
                float iy = fma_mode ? fmaf(gy + 1.f, (float)H, -1.f) / 2 : ((gy + 1.f) * H - 1) / 2;
                ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
                iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
                const float fx = floorf(ix), fy = floorf(iy);
                const int ix_nw = (int)fx, iy_nw = (int)fy;
                const int ix_se = ix_nw + 1, iy_se = iy_nw + 1;
                const float nw = ((float)ix_se - ix) * ((float)iy_se - iy);
                const float ne = (ix - (float)ix_nw) * ((float)iy_se - iy);
                const float sw = ((float)ix_se - ix) * (iy - (float)iy_nw);
                const float se = (ix - (float)ix_nw) * (iy - (float)iy_nw);
                for (int c = 0; c < C; ++c) {
                    float acc = fmaf(img[IDX4(b, c, iy_nw, ix_nw, C, H, W)], nw, 0.f);     /* nw is always inside after the clip */
                    if (ix_se < W) acc = fmaf(img[IDX4(b, c, iy_nw, ix_se, C, H, W)], ne, acc);
                    if (iy_se < H) acc = fmaf(img[IDX4(b, c, iy_se, ix_nw, C, H, W)], sw, acc);
                    if (iy_se < H && ix_se < W) acc = fmaf(img[IDX4(b, c, iy_se, ix_se, C, H, W)], se, acc);
                    out[IDX4(b, c, y, x, C, H, W)] = round16(acc, dtype);
                }
            }
}

/* ------------------------------------------------------------------------------------------ */
/* Model of the product's ChannelNorm-backward quotient (ir2rgb_b200/csrc/cnorm.cu)             */
/* ------------------------------------------------------------------------------------------ */
/* Not a restatement of the reference: a CPU model of the arithmetic the CUDA kernel uses INSTEAD of the reference's
 * fp64 divide, so that its exactness argument can be tested without a GPU.  The kernel starts from a hardware
 * reciprocal seed r0 = (1/d)(1 + delta) -- delta is an input here, the test sweeps |delta| up to 2^-16, far worse than
 * MUFU.RCP64H -- applies two Newton steps, forms q = p*r, corrects it with one residual step and rounds to float;
 * `want` is what the reference computes, (float)((double)prod / d) (channelnorm_kernel.cu:93). */
void oracle_quotient_model(const float *prod, const double *d, const double *delta, float *got, float *want, long long n)
{
    for (long long i = 0; i < n; ++i) {
        const double p = (double)prod[i];
        double r = (1.0 / d[i]) * (1.0 + delta[i]);
        r = fma(r, fma(-d[i], r, 1.0), r);
        r = fma(r, fma(-d[i], r, 1.0), r);
        double q = p * r;
        q = fma(fma(-q, d[i], p), r, q);
        got[i] = copysignf((float)q, prod[i]);
        want[i] = (float)(p / d[i]);
    }
}

int oracle_abi_version(void) { return 2; }
