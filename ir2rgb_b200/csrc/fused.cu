// fused.cu -- FlowNet2 glue kernels: the operator chains that surround
// the conv body in reference models/flownet2_pytorch/models.py:109-150 and models/flownet.py:50, each
// collapsed into one pass over HBM.  Additive API: results are bit-identical to running the separate
// Resample2d / subtract / ChannelNorm operators.
//
//   warp_diff_norm : warped = Resample2d(img1, flow); norm = ChannelNorm(img0 - warped)
//       unfused traffic per pixel (C = 3): warp 8 + .contiguous() copy of the x[:,3:] slice 6 +
//       subtract 9 + norm 4 = 27 floats;  fused: 3 + 3 + 2 read, 3 + 1 written = 12 floats.
//       Inputs may be channel slices of a wider tensor (batch stride given), outputs may be channel
//       slices of the concat buffer the next sub-network reads (models.py:114,126).
//   warp_conf      : conf = (sum_c (im1 - warp(im2, flow))^2 < thresh) as 0/1 floats (flownet.py:50,56-57):
//       3 + 3 + 2 read, 1 written, instead of five elementwise kernels plus the grid_sample chain.  NOTE: in the
//       reference `self.resample` inside FlowNet resolves to the METHOD Model.resample (base_model.py:129, the
//       grid_sample warp), not to the Resample2d submodule assigned at flownet.py:17 -- a class attribute shadows
//       an nn.Module submodule of the same name -- so the as-run mask uses mode GRIDSAMPLE; both modes exist here.
#include "warp_rows.cuh"

namespace flowops {

template <int CT, bool WRITE_WARPED>
__global__ void __launch_bounds__(256) warp_diff_norm_kernel(const float *__restrict__ img0, const float *__restrict__ img1,
                                                             size_t img_bs, const float *__restrict__ flow,
                                                             float *__restrict__ warped, size_t warped_bs,
                                                             float *__restrict__ norm, size_t norm_bs,
                                                             int B, int C, int H, int W)
{
    const int c_n = CT > 0 ? CT : C;
    const size_t hw = (size_t)H * W;
    const size_t total = (size_t)B * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw;
        const int p = (int)(i - b * hw);
        const int y = p / W, x = p - y * W;
        const float dx = ldg_stream(flow + (b * 2) * hw + p);
        const float dy = ldg_stream(flow + (b * 2 + 1) * hw + p);
        float xf, yf; Corners k;
        r2d_coords(x, y, dx, dy, H, W, xf, yf, k);
        const R2dWeights w = r2d_weights(xf, yf);
        const float *src = img1 + b * img_bs;
        const float *ref = img0 + b * img_bs + p;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < c_n; ++c) {
            const float *pl = src + (size_t)c * hw;
            const float val = r2d_blend(w, __ldg(pl + k.o_tl), __ldg(pl + k.o_tr), __ldg(pl + k.o_bl), __ldg(pl + k.o_br));
            if (WRITE_WARPED) stg_stream(warped + b * warped_bs + (size_t)c * hw + p, val);
            const float d = __fsub_rn(ldg_stream(ref + (size_t)c * hw), val);   // img0 - warped (models.py:110)
            acc = __fmaf_rn(d, d, acc);                                        // channelnorm_kernel.cu:55-56
        }
        stg_stream(norm + b * norm_bs + p, __fsqrt_rn(acc));
    }
}

template <int MODE, int CT>
__global__ void __launch_bounds__(256) warp_conf_kernel(const float *__restrict__ im1, const float *__restrict__ im2,
                                                        const float *__restrict__ flow, float *__restrict__ conf,
                                                        float thresh, int B, int C, int H, int W,
                                                        const float *__restrict__ lin_x, const float *__restrict__ lin_y,
                                                        float invx, float invy)
{
    const int c_n = CT > 0 ? CT : C;
    const size_t hw = (size_t)H * W;
    const size_t total = (size_t)B * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw;
        const int p = (int)(i - b * hw);
        const int y = p / W, x = p - y * W;
        const float dx = ldg_stream(flow + (b * 2) * hw + p);
        const float dy = ldg_stream(flow + (b * 2 + 1) * hw + p);
        const float *src = im2 + b * c_n * hw;
        const float *ref = im1 + b * c_n * hw + p;
        float xf, yf; Corners k; R2dWeights w; GsWeights gw;
        if (MODE == FLOWOPS_WARP_RESAMPLE2D) {
            r2d_coords(x, y, dx, dy, H, W, xf, yf, k);
            w = r2d_weights(xf, yf);
        } else {
            gw = gs_weights(gs_coords(x, y, dx, dy, H, W, lin_x, lin_y, invx, invy), H, W);
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < c_n; ++c) {
            const float *pl = src + (size_t)c * hw;
            const float val = MODE == FLOWOPS_WARP_RESAMPLE2D
                                  ? r2d_blend(w, __ldg(pl + k.o_tl), __ldg(pl + k.o_tr), __ldg(pl + k.o_bl), __ldg(pl + k.o_br))
                                  : gs_gather(gw, pl, W);
            const float d = __fsub_rn(ldg_stream(ref + (size_t)c * hw), val);
            // torch.sum(t*t, dim=1): products are rounded before they are added (flownet.py:56-57)
            acc = c == 0 ? __fmul_rn(d, d) : __fadd_rn(acc, __fmul_rn(d, d));
        }
        stg_stream(conf + i, acc < thresh ? 1.f : 0.f);
    }
}

static inline unsigned fused_grid(size_t total)
{
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_warp_diff_norm_fwd(const float *img0, const float *img1, size_t img_batch_stride,
                                          const float *flow, float *warped, size_t warped_batch_stride,
                                          float *norm, size_t norm_batch_stride,
                                          int B, int C, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(img0 && img1 && flow && norm, FLOWOPS_EINVAL, "warp_diff_norm_fwd: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "warp_diff_norm_fwd: bad shape %dx%dx%dx%d", B, C, H, W);
    const size_t hw = (size_t)H * W;
    FLOWOPS_REQUIRE(img_batch_stride >= (size_t)C * hw && norm_batch_stride >= hw &&
                    (!warped || warped_batch_stride >= (size_t)C * hw), FLOWOPS_EINVAL,
                    "warp_diff_norm_fwd: batch stride smaller than one item");
    FLOWOPS_REQUIRE((size_t)C * hw < (1ull << 31), FLOWOPS_EUNSUPPORTED, "warp_diff_norm_fwd: C*H*W exceeds int32 indexing");
    FLOWOPS_REQUIRE(H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED, "warp_diff_norm_fwd: H, W above 2^22 are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 3) {
        // row-pipelined gather (warp_rows.cuh) with the subtract + ChannelNorm epilogue
        WarpArgs a{};
        a.img = img1; a.img_bs = img_batch_stride; a.flow = flow; a.ref = img0; a.ref_bs = img_batch_stride;
        a.out = warped; a.out_bs = warped_batch_stride; a.aux = norm; a.aux_bs = norm_batch_stride;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W);
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        if (warp_impl_flags() & 2) {       // tolerance mode: fp32 bilinear weights (flowops_warp_set_impl)
            if (warped) launch_warp_rows<kWarpResample2dF32, EPI_DIFF_NORM, true>(a, st);
            else launch_warp_rows<kWarpResample2dF32, EPI_DIFF_NORM, false>(a, st);
        } else if (warped) launch_warp_rows<FLOWOPS_WARP_RESAMPLE2D, EPI_DIFF_NORM, true>(a, st);
        else launch_warp_rows<FLOWOPS_WARP_RESAMPLE2D, EPI_DIFF_NORM, false>(a, st);
        return check_launch("warp_diff_norm_fwd");
    }
    const unsigned grid = fused_grid((size_t)B * hw);
#define LAUNCH(CT)                                                                                              \
    do {                                                                                                        \
        if (warped) warp_diff_norm_kernel<CT, true><<<grid, 256, 0, st>>>(img0, img1, img_batch_stride, flow,  \
                        warped, warped_batch_stride, norm, norm_batch_stride, B, C, H, W);                     \
        else warp_diff_norm_kernel<CT, false><<<grid, 256, 0, st>>>(img0, img1, img_batch_stride, flow,        \
                        warped, warped_batch_stride, norm, norm_batch_stride, B, C, H, W);                     \
    } while (0)
    LAUNCH(0);
#undef LAUNCH
    return check_launch("warp_diff_norm_fwd");
}

extern "C" int flowops_warp_conf_fwd(const float *im1, const float *im2, const float *flow, float *conf,
                                     float thresh, int B, int C, int H, int W, int mode,
                                     const float *lin_x, const float *lin_y, void *stream)
{
    FLOWOPS_REQUIRE(im1 && im2 && flow && conf, FLOWOPS_EINVAL, "warp_conf_fwd: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "warp_conf_fwd: bad shape %dx%dx%dx%d", B, C, H, W);
    FLOWOPS_REQUIRE((size_t)C * H * W < (1ull << 31), FLOWOPS_EUNSUPPORTED, "warp_conf_fwd: C*H*W exceeds int32 indexing");
    FLOWOPS_REQUIRE(mode == FLOWOPS_WARP_RESAMPLE2D || mode == FLOWOPS_WARP_GRIDSAMPLE, FLOWOPS_EINVAL, "warp_conf_fwd: unknown mode %d", mode);
    FLOWOPS_REQUIRE(H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED, "warp_conf_fwd: H, W above 2^22 are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = fused_grid((size_t)B * H * W);
    float invx = 0.f, invy = 0.f, mulx, muly;
    if (mode == FLOWOPS_WARP_GRIDSAMPLE) {
        FLOWOPS_REQUIRE(lin_x && lin_y && H > 1 && W > 1, FLOWOPS_EINVAL, "warp_conf_fwd: GRIDSAMPLE mode needs the linspace tables and H, W > 1");
        gs_scales(H, W, invx, invy, mulx, muly);
    }
    if (C <= 3) {
        // row-pipelined gather (warp_rows.cuh) with the sum-of-squares threshold epilogue
        WarpArgs a{};
        const size_t chw = (size_t)C * H * W;
        a.img = im2; a.img_bs = chw; a.flow = flow; a.ref = im1; a.ref_bs = chw;
        a.aux = conf; a.aux_bs = (size_t)H * W;
        a.B = B; a.C = C; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W);
        a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
        a.lin_x = lin_x; a.lin_y = lin_y; a.invx = invx; a.invy = invy; a.thresh = thresh;
        if (mode == FLOWOPS_WARP_GRIDSAMPLE) launch_warp_rows<FLOWOPS_WARP_GRIDSAMPLE, EPI_CONF, false>(a, st);
        else if (warp_impl_flags() & 2) launch_warp_rows<kWarpResample2dF32, EPI_CONF, false>(a, st);
        else launch_warp_rows<FLOWOPS_WARP_RESAMPLE2D, EPI_CONF, false>(a, st);
        return check_launch("warp_conf_fwd");
    }
    if (mode == FLOWOPS_WARP_GRIDSAMPLE)
        warp_conf_kernel<FLOWOPS_WARP_GRIDSAMPLE, 0><<<grid, 256, 0, st>>>(im1, im2, flow, conf, thresh, B, C, H, W, lin_x, lin_y, invx, invy);
    else
        warp_conf_kernel<FLOWOPS_WARP_RESAMPLE2D, 0><<<grid, 256, 0, st>>>(im1, im2, flow, conf, thresh, B, C, H, W, nullptr, nullptr, 0.f, 0.f);
    return check_launch("warp_conf_fwd");
}

extern "C" int flowops_warp_diff_norm_concat_nhwc(const float *x, const float *flow, float div_flow, float *out, int c_dst,
                                                  int B, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(x && flow && out, FLOWOPS_EINVAL, "warp_diff_norm_concat_nhwc: null pointer");
    FLOWOPS_REQUIRE(B > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "warp_diff_norm_concat_nhwc: bad shape %dx%dx%d", B, H, W);
    FLOWOPS_REQUIRE(c_dst >= 12 && (c_dst & 3) == 0 && aligned16(out), FLOWOPS_EINVAL,
                    "warp_diff_norm_concat_nhwc: c_dst must be a multiple of 4 and at least 12, out 16-byte aligned");
    FLOWOPS_REQUIRE((size_t)6 * H * W < (1ull << 31) && H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED,
                    "warp_diff_norm_concat_nhwc: frame too large for int32 indexing");
    const size_t hw = (size_t)H * W;
    WarpArgs a{};
    a.img = x + 3 * hw; a.img_bs = 6 * hw; a.flow = flow; a.ref = x; a.ref_bs = 6 * hw;
    a.aux = out; a.aux_bs = hw * c_dst; a.c_dst = c_dst; a.inv_div_flow = 1.0f / div_flow;
    a.B = B; a.C = 3; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W);
    a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
    if (warp_impl_flags() & 2) launch_warp_rows<kWarpResample2dF32, EPI_CONCAT, false>(a, (cudaStream_t)stream);
    else launch_warp_rows<FLOWOPS_WARP_RESAMPLE2D, EPI_CONCAT, false>(a, (cudaStream_t)stream);
    return check_launch("warp_diff_norm_concat_nhwc");
}

// The same concat with the x4 bilinear upsampling and the scaling of the previous sub-network's flow folded in
// (models.py:106,118: upsample(flownet(x)[0] * div_flow)): the kernel forms the full-resolution flow from the
// quarter-resolution field while it walks the rows, so the full-resolution flow tensor is never written or read.
extern "C" int flowops_warp_diff_norm_concat_up4_nhwc(const float *x, const float *flow_lo, float flow_mul, float div_flow,
                                                      float *out, int c_dst, int B, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(x && flow_lo && out, FLOWOPS_EINVAL, "warp_diff_norm_concat_up4_nhwc: null pointer");
    FLOWOPS_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 3) == 0 && (W & 3) == 0, FLOWOPS_EINVAL,
                    "warp_diff_norm_concat_up4_nhwc: bad shape %dx%dx%d (H, W must be multiples of 4)", B, H, W);
    FLOWOPS_REQUIRE(c_dst >= 12 && (c_dst & 3) == 0 && aligned16(out), FLOWOPS_EINVAL,
                    "warp_diff_norm_concat_up4_nhwc: c_dst must be a multiple of 4 and at least 12, out 16-byte aligned");
    FLOWOPS_REQUIRE((size_t)6 * H * W < (1ull << 31) && H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED,
                    "warp_diff_norm_concat_up4_nhwc: frame too large for int32 indexing");
    const size_t hw = (size_t)H * W;
    WarpArgs a{};
    a.img = x + 3 * hw; a.img_bs = 6 * hw; a.flow = nullptr; a.ref = x; a.ref_bs = 6 * hw;
    a.flow_lo = flow_lo; a.flow_mul = flow_mul;
    a.aux = out; a.aux_bs = hw * c_dst; a.c_dst = c_dst; a.inv_div_flow = 1.0f / div_flow;
    a.B = B; a.C = 3; a.H = H; a.W = W; a.rows = warp_rows_pick(B, H, W);
    a.wm1 = (float)(W - 1); a.hm1 = (float)(H - 1);
    if (warp_impl_flags() & 2) launch_warp_rows<kWarpResample2dF32, EPI_CONCAT, false, true>(a, (cudaStream_t)stream);
    else launch_warp_rows<FLOWOPS_WARP_RESAMPLE2D, EPI_CONCAT, false, true>(a, (cudaStream_t)stream);
    return check_launch("warp_diff_norm_concat_up4_nhwc");
}

// ---------------------------------------------------------------------------------------------
// Input of the fusion network (models.py:129-152), one pass, channels-last:
//   flow_s2 = upsample4(flow2_s2 * div_flow)      flow_sd = upsample3(flow2_sd / div_flow)        (nearest, x4)
//   concat3 = (frame 0, flow_sd, flow_s2, ChannelNorm(flow_sd), ChannelNorm(flow_s2),
//              ChannelNorm(frame 0 - Resample2d(frame 1, flow_sd)), ChannelNorm(frame 0 - Resample2d(frame 1, flow_s2)))
// The reference runs 2 scalings, 2 upsamplings, 2 ChannelNorms, 2 Resample2d + subtract + ChannelNorm chains and a
// torch.cat, and cuDNN then converts the 11-channel NCHW result to padded NHWC; here a thread walks `rows` rows of one
// column, reads the two quarter-resolution flows (L1/L2-resident), gathers frame 1 twice and writes the pixel's 11
// values plus zero padding as float4s.  Same arithmetic, operation for operation (pix_prep / pix_blend of
// warp_rows.cuh; FFMA chain + IEEE sqrt of cnorm.cu), so the values equal the operator chain bit for bit.
// ---------------------------------------------------------------------------------------------
namespace flowops {

struct FusionInputArgs {
    WarpArgs g;                          // geometry (B, H, W, rows, wm1, hm1); tensor pointers unused
    const float *x;                      // [B,6,H,W] planar frame stack
    const float *lo_s2, *lo_sd;          // [B,2,H/4,W/4] flow2 outputs of FlowNetS2 / FlowNetSD (network units)
    float mul_s2, mul_sd;                // div_flow and the fp32 reciprocal of div_flow (models.py:130,143)
    float *out;                          // [B,H,W,c_dst]
    int c_dst;
};

#ifndef FLOWOPS_TUNE_FUSION_MINBLOCKS       // variant builds: python -m ir2rgb_b200.build --out ... -DFLOWOPS_TUNE_FUSION_MINBLOCKS=4
#define FLOWOPS_TUNE_FUSION_MINBLOCKS 3     // 80 registers, 3 CTAs per SM (ncu: latency-bound at 34 % occupancy)
#endif
template <int MODE>      // FLOWOPS_WARP_RESAMPLE2D (the reference's mixed fp64 / fp32 blend, bit-exact) or kWarpResample2dF32
__global__ void __launch_bounds__(256, FLOWOPS_TUNE_FUSION_MINBLOCKS) fusion_input_kernel(const __grid_constant__ FusionInputArgs a)
{
    const WarpArgs &g = a.g;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * g.rows;
    if (x >= g.W || y0 >= g.H) return;
    const int y1 = min(y0 + g.rows, g.H);
    const unsigned hw = (unsigned)g.H * g.W, W = (unsigned)g.W;
    const unsigned Wl = W >> 2, hwl = (unsigned)(g.H >> 2) * Wl;
    const size_t b = blockIdx.z;
    const float *frame0 = a.x + b * 6 * hw;
    const float *src = frame0 + 3 * (size_t)hw;                       // frame 1: the gather source
    asm("" : "+l"(src));
    const float *ls2 = a.lo_s2 + b * 2 * hwl + (unsigned)(x >> 2);
    const float *lsd = a.lo_sd + b * 2 * hwl + (unsigned)(x >> 2);
    const float xfl = small_int_as_float(x);
    for (int y = y0; y < y1; ++y) {
        const unsigned p = (unsigned)y * W + (unsigned)x, pl = (unsigned)(y >> 2) * Wl;
        // nearest x4 upsampling: source index floor(dst * 0.25) (ATen upsample_nearest2d with scale_factor 4)
        const float s2x = __fmul_rn(__ldg(ls2 + pl), a.mul_s2), s2y = __fmul_rn(__ldg(ls2 + pl + hwl), a.mul_s2);
        const float sdx = __fmul_rn(__ldg(lsd + pl), a.mul_sd), sdy = __fmul_rn(__ldg(lsd + pl + hwl), a.mul_sd);
        const float yfl = small_int_as_float(y);
        PixPrep<MODE> qd, q2;
        pix_prep(qd, g, xfl, yfl, x, y, sdx, sdy);
        pix_prep(q2, g, xfl, yfl, x, y, s2x, s2y);
        PixVals<3> vd, v2;
        pix_gather<3, true, false>(vd, qd, src, frame0 + p, nullptr, hw);      // also loads frame 0 at this pixel (vd.r)
        pix_gather<3, false, false>(v2, q2, src, nullptr, nullptr, hw);
        float ed = 0.f, e2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float dd = __fsub_rn(vd.r[c], pix_blend(qd, vd.v[c]));       // models.py:110
            const float d2 = __fsub_rn(vd.r[c], pix_blend(q2, v2.v[c]));
            ed = __fmaf_rn(dd, dd, ed);                                        // channelnorm_kernel.cu:55-56
            e2 = __fmaf_rn(d2, d2, e2);
        }
        const float nd = __fsqrt_rn(__fmaf_rn(sdy, sdy, __fmaf_rn(sdx, sdx, 0.f)));
        const float n2 = __fsqrt_rn(__fmaf_rn(s2y, s2y, __fmaf_rn(s2x, s2x, 0.f)));
        float4 *dst = reinterpret_cast<float4 *>(a.out + (b * hw + p) * (size_t)a.c_dst);
        dst[0] = make_float4(vd.r[0], vd.r[1], vd.r[2], sdx);
        dst[1] = make_float4(sdy, s2x, s2y, nd);
        dst[2] = make_float4(n2, __fsqrt_rn(ed), __fsqrt_rn(e2), 0.f);
        for (int q4 = 3; q4 < a.c_dst / 4; ++q4) dst[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace flowops

extern "C" int flowops_flownet2_fusion_input_nhwc(const float *x, const float *flow2_s2, const float *flow2_sd, float div_flow,
                                                  float *out, int c_dst, int B, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(x && flow2_s2 && flow2_sd && out, FLOWOPS_EINVAL, "flownet2_fusion_input_nhwc: null pointer");
    FLOWOPS_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 3) == 0 && (W & 3) == 0, FLOWOPS_EINVAL,
                    "flownet2_fusion_input_nhwc: bad shape %dx%dx%d (H and W must be multiples of 4)", B, H, W);
    FLOWOPS_REQUIRE(c_dst >= 12 && (c_dst & 3) == 0 && aligned16(out), FLOWOPS_EINVAL,
                    "flownet2_fusion_input_nhwc: c_dst must be a multiple of 4 and at least 12, out 16-byte aligned");
    FLOWOPS_REQUIRE((size_t)6 * H * W < (1ull << 31) && H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED,
                    "flownet2_fusion_input_nhwc: frame too large for int32 indexing");
    FusionInputArgs a{};
    a.x = x; a.lo_s2 = flow2_s2; a.lo_sd = flow2_sd; a.mul_s2 = div_flow; a.mul_sd = 1.0f / div_flow;
    a.out = out; a.c_dst = c_dst;
    a.g.B = B; a.g.C = 3; a.g.H = H; a.g.W = W; a.g.rows = warp_rows_pick(B, H, W);
    a.g.wm1 = (float)(W - 1); a.g.hm1 = (float)(H - 1);
    const size_t hw = (size_t)H * W;
    for (int b0 = 0; b0 < B; b0 += 65535) {
        FusionInputArgs c = a;
        c.g.B = B - b0 < 65535 ? B - b0 : 65535;
        c.x = x + (size_t)b0 * 6 * hw; c.out = out + (size_t)b0 * hw * c_dst;
        c.lo_s2 = flow2_s2 + (size_t)b0 * 2 * (hw / 16); c.lo_sd = flow2_sd + (size_t)b0 * 2 * (hw / 16);
        dim3 grid, block;
        warp_rows_shape(c.g.B, H, W, c.g.rows, grid, block);
        if (warp_impl_flags() & 2) fusion_input_kernel<kWarpResample2dF32><<<grid, block, 0, (cudaStream_t)stream>>>(c);
        else fusion_input_kernel<FLOWOPS_WARP_RESAMPLE2D><<<grid, block, 0, (cudaStream_t)stream>>>(c);
    }
    return check_launch("flownet2_fusion_input_nhwc");
}

// ---------------------------------------------------------------------------------------------
// FlowNet2 input preparation (models.py:97-101): x = (inputs - rgb_mean) / rgb_max, frames stacked along
// channels -- written once in every layout its consumers want: planar [B,6,H,W] for the warp kernels, and
// channels-last copies padded to 4 / 4 / 8 channels for the first convolutions of FlowNetC (per frame) and
// FlowNetSD (both frames), which otherwise each trigger a slice copy, a layout conversion and cuDNN's own
// channel padding.
// ---------------------------------------------------------------------------------------------
namespace flowops {

// S2D: xa / xb are written "space to depth" instead of as 4-channel frames -- [B, H/2 + 1, W/2 + 1, 16], the four pixels
// of a 2 x 2 block side by side (channel (py*2+px)*4 + c), shifted down and right by one block (block row 0 and block
// column 0 are a zero border the caller provides).  A 7 x 7 stride-2 convolution of the frame is then a 4 x 4 stride-1
// convolution of this tensor with padding 1 (see FlowNetC.conv1_s2d): 16 input channels instead of 4 puts FlowNetC's
// first layer on cuDNN's tensor-op kernels.
template <bool S2D>
__global__ void __launch_bounds__(256) flownet2_prep_kernel(const float *__restrict__ in, const float *__restrict__ mean, float inv_rgb_max,
                                                            float *__restrict__ xp, float4 *__restrict__ xa, float4 *__restrict__ xb,
                                                            float4 *__restrict__ x8, unsigned x8_quads, unsigned hw, size_t total, unsigned W)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw;
        const unsigned p = (unsigned)(i - b * hw);
        float v[6];                                          // v[f*3 + c]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float m = __ldg(mean + b * 3 + c);
#pragma unroll
            for (int f = 0; f < 2; ++f)                      // inputs[b][c][f][p]
                v[f * 3 + c] = __fmul_rn(__fsub_rn(ldg_stream(in + ((b * 3 + c) * 2 + f) * hw + p), m), inv_rgb_max);
        }
        if (xp) {
#pragma unroll
            for (int k = 0; k < 6; ++k) xp[(b * 6 + k) * hw + p] = v[k];
        }
        size_t ia = i;
        if (S2D) {
            const unsigned y = p / W, x = p - y * W, W2 = (W >> 1) + 1, H2 = (hw / W >> 1) + 1;
            ia = ((b * H2 + (y >> 1) + 1) * W2 + (x >> 1) + 1) * 4 + (y & 1) * 2 + (x & 1);
        }
        if (xa) xa[ia] = make_float4(v[0], v[1], v[2], 0.f);
        if (xb) xb[ia] = make_float4(v[3], v[4], v[5], 0.f);
        // x8_quads float4s per pixel: 2 = a dense 8-channel tensor; more = the first 8 channels of a wider, pre-zeroed one
        if (x8) { x8[x8_quads * i] = make_float4(v[0], v[1], v[2], v[3]); x8[x8_quads * i + 1] = make_float4(v[4], v[5], 0.f, 0.f); }
    }
}

}  // namespace flowops

static int flownet2_prep_launch(const char *what, bool s2d, const float *inputs, const float *rgb_mean, float rgb_max,
                                float *x_planar, float *xa, float *xb, float *x_packed, int packed_channels,
                                int B, int H, int W, void *stream)
{
    FLOWOPS_REQUIRE(inputs && rgb_mean, FLOWOPS_EINVAL, "%s: null pointer", what);
    FLOWOPS_REQUIRE(x_planar || xa || xb || x_packed, FLOWOPS_EINVAL, "%s: no output requested", what);
    FLOWOPS_REQUIRE(!s2d || (xa && xb), FLOWOPS_EINVAL, "%s: null pointer", what);
    FLOWOPS_REQUIRE(B > 0 && H > 0 && W > 0 && (size_t)H * W < (1ull << 31), FLOWOPS_EINVAL, "%s: bad shape %dx%dx%d", what, B, H, W);
    FLOWOPS_REQUIRE(!s2d || ((H & 1) == 0 && (W & 1) == 0), FLOWOPS_EINVAL, "%s: bad shape %dx%dx%d (H, W must be even)", what, B, H, W);
    FLOWOPS_REQUIRE(packed_channels >= 8 && packed_channels % 4 == 0, FLOWOPS_EINVAL, "%s: packed_channels %d (a multiple of 4, >= 8)",
                    what, packed_channels);
    FLOWOPS_REQUIRE(aligned16(xa) && aligned16(xb) && aligned16(x_packed), FLOWOPS_EINVAL, "%s: outputs must be 16-byte aligned", what);
    const size_t total = (size_t)B * H * W;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    // tensor / python-scalar on CUDA is tensor * (1.0f / scalar) in ATen; reproduced so that x matches models.py:98 bit for bit
    if (s2d)
        flownet2_prep_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(inputs, rgb_mean, 1.0f / rgb_max, x_planar,
            reinterpret_cast<float4 *>(xa), reinterpret_cast<float4 *>(xb), reinterpret_cast<float4 *>(x_packed),
            (unsigned)packed_channels / 4, (unsigned)((size_t)H * W), total, (unsigned)W);
    else
        flownet2_prep_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(inputs, rgb_mean, 1.0f / rgb_max, x_planar,
            reinterpret_cast<float4 *>(xa), reinterpret_cast<float4 *>(xb), reinterpret_cast<float4 *>(x_packed),
            (unsigned)packed_channels / 4, (unsigned)((size_t)H * W), total, (unsigned)W);
    return check_launch(what);
}

extern "C" int flowops_flownet2_prep(const float *inputs, const float *rgb_mean, float rgb_max,
                                     float *x_planar, float *xa_nhwc4, float *xb_nhwc4, float *x_nhwc8,
                                     int B, int H, int W, void *stream)
{
    return flownet2_prep_launch("flownet2_prep", false, inputs, rgb_mean, rgb_max, x_planar, xa_nhwc4, xb_nhwc4, x_nhwc8, 8, B, H, W, stream);
}

extern "C" int flowops_flownet2_prep_s2d(const float *inputs, const float *rgb_mean, float rgb_max,
                                         float *x_planar, float *xa_s2d, float *xb_s2d, float *x_nhwc8,
                                         int B, int H, int W, void *stream)
{
    return flownet2_prep_launch("flownet2_prep_s2d", true, inputs, rgb_mean, rgb_max, x_planar, xa_s2d, xb_s2d, x_nhwc8, 8, B, H, W, stream);
}

extern "C" int flowops_flownet2_prep_pitched(const float *inputs, const float *rgb_mean, float rgb_max,
                                             float *x_planar, float *xa, float *xb, float *x_packed, int packed_channels, int s2d,
                                             int B, int H, int W, void *stream)
{
    return flownet2_prep_launch("flownet2_prep_pitched", s2d != 0, inputs, rgb_mean, rgb_max, x_planar, xa, xb, x_packed, packed_channels,
                                B, H, W, stream);
}
