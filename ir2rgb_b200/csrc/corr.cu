// corr.cu -- C-ABI entry points of the Correlation operator and the fast/generic dispatch.
#include <math.h>

#include "corr.cuh"

namespace flowops {

int corr_geometry(CorrGeom &g, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "corr: bad shape %dx%dx%dx%d", B, C, H, W);
    FLOWOPS_REQUIRE(k >= 1 && (k & 1) && s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0, FLOWOPS_EINVAL,
                    "corr: bad parameters pad=%d k=%d md=%d s1=%d s2=%d", pad, k, md, s1, s2);
    g.B = B; g.C = C; g.H = H; g.W = W;
    g.pad = pad; g.k = k; g.md = md; g.s1 = s1; g.s2 = s2;
    g.kr = (k - 1) / 2;
    g.dr = md / s2;
    g.D = 2 * g.dr + 1;
    const int br = g.kr + md;
    g.oC = g.D * g.D;
    // correlation_cuda.cc:31-32 (float ceil of a float quotient)
    g.oH = (int)ceilf((float)(H + 2 * pad - 2 * br) / (float)s1);
    g.oW = (int)ceilf((float)(W + 2 * pad - 2 * br) / (float)s1);
    FLOWOPS_REQUIRE(g.oH > 0 && g.oW > 0, FLOWOPS_EINVAL, "corr: empty output (%d x %d)", g.oH, g.oW);
    FLOWOPS_REQUIRE((size_t)B * g.oC * g.oH * g.oW < (1ull << 40) && (size_t)C * H * W < (1ull << 31),
                    FLOWOPS_EUNSUPPORTED, "corr: tensor too large");
    return 0;
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_corr_out_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                                      int *oC, int *oH, int *oW)
{
    CorrGeom g;
    const int rc = corr_geometry(g, 1, 1, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    if (oC) *oC = g.oC;
    if (oH) *oH = g.oH;
    if (oW) *oW = g.oW;
    return 0;
}

extern "C" size_t flowops_corr_fwd_workspace_bytes(int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    CorrGeom g;
    if (corr_geometry(g, B, C, H, W, pad, k, md, s1, s2)) return 0;
    if (!corr_fast_supported(g)) return 0;
    size_t n = corr_fast_fwd_workspace(g);             // also what the 16-bit entry point uses
    if (corr_tc_supported(g)) { const size_t t = corr_tc_fwd_workspace(g, true); if (t > n) n = t; }
    return n;
}

extern "C" size_t flowops_corr_bwd_workspace_bytes(int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    CorrGeom g;
    if (corr_geometry(g, B, C, H, W, pad, k, md, s1, s2)) return 0;
    if (!corr_fast_supported(g)) return 0;
    size_t n = corr_fast_bwd_workspace(g);
    if (corr_tc_bwd_supported(g)) { const size_t t = corr_tc_bwd_workspace(g); if (t > n) n = t; }
    return n;
}

extern "C" int flowops_corr_fwd(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                                int pad, int k, int md, int s1, int s2, int in_layout,
                                void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(in_layout == FLOWOPS_LAYOUT_NCHW || in_layout == FLOWOPS_LAYOUT_NHWC, FLOWOPS_EINVAL,
                    "corr_fwd: unknown input layout %d", in_layout);
    FLOWOPS_REQUIRE(in1 && in2 && out, FLOWOPS_EINVAL, "corr_fwd: null pointer");
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    FLOWOPS_REQUIRE(pad >= md + g.kr || k == 1, FLOWOPS_EUNSUPPORTED,
                    "corr_fwd: pad_size < max_displacement + kernel_radius reads outside the padded scratch in the reference");
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tc_supported(g)) {                         // tcgen05 path: layout pass -> UMMA kernel -> NCHW store pass
        int r2 = in_layout == FLOWOPS_LAYOUT_NCHW ? corr_tc_planes_nchw(in1, in2, g, workspace, workspace_bytes, st)
                                                  : corr_tc_planes_nhwc(in1, 0, g, nullptr, 1.f, nullptr, workspace, workspace_bytes, st);
        if (!r2 && in_layout == FLOWOPS_LAYOUT_NHWC)
            r2 = corr_tc_planes_nhwc(in2, 1, g, nullptr, 1.f, nullptr, workspace, workspace_bytes, st);
        if (r2) return r2;
        return corr_tc_main(out, g, workspace, workspace_bytes, st, true, 0, 0, 1.f);
    }
    if (corr_fast_supported(g)) return corr_fast_fwd_launch(in1, in2, out, g, in_layout, workspace, workspace_bytes, st);
    FLOWOPS_REQUIRE(in_layout == FLOWOPS_LAYOUT_NCHW, FLOWOPS_EUNSUPPORTED,
                    "corr_fwd: channels-last inputs are only taken by the FlowNetC configuration");
    return corr_fwd_generic_launch(in1, in2, out, g, st);
}

extern "C" int flowops_corr_fwd_16(const void *in1, const void *in2, void *out, int B, int C, int H, int W,
                                   int pad, int k, int md, int s1, int s2, int dtype,
                                   void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(in1 && in2 && out, FLOWOPS_EINVAL, "corr_fwd_16: null pointer");
    FLOWOPS_REQUIRE(dtype == FLOWOPS_DTYPE_F16 || dtype == FLOWOPS_DTYPE_BF16, FLOWOPS_EINVAL,
                    "corr_fwd_16: dtype must be FLOWOPS_DTYPE_F16 or FLOWOPS_DTYPE_BF16, got %d", dtype);
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    FLOWOPS_REQUIRE(corr_fast_supported(g), FLOWOPS_EUNSUPPORTED,
                    "corr_fwd_16: FlowNetC configuration only (cast and call flowops_corr_fwd for other parameters)");
    return corr_fast_fwd_launch(static_cast<const float *>(in1), static_cast<const float *>(in2), static_cast<float *>(out),
                                g, FLOWOPS_LAYOUT_NCHW, workspace, workspace_bytes, (cudaStream_t)stream, dtype);
}

extern "C" int flowops_corr_planes_from_conv(const float *y, const float *bias, float slope, float *act, int which,
                                             int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                             void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(y && bias, FLOWOPS_EINVAL, "corr_planes_from_conv: null pointer");
    FLOWOPS_REQUIRE(which == 0 || which == 1, FLOWOPS_EINVAL, "corr_planes_from_conv: input slot must be 0 or 1");
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    FLOWOPS_REQUIRE(corr_fast_supported(g), FLOWOPS_EUNSUPPORTED, "corr_planes_from_conv: FlowNetC configuration only");
    if (corr_tc_supported(g))
        return corr_tc_planes_nhwc(y, which, g, bias, slope, act, workspace, workspace_bytes, (cudaStream_t)stream);
    return corr_fast_planes_nhwc(which == 0 ? y : nullptr, which == 1 ? y : nullptr, g, which, bias, slope, act,
                                 workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int flowops_corr_fwd_planes(float *out, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                       void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(out, FLOWOPS_EINVAL, "corr_fwd_planes: null pointer");
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    FLOWOPS_REQUIRE(corr_fast_supported(g), FLOWOPS_EUNSUPPORTED, "corr_fwd_planes: FlowNetC configuration only");
    if (corr_tc_supported(g)) return corr_tc_main(out, g, workspace, workspace_bytes, (cudaStream_t)stream, true, 0, 0, 1.f);
    return corr_fast_main(out, g, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int flowops_corr_bwd(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                                int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(in1 && in2 && gout, FLOWOPS_EINVAL, "corr_bwd: null pointer");
    FLOWOPS_REQUIRE(gin1 || gin2, FLOWOPS_EINVAL, "corr_bwd: both gradient outputs are null");
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    // the reference backward indexes gradInput with blockIdx * stride1 and writes out of bounds for
    // stride1 > 1 (correlation_cuda_kernel.cu:164-165,238); it is never used that way.
    FLOWOPS_REQUIRE(s1 == 1, FLOWOPS_EUNSUPPORTED, "corr_bwd: stride1 != 1 is not defined by the reference backward");
    cudaStream_t st = (cudaStream_t)stream;
    if (corr_tc_bwd_supported(g)) return corr_tc_bwd_launch(in1, in2, gout, gin1, gin2, g, workspace, workspace_bytes, st);
    if (corr_fast_supported(g)) return corr_fast_bwd_launch(in1, in2, gout, gin1, gin2, g, workspace, workspace_bytes, st);
    return corr_bwd_generic_launch(in1, in2, gout, gin1, gin2, g, st);
}

extern "C" int flowops_corr_fwd_planes_nhwc(float *out, int c_dst, int c_off, float lrelu_slope,
                                            int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                            void *workspace, size_t workspace_bytes, void *stream)
{
    FLOWOPS_REQUIRE(out, FLOWOPS_EINVAL, "corr_fwd_planes_nhwc: null pointer");
    CorrGeom g;
    const int rc = corr_geometry(g, B, C, H, W, pad, k, md, s1, s2);
    if (rc) return rc;
    FLOWOPS_REQUIRE(corr_fast_supported(g), FLOWOPS_EUNSUPPORTED, "corr_fwd_planes_nhwc: FlowNetC configuration only");
    FLOWOPS_REQUIRE(c_off >= 0 && c_off + g.D * g.D <= c_dst, FLOWOPS_EINVAL,
                    "corr_fwd_planes_nhwc: channels [%d, %d) do not fit in %d", c_off, c_off + g.D * g.D, c_dst);
    if (corr_tc_supported(g))
        return corr_tc_main(out, g, workspace, workspace_bytes, (cudaStream_t)stream, false, c_dst, c_off, lrelu_slope);
    return corr_fast_main(out, g, workspace, workspace_bytes, (cudaStream_t)stream, true, c_dst, c_off, lrelu_slope);
}
