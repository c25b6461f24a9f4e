// corr_generic.cu -- Correlation forward/backward for ANY (pad, k, md, s1, s2): the catch-all
// behind the FlowNetC fast path (corr_fast.cu).  Plain gather kernels, one thread per output
// element, x fastest (coalesced), inputs read in place from NCHW -- no padded NHWC scratch copies
// (reference correlation_cuda_kernel.cu:47-70) and no per-batch-item launches (:522-554).
//
// Semantics follow the reference kernels exactly (correlation_cuda_kernel.cu:74-334), including
// the C truncating divisions by stride1 in the backward; reads that fall outside the reference's
// padded scratch array (only possible for kernel_size > 1, where the reference reads out of
// bounds) are defined as zero.
#include "common.cuh"
#include "corr.cuh"

namespace flowops {

__device__ __forceinline__ float padded_at(const float *__restrict__ in, size_t img_base, int ch, int H, int W,
                                           int pad, int yy, int xx)
{
    const int y = yy - pad, x = xx - pad;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
    return __ldg(in + img_base + ((size_t)ch * H + y) * W + x);
}

__global__ void __launch_bounds__(256) corr_fwd_generic(const float *__restrict__ in1, const float *__restrict__ in2,
                                                        float *__restrict__ out, CorrGeom g, size_t total)
{
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % g.oW);
        const int oy = (int)((idx / g.oW) % g.oH);
        const int tc = (int)((idx / ((size_t)g.oW * g.oH)) % g.oC);
        const int n = (int)(idx / ((size_t)g.oW * g.oH * g.oC));
        const int tj = tc / g.D - g.dr, ti = tc % g.D - g.dr;
        const int y1 = oy * g.s1 + g.md, x1 = ox * g.s1 + g.md;
        const int y2 = y1 + tj * g.s2, x2 = x1 + ti * g.s2;
        const size_t base = (size_t)n * g.C * g.H * g.W;
        float acc = 0.f;
        for (int j = -g.kr; j <= g.kr; ++j)
            for (int i = -g.kr; i <= g.kr; ++i)
                for (int ch = 0; ch < g.C; ++ch)
                    acc = __fmaf_rn(padded_at(in1, base, ch, g.H, g.W, g.pad, y1 + j, x1 + i),
                                    padded_at(in2, base, ch, g.H, g.W, g.pad, y2 + j, x2 + i), acc);
        out[idx] = acc / (float)(g.k * g.k * g.C);
    }
}

// gin1 / gin2 in one launch: blockIdx.y selects which.
__global__ void __launch_bounds__(256) corr_bwd_generic(const float *__restrict__ in1, const float *__restrict__ in2,
                                                        const float *__restrict__ gout,
                                                        float *__restrict__ gin1, float *__restrict__ gin2,
                                                        CorrGeom g, size_t total)
{
    const bool second = blockIdx.y == 1;
    float *dst = second ? gin2 : gin1;
    if (!dst) return;
    const float nelems = (float)(g.k * g.k * g.C);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int bx = (int)(idx % g.W);
        const int by = (int)((idx / g.W) % g.H);
        const int c = (int)((idx / ((size_t)g.W * g.H)) % g.C);
        const int n = (int)(idx / ((size_t)g.W * g.H * g.C));
        const int y = by * g.s1 + g.pad, x = bx * g.s1 + g.pad;
        const size_t base = (size_t)n * g.C * g.H * g.W;
        const float *go = gout + (size_t)n * g.oC * g.oH * g.oW;
        float sum = 0.f;
        if (!second) {
            int xmin = (x - g.kr - g.md) / g.s1, ymin = (y - g.kr - g.md) / g.s1;
            int xmax = (x + g.kr - g.md) / g.s1, ymax = (y + g.kr - g.md) / g.s1;
            const bool skip = xmax < 0 || ymax < 0 || xmin >= g.oW || ymin >= g.oH || xmin > xmax || ymin > ymax;
            if (!skip) {
                xmin = max(0, xmin); xmax = min(g.oW - 1, xmax);
                ymin = max(0, ymin); ymax = min(g.oH - 1, ymax);
                for (int tc = 0; tc < g.oC; ++tc) {
                    const int i2 = (tc % g.D - g.dr) * g.s2, j2 = (tc / g.D - g.dr) * g.s2;
                    const float v2 = padded_at(in2, base, c, g.H, g.W, g.pad, y + j2, x + i2);
                    for (int j = ymin; j <= ymax; ++j)
                        for (int i = xmin; i <= xmax; ++i)
                            sum = __fmaf_rn(__ldg(go + ((size_t)tc * g.oH + j) * g.oW + i), v2, sum);
                }
            }
        } else {
            for (int tc = 0; tc < g.oC; ++tc) {
                const int i2 = (tc % g.D - g.dr) * g.s2, j2 = (tc / g.D - g.dr) * g.s2;
                int xmin = (x - g.kr - g.md - i2) / g.s1, ymin = (y - g.kr - g.md - j2) / g.s1;
                int xmax = (x + g.kr - g.md - i2) / g.s1, ymax = (y + g.kr - g.md - j2) / g.s1;
                if (xmax < 0 || ymax < 0 || xmin >= g.oW || ymin >= g.oH) continue;
                if (xmin > xmax || ymin > ymax) continue;
                xmin = max(0, xmin); xmax = min(g.oW - 1, xmax);
                ymin = max(0, ymin); ymax = min(g.oH - 1, ymax);
                const float v1 = padded_at(in1, base, c, g.H, g.W, g.pad, y - j2, x - i2);
                for (int j = ymin; j <= ymax; ++j)
                    for (int i = xmin; i <= xmax; ++i)
                        sum = __fmaf_rn(__ldg(go + ((size_t)tc * g.oH + j) * g.oW + i), v1, sum);
            }
        }
        dst[idx] = sum / nelems;
    }
}

int corr_fwd_generic_launch(const float *in1, const float *in2, float *out, const CorrGeom &g, cudaStream_t st)
{
    const size_t total = (size_t)g.B * g.oC * g.oH * g.oW;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)kNumSMs * 8 * 32) blocks = (size_t)kNumSMs * 8 * 32;
    corr_fwd_generic<<<(unsigned)blocks, 256, 0, st>>>(in1, in2, out, g, total);
    return check_launch("corr_fwd(generic)");
}

int corr_bwd_generic_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                            const CorrGeom &g, cudaStream_t st)
{
    const size_t total = (size_t)g.B * g.C * g.H * g.W;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)kNumSMs * 8 * 32) blocks = (size_t)kNumSMs * 8 * 32;
    corr_bwd_generic<<<dim3((unsigned)blocks, 2), 256, 0, st>>>(in1, in2, gout, gin1, gin2, g, total);
    return check_launch("corr_bwd(generic)");
}

}  // namespace flowops
