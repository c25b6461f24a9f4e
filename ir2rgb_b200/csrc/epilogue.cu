// epilogue.cu -- conv-body glue for FlowNet2 inference: per-channel bias add
// + LeakyReLU in ONE in-place pass over the convolution output.
//
// The stock path (reference submodules.py:7-38 -> cuDNN) runs three kernels per conv layer: the
// convolution, an elementwise bias add and an in-place LeakyReLU -- 19 % + 7 % of a channels_last FlowNet2
// forward on B200 (profiles/torchprof_flownet_cl_r01.txt).  This kernel does `t = y + b[c];
// y = t > 0 ? t : t * slope` with 128-bit accesses, bit-identical to the two torch kernels it replaces.
// HBM-bound: 8 bytes per element (one read, one write).
#include "common.cuh"

namespace flowops {

__device__ __forceinline__ float lrelu(float t, float slope) { return t > 0.f ? t : __fmul_rn(t, slope); }

// channels-last: y[n][hw][c], C % 4 == 0
__global__ void __launch_bounds__(256) bias_lrelu_nhwc(float *__restrict__ y, const float *__restrict__ bias,
                                                       size_t total4, unsigned c4n, float slope)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned c4 = (unsigned)(i % c4n);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        float4 v = *reinterpret_cast<float4 *>(y + i * 4);
        v.x = lrelu(__fadd_rn(v.x, b.x), slope); v.y = lrelu(__fadd_rn(v.y, b.y), slope);
        v.z = lrelu(__fadd_rn(v.z, b.z), slope); v.w = lrelu(__fadd_rn(v.w, b.w), slope);
        *reinterpret_cast<float4 *>(y + i * 4) = v;
    }
}

// NCHW: y[n][c][hw], HW % 4 == 0
__global__ void __launch_bounds__(256) bias_lrelu_nchw(float *__restrict__ y, const float *__restrict__ bias,
                                                       size_t total4, unsigned hw4, unsigned C, float slope)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const float b = __ldg(bias + (unsigned)((i / hw4) % C));
        float4 v = *reinterpret_cast<float4 *>(y + i * 4);
        v.x = lrelu(__fadd_rn(v.x, b), slope); v.y = lrelu(__fadd_rn(v.y, b), slope);
        v.z = lrelu(__fadd_rn(v.z, b), slope); v.w = lrelu(__fadd_rn(v.w, b), slope);
        *reinterpret_cast<float4 *>(y + i * 4) = v;
    }
}

// channels-last, out of place into a channel slice of a wider pixel: dst[pix][c_off + c] = lrelu(y[pix][c] + b[c]).
// The epilogue of a decoder deconvolution writes straight into the concat buffer the next layers read
// (reference FlowNetS.py:74-76: torch.cat((out_conv5, out_deconv5, flow6_up), 1)), which saves the concat's own
// read + write of that tensor.  V = float4 when every channel count / offset is a multiple of 4, else float.
// `also` (nullable, may alias y): a dense copy of the result as well -- an encoder layer whose output is both the next
// layer's input and a decoder skip connection is written to both places from one read.
template <typename V>
__global__ void __launch_bounds__(256) bias_lrelu_nhwc_to(const V *y, const V *__restrict__ bias, V *__restrict__ dst,
                                                          size_t total, unsigned src_v, unsigned dst_v, unsigned off_v, float slope,
                                                          V *also)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / src_v;
        const unsigned c = (unsigned)(i - pix * src_v);
        if constexpr (sizeof(V) == 16) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c);
            float4 v = *reinterpret_cast<const float4 *>(y + i);     // plain load: `also` may alias y
            v.x = lrelu(__fadd_rn(v.x, b.x), slope); v.y = lrelu(__fadd_rn(v.y, b.y), slope);
            v.z = lrelu(__fadd_rn(v.z, b.z), slope); v.w = lrelu(__fadd_rn(v.w, b.w), slope);
            reinterpret_cast<float4 *>(dst)[pix * dst_v + off_v + c] = v;
            if (also) reinterpret_cast<float4 *>(also)[i] = v;
        } else {
            const float v = lrelu(__fadd_rn(reinterpret_cast<const float *>(y)[i], __ldg(reinterpret_cast<const float *>(bias) + c)), slope);
            reinterpret_cast<float *>(dst)[pix * dst_v + off_v + c] = v;
            if (also) reinterpret_cast<float *>(also)[i] = v;
        }
    }
}

// dst[pix][c_off .. c_off + c_n) = value  (the zero pad channels that round a concat buffer up to a multiple of 8)
__global__ void __launch_bounds__(256) fill_channels_nhwc(float *__restrict__ dst, size_t total, unsigned c_n, unsigned c_dst,
                                                          unsigned c_off, float value)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / c_n;
        dst[pix * c_dst + c_off + (unsigned)(i - pix * c_n)] = value;
    }
}

// any shape / alignment
__global__ void __launch_bounds__(256) bias_lrelu_scalar(float *__restrict__ y, const float *__restrict__ bias,
                                                         size_t total, size_t inner, unsigned C, int channels_last, float slope)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned c = channels_last ? (unsigned)(i % C) : (unsigned)((i / inner) % C);
        y[i] = lrelu(__fadd_rn(y[i], __ldg(bias + c)), slope);
    }
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_bias_lrelu(float *y, const float *bias, int N, int C, int HW, int channels_last,
                                  float slope, void *stream)
{
    FLOWOPS_REQUIRE(y && bias, FLOWOPS_EINVAL, "bias_lrelu: null pointer");
    FLOWOPS_REQUIRE(N > 0 && C > 0 && HW > 0, FLOWOPS_EINVAL, "bias_lrelu: bad shape %d x %d x %d", N, C, HW);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)N * C * HW;
    auto grid_for = [](size_t items) {
        size_t blocks = (items + 255) / 256;
        const size_t cap = (size_t)kNumSMs * 8 * 16;
        return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
    };
    if (channels_last && (C & 3) == 0 && aligned16(y) && aligned16(bias))
        bias_lrelu_nhwc<<<grid_for(total / 4), 256, 0, st>>>(y, bias, total / 4, (unsigned)(C / 4), slope);
    else if (!channels_last && (HW & 3) == 0 && aligned16(y))
        bias_lrelu_nchw<<<grid_for(total / 4), 256, 0, st>>>(y, bias, total / 4, (unsigned)(HW / 4), (unsigned)C, slope);
    else
        bias_lrelu_scalar<<<grid_for(total), 256, 0, st>>>(y, bias, total, (size_t)HW, (unsigned)C, channels_last, slope);
    return check_launch("bias_lrelu");
}

extern "C" int flowops_bias_lrelu_nhwc_to(const float *y, const float *bias, float *dst, size_t n_pixels,
                                          int C, int c_dst, int c_off, float slope, float *also, void *stream)
{
    FLOWOPS_REQUIRE(y && bias && dst, FLOWOPS_EINVAL, "bias_lrelu_nhwc_to: null pointer");
    FLOWOPS_REQUIRE(n_pixels > 0 && C > 0 && c_off >= 0 && c_off + C <= c_dst, FLOWOPS_EINVAL,
                    "bias_lrelu_nhwc_to: bad channel range %d + %d of %d", c_off, C, c_dst);
    cudaStream_t st = (cudaStream_t)stream;
    auto grid_for = [](size_t items) {
        size_t blocks = (items + 255) / 256;
        const size_t cap = (size_t)kNumSMs * 8 * 16;
        return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
    };
    const size_t total = n_pixels * (size_t)C;
    if (((C | c_dst | c_off) & 3) == 0 && aligned16(y) && aligned16(dst) && aligned16(bias) && aligned16(also))
        bias_lrelu_nhwc_to<float4><<<grid_for(total / 4), 256, 0, st>>>(reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(bias),
                                                                        reinterpret_cast<float4 *>(dst), total / 4, C / 4, c_dst / 4, c_off / 4, slope,
                                                                        reinterpret_cast<float4 *>(also));
    else
        bias_lrelu_nhwc_to<float><<<grid_for(total), 256, 0, st>>>(y, bias, dst, total, C, c_dst, c_off, slope, also);
    return check_launch("bias_lrelu_nhwc_to");
}

extern "C" int flowops_fill_channels_nhwc(float *dst, size_t n_pixels, int c_dst, int c_off, int c_n, float value, void *stream)
{
    FLOWOPS_REQUIRE(dst, FLOWOPS_EINVAL, "fill_channels_nhwc: null pointer");
    FLOWOPS_REQUIRE(n_pixels > 0 && c_n > 0 && c_off >= 0 && c_off + c_n <= c_dst, FLOWOPS_EINVAL,
                    "fill_channels_nhwc: bad channel range %d + %d of %d", c_off, c_n, c_dst);
    const size_t total = n_pixels * (size_t)c_n;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    fill_channels_nhwc<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, total, c_n, c_dst, c_off, value);
    return check_launch("fill_channels_nhwc");
}

// ---------------------------------------------------------------------------------------------
// channel concatenation of channels-last tensors (the decoder skip concats of FlowNetC/S/SD/Fusion,
// e.g. reference FlowNetS.py:74-90).  torch.cat's generic batched copy reaches ~1 TB/s on these shapes
// (profiles/torchprof_flownet_cl_r01.txt); this is a plain strided row copy: every pixel's C_src channels are
// contiguous in the source and land at a channel offset inside the C_dst-wide destination pixel.
// ---------------------------------------------------------------------------------------------
namespace flowops {

template <typename V>
__global__ void __launch_bounds__(256) copy_channels_nhwc(const V *__restrict__ src, V *__restrict__ dst,
                                                          size_t total, unsigned src_v, unsigned dst_v, unsigned off_v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / src_v;
        const unsigned c = (unsigned)(i - pix * src_v);
        dst[pix * dst_v + off_v + c] = src[i];
    }
}

}  // namespace flowops

extern "C" int flowops_concat_nhwc(const float *src, float *dst, size_t n_pixels, int c_src, int c_dst, int c_off, void *stream)
{
    FLOWOPS_REQUIRE(src && dst, FLOWOPS_EINVAL, "concat_nhwc: null pointer");
    FLOWOPS_REQUIRE(n_pixels > 0 && c_src > 0 && c_off >= 0 && c_off + c_src <= c_dst, FLOWOPS_EINVAL,
                    "concat_nhwc: bad channel range %d + %d of %d", c_off, c_src, c_dst);
    cudaStream_t st = (cudaStream_t)stream;
    auto grid_for = [](size_t items) {
        size_t blocks = (items + 255) / 256;
        const size_t cap = (size_t)kNumSMs * 8 * 16;
        return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
    };
    const size_t total = n_pixels * (size_t)c_src;
    if (((c_src | c_dst | c_off) & 3) == 0 && aligned16(src) && aligned16(dst))
        copy_channels_nhwc<float4><<<grid_for(total / 4), 256, 0, st>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst),
                                                                        total / 4, c_src / 4, c_dst / 4, c_off / 4);
    else if (((c_src | c_dst | c_off) & 1) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 7) == 0)
        copy_channels_nhwc<float2><<<grid_for(total / 2), 256, 0, st>>>(reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst),
                                                                        total / 2, c_src / 2, c_dst / 2, c_off / 2);
    else
        copy_channels_nhwc<float><<<grid_for(total), 256, 0, st>>>(src, dst, total, c_src, c_dst, c_off);
    return check_launch("concat_nhwc");
}


// ---------------------------------------------------------------------------------------------
// The decoders' flow upsamplers: ConvTranspose2d(2, 2, kernel 4, stride 2, padding 1) on the 2-channel flow
// (FlowNetS.py:46-49,75-88 `upsampled_flow6_to_5` ...), written straight into its 2-channel slice of the next level's
// channels-last concat buffer.  16 multiply-adds per output pixel: cuDNN answers this layer with a strided-dgrad GEMM
// kernel plus channel-padding kernels on both sides, followed here by a bias pass and a copy into the concat buffer --
// five launches for what is one small gather.
//   out[b, 2m+py, 2n+px, c_off+co] = bias[co] + sum_ci sum_{t,u in 0..1} in[b, m-1+py+t, n-1+px+u, ci] * w[ci, co, ky(py,t), kx(px,u)]
//   with ky(0,.) = (3, 1), ky(1,.) = (2, 0)  (see cudnn_fused.py for the derivation); out-of-range input reads as zero.
// Accumulation order: ci outer, taps inner, fp32 FMAs (cuDNN's order for this layer is unspecified; the operator
// tolerance applies, tests/test_flownet_gpu.py).
// ---------------------------------------------------------------------------------------------
namespace flowops {

// one output pixel of ConvTranspose2d(2, 2, k4, s2, p1): sw = [ci][co][ky][kx] weights in shared memory
__device__ __forceinline__ float2 flow_deconv_pixel(const float2 *__restrict__ in, const float *sw, float b0, float b1,
                                                    int b, int oy, int ox, int h, int w)
{
    const int py = oy & 1, px = ox & 1, m = oy >> 1, n = ox >> 1;
    float a0 = b0, a1 = b1;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int iy = m - 1 + py + t, ky = py ? (t ? 0 : 2) : (t ? 1 : 3);
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int ix = n - 1 + px + u, kx = px ? (u ? 0 : 2) : (u ? 1 : 3);
            if (ix < 0 || ix >= w) continue;
            const float2 v = __ldg(in + ((size_t)b * h + iy) * w + ix);
            const int k = ky * 4 + kx;
            a0 = __fmaf_rn(v.x, sw[k], a0);      a1 = __fmaf_rn(v.x, sw[16 + k], a1);        // ci = 0 -> co = 0, 1
            a0 = __fmaf_rn(v.y, sw[32 + k], a0); a1 = __fmaf_rn(v.y, sw[48 + k], a1);        // ci = 1
        }
    }
    return make_float2(a0, a1);
}

__global__ void __launch_bounds__(256) flow_deconv_nhwc_kernel(const float2 *__restrict__ in, const float *__restrict__ wgt,
                                                               const float *__restrict__ bias, float *__restrict__ dst,
                                                               int B, int h, int w, unsigned c_dst, unsigned c_off, unsigned tail_zero)
{
    __shared__ float sw[64];                      // [ci][co][ky][kx]
    if (threadIdx.x < 64) sw[threadIdx.x] = wgt[threadIdx.x];
    __syncthreads();
    const int W2 = 2 * w, H2 = 2 * h;
    const size_t total = (size_t)B * H2 * W2;
    const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % W2);
        const size_t r = i / W2;
        const int oy = (int)(r % H2), b = (int)(r / H2);
        const float2 f = flow_deconv_pixel(in, sw, b0, b1, b, oy, ox, h, w);
        float *d = dst + i * c_dst + c_off;
        if (tail_zero == 0) {
            *reinterpret_cast<float2 *>(d) = f;
        } else {                                  // whole sectors: see bias_lrelu_d2s_flowup_kernel
            *reinterpret_cast<float4 *>(d) = make_float4(f.x, f.y, 0.f, 0.f);
            if (tail_zero == 6) *reinterpret_cast<float4 *>(d + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

}  // namespace flowops

extern "C" int flowops_flow_deconv_nhwc_to(const float *flow, const float *weight, const float *bias, float *dst,
                                           int B, int h, int w, int c_dst, int c_off, int tail_zero, void *stream)
{
    FLOWOPS_REQUIRE(flow && weight && dst, FLOWOPS_EINVAL, "flow_deconv_nhwc_to: null pointer");
    FLOWOPS_REQUIRE(B > 0 && h > 0 && w > 0 && c_off >= 0 && c_off + 2 <= c_dst && ((c_off | c_dst) & 1) == 0, FLOWOPS_EINVAL,
                    "flow_deconv_nhwc_to: bad shape / channel range (%d, %d x %d, channels %d + 2 of %d; offsets must be even)", B, h, w, c_off, c_dst);
    FLOWOPS_REQUIRE((((uintptr_t)flow | (uintptr_t)dst) & 7) == 0, FLOWOPS_EINVAL, "flow_deconv_nhwc_to: 8-byte alignment required");
    FLOWOPS_REQUIRE(tail_zero == 0 || ((tail_zero == 2 || tail_zero == 6) && c_off + 2 + tail_zero <= c_dst && ((c_off | c_dst) & 3) == 0 &&
                                       aligned16(dst)), FLOWOPS_EINVAL,
                    "flow_deconv_nhwc_to: tail_zero %d (0, or 2 / 6 zero channels behind the flow inside the %d channels, with c_off, c_dst "
                    "multiples of 4 and a 16-byte aligned dst)", tail_zero, c_dst);
    const size_t total = (size_t)B * 4 * h * w;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 8;
    if (blocks > cap) blocks = cap;
    flow_deconv_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(flow), weight, bias, dst,
                                                                                  B, h, w, (unsigned)c_dst, (unsigned)c_off, (unsigned)tail_zero);
    return check_launch("flow_deconv_nhwc_to");
}


// ---------------------------------------------------------------------------------------------
// Epilogue of a ConvTranspose2d(k4, s2, p1) that was computed as a 3x3 convolution with 4*C output channels at the INPUT
// resolution (one group of C channels per output parity; submodules.deconv_as_conv3): bias + LeakyReLU + depth-to-space,
// written into channels [c_off, c_off + C) of the channels-last concat buffer at twice the resolution.
//   dst[b, 2m+py, 2n+px, c_off + co] = lrelu(y4[b, m, n, (py*2+px)*C + co] + bias[co])
// A thread moves one float4: reads are fully coalesced, writes are C-float runs per output pixel.
// ---------------------------------------------------------------------------------------------
namespace flowops {

__global__ void __launch_bounds__(256) bias_lrelu_d2s_kernel(const float4 *__restrict__ y4, const float *__restrict__ bias,
                                                             float *__restrict__ dst, size_t total_q, unsigned h, unsigned w,
                                                             unsigned cq, unsigned c_dst, unsigned c_off, float slope)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_q; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned q = (unsigned)(i % cq);                 // float4 index inside the parity's C channels
        size_t r = i / cq;
        const unsigned par = (unsigned)(r & 3); r >>= 2;
        const unsigned n = (unsigned)(r % w); r /= w;
        const unsigned m = (unsigned)(r % h);
        const size_t b = r / h;
        const float4 v = y4[i];
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias) + q);
        float4 o;
        o.x = __fadd_rn(v.x, bb.x); o.x = o.x > 0.f ? o.x : __fmul_rn(o.x, slope);
        o.y = __fadd_rn(v.y, bb.y); o.y = o.y > 0.f ? o.y : __fmul_rn(o.y, slope);
        o.z = __fadd_rn(v.z, bb.z); o.z = o.z > 0.f ? o.z : __fmul_rn(o.z, slope);
        o.w = __fadd_rn(v.w, bb.w); o.w = o.w > 0.f ? o.w : __fmul_rn(o.w, slope);
        const size_t pix = (b * (2 * h) + 2 * m + (par >> 1)) * (size_t)(2 * w) + 2 * n + (par & 1);
        *reinterpret_cast<float4 *>(dst + pix * c_dst + c_off + 4 * q) = o;
    }
}

}  // namespace flowops

extern "C" int flowops_bias_lrelu_d2s_nhwc_to(const float *y4, const float *bias, float *dst, int B, int h, int w, int C,
                                              int c_dst, int c_off, float slope, void *stream)
{
    FLOWOPS_REQUIRE(y4 && bias && dst, FLOWOPS_EINVAL, "bias_lrelu_d2s_nhwc_to: null pointer");
    FLOWOPS_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && (C & 3) == 0 && c_off >= 0 && c_off + C <= c_dst && ((c_off | c_dst) & 3) == 0,
                    FLOWOPS_EINVAL, "bias_lrelu_d2s_nhwc_to: bad shape / channel range (C %d, channels %d + C of %d; multiples of 4)", C, c_off, c_dst);
    FLOWOPS_REQUIRE(aligned16(y4) && aligned16(dst) && aligned16(bias), FLOWOPS_EINVAL, "bias_lrelu_d2s_nhwc_to: 16-byte alignment required");
    const size_t total_q = (size_t)B * h * w * C;            // = B*h*w*4*C / 4 float4s
    size_t blocks = (total_q + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    bias_lrelu_d2s_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(y4), bias, dst, total_q,
                                                                                 (unsigned)h, (unsigned)w, (unsigned)(C / 4), (unsigned)c_dst,
                                                                                 (unsigned)c_off, slope);
    return check_launch("bias_lrelu_d2s_nhwc_to");
}


// ---------------------------------------------------------------------------------------------
// The same epilogue with the decoder level's 2-channel flow upsampler folded in: channels [c_off, c_off + C) as above and
// channels [c_off + C, c_off + C + 2) = ConvTranspose2d(2, 2, k4, s2, p1)(flow) + its bias, for the flow at the INPUT
// resolution [B, h, w, 2] (FlowNetFusion.py:52-60: torch.cat((skip, deconv(x), upsampled_flow(flow)), 1)).  One more "quad"
// per output pixel writes the two flow channels right behind the pixel's run of deconvolution channels, instead of a second
// kernel writing 8 bytes per pixel at the concat buffer's channel pitch (309 us per 16 pairs at 512 x 1024).
// ---------------------------------------------------------------------------------------------
namespace flowops {

__global__ void __launch_bounds__(256) bias_lrelu_d2s_flowup_kernel(const float4 *__restrict__ y4, const float *__restrict__ bias,
                                                                    float *__restrict__ dst, size_t total_q, unsigned h, unsigned w,
                                                                    unsigned cq, unsigned c_dst, unsigned c_off, float slope,
                                                                    const float2 *__restrict__ flow, const float *__restrict__ fwgt,
                                                                    const float *__restrict__ fbias, unsigned tail_zero)
{
    __shared__ float sw[64];
    if (threadIdx.x < 64) sw[threadIdx.x] = fwgt[threadIdx.x];
    __syncthreads();
    const float fb0 = fbias ? __ldg(fbias) : 0.f, fb1 = fbias ? __ldg(fbias + 1) : 0.f;
    const unsigned cq1 = cq + 1;                               // quads per output pixel: C / 4 of the deconvolution + 1 for the flow
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_q; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned q = (unsigned)(i % cq1);
        size_t r = i / cq1;
        const unsigned par = (unsigned)(r & 3); r >>= 2;
        const unsigned n = (unsigned)(r % w); r /= w;
        const unsigned m = (unsigned)(r % h);
        const size_t b = r / h;
        const unsigned oy = 2 * m + (par >> 1), ox = 2 * n + (par & 1);
        const size_t pix = (b * (2 * h) + oy) * (size_t)(2 * w) + ox;
        if (q < cq) {
            const float4 v = y4[(((b * h + m) * w + n) * 4 + par) * cq + q];
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias) + q);
            float4 o;
            o.x = __fadd_rn(v.x, bb.x); o.x = o.x > 0.f ? o.x : __fmul_rn(o.x, slope);
            o.y = __fadd_rn(v.y, bb.y); o.y = o.y > 0.f ? o.y : __fmul_rn(o.y, slope);
            o.z = __fadd_rn(v.z, bb.z); o.z = o.z > 0.f ? o.z : __fmul_rn(o.z, slope);
            o.w = __fadd_rn(v.w, bb.w); o.w = o.w > 0.f ? o.w : __fmul_rn(o.w, slope);
            *reinterpret_cast<float4 *>(dst + pix * c_dst + c_off + 4 * q) = o;
        } else {
            const float2 f = flow_deconv_pixel(flow, sw, fb0, fb1, (int)b, (int)oy, (int)ox, (int)h, (int)w);
            float *d = dst + pix * c_dst + c_off + 4 * cq;
            if (tail_zero == 0) {
                *reinterpret_cast<float2 *>(d) = f;
            } else {
                // the zero pad channels behind the flow are (re)written with it: 16-byte stores that complete the pixel's last
                // 32-byte sector instead of an 8-byte store into it -- L2 answers a partially written sector with a DRAM
                // read-modify-write, which made the two flow channels cost as much as the 16 deconvolution channels
                *reinterpret_cast<float4 *>(d) = make_float4(f.x, f.y, 0.f, 0.f);
                if (tail_zero == 6) *reinterpret_cast<float4 *>(d + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

}  // namespace flowops

extern "C" int flowops_bias_lrelu_d2s_flowup_nhwc_to(const float *y4, const float *bias, float *dst, int B, int h, int w, int C,
                                                     int c_dst, int c_off, float slope, const float *flow, const float *flow_weight,
                                                     const float *flow_bias, int tail_zero, void *stream)
{
    FLOWOPS_REQUIRE(y4 && bias && dst && flow && flow_weight, FLOWOPS_EINVAL, "bias_lrelu_d2s_flowup_nhwc_to: null pointer");
    FLOWOPS_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && (C & 3) == 0 && c_off >= 0 && c_off + C + 2 <= c_dst && ((c_off | c_dst) & 3) == 0,
                    FLOWOPS_EINVAL, "bias_lrelu_d2s_flowup_nhwc_to: bad shape / channel range (C %d + 2, channels from %d of %d; multiples of 4)", C, c_off, c_dst);
    FLOWOPS_REQUIRE(aligned16(y4) && aligned16(dst) && aligned16(bias) && (((uintptr_t)flow) & 7) == 0, FLOWOPS_EINVAL,
                    "bias_lrelu_d2s_flowup_nhwc_to: 16-byte alignment required (8 for the flow)");
    FLOWOPS_REQUIRE((tail_zero == 0 || tail_zero == 2 || tail_zero == 6) && c_off + C + 2 + tail_zero <= c_dst, FLOWOPS_EINVAL,
                    "bias_lrelu_d2s_flowup_nhwc_to: tail_zero %d (0, 2 or 6 zero channels behind the flow, inside the %d channels)", tail_zero, c_dst);
    const size_t total_q = (size_t)B * h * w * 4 * (C / 4 + 1);
    size_t blocks = (total_q + 255) / 256;
    const size_t cap = (size_t)kNumSMs * 8 * 16;
    if (blocks > cap) blocks = cap;
    bias_lrelu_d2s_flowup_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(y4), bias, dst, total_q, (unsigned)h, (unsigned)w, (unsigned)(C / 4), (unsigned)c_dst, (unsigned)c_off,
        slope, reinterpret_cast<const float2 *>(flow), flow_weight, flow_bias, (unsigned)tail_zero);
    return check_launch("bias_lrelu_d2s_flowup_nhwc_to");
}
