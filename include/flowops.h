/*
 * flowops.h -- C ABI of libflowops.so: the B200 (sm_100a) flow hot path of ir2rgb / vid2vid.
 *
 * This is the drop-in boundary.  Each entry point replaces one pybind11/ATen entry point (or one
 * ATen call) of the reference; paths are relative to
 * /root/reference/models/flownet2_pytorch/networks/ unless they start with models/.
 *
 * Conventions
 *   - every tensor is fp32, contiguous NCHW, resident on the device the stream belongs to (the *_16 entry points at
 *     the end take fp16 / bf16 storage instead and say so);
 *   - the library never allocates, frees or retains a pointer; scratch space is passed in as
 *     `workspace` (query the size first), outputs are fully written (callers need not zero them);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*, NULL = legacy default
 *     stream), performs no host synchronisation and is CUDA-graph capturable;
 *   - return value: 0 on success, a positive cudaError_t, or a negative FLOWOPS_E* code;
 *     flowops_last_error() returns a thread-local message for the last non-zero return.
 */
#ifndef FLOWOPS_H
#define FLOWOPS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOWOPS_VERSION 1

#define FLOWOPS_EINVAL (-1)        /* bad argument (null pointer, non-positive size, ...)            */
#define FLOWOPS_EUNSUPPORTED (-2)  /* parameter combination the reference never defines/uses        */
#define FLOWOPS_EWORKSPACE (-3)    /* workspace missing, misaligned or too small                    */

/* memory layout of a [B,C,H,W] input */
#define FLOWOPS_LAYOUT_NCHW 0      /* contiguous, what the reference extensions require                */
#define FLOWOPS_LAYOUT_NHWC 1      /* torch.channels_last: element (n,c,y,x) at ((n*H+y)*W+x)*C+c      */

/* storage type of the *_16 entry points ("16-bit storage, fp32 math") */
#define FLOWOPS_DTYPE_F16 1        /* IEEE binary16 (torch.float16)                                    */
#define FLOWOPS_DTYPE_BF16 2       /* bfloat16 (torch.bfloat16)                                        */

/* warp coordinate conventions */
#define FLOWOPS_WARP_RESAMPLE2D 0  /* resample2d_package: sample at (x+dx, y+dy), clamp corners      */
#define FLOWOPS_WARP_GRIDSAMPLE 1  /* models/networks.py:93-100: vid2vid grid + F.grid_sample
                                      (bilinear, border, align_corners=False)                       */

int flowops_version(void);
const char *flowops_last_error(void);

/* ---- ChannelNorm --------------------------------------------------------------------------- */

/* Replaces channelnorm_cuda.forward (channelnorm_cuda.cc:6-15 -> channelnorm_kernel.cu:98-127).
 * y[b,0,h,w] = sqrt(sum_c x[b,c,h,w]^2); x is [B,C,H,W], y is [B,1,H,W].  Bit-identical to the
 * reference kernel (FFMA chain in channel order, IEEE sqrt).  norm_deg is ignored there too. */
int flowops_cnorm_fwd(const float *x, float *y, int B, int C, int H, int W, void *stream);

/* Replaces channelnorm_cuda.backward (channelnorm_cuda.cc:17-25 -> channelnorm_kernel.cu:129-177).
 * gx[b,c,h,w] = gy[b,0,h,w] * x[b,c,h,w] / (y[b,0,h,w] + 1e-9), divide in fp64. */
int flowops_cnorm_bwd(const float *x, const float *y, const float *gy, float *gx,
                      int B, int C, int H, int W, void *stream);

/* ---- Flow warp ----------------------------------------------------------------------------- */

/* Replaces resample2d_cuda.forward (resample2d_cuda.cc:6-14 -> resample2d_kernel.cu:192-233) for
 * mode RESAMPLE2D (kernel_size 1), and the get_grid + normalise + F.grid_sample chain of
 * models/networks.py:15-28,89-100 (== models/base_model.py:123-136) for mode GRIDSAMPLE.
 * img, out: [B,C,H,W]; flow: [B,2,H,W] in pixels (ch0 = dx, ch1 = dy).
 * lin_x[W], lin_y[H]: torch.linspace(-1,1,W|H) tables (device), required for GRIDSAMPLE only. */
int flowops_warp_fwd(const float *img, const float *flow, float *out,
                     int B, int C, int H, int W, int mode,
                     const float *lin_x, const float *lin_y, void *stream);

/* Replaces resample2d_cuda.backward (resample2d_cuda.cc:16-26 -> resample2d_kernel.cu:235-310) and
 * autograd through the models/networks.py chain.  gimg ([B,C,H,W]) and gflow ([B,2,H,W]) may each
 * be NULL when that gradient is not needed (discriminator.py:120,137).  gimg is zero-filled by the
 * callee and then accumulated with warp-aggregated atomics. */
int flowops_warp_bwd(const float *img, const float *flow, const float *gout,
                     float *gimg, float *gflow,
                     int B, int C, int H, int W, int mode,
                     const float *lin_x, const float *lin_y, void *stream);

/* Process-wide switches of the warp kernels (read at every call; environment variable FLOWOPS_WARP_IMPL sets the
 * initial value):
 *   bit 0 (default 0)  flowops_warp_bwd accumulates the image gradient in per-warp shared-memory windows that each warp
 *                      owns (plain adds, dense vector reductions on flush) instead of one L2 reduction per contribution;
 *                      same values up to the summation order (backward tolerance 1e-4).  Needs C <= 3, W % 4 == 0.
 *                      Faster on incoherent flows (per-pixel noise), slower on smooth ones -- hence off by default.
 *   bit 1 (default 0)  tolerance mode of the RESAMPLE2D forward (flowops_warp_fwd with C <= 3 and the fused glue kernels
 *                      below): bilinear weights and blend in fp32.  The reference forms three of its four weight products
 *                      in fp64 by accident of a `1.` literal (resample2d_kernel.cu:55-58); the default path reproduces
 *                      that bit for bit at the price of 20 fp64 conversions per pixel.  The fp32 blend differs from it
 *                      by ~1e-7 max-relative (tolerance 1e-5).
 *   bit 2 (default 0)  flowops_warp_bwd accumulates the image gradient per tile in shared memory in fixed point (exact to
 *                      2^-28 of the tile's largest gradient); faster on per-pixel-random flows only.
 *   bit 3 (default 0)  GRIDSAMPLE forward of frames with C <= 3: use the row-walking kernel instead of the variant that keeps
 *                      two rows of corner gathers in flight per thread (bit-identical results; A/B timing only).
 * No reference counterpart. */
int flowops_warp_set_impl(int flags);
int flowops_warp_get_impl(void);

/* ---- Correlation --------------------------------------------------------------------------- */

/* Output shape, correlation_cuda.cc:19-34. */
int flowops_corr_out_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                           int *oC, int *oH, int *oW);

/* Scratch the fast path needs (the role of rInput1/rInput2 in correlation.py:16-17); 0 when the
 * generic kernel is used.  Alignment requirement on `workspace`: 256 bytes. */
size_t flowops_corr_fwd_workspace_bytes(int B, int C, int H, int W,
                                        int pad, int k, int md, int s1, int s2);
size_t flowops_corr_bwd_workspace_bytes(int B, int C, int H, int W,
                                        int pad, int k, int md, int s1, int s2);

/* Replaces correlation_cuda.forward (correlation_cuda.cc:10-87 -> correlation_cuda_kernel.cu:336-427).
 * in1, in2: [B,C,H,W] in `in_layout` (FLOWOPS_LAYOUT_NCHW like the reference, or FLOWOPS_LAYOUT_NHWC so that a
 * channels_last conv body can hand its features over without a layout copy; FlowNetC configuration only);
 * out: [B,oC,oH,oW], always NCHW.  corr_multiply is ignored by the reference kernels and is not part of this ABI. */
int flowops_corr_fwd(const float *in1, const float *in2, float *out,
                     int B, int C, int H, int W,
                     int pad, int k, int md, int s1, int s2, int in_layout,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Split form of flowops_corr_fwd for a channels_last FlowNetC (FlowNetC.py:75-89), FlowNetC configuration only:
 *   flowops_corr_planes_from_conv  is the epilogue of the conv3 convolution of frame `which` (0 or 1): it applies
 *       bias + LeakyReLU(slope) to the bias-free NHWC conv output y [B,C,H,W], writes the result into the
 *       correlation's internal layout inside `workspace`, and -- when act is not NULL (may alias y) -- also back in
 *       NHWC order for other consumers (conv_redir reads frame 0's features, nothing else reads frame 1's);
 *   flowops_corr_fwd_planes        then runs the correlation proper on what the workspace holds.
 * Together they equal  bias+LeakyReLU (x2)  ->  flowops_corr_fwd  bit for bit, with two passes over each feature
 * tensor less. */
int flowops_corr_planes_from_conv(const float *y, const float *bias, float slope, float *act, int which,
                                  int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                  void *workspace, size_t workspace_bytes, void *stream);
int flowops_corr_fwd_planes(float *out, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Channels-last form of flowops_corr_fwd_planes for a channels_last FlowNetC: the 441-channel cost volume goes to
 * channels [c_off, c_off + 441) of out, a channels-last [B, H, W, c_dst] tensor, with LeakyReLU(lrelu_slope) applied
 * (1.0f = none).  That is `corr_activation` and the concat with conv_redir of FlowNetC.py:89-94 folded into the
 * correlation's store -- same values as correlation -> LeakyReLU -> torch.cat, bit for bit. */
int flowops_corr_fwd_planes_nhwc(float *out, int c_dst, int c_off, float lrelu_slope,
                                 int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                 void *workspace, size_t workspace_bytes, void *stream);

/* Which kernel family flowops_corr_fwd* uses for the FlowNetC configuration (process-wide switch, read at every call):
 *   bit 0 = 1  the tcgen05 tensor-core kernel (3xTF32, csrc/corr_tc.cu) when the shape allows it (C % 32 == 0, H and W
 *              even) -- the default; max-relative error ~1e-6 against the reference kernel
 *   bit 0 = 0  the FP32-FMA kernel (csrc/corr_fast.cu; also selected by the environment variable FLOWOPS_CORR_IMPL=ffma)
 * bits 1-2 are debugging aids of the tensor-core kernel (2: rewrite the hi operand tile truncated, 4: single TF32
 * product -- NOT within tolerance).  The two halves of the split form (flowops_corr_planes_from_conv, then
 * flowops_corr_fwd_planes*) must run under the same setting: they share the workspace layout.  No reference
 * counterpart. */
int flowops_corr_set_impl(int flags);
int flowops_corr_get_impl(void);
/* Debugging aid of the tensor-core kernel: a device buffer of 8 x (number of SMs) uint64 that subsequent launches fill
 * with per-CTA wait / work cycle counters (see csrc/corr_tc.cu); NULL switches it off.  Not thread-safe. */
int flowops_corr_tc_trace(void *device_buffer);

/* Replaces correlation_cuda.backward (correlation_cuda.cc:89-167 -> correlation_cuda_kernel.cu:429-564).
 * gout: [B,oC,oH,oW]; gin1, gin2: [B,C,H,W] (either may be NULL). */
int flowops_corr_bwd(const float *in1, const float *in2, const float *gout,
                     float *gin1, float *gin2,
                     int B, int C, int H, int W,
                     int pad, int k, int md, int s1, int s2,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ---- Fused FlowNet2 glue (additive; bit-identical to chaining the operators above) ------------ */

/* warped = Resample2d(img1, flow); norm = ChannelNorm(img0 - warped)
 * (models/flownet2_pytorch/models.py:109-111,121-123,133-137,146-150) in one pass.
 * img0/img1 are [B,C,H,W] views with a common batch stride (in floats) -- e.g. x[:, :3] and x[:, 3:]
 * of the 6-channel frame stack, which the reference copies with .contiguous() (resample2d.py:45).
 * warped ([B,C,H,W], may be NULL when only the error magnitude is needed, models.py:133-137) and
 * norm ([B,1,H,W]) take their own batch strides so that they can be channel slices of the concat
 * buffer of models.py:114. */
int flowops_warp_diff_norm_fwd(const float *img0, const float *img1, size_t img_batch_stride,
                               const float *flow, float *warped, size_t warped_batch_stride,
                               float *norm, size_t norm_batch_stride,
                               int B, int C, int H, int W, void *stream);

/* conf = (sum_c (im1 - warp(im2, flow))^2 < thresh) ? 1 : 0   (models/flownet.py:50,56-57).
 * im1, im2: [B,C,H,W] contiguous; conf: [B,1,H,W].  mode / lin_x / lin_y as in flowops_warp_fwd.  As run, the
 * reference's `self.resample` at flownet.py:50 is Model.resample (base_model.py:129, mode GRIDSAMPLE): the method
 * shadows the Resample2d submodule of the same name. */
int flowops_warp_conf_fwd(const float *im1, const float *im2, const float *flow, float *conf,
                          float thresh, int B, int C, int H, int W, int mode,
                          const float *lin_x, const float *lin_y, void *stream);

/* The whole concat of models.py:112-114 / 124-126 in one pass, channels-last:
 *   out[b,y,x, 0:12] = (frame 0, frame 1, Resample2d(frame 1, flow), flow / div_flow, ChannelNorm(frame 0 - warped)),
 *   out[b,y,x, 12:c_dst] = 0
 * x: [B,6,H,W] planar stack of both frames, flow: [B,2,H,W], out: [B,H,W,c_dst] (c_dst a multiple of 4, >= 12).
 * Same values as the separate operators followed by torch.cat, bit for bit (flow / div_flow is evaluated as ATen does for a
 * CUDA tensor and a Python scalar: a multiply by the fp32 reciprocal; the same holds for rgb_max below); the layout and the padded channel count
 * are what the first convolution of the next FlowNetS (FlowNetS.py:20) wants on a channels_last body. */
int flowops_warp_diff_norm_concat_nhwc(const float *x, const float *flow, float div_flow, float *out, int c_dst,
                                       int B, int H, int W, void *stream);

/* flowops_warp_diff_norm_concat_nhwc with the producer of the flow folded in: flow = nn.Upsample(scale_factor=4,
 * mode='bilinear')(flow_lo * flow_mul) (models.py:106,118 -- the previous sub-network's quarter-resolution flow2 times
 * div_flow) is formed inside the kernel from flow_lo [B,2,H/4,W/4]; the full-resolution flow is never materialised.
 * Same values as upsampling with torch first and calling the function above, up to the FMA contraction of the
 * bilinear blend (<= 1e-6 max-relative on the flow channels; tests/test_flownet_gpu.py).  H, W multiples of 4. */
int flowops_warp_diff_norm_concat_up4_nhwc(const float *x, const float *flow_lo, float flow_mul, float div_flow,
                                           float *out, int c_dst, int B, int H, int W, void *stream);

/* The input of the fusion network, models.py:129-152, in one pass, channels-last:
 *   out[b,y,x, 0:11] = (frame 0, flow_sd, flow_s2, ChannelNorm(flow_sd), ChannelNorm(flow_s2),
 *                       ChannelNorm(frame 0 - Resample2d(frame 1, flow_sd)), ChannelNorm(frame 0 - Resample2d(frame 1, flow_s2))),
 *   out[b,y,x, 11:c_dst] = 0,
 * with flow_s2 = nearest-x4(flow2_s2 * div_flow) (models.py:130) and flow_sd = nearest-x4(flow2_sd / div_flow)
 * (models.py:143, the division reproduced as is, evaluated like ATen: a multiply by the fp32 reciprocal).
 * x: [B,6,H,W] planar stack of both frames; flow2_s2, flow2_sd: [B,2,H/4,W/4], the quarter-resolution outputs of
 * FlowNetS2 and FlowNetSD; out: [B,H,W,c_dst] (c_dst a multiple of 4, >= 12).  Same values as the two scalings, two
 * nn.Upsample, two ChannelNorm, two Resample2d -> subtract -> ChannelNorm chains and the torch.cat, bit for bit. */
int flowops_flownet2_fusion_input_nhwc(const float *x, const float *flow2_s2, const float *flow2_sd, float div_flow,
                                       float *out, int c_dst, int B, int H, int W, void *stream);

/* FlowNet2 input preparation (models.py:97-101): x = (inputs - rgb_mean) / rgb_max with the two frames stacked
 * along channels.  inputs: [B,3,2,H,W]; rgb_mean: [B,3] (the caller computes the mean).  Outputs, each optional:
 * x_planar [B,6,H,W]; channels-last copies padded with zero channels -- xa_nhwc4 / xb_nhwc4 [B,H,W,4] (frame 0 / 1,
 * FlowNetC's tower inputs, FlowNetC.py:75-76) and x_nhwc8 [B,H,W,8] (both frames, FlowNetSD's conv0). */
int flowops_flownet2_prep(const float *inputs, const float *rgb_mean, float rgb_max,
                          float *x_planar, float *xa_nhwc4, float *xb_nhwc4, float *x_nhwc8,
                          int B, int H, int W, void *stream);

/* flowops_flownet2_prep with frame 0 / frame 1 written "space to depth" for FlowNetC's first layer: xa_s2d, xb_s2d are
 * channels-last [B, H/2 + 1, W/2 + 1, 16] -- channel (py*2+px)*4 + c holds channel c (c < 3; c = 3 is zero) of pixel
 * (2(Y-1)+py, 2(X-1)+px); block row 0 and block column 0 are a zero border that the CALLER zeroes once (the kernel never
 * writes it).  The 7x7 stride-2 convolution of FlowNetC.py:18 on the 3-channel frame equals a 4x4 stride-1 convolution
 * with padding 1 on this tensor with weights W'[co][(py*2+px)*4+c][t][u] = W[co][c][2t+py-1][2u+px-1] (zero where an index
 * is -1): 16 input channels instead of 3 (4) put the layer on cuDNN's tensor-op kernels.  H, W even. */
int flowops_flownet2_prep_s2d(const float *inputs, const float *rgb_mean, float rgb_max,
                              float *x_planar, float *xa_s2d, float *xb_s2d, float *x_nhwc8,
                              int B, int H, int W, void *stream);

/* Either of the two above (s2d = 0 / 1) with the both-frames tensor at a wider channel pitch: x_packed is channels-last
 * [B,H,W,packed_channels] (a multiple of 4, >= 8); the kernel writes channels 0..7 (six frame channels + two zeros) of
 * every pixel and never touches the rest, which the CALLER zeroes once.  cuDNN's fused conv + bias + LeakyReLU engine
 * for FlowNetSD's conv0 (FlowNetSD.py:18, 3x3, 6 -> 64) is 1.4x faster on a 16-channel input than on the 8-channel one
 * (tools/conv_pad_probe.py: 678 vs 967 us per 16 pairs at 512 x 1024). */
int flowops_flownet2_prep_pitched(const float *inputs, const float *rgb_mean, float rgb_max,
                                  float *x_planar, float *xa, float *xb, float *x_packed, int packed_channels, int s2d,
                                  int B, int H, int W, void *stream);

/* ---- 16-bit storage variants (fp16 / bf16 in HBM, fp32 arithmetic) ----------------------------------------
 * The reference's fp16 mode is "fp16 storage, fp32 math" (flownet2_pytorch/main.py:59).  As run it reaches the
 * operators in three ways, and each entry point below reproduces exactly one of them in a single pass, with half the
 * bytes of the fp32 operator and without the separate cast kernels:
 *   - ChannelNorm is called on half tensors and dispatches its kernels for at::Half
 *     (channelnorm_kernel.cu:111,152): own arithmetic, see flowops_cnorm_fwd_16;
 *   - Resample2d and Correlation are fp32-only there and are wrapped in casts:
 *     `resample(a.float(), b.float()).half()` (models.py:22-28), `corr(a.float(), b.float()).half()`
 *     (FlowNetC.py:86-87);
 *   - Model.resample builds its sampling grid in the flow's dtype and casts around F.grid_sample
 *     (models/base_model.py:123-136).
 * `dtype` is FLOWOPS_DTYPE_F16 or FLOWOPS_DTYPE_BF16 and applies to every `void *` tensor of the call; all of them
 * are contiguous NCHW.  float -> 16-bit conversions round to nearest even (what `.half()` / `.bfloat16()` do).
 * Backward passes of the cast-wrapped operators are the fp32 entry points above between two casts, which is what
 * autograd makes of the reference's chain; only ChannelNorm has a 16-bit backward kernel of its own. */

/* kernel_channelnorm_update_output<at::Half> (channelnorm_kernel.cu:19-60): the square of each element is rounded to
 * the storage type before it is added (`val * val` on two at::Half values), the sum and the sqrt are fp32, the result is
 * rounded to the storage type.  Not the same values as cast -> flowops_cnorm_fwd -> cast. */
int flowops_cnorm_fwd_16(const void *x, void *y, int B, int C, int H, int W, int dtype, void *stream);

/* kernel_channelnorm_backward_input1<at::Half> (channelnorm_kernel.cu:64-96):
 * gx = T( float( (double)(float(gy) * float(x)) / ((double)float(y) + 1e-9) ) ). */
int flowops_cnorm_bwd_16(const void *x, const void *y, const void *gy, void *gx,
                         int B, int C, int H, int W, int dtype, void *stream);

/* Mode RESAMPLE2D: fp16_resample2d (models.py:22-28) in one pass -- flowops_warp_fwd's fp32 arithmetic on the
 * widened inputs, result rounded to the storage type; bit-identical to the cast chain.
 * Mode GRIDSAMPLE: Model.resample with `opt['fp16']` on 16-bit tensors (models/base_model.py:123-136): the grid
 * get_grid(..., dtype=flow.dtype), the normalised flow `flow / ((w-1)/2)` and their sum are each rounded to the
 * storage type (that is where the reference evaluates them), F.grid_sample then runs in fp32 on the widened grid and
 * image and its result is rounded to the storage type.  lin_x[W], lin_y[H]: fp32 tables holding
 * linspace(-1,1,n) ROUNDED to the storage type. */
int flowops_warp_fwd_16(const void *img, const void *flow, void *out, int B, int C, int H, int W, int mode,
                        const float *lin_x, const float *lin_y, int dtype, void *stream);

/* `corr(a.float(), b.float()).half()` (FlowNetC.py:86-87) in one pass: the layout pre-pass widens the 16-bit features
 * into the fp32 workspace planes, the correlation proper is flowops_corr_fwd's, the store rounds to the storage type.
 * Workspace: flowops_corr_fwd_workspace_bytes.  Parameter sets outside the FlowNetC configuration return
 * FLOWOPS_EUNSUPPORTED (cast and call flowops_corr_fwd). */
int flowops_corr_fwd_16(const void *in1, const void *in2, void *out, int B, int C, int H, int W,
                        int pad, int k, int md, int s1, int s2, int dtype,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- Conv-body epilogue (FlowNet2 inference glue, not an operator of the reference's native surface) ---- */

/* In place: t = y + bias[c]; y = t > 0 ? t : t * slope.  Replaces the separate bias-add and LeakyReLU kernels
 * that follow every convolution built by submodules.py:7-38 (conv / deconv).  y is [N,C,H,W] stored NCHW
 * (channels_last = 0) or NHWC (channels_last = 1); HW = H*W. */
int flowops_bias_lrelu(float *y, const float *bias, int N, int C, int HW, int channels_last,
                       float slope, void *stream);

/* Out-of-place form for channels-last tensors: dst[pix][c_off + c] = lrelu(y[pix][c] + bias[c]) with y dense
 * [n_pixels][C] and dst [n_pixels][c_dst].  Lets a decoder deconvolution (submodules.py:34-38) write its activated
 * output directly into the concat buffer of torch.cat((skip, deconv, flow_up), 1) (e.g. FlowNetS.py:74-76).
 * `also` (may be NULL, may alias y): a dense [n_pixels][C] copy of the result as well, for an encoder layer whose
 * output is both the next layer's input and a skip connection (FlowNetS.py:63-67 -> :74). */
int flowops_bias_lrelu_nhwc_to(const float *y, const float *bias, float *dst, size_t n_pixels,
                               int C, int c_dst, int c_off, float slope, float *also, void *stream);

/* dst[pix][c_off .. c_off + c_n) = value for every pixel of a channels-last [n_pixels][c_dst] tensor: the zero pad
 * channels that round a concat buffer up to a multiple of 8 channels (cuDNN otherwise re-pads odd channel counts
 * -- 1026, 770, 386, 194, 473 ... -- with a kernel of its own in front of every convolution that reads them). */
int flowops_fill_channels_nhwc(float *dst, size_t n_pixels, int c_dst, int c_off, int c_n, float value, void *stream);

/* Channel concatenation of channels-last tensors: copies src ([n_pixels][c_src], dense) into channels
 * [c_off, c_off + c_src) of dst ([n_pixels][c_dst]).  One call per concatenated tensor (torch.cat of the
 * decoder skip connections, e.g. FlowNetS.py:74). */
int flowops_concat_nhwc(const float *src, float *dst, size_t n_pixels, int c_src, int c_dst, int c_off, void *stream);

/* The decoders' flow upsamplers -- ConvTranspose2d(2, 2, kernel 4, stride 2, padding 1[, bias]) on a dense channels-last
 * 2-channel flow [B, h, w, 2] (FlowNetS.py:46-49 `upsampled_flow6_to_5` ...; weight [2, 2, 4, 4] contiguous, bias [2] or
 * NULL) -- written into channels [c_off, c_off + 2) of the channels-last concat buffer dst [B, 2h, 2w, c_dst] (c_off, c_dst
 * even).  Replaces cuDNN's strided-dgrad launch with its channel-padding kernels, the bias pass and the copy into the
 * buffer.  tail_zero (0, 2 or 6; needs c_off, c_dst multiples of 4): that many channels behind the flow are written as
 * zeros in the same 16-byte stores -- the zero pad channels that end the buffer's pixel record, so that its last 32-byte
 * sector is written whole instead of receiving an 8-byte partial write. */
int flowops_flow_deconv_nhwc_to(const float *flow, const float *weight, const float *bias, float *dst,
                                int B, int h, int w, int c_dst, int c_off, int tail_zero, void *stream);

/* Epilogue of a ConvTranspose2d(kernel 4, stride 2, padding 1) + LeakyReLU (submodules.py:28-31 `deconv`) whose sums were
 * formed by a 3x3 convolution with 4*C output channels at the input resolution, one group of C per output parity
 * (ir2rgb_b200 submodules.deconv_as_conv3): y4 [B, h, w, 4*C] channels-last ->
 *   dst[b, 2m+py, 2n+px, c_off + co] = LeakyReLU(y4[b, m, n, (py*2+px)*C + co] + bias[co])
 * i.e. bias, activation and depth-to-space in one pass into the concat buffer dst [B, 2h, 2w, c_dst].  C, c_off, c_dst
 * multiples of 4. */
int flowops_bias_lrelu_d2s_nhwc_to(const float *y4, const float *bias, float *dst, int B, int h, int w, int C,
                                   int c_dst, int c_off, float slope, void *stream);

/* The same epilogue with the decoder level's flow upsampler folded in (FlowNetFusion.py:52-60,
 * torch.cat((skip, deconv(x), upsampled_flow(flow)), 1)): channels [c_off, c_off + C) as above and channels
 * [c_off + C, c_off + C + 2) = ConvTranspose2d(2, 2, 4, 2, 1)(flow) + flow_bias for the dense channels-last flow
 * [B, h, w, 2] at the input resolution, flow_weight [2, 2, 4, 4] contiguous, flow_bias [2] or NULL -- the arithmetic of
 * flowops_flow_deconv_nhwc_to, bit for bit.  tail_zero (0, 2 or 6): that many channels behind the flow are written as
 * zeros in the same 16-byte stores -- for the zero pad channels that end a concat buffer's pixel record, so that the
 * record's last 32-byte sector is written whole (an 8-byte store into it costs a DRAM read-modify-write). */
int flowops_bias_lrelu_d2s_flowup_nhwc_to(const float *y4, const float *bias, float *dst, int B, int h, int w, int C,
                                          int c_dst, int c_off, float slope, const float *flow, const float *flow_weight,
                                          const float *flow_bias, int tail_zero, void *stream);

/* The decoders' flow heads -- predict_flow = nn.Conv2d(C, 2, kernel_size 3, stride 1, padding 1) (networks/submodules.py:40-41;
 * FlowNetS.py:36-40, FlowNetFusion.py:32-34, ...) -- as one direct FP32 convolution with the bias included:
 *   out[b, y, x, co] = bias[co] + sum_{dy, dx, c < cin} x[b, y+dy-1, x+dx-1, c] * w[co, c, dy, dx]        (zero padding)
 * x: channels-last [B, H, W, c_pitch], of which the first cin channels are read (cin, c_pitch multiples of 4; the weights of
 * zero pad channels are zero); out: channels-last [B, H, W, 2], dense; bias [2] or NULL.  w_packed: the filter in the order the
 * kernel reads it, [ceil(cin / 16)][dy * 3 + dx][(c % 16) / 4][c % 4][co], zero beyond the layer's real channels
 * (ir2rgb_b200.functional.pack_flow_head_weight).  FP32 multiply-adds (FFMA2), i.e. more exact than the TF32 convolution it
 * stands in for; replaces cuDNN's convolution + its two filter / output channel-padding kernels + the bias pass. */
int flowops_flow_head_nhwc(const float *x, int c_pitch, int cin, const float *w_packed, const float *bias, float *out,
                           int B, int H, int W, void *stream);

/* ---- Measurement helper (not part of the reference surface) -------------------------------- */

/* Launches a register-resident FFMA chain kernel on every SM: `iters` loop trips of 64 independent
 * FFMAs per thread.  Returns (through *flops) the FLOP count of one launch so that the caller can
 * time it with CUDA events and derive the FP32-FMA pipe peak the Correlation roofline is quoted
 * against.  sink: device buffer of at least 4 bytes. */
int flowops_bench_ffma(float *sink, int iters, double *flops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWOPS_H */
