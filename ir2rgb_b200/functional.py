"""Tensor-level calls into libflowops.so: argument checking, output allocation, stream and device
handling.  PyTorch is plumbing here (device memory, streams); all arithmetic happens in the library.

Every function requires CUDA tensors and raises otherwise -- by design there is no CPU path.  fp32 is the
operators' dtype (as in the reference); ChannelNorm, the forward warps and the Correlation forward also take
fp16 / bf16 tensors and then run the library's "16-bit storage, fp32 math" entry points (flowops_*_16), which
reproduce what the reference's fp16 mode computes (include/flowops.h).
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import WARP_GRIDSAMPLE, WARP_RESAMPLE2D, check

_DTYPE16 = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _require(t, name, ndim=4, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the flow hot path has no CPU implementation "
                           "(got device %s)" % (name, t.device))
    if t.dtype != dtype:
        if dtype == torch.float32:
            raise TypeError("%s must be float32 (the reference ops are fp32-only on this path; cast like "
                            "FlowNetC.py:86-87 does), got %s" % (name, t.dtype))
        raise TypeError("%s must be %s like the other tensors of the call, got %s" % (name, dtype, t.dtype))
    if t.dim() != ndim:
        raise ValueError("%s must be %d-D, got shape %s" % (name, ndim, tuple(t.shape)))
    return t


def _io_dtype(t):
    """fp32, or one of the 16-bit storage types the *_16 entry points take."""
    if isinstance(t, torch.Tensor) and t.dtype in _DTYPE16:
        return t.dtype
    return torch.float32


def _is_nhwc(t):
    return (not t.is_contiguous()) and t.is_contiguous(memory_format=torch.channels_last)


def _is_nhwc_view(t):
    """Dense channels_last, or a channel slice of a dense channels_last tensor (pixel stride > C): what the cuDNN
    graph path (cudnn_fused.py) reads and writes in place of a copy."""
    if t.dim() != 4 or t.stride(1) != 1 or t.shape[1] == 1:
        return False
    ps = t.stride(3)
    return ps >= t.shape[1] and t.stride(2) == t.shape[3] * ps and t.stride(0) == t.shape[2] * t.stride(2)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------
# ChannelNorm
# ---------------------------------------------------------------------------------------------
def channelnorm_forward(x):
    """fp32, or fp16 / bf16: the reference kernel as instantiated for at::Half (channelnorm_kernel.cu:111)."""
    dt = _io_dtype(x)
    x = _require(x, "input1", dtype=dt).contiguous()
    B, C, H, W = x.shape
    with torch.cuda.device_of(x):
        y = torch.empty((B, 1, H, W), device=x.device, dtype=dt)
        if x.numel():
            if dt == torch.float32:
                check(_lib.load().flowops_cnorm_fwd(_p(x), _p(y), B, C, H, W, _stream()), "cnorm_fwd")
            else:
                check(_lib.load().flowops_cnorm_fwd_16(_p(x), _p(y), B, C, H, W, _DTYPE16[dt], _stream()), "cnorm_fwd_16")
    return y


def channelnorm_backward(x, y, gy):
    dt = _io_dtype(x)
    x = _require(x, "input1", dtype=dt).contiguous()
    y = _require(y, "output", dtype=dt).contiguous()
    gy = _require(gy, "grad_output", dtype=dt).contiguous()
    B, C, H, W = x.shape
    with torch.cuda.device_of(x):
        gx = torch.empty_like(x)
        if x.numel():
            if dt == torch.float32:
                check(_lib.load().flowops_cnorm_bwd(_p(x), _p(y), _p(gy), _p(gx), B, C, H, W, _stream()), "cnorm_bwd")
            else:
                check(_lib.load().flowops_cnorm_bwd_16(_p(x), _p(y), _p(gy), _p(gx), B, C, H, W, _DTYPE16[dt], _stream()),
                      "cnorm_bwd_16")
    return gx


# ---------------------------------------------------------------------------------------------
# warps
# ---------------------------------------------------------------------------------------------
_LIN_CACHE = {}


def _lin_tables(H, W, device, dtype=torch.float32):
    """torch.linspace(-1, 1, n) on the host in fp32, moved to the device: exactly the values
    get_grid produces (reference models/networks.py:15-28).  Cached per (H, W, device) -- the
    counterpart of the reference's `self.grid` cache (networks.py:95-96), 4*(H+W) bytes instead
    of a [b,2,h,w] tensor.  For a 16-bit flow the reference builds the grid in that dtype
    (`get_grid(..., dtype=flow.dtype)`, networks.py:96): the tables then hold the rounded values (as fp32)."""
    key = (H, W, device, dtype)
    t = _LIN_CACHE.get(key)
    if t is None:
        t = tuple(torch.linspace(-1.0, 1.0, n).to(dtype).float().to(device) for n in (W, H))
        _LIN_CACHE[key] = t
    return t


def _warp_args(img, flow, mode, dtype=torch.float32):
    img = _require(img, "image", dtype=dtype)
    flow = _require(flow, "flow", dtype=dtype)
    B, C, H, W = img.shape
    if flow.shape != (B, 2, H, W):
        raise ValueError("flow must be [B,2,H,W] matching the image %s, got %s" % (tuple(img.shape), tuple(flow.shape)))
    if flow.device != img.device:
        raise RuntimeError("image and flow must be on the same device")
    if mode == WARP_GRIDSAMPLE:
        lx, ly = _lin_tables(H, W, img.device, dtype)
    else:
        lx = ly = None
    return img.contiguous(), flow.contiguous(), B, C, H, W, lx, ly


class warp_tolerance_mode:
    """Context manager: inside it the RESAMPLE2D forward (Resample2d with up to 3 channels and the fused FlowNet2 glue)
    blends with fp32 weights instead of reproducing the reference's accidental fp64 weight products bit for bit
    (resample2d_kernel.cu:55-58) -- ~1e-7 max-relative away from the reference kernel (tolerance 1e-5), without the 20
    fp64 conversions per pixel.  Process-wide switch of the library (flowops_warp_set_impl bit 1), restored on exit."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        lib = _lib.load()
        self.prev = lib.flowops_warp_get_impl()
        lib.flowops_warp_set_impl((self.prev | 2) if self.on else (self.prev & ~2))
        return self

    def __exit__(self, *exc):
        _lib.load().flowops_warp_set_impl(self.prev)
        return False


def warp_forward(img, flow, mode=WARP_RESAMPLE2D):
    """fp32 tensors, or image and flow both fp16 / both bf16 (1..3 channels): then one kernel does what the reference's
    fp16 mode spreads over casts -- fp16_resample2d (models.py:22-28) for RESAMPLE2D, Model.resample with opt['fp16']
    (base_model.py:123-136) for GRIDSAMPLE."""
    dt = _io_dtype(img)
    img, flow, B, C, H, W, lx, ly = _warp_args(img, flow, mode, dt)
    with torch.cuda.device_of(img):
        out = torch.empty_like(img)
        if img.numel():
            if dt == torch.float32:
                check(_lib.load().flowops_warp_fwd(_p(img), _p(flow), _p(out), B, C, H, W, mode, _p(lx), _p(ly), _stream()),
                      "warp_fwd")
            else:
                check(_lib.load().flowops_warp_fwd_16(_p(img), _p(flow), _p(out), B, C, H, W, mode, _p(lx), _p(ly),
                                                      _DTYPE16[dt], _stream()), "warp_fwd_16")
    return out


def warp_backward(img, flow, gout, need_img=True, need_flow=True, mode=WARP_RESAMPLE2D):
    img, flow, B, C, H, W, lx, ly = _warp_args(img, flow, mode)
    gout = _require(gout, "grad_output").contiguous()
    if gout.shape != img.shape:
        raise ValueError("grad_output shape %s does not match the image %s" % (tuple(gout.shape), tuple(img.shape)))
    with torch.cuda.device_of(img):
        gimg = torch.empty_like(img) if need_img else None
        gflow = torch.empty_like(flow) if need_flow else None
        if img.numel() and (need_img or need_flow):
            check(_lib.load().flowops_warp_bwd(_p(img), _p(flow), _p(gout), _p(gimg), _p(gflow), B, C, H, W, mode,
                                               _p(lx), _p(ly), _stream()), "warp_bwd")
    return gimg, gflow


# ---------------------------------------------------------------------------------------------
# fused FlowNet2 glue (forward only; used under no_grad)
# ---------------------------------------------------------------------------------------------
def _plane_ptr(t, ch):
    return ctypes.c_void_p(t.data_ptr() + 4 * ch * t.stride(1))


def warp_diff_norm_forward(x, flow, out=None, need_warped=True):
    """x: [B,6,H,W] stack of (frame 0, frame 1), contiguous.  Returns (warped frame 1 or None,
    |frame 0 - warped| as [B,1,H,W]).  With ``out=(buf, ch_warped, ch_norm)`` the results are written
    into channels of the preallocated concat buffer ``buf`` instead of fresh tensors."""
    x = _require(x, "x")
    flow = _require(flow, "flow").contiguous()
    B, C6, H, W = x.shape
    if C6 != 6 or flow.shape != (B, 2, H, W):
        raise ValueError("expected x [B,6,H,W] and flow [B,2,H,W], got %s and %s" % (tuple(x.shape), tuple(flow.shape)))
    x = x.contiguous()
    with torch.cuda.device_of(x):
        if out is None:
            warped = torch.empty((B, 3, H, W), device=x.device, dtype=torch.float32) if need_warped else None
            norm = torch.empty((B, 1, H, W), device=x.device, dtype=torch.float32)
            wp, wbs = _p(warped), 3 * H * W
            npt, nbs = _p(norm), H * W
        else:
            buf, ch_w, ch_n = out
            assert buf.is_contiguous() and buf.shape[0] == B and tuple(buf.shape[2:]) == (H, W)
            warped = buf[:, ch_w:ch_w + 3] if need_warped else None
            norm = buf[:, ch_n:ch_n + 1]
            wp, wbs = (_plane_ptr(buf, ch_w) if need_warped else _p(None)), buf.stride(0)
            npt, nbs = _plane_ptr(buf, ch_n), buf.stride(0)
        if x.numel():
            check(_lib.load().flowops_warp_diff_norm_fwd(_plane_ptr(x, 0), _plane_ptr(x, 3), x.stride(0), _p(flow),
                                                         wp, wbs, npt, nbs, B, 3, H, W, _stream()), "warp_diff_norm_fwd")
    return warped, norm


def warp_diff_norm_concat(x, flow, div_flow, c_pad=16):
    """The concat of models.py:112-114 in one pass: returns a channels_last [B, c_pad, H, W] tensor whose first 12
    channels are (frame 0, frame 1, Resample2d(frame 1, flow), flow / div_flow, ChannelNorm(frame 0 - warped)) and
    whose remaining channels are zero.  x: [B,6,H,W] contiguous (planar), flow: [B,2,H,W]."""
    x = _require(x, "x").contiguous()
    flow = _require(flow, "flow").contiguous()
    B, C6, H, W = x.shape
    if C6 != 6 or flow.shape != (B, 2, H, W):
        raise ValueError("expected x [B,6,H,W] and flow [B,2,H,W], got %s and %s" % (tuple(x.shape), tuple(flow.shape)))
    with torch.cuda.device_of(x):
        out = torch.empty((B, c_pad, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        if x.numel():
            check(_lib.load().flowops_warp_diff_norm_concat_nhwc(_p(x), _p(flow), ctypes.c_float(div_flow), _p(out), c_pad,
                                                                 B, H, W, _stream()), "warp_diff_norm_concat_nhwc")
    return out


def warp_diff_norm_concat_up4(x, flow_lo, flow_mul, div_flow, c_pad=16):
    """warp_diff_norm_concat for flow = nn.Upsample(scale_factor=4, mode='bilinear')(flow_lo * flow_mul) (models.py:106,118)
    without materialising that flow: flow_lo is the previous sub-network's quarter-resolution flow2 [B,2,H/4,W/4]."""
    x = _require(x, "x").contiguous()
    flow_lo = _require(flow_lo, "flow_lo").contiguous()
    B, C6, H, W = x.shape
    if C6 != 6 or H % 4 or W % 4 or tuple(flow_lo.shape) != (B, 2, H // 4, W // 4):
        raise ValueError("expected x [B,6,H,W] with H, W multiples of 4 and flow_lo [B,2,H/4,W/4], got %s and %s"
                         % (tuple(x.shape), tuple(flow_lo.shape)))
    with torch.cuda.device_of(x):
        out = torch.empty((B, c_pad, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        if x.numel():
            check(_lib.load().flowops_warp_diff_norm_concat_up4_nhwc(_p(x), _p(flow_lo), ctypes.c_float(flow_mul), ctypes.c_float(div_flow),
                                                                     _p(out), c_pad, B, H, W, _stream()), "warp_diff_norm_concat_up4_nhwc")
    return out


def flownet2_fusion_input(x, flow2_s2, flow2_sd, div_flow, c_pad=16):
    """concat3 of models.py:129-152 in one pass: a channels_last [B, c_pad, H, W] tensor whose first 11 channels are
    (frame 0, flow_sd, flow_s2, |flow_sd|, |flow_s2|, warp error under flow_sd, warp error under flow_s2), the rest zero.
    x: [B,6,H,W] planar; flow2_s2 / flow2_sd: the quarter-resolution outputs of FlowNetS2 / FlowNetSD in network units
    (the scaling by div_flow and the nearest x4 upsampling happen inside)."""
    x = _require(x, "x").contiguous()
    flow2_s2 = _require(flow2_s2, "flow2_s2").contiguous()
    flow2_sd = _require(flow2_sd, "flow2_sd").contiguous()
    B, C6, H, W = x.shape
    if C6 != 6 or H % 4 or W % 4 or tuple(flow2_s2.shape) != (B, 2, H // 4, W // 4) or flow2_sd.shape != flow2_s2.shape:
        raise ValueError("expected x [B,6,H,W] with H, W multiples of 4 and two [B,2,H/4,W/4] flows, got %s, %s, %s"
                         % (tuple(x.shape), tuple(flow2_s2.shape), tuple(flow2_sd.shape)))
    with torch.cuda.device_of(x):
        out = torch.empty((B, c_pad, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        if x.numel():
            check(_lib.load().flowops_flownet2_fusion_input_nhwc(_p(x), _p(flow2_s2), _p(flow2_sd), ctypes.c_float(div_flow), _p(out),
                                                                 c_pad, B, H, W, _stream()), "flownet2_fusion_input_nhwc")
    return out


_PREP_CACHE = {}


def _prep_packed(B, H, W, device, channels):
    """The both-frames channels_last tensor FlowNetSD's conv0 reads.  8 channels: a fresh dense tensor, written whole by
    the kernel.  More: the kernel writes channels 0..7 only, so the tensor is allocated and zeroed once per shape and
    device and REUSED by later calls (like the space-to-depth frames below)."""
    cl = dict(device=device, dtype=torch.float32, memory_format=torch.channels_last)
    if channels == 8:
        return torch.empty((B, 8, H, W), **cl)
    key = ("packed", B, H, W, device, channels)
    if key not in _PREP_CACHE:
        _PREP_CACHE[key] = torch.empty((B, channels, H, W), **cl).zero_()
    return _PREP_CACHE[key]


def flownet2_prep(inputs, rgb_mean, rgb_max, sd_channels=8):
    """x = (inputs - rgb_mean) / rgb_max for inputs [B,3,2,H,W] (models.py:97-101), in the four layouts its consumers
    read: (x planar [B,6,H,W], frame 0 and frame 1 as channels_last [B,4,H,W], both frames as channels_last
    [B,sd_channels,H,W]); the extra channels are zero.  sd_channels > 8: see _prep_packed (a cached, reused tensor)."""
    inputs = _require(inputs, "inputs", ndim=5).contiguous()
    B, C, F2, H, W = inputs.shape
    if C != 3 or F2 != 2:
        raise ValueError("inputs must be [B,3,2,H,W], got %s" % (tuple(inputs.shape),))
    rgb_mean = rgb_mean.reshape(B, 3).contiguous()
    with torch.cuda.device_of(inputs):
        cl = dict(device=inputs.device, dtype=torch.float32, memory_format=torch.channels_last)
        x = torch.empty((B, 6, H, W), device=inputs.device, dtype=torch.float32)
        xa, xb = torch.empty((B, 4, H, W), **cl), torch.empty((B, 4, H, W), **cl)
        x8 = _prep_packed(B, H, W, inputs.device, sd_channels)
        if inputs.numel():
            check(_lib.load().flowops_flownet2_prep_pitched(_p(inputs), _p(rgb_mean), ctypes.c_float(rgb_max), _p(x), _p(xa), _p(xb),
                                                            _p(x8), sd_channels, 0, B, H, W, _stream()), "flownet2_prep")
    return x, xa, xb, x8


def flownet2_prep_s2d(inputs, rgb_mean, rgb_max, sd_channels=8):
    """flownet2_prep with the two frames in the space-to-depth layout FlowNetC's first layer reads on the fast path
    (include/flowops.h): returns (x planar, xa_s2d, xb_s2d, x8).  The s2d tensors are [B, 16, H/2+1, W/2+1] channels_last
    with a zero first row / column; they are allocated (and zeroed) once per shape and device and REUSED by later calls."""
    inputs = _require(inputs, "inputs", ndim=5).contiguous()
    B, C, F2, H, W = inputs.shape
    if C != 3 or F2 != 2 or H % 2 or W % 2:
        raise ValueError("inputs must be [B,3,2,H,W] with even H, W, got %s" % (tuple(inputs.shape),))
    rgb_mean = rgb_mean.reshape(B, 3).contiguous()
    with torch.cuda.device_of(inputs):
        cl = dict(device=inputs.device, dtype=torch.float32, memory_format=torch.channels_last)
        key = ("s2d", B, H, W, inputs.device)
        if key not in _PREP_CACHE:
            _PREP_CACHE[key] = tuple(torch.empty((B, 16, H // 2 + 1, W // 2 + 1), **cl).zero_() for _ in range(2))
        xa, xb = _PREP_CACHE[key]
        x = torch.empty((B, 6, H, W), device=inputs.device, dtype=torch.float32)
        x8 = _prep_packed(B, H, W, inputs.device, sd_channels)
        check(_lib.load().flowops_flownet2_prep_pitched(_p(inputs), _p(rgb_mean), ctypes.c_float(rgb_max), _p(x), _p(xa), _p(xb),
                                                        _p(x8), sd_channels, 1, B, H, W, _stream()), "flownet2_prep_s2d")
    return x, xa, xb, x8


def warp_conf_forward(im1, im2, flow, thresh=0.02, mode=WARP_GRIDSAMPLE):
    im1 = _require(im1, "im1").contiguous()
    im2 = _require(im2, "im2").contiguous()
    flow = _require(flow, "flow").contiguous()
    B, C, H, W = im1.shape
    if im2.shape != im1.shape or flow.shape != (B, 2, H, W):
        raise ValueError("shape mismatch: %s %s %s" % (tuple(im1.shape), tuple(im2.shape), tuple(flow.shape)))
    lx, ly = _lin_tables(H, W, im1.device) if mode == WARP_GRIDSAMPLE else (None, None)
    with torch.cuda.device_of(im1):
        conf = torch.empty((B, 1, H, W), device=im1.device, dtype=torch.float32)
        if im1.numel():
            check(_lib.load().flowops_warp_conf_fwd(_p(im1), _p(im2), _p(flow), _p(conf), ctypes.c_float(thresh),
                                                    B, C, H, W, mode, _p(lx), _p(ly), _stream()), "warp_conf_fwd")
    return conf


def pack_flow_head_weight(weight, cin):
    """The [2, C, 3, 3] filter of a flow head (C <= cin) in the order flowops_flow_head_nhwc reads it:
    [ceil(cin / 16)][dy * 3 + dx][(c % 16) / 4][c % 4][co], zero for c >= C (include/flowops.h)."""
    co, c, kh, kw = weight.shape
    if co != 2 or kh != 3 or kw != 3 or c > cin:
        raise ValueError("pack_flow_head_weight: expected a [2, C <= %d, 3, 3] filter, got %s" % (cin, tuple(weight.shape)))
    n_chunks = -(-cin // 16)
    wp = weight.new_zeros(n_chunks * 16, 9, 2)                       # [c][tap][co]
    wp[:c] = weight.detach().float().reshape(2, c, 9).permute(1, 2, 0)
    return wp.view(n_chunks, 4, 4, 9, 2).permute(0, 3, 1, 2, 4).contiguous()      # [chunk][tap][quad][ch][co]


def flow_head(x, w_packed, bias):
    """conv2d(x, w, bias, stride 1, padding 1) for a 2-filter 3x3 layer on a channels_last (view) x, as one FP32 kernel
    (flowops_flow_head_nhwc); w_packed from pack_flow_head_weight(w, C_x).  Returns channels_last [B, 2, H, W]."""
    x = _require(x, "x")
    if not (_is_nhwc_view(x) or (x.dim() == 4 and x.shape[1] == 1 and x.is_contiguous())):
        raise ValueError("flow_head: x must be channels_last (or a channel slice of a channels_last tensor)")
    B, C, H, W = x.shape
    if C % 4 or x.stride(3) % 4 or w_packed.numel() != -(-C // 16) * 288:
        raise ValueError("flow_head: channel count %d / pitch %d must be multiples of 4 and match the packed filter" % (C, x.stride(3)))
    with torch.cuda.device_of(x):
        out = torch.empty((B, 2, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        if out.numel():
            check(_lib.load().flowops_flow_head_nhwc(_p(x), x.stride(3), C, _p(w_packed), _p(bias), _p(out), B, H, W, _stream()),
                  "flow_head_nhwc")
    return out


def bias_lrelu_(y, bias, slope):
    """In-place per-channel bias add + LeakyReLU on a conv output (NCHW- or channels_last-contiguous)."""
    y = _require(y, "y")
    N, C, H, W = y.shape
    if y.is_contiguous():
        cl = 0
    elif y.is_contiguous(memory_format=torch.channels_last):
        cl = 1
    else:
        raise ValueError("bias_lrelu_: tensor must be dense in NCHW or channels_last order")
    if bias.shape != (C,) or bias.dtype != torch.float32 or bias.device != y.device or not bias.is_contiguous():
        raise ValueError("bias_lrelu_: bias must be a contiguous fp32 [C] tensor on the same device")
    if y.numel():
        with torch.cuda.device_of(y):
            check(_lib.load().flowops_bias_lrelu(_p(y), _p(bias), N, C, H * W, cl, ctypes.c_float(slope), _stream()),
                  "bias_lrelu")
    return y


def _cat_fast(tensors):
    t0 = tensors[0]
    return (not torch.is_grad_enabled() and t0.numel() > 0 and
            all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4
                and t.shape[0] == t0.shape[0] and t.shape[2:] == t0.shape[2:] and t.device == t0.device and _is_nhwc(t)
                for t in tensors))


D2S_WRITE_PAD = os.environ.get("FLOWOPS_D2S_WRITE_PAD", "1") != "0"      # A/B switch of ConcatBuffer.bias_lrelu_d2s_in's tail write


class ConcatBuffer:
    """A channels_last [B, C_pad, H, W] concat target that producers fill slice by slice (inference only).
    C_pad rounds the channel count up to a multiple of `pad_to`; the pad channels are zero, and the layers that
    read the buffer use weights zero-padded to match (networks/submodules.py), so results are those of the
    unpadded concat."""

    def __init__(self, like, c_total, pad_to=8, shape=None):
        """like: a tensor giving batch, spatial size and device -- or, with shape=(B, H, W), only the device."""
        B, _, H, W = like.shape if shape is None else (shape[0], 0, shape[1], shape[2])
        self.c_total = c_total
        self.c_pad = -(-c_total // pad_to) * pad_to
        self.n_pixels = B * H * W
        with torch.cuda.device_of(like):
            self.tensor = torch.empty((B, self.c_pad, H, W), device=like.device, dtype=torch.float32,
                                      memory_format=torch.channels_last)
            if self.c_pad > c_total:
                check(_lib.load().flowops_fill_channels_nhwc(_p(self.tensor), self.n_pixels, self.c_pad, c_total,
                                                             self.c_pad - c_total, ctypes.c_float(0.0), _stream()), "fill_channels_nhwc")

    @classmethod
    def cached(cls, owner, tag, like, c_total, pad_to=8, shape=None):
        """The concat buffer of one place in the network, allocated (and its zero pad channels filled) once per shape and
        reused by every later forward: saves the allocation and the fill launch per call, and under CUDA-graph replay the
        fill is not replayed.  Inference-only, one forward at a time per module (the buffers are internal to a forward:
        every channel other than the padding is rewritten by the producers of the next one).  `owner` is the nn.Module
        the buffer belongs to; ConcatBuffer.drop_cached(module) frees them."""
        B, H, W = (like.shape[0], like.shape[2], like.shape[3]) if shape is None else shape
        key = (tag, B, H, W, c_total, pad_to, like.device)
        store = owner.__dict__.setdefault("_flowops_cbuf", {})
        buf = store.get(key)
        if buf is None:
            buf = store[key] = cls(like, c_total, pad_to, shape=shape)
        return buf

    @staticmethod
    def drop_cached(module):
        for m in module.modules():
            m.__dict__.pop("_flowops_cbuf", None)

    def copy_in(self, t, c_off):
        with torch.cuda.device_of(t):
            check(_lib.load().flowops_concat_nhwc(_p(t), _p(self.tensor), self.n_pixels, t.shape[1], self.c_pad, c_off, _stream()),
                  "concat_nhwc")
        return c_off + t.shape[1]

    def flow_deconv_in(self, flow, weight, bias, c_off):
        """dst[:, c_off : c_off + 2] = conv_transpose2d(flow, weight, bias, stride 2, padding 1) for a dense channels_last
        2-channel flow at half the buffer's resolution and a contiguous [2, 2, 4, 4] weight (the decoders' flow upsamplers)."""
        B, _, h, w = flow.shape
        with torch.cuda.device_of(flow):
            check(_lib.load().flowops_flow_deconv_nhwc_to(_p(flow), _p(weight), _p(bias), _p(self.tensor), B, h, w, self.c_pad, c_off,
                                                          self._pad_tail(c_off + 2, c_off), _stream()), "flow_deconv_nhwc_to")
        return c_off + 2

    def _pad_tail(self, c_end, c_off):
        """How many of the buffer's zero pad channels a 2-channel flow slice ending at channel c_end rewrites with it (2 or 6:
        whole 16-byte stores that complete the pixel record's last sector, include/flowops.h), or 0."""
        tail = self.c_pad - self.c_total if (D2S_WRITE_PAD and c_end == self.c_total and c_off % 4 == 0 and self.c_pad % 4 == 0) else 0
        return tail if tail in (2, 6) else 0

    def bias_lrelu_d2s_in(self, y4, bias, slope, c_off, flow_up=None):
        """dst[b, 2m+py, 2n+px, c_off + co] = LeakyReLU(y4[b, m, n, (py*2+px)*C + co] + bias[co]): the epilogue of a k4 s2 p1
        transposed convolution computed as a 3x3 convolution with 4*C output channels at the input resolution (one group
        of C per output parity) -- bias, activation and depth-to-space in one pass into the concat slice.  With
        flow_up = (flow, weight, bias) the level's 2-channel flow upsampler (see flow_deconv_in) is written by the same
        kernel into the two channels behind the slice."""
        B, C4, h, w = y4.shape
        with torch.cuda.device_of(y4):
            if flow_up is None:
                check(_lib.load().flowops_bias_lrelu_d2s_nhwc_to(_p(y4), _p(bias), _p(self.tensor), B, h, w, C4 // 4, self.c_pad, c_off,
                                                                 ctypes.c_float(slope), _stream()), "bias_lrelu_d2s_nhwc_to")
                return c_off + C4 // 4
            flow, fw, fb = flow_up
            if tuple(flow.shape) != (B, 2, h, w):
                raise ValueError("bias_lrelu_d2s_in: the flow must be [B, 2, h, w] at the deconvolution's input resolution")
            # the slice ends the buffer's real channels: its zero pad channels are rewritten together with the flow, which
            # completes the pixel record's last sector (include/flowops.h)
            tail = self._pad_tail(c_off + C4 // 4 + 2, c_off)
            check(_lib.load().flowops_bias_lrelu_d2s_flowup_nhwc_to(_p(y4), _p(bias), _p(self.tensor), B, h, w, C4 // 4, self.c_pad, c_off,
                                                                    ctypes.c_float(slope), _p(flow), _p(fw), _p(fb), tail, _stream()),
                  "bias_lrelu_d2s_flowup_nhwc_to")
        return c_off + C4 // 4 + 2

    def bias_lrelu_in(self, y, bias, slope, c_off, in_place_too=False):
        """dst[:, c_off : c_off + C] = LeakyReLU(y + bias) for a dense channels_last conv output y; with in_place_too
        y itself receives the result as well (one read, two writes)."""
        if not _is_nhwc(y) or y.shape[0] * y.shape[2] * y.shape[3] != self.n_pixels:
            raise ValueError("bias_lrelu_in: y must be dense channels_last with the buffer's batch and spatial shape")
        with torch.cuda.device_of(y):
            check(_lib.load().flowops_bias_lrelu_nhwc_to(_p(y), _p(bias), _p(self.tensor), self.n_pixels, y.shape[1], self.c_pad,
                                                         c_off, ctypes.c_float(slope), _p(y if in_place_too else None), _stream()),
                  "bias_lrelu_nhwc_to")
        return c_off + y.shape[1]


def cat_channels(tensors, pad_to=1):
    """torch.cat(tensors, dim=1).  When autograd is off and every input is a dense channels_last fp32 CUDA
    tensor the copy runs as one libflowops kernel per input; anything else goes to torch.cat.  With pad_to > 1 the
    fast path returns a tensor whose channel count is rounded up to a multiple of pad_to (zero channels)."""
    if not _cat_fast(tensors):
        return torch.cat(tensors, 1)
    buf = ConcatBuffer(tensors[0], sum(t.shape[1] for t in tensors), pad_to)
    off = 0
    for t in tensors:
        off = buf.copy_in(t, off)
    return buf.tensor


# ---------------------------------------------------------------------------------------------
# Correlation
# ---------------------------------------------------------------------------------------------
def correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(_lib.load().flowops_corr_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                             ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)), "corr_out_shape")
    return oc.value, oh.value, ow.value


def _workspace(nbytes, device):
    # torch.empty on CUDA is 512-byte aligned; the library asks for 256
    return torch.empty((max(nbytes, 1),), device=device, dtype=torch.uint8)


def correlation_has_16bit_path(in1, pad_size, kernel_size, max_displacement, stride1, stride2):
    """True when flowops_corr_fwd_16 takes this shape and parameter set (the FlowNetC configuration)."""
    B, C, H, W = in1.shape
    return _lib.load().flowops_corr_fwd_workspace_bytes(B, C, H, W, int(pad_size), int(kernel_size), int(max_displacement),
                                                        int(stride1), int(stride2)) > 0


def correlation_forward(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2):
    """fp32 tensors, or both fp16 / both bf16 in the FlowNetC configuration: `corr(a.float(), b.float()).half()`
    (FlowNetC.py:86-87) as one operator call."""
    dt = _io_dtype(in1)
    in1, in2 = _require(in1, "input1", dtype=dt), _require(in2, "input2", dtype=dt)
    if in1.shape != in2.shape or in1.device != in2.device:
        raise ValueError("input1 %s and input2 %s must have the same shape and device" % (tuple(in1.shape), tuple(in2.shape)))
    B, C, H, W = in1.shape
    params = (int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
    lib = _lib.load()
    if dt != torch.float32:
        in1, in2 = in1.contiguous(), in2.contiguous()
        oc, oh, ow = correlation_out_shape(H, W, *params)
        with torch.cuda.device_of(in1):
            out = torch.empty((B, oc, oh, ow), device=in1.device, dtype=dt)
            nbytes = lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *params)
            ws = _workspace(nbytes, in1.device)
            if in1.numel():
                check(lib.flowops_corr_fwd_16(_p(in1), _p(in2), _p(out), B, C, H, W, *params, _DTYPE16[dt], _p(ws), nbytes,
                                              _stream()), "corr_fwd_16")
        return out
    # channels_last features (a channels_last conv body) are taken as they are by the FlowNetC fast path
    layout = 1 if (_is_nhwc(in1) and _is_nhwc(in2) and lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *params) > 0) else 0
    if layout == 0:
        in1, in2 = in1.contiguous(), in2.contiguous()
    oc, oh, ow = correlation_out_shape(H, W, *params)
    with torch.cuda.device_of(in1):
        out = torch.empty((B, oc, oh, ow), device=in1.device, dtype=torch.float32)
        nbytes = lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *params)
        ws = _workspace(nbytes, in1.device)       # the role of rbot1/rbot2 (correlation.py:16-17)
        if in1.numel():
            check(lib.flowops_corr_fwd(_p(in1), _p(in2), _p(out), B, C, H, W, *params, layout, _p(ws), nbytes, _stream()),
                  "corr_fwd")
    return out


class CorrelationPlanes:
    """Workspace of one FlowNetC correlation whose input planes are written by the conv3 epilogues
    (flowops_corr_planes_from_conv) before `forward()` runs the correlation proper."""

    def __init__(self, shape, device, params=(20, 1, 20, 1, 2)):
        self.shape, self.params, self.device = tuple(shape), tuple(int(p) for p in params), device
        B, C, H, W = self.shape
        self.nbytes = _lib.load().flowops_corr_fwd_workspace_bytes(B, C, H, W, *self.params)
        if self.nbytes == 0:
            raise NotImplementedError("CorrelationPlanes: FlowNetC configuration only")
        with torch.cuda.device(device):
            self.ws = _workspace(self.nbytes, device)

    def fill_from_conv_(self, y, bias, slope, which, write_act):
        """y: bias-free conv output, dense channels_last.  Applies bias + LeakyReLU, fills input slot `which`;
        with write_act the activated features also replace y in place.  Returns y (activated) or None."""
        y = _require(y, "y")
        if tuple(y.shape) != self.shape or not _is_nhwc(y):
            raise ValueError("fill_from_conv_: expected a dense channels_last tensor of shape %s" % (self.shape,))
        B, C, H, W = self.shape
        with torch.cuda.device_of(y):
            check(_lib.load().flowops_corr_planes_from_conv(_p(y), _p(bias), ctypes.c_float(slope), _p(y if write_act else None),
                                                            which, B, C, H, W, *self.params, _p(self.ws), self.nbytes, _stream()),
                  "corr_planes_from_conv")
        return y if write_act else None

    def forward(self):
        B, C, H, W = self.shape
        oc, oh, ow = correlation_out_shape(H, W, *self.params)
        with torch.cuda.device(self.device):
            out = torch.empty((B, oc, oh, ow), device=self.device, dtype=torch.float32)
            check(_lib.load().flowops_corr_fwd_planes(_p(out), B, C, H, W, *self.params, _p(self.ws), self.nbytes, _stream()),
                  "corr_fwd_planes")
        return out


def correlation_planes_forward_into(planes, buf, c_off, slope):
    """The correlation proper, written channels-last into channels [c_off, c_off + 441) of the ConcatBuffer `buf`
    with LeakyReLU(slope) applied (module-level so that bench.py can time it)."""
    B, C, H, W = planes.shape
    if buf.n_pixels != B * H * W:
        raise ValueError("correlation_planes_forward_into: buffer shape does not match the correlation output")
    with torch.cuda.device(planes.device):
        check(_lib.load().flowops_corr_fwd_planes_nhwc(_p(buf.tensor), buf.c_pad, c_off, ctypes.c_float(slope), B, C, H, W,
                                                       *planes.params, _p(planes.ws), planes.nbytes, _stream()), "corr_fwd_planes_nhwc")
    return c_off + correlation_out_shape(H, W, *planes.params)[0]


def correlation_planes_forward(planes):
    """Module-level entry (so that bench.py can time the correlation proper)."""
    return planes.forward()


def correlation_backward(in1, in2, gout, pad_size, kernel_size, max_displacement, stride1, stride2,
                         need1=True, need2=True):
    in1 = _require(in1, "input1").contiguous()
    in2 = _require(in2, "input2").contiguous()
    gout = _require(gout, "grad_output").contiguous()
    B, C, H, W = in1.shape
    params = (int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
    lib = _lib.load()
    if tuple(gout.shape) != (B,) + correlation_out_shape(H, W, *params):
        raise ValueError("grad_output has shape %s" % (tuple(gout.shape),))
    with torch.cuda.device_of(in1):
        g1 = torch.empty_like(in1) if need1 else None
        g2 = torch.empty_like(in2) if need2 else None
        nbytes = lib.flowops_corr_bwd_workspace_bytes(B, C, H, W, *params)
        ws = _workspace(nbytes, in1.device)
        if in1.numel() and (need1 or need2):
            check(lib.flowops_corr_bwd(_p(in1), _p(in2), _p(gout), _p(g1), _p(g2), B, C, H, W, *params,
                                       _p(ws), nbytes, _stream()), "corr_bwd")
    return g1, g2


# ---------------------------------------------------------------------------------------------
# measurement helper
# ---------------------------------------------------------------------------------------------
def ffma_peak_tflops(iters=20000, reps=5):
    """Measured FP32-FMA pipe throughput of this GPU (TFLOP/s): the Correlation roofline denominator."""
    lib = _lib.load()
    sink = torch.zeros(4, device="cuda")
    flops = ctypes.c_double()
    best = 0.0
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.flowops_bench_ffma(_p(sink), iters, ctypes.byref(flops), _stream()), "bench_ffma")
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
