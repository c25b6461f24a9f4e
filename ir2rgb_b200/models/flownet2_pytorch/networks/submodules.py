"""Layer factories of the FlowNet2 family (counterpart of the reference's networks/submodules.py:7-38).

The conv body of FlowNet2 is stock cuDNN work and is NOT part of the hand-written hot path; it is
restated here only because BASELINE config 4 (FlowNet2 frame pairs/s) needs the network around the
three custom operators.  nn.Sequential indices are kept (``conv1.0.weight`` ...) so that a FlowNet2
checkpoint of the reference loads unchanged.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import cudnn_fused as _cf
from .... import functional as _F

LEAK = 0.1
FUSE_EPILOGUE = True      # inference only: conv -> one libflowops pass for bias + LeakyReLU


PAD_CHANNELS = 8          # inference only: concat buffers are rounded up to this many channels (zero-filled)


def padded_weight(conv, cin):
    """conv.weight zero-padded along its input-channel axis to `cin` channels, cached on the module.  Lets a layer
    read a concat buffer whose channel count was rounded up (cuDNN pads odd channel counts -- 1026, 770, 386, 194,
    473 ... -- with a kernel of its own on every call; the extra zero channels add exact zeros to each sum)."""
    w = conv.weight
    axis = 0 if isinstance(conv, nn.ConvTranspose2d) else 1
    if w.shape[axis] == cin:
        return w
    if torch.is_grad_enabled() and w.requires_grad:
        # differentiable zero-padding: the cached copy below is detached and would give conv.weight no gradient
        shape = list(w.shape)
        shape[axis] = cin - w.shape[axis]
        return torch.cat((w, w.new_zeros(shape)), axis)
    # the cache is keyed on the tensor's version counter: in-place updates through `.data` (w.data.copy_(), the
    # reference's init_deconv_bilinear) do NOT bump it -- call reset_padded_weights(net) after such an update
    # (load_state_dict goes through Tensor.copy_ and is seen)
    key = (w.data_ptr(), w._version, cin)
    cache = conv.__dict__.get("_flowops_wpad")
    if cache is None or cache[0] != key:
        shape = list(w.shape)
        shape[axis] = cin - w.shape[axis]
        wp = torch.cat((w.detach(), w.new_zeros(shape)), axis)
        if w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous():
            wp = wp.contiguous(memory_format=torch.channels_last)
        cache = (key, wp)
        conv.__dict__["_flowops_wpad"] = cache
    return cache[1]


FLOW_HEAD_KERNEL = os.environ.get("FLOWOPS_FLOW_HEAD", "1") != "0"      # inference, TF32 convolutions allowed: predict_flow layers
                                                                      # through flowops_flow_head_nhwc (the variable: A/B timing)


# The kernel wins where the head is bandwidth-bound -- the fusion network's full- and half-resolution heads (16 / 32 channels:
# 200 vs 387 us and 108 vs 131 us per 16 pairs) -- and loses on the wide, low-resolution decoder levels, where it is bound by
# the shared-memory reads of its weights and cuDNN's tensor-op kernels are not (194 channels: 222 vs 160 us;
# profiles/step_probe_flow_heads_r02.json, profiles/ncu_flow_head_c16_r02.txt).
FLOW_HEAD_MAX_CHANNELS = 32


def _flow_head_ok(conv, x):
    return (isinstance(conv, nn.Conv2d) and conv.out_channels == 2 and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1)
            and tuple(conv.padding) == (1, 1) and tuple(conv.dilation) == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros"
            and x.shape[1] % 4 == 0 and conv.in_channels <= x.shape[1] <= FLOW_HEAD_MAX_CHANNELS and _F._is_nhwc_view(x) and x.stride(3) % 4 == 0
            and x.data_ptr() % 16 == 0 and x.shape[0] <= 65535)


def flow_head_weight(conv, cin):
    """conv.weight of a flow head packed for flowops_flow_head_nhwc (zero for the pad channels of a `cin`-channel input),
    cached on the module under the same key -- and with the same `.data` caveat -- as padded_weight."""
    w = conv.weight
    key = (w.data_ptr(), w._version, cin)
    cache = conv.__dict__.get("_flowops_whead")
    if cache is None or cache[0] != key:
        cache = (key, _F.pack_flow_head_weight(w, cin))
        conv.__dict__["_flowops_whead"] = cache
    return cache[1]


def reset_padded_weights(net):
    """Drop every cached tensor derived from the weights of `net` (needed after weight updates made through `.data`)
    and the cached concat buffers."""
    for m in net.modules():
        for k in ("_flowops_wpad", "_flowops_wdense", "_flowops_cbuf", "_flowops_conv1_s2d", "_flowops_wconv3", "_flowops_whead"):
            m.__dict__.pop(k, None)
        if isinstance(m.__dict__.get("_sd_warm"), set):
            m._sd_warm.clear()


CACHE_CONCAT_BUFFERS = True      # inference: one concat buffer per place and shape, reused across forwards


def _new_buffer(owner, tag, like, c_total, shape=None):
    if CACHE_CONCAT_BUFFERS:
        return _F.ConcatBuffer.cached(owner, tag, like, c_total, PAD_CHANNELS, shape=shape)
    return _F.ConcatBuffer(like, c_total, PAD_CHANNELS, shape=shape)


def _conv_out_hw(conv, x):
    return tuple((x.shape[2 + i] + 2 * conv.padding[i] - conv.dilation[i] * (conv.kernel_size[i] - 1) - 1) // conv.stride[i] + 1
                 for i in range(2))


def _raw_conv(conv, x, bias):
    w = padded_weight(conv, x.shape[1])
    if isinstance(conv, nn.ConvTranspose2d):
        return F.conv_transpose2d(x, w, bias, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
    return F.conv2d(x, w, bias, conv.stride, conv.padding, conv.dilation, conv.groups)


def apply_conv(mod, x):
    """mod(x) where x may carry zero pad channels beyond mod's in_channels (mod: ConvAct, Conv2d, ConvTranspose2d,
    or an nn.Sequential starting with one of the two)."""
    if isinstance(mod, ConvAct):
        return mod(x)
    conv = mod[0] if isinstance(mod, nn.Sequential) else mod
    cin = conv.weight.shape[0 if isinstance(conv, nn.ConvTranspose2d) else 1] * conv.groups
    if (FUSE_EPILOGUE and conv.bias is not None and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32
            and isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d)) and not conv.__dict__.get("_flowops_stock", False)):
        # a bare convolution with bias (the 2-channel flow heads, the flow upsamplers, inter_conv): ATen adds the bias
        # with a generic strided elementwise kernel after cuDNN's bias-free convolution; the library's epilogue with
        # slope 1 (t > 0 ? t : t * 1 == t) is the same add, bit for bit, in one vectorised pass.  (cuDNN's fused
        # conv -> bias engines lose on every one of these layers -- narrow outputs, wide inputs: the fusion network's inter_conv0
        # 3059 us fused vs 961 us, its flow head 592 vs 382; profiles/step_probe_r02.json -- so they are not tried here.)
        if FLOW_HEAD_KERNEL and torch.backends.cudnn.allow_tf32 and _flow_head_ok(conv, x):
            # the 2-channel flow heads as one direct FP32 convolution with the bias included (csrc/flowhead.cu) instead of
            # cuDNN's convolution between two channel-padding kernels + the bias pass.  Exact FP32 sums, so it stands in for
            # the TF32 convolution only; with TF32 off the cuDNN fp32 convolution below keeps the round-1 summation order
            y = _F.flow_head(x, flow_head_weight(conv, x.shape[1]), conv.bias)
        else:
            y = _raw_conv(conv, x, None)
            if y.is_contiguous() or _F._is_nhwc(y):
                _F.bias_lrelu_(y, conv.bias, 1.0)
            else:
                y += conv.bias.view(1, -1, 1, 1)
    elif x.shape[1] == cin:
        return mod(x)
    else:
        y = _raw_conv(conv, x, conv.bias)
    if isinstance(mod, nn.Sequential):
        for layer in list(mod)[1:]:
            y = layer(y)
    return y


class ConvAct(nn.Sequential):
    """(Conv2d | ConvTranspose2d) + LeakyReLU with the reference's nn.Sequential indexing (so `conv1.0.weight`
    keeps its name).  Under no_grad on CUDA fp32 the bias add and the activation run as one libflowops kernel
    after a bias-free convolution -- bit-identical to conv(+bias) -> LeakyReLU -- either in place, or (`into`)
    straight into a channel slice of the concat buffer the next layers read."""

    def fusable(self, x):
        return (FUSE_EPILOGUE and len(self) == 2 and self[0].bias is not None and not torch.is_grad_enabled()
                and x.is_cuda and x.dtype == torch.float32)

    def forward(self, x, into=None, skip=None, flow_up=None):
        conv = self[0]
        if self.fusable(x):
            sbuf = None
            if isinstance(conv, nn.Conv2d) and torch.backends.cudnn.allow_tf32 and _F._is_nhwc_view(x) and _cf.available():
                # convolution + bias + LeakyReLU as ONE cuDNN runtime-fusion launch (TF32 math, as the unfused cuDNN
                # convolution under the same setting) where that is faster than the two-kernel path -- decided per layer
                # by timing both once; with TF32 off the bit-exact two-kernel path below runs
                slope = self[1].negative_slope
                out = None
                if into is not None:
                    buf, c_off = into
                    out = buf.tensor[:, c_off:c_off + conv.out_channels] if c_off % 4 == 0 else None
                if into is None or out is not None:
                    if skip is not None:
                        # also a decoder skip connection: the fused launch writes straight into the level's concat
                        # buffer, and the layers that consume this output read that channel slice in place (the cuDNN
                        # graph path takes strided tensors; a stock torch convolution makes its own dense copy)
                        sbuf = _new_buffer(self, "skip", x, skip.c_total(conv.out_channels),
                                           shape=(x.shape[0],) + _conv_out_hw(conv, x))
                        out = sbuf.tensor[:, :conv.out_channels]

                    def unfused():
                        t = _raw_conv(conv, x, None)
                        if into is not None:
                            into[0].bias_lrelu_in(t, conv.bias, slope, into[1])
                        elif skip is not None:
                            sbuf.bias_lrelu_in(t, conv.bias, slope, 0, in_place_too=True)
                        else:
                            _F.bias_lrelu_(t, conv.bias, slope)
                    y = _cf.conv_bias_lrelu(conv, x, padded_weight(conv, x.shape[1]), slope, out, unfused)
                    if y is not NotImplemented:
                        if skip is not None:
                            skip.buf = sbuf
                        return None if into is not None else y
            if (isinstance(conv, nn.ConvTranspose2d) and into is not None and DECONV_AS_CONV3 and torch.backends.cudnn.allow_tf32
                    and _F._is_nhwc(x) and _deconv_as_conv3_ok(conv, into[1])):
                # narrow transposed convolution (<= 32 output channels: cuDNN's strided-dgrad kernel runs it at a fraction
                # of the tensor-op rate) as a 3x3 convolution with 4 * C output channels + depth-to-space epilogue, where
                # timing both once says it is faster
                if _deconv_conv3_run(self, conv, x, into, flow_up):
                    return None
            y = _raw_conv(conv, x, None)
            if into is not None:
                buf, c_off = into
                buf.bias_lrelu_in(y, conv.bias, self[1].negative_slope, c_off)
                return None
            if skip is not None and _F._is_nhwc(y):
                # this output is also a decoder skip connection: allocate that level's concat buffer now and write the
                # activated features to both places from one read
                skip.buf = sbuf if sbuf is not None else _new_buffer(self, "skip", y, skip.c_total(y.shape[1]))
                skip.buf.bias_lrelu_in(y, conv.bias, self[1].negative_slope, 0, in_place_too=True)
                return y
            return _F.bias_lrelu_(y, conv.bias, self[1].negative_slope)
        assert into is None
        if x.shape[1] != conv.weight.shape[0 if isinstance(conv, nn.ConvTranspose2d) else 1] * conv.groups:
            y = _raw_conv(conv, x, conv.bias)          # zero pad channels in x: weights padded to match
            for layer in list(self)[1:]:
                y = layer(y)
            return y
        return super().forward(x)


DECONV_AS_CONV3 = True


def _deconv_as_conv3_ok(conv, c_off):
    return (conv.groups == 1 and tuple(conv.kernel_size) == (4, 4) and tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (1, 1)
            and tuple(conv.output_padding) == (0, 0) and tuple(conv.dilation) == (1, 1) and conv.out_channels <= 32
            and conv.out_channels % 4 == 0 and c_off % 4 == 0 and conv.bias is not None)


def deconv_as_conv3_weight(conv, w):
    """ConvTranspose2d(k4, s2, p1) weight w [Cin, C, 4, 4] (possibly zero-padded along Cin) -> the [4*C, Cin, 3, 3] weight of
    the 3x3 convolution that forms all four output parities at the input resolution:
        out[2m+py] = in[m-1] w[3] + in[m] w[1]  (py = 0),   in[m] w[2] + in[m+1] w[0]  (py = 1)        (same along x)
    so parity (py, px), channel co is output channel (py*2+px)*C + co with taps (dr, ky) in {(-1,3),(0,1)} / {(0,2),(1,0)};
    the five taps a parity does not use are zero.  Cached on the module (channels_last)."""
    key = (w.data_ptr(), w._version, tuple(w.shape))
    cache = conv.__dict__.get("_flowops_wconv3")
    if cache is None or cache[0] != key:
        cin, c = w.shape[0], w.shape[1]
        w3 = w.new_zeros(4 * c, cin, 3, 3)
        taps = {0: ((0, 3), (1, 1)), 1: ((1, 2), (2, 0))}          # py -> ((dr + 1, ky), ...)
        wt = w.detach().permute(1, 0, 2, 3)                        # [C, Cin, 4, 4]
        for py in (0, 1):
            for px in (0, 1):
                g = (py * 2 + px) * c
                for r, ky in taps[py]:
                    for q, kx in taps[px]:
                        w3[g:g + c, :, r, q] = wt[:, :, ky, kx]
        cache = (key, w3.contiguous(memory_format=torch.channels_last))
        conv.__dict__["_flowops_wconv3"] = cache
    return cache[1]


def _deconv_conv3_run(act, conv, x, into, flow_up=None):
    """Run (and, the first time per shape, time against strided dgrad + epilogue) the 3x3-convolution form; False when the
    plain path is faster for this layer.  flow_up: {"flow", "weight", "bias", "done"} of the level's flow upsampler -- the
    depth-to-space epilogue then writes the two upsampled flow channels behind the slice too and sets "done"."""
    buf, c_off = into
    slope = act[1].negative_slope
    w3 = deconv_as_conv3_weight(conv, padded_weight(conv, x.shape[1]))
    fu = None
    if flow_up is not None and tuple(flow_up["flow"].shape[2:]) == tuple(x.shape[2:]) and (c_off + conv.out_channels) % 2 == 0:
        fu = (flow_up["flow"], flow_up["weight"], flow_up["bias"])

    def as_conv3():
        buf.bias_lrelu_d2s_in(F.conv2d(x, w3, None, 1, 1), conv.bias, slope, c_off, fu)
    key = (tuple(x.shape), buf.c_pad, c_off)
    choice = conv.__dict__.setdefault("_flowops_conv3_choice", {})
    if key not in choice:
        if torch.cuda.is_current_stream_capturing():
            return False

        def plain():
            buf.bias_lrelu_in(_raw_conv(conv, x, None), conv.bias, slope, c_off)
        choice[key] = _cf._time3(as_conv3) < 0.97 * _cf._time3(plain)
    if not choice[key]:
        return False
    as_conv3()
    if fu is not None:
        flow_up["done"] = True
    return True


class Skip:
    """Hand-over of a decoder level's concat buffer from the encoder layer that produces the skip tensor
    (ConvAct.forward(skip=...)) to refine()."""

    def __init__(self, net, lv):
        self.extra = getattr(net, "deconv%d" % lv)[0].out_channels + 2      # deconv features + upsampled flow
        self.buf = None

    def c_total(self, c_skip):
        return c_skip + self.extra


def conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1):
    layers = [nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, bias=not batchNorm)]
    if batchNorm:
        layers.append(nn.BatchNorm2d(out_planes))
    layers.append(nn.LeakyReLU(LEAK, inplace=True))
    return ConvAct(*layers)


def i_conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1, bias=True):
    layers = [nn.Conv2d(in_planes, out_planes, kernel_size, stride, (kernel_size - 1) // 2, bias=bias)]
    if batchNorm:
        layers.append(nn.BatchNorm2d(out_planes))
    return nn.Sequential(*layers)


def predict_flow(in_planes):
    return nn.Conv2d(in_planes, 2, 3, 1, 1, bias=True)


def deconv(in_planes, out_planes):
    return ConvAct(nn.ConvTranspose2d(in_planes, out_planes, 4, 2, 1, bias=True), nn.LeakyReLU(LEAK, inplace=True))


def flow_upsampler(bias=True):
    return nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=bias)


def reference_init(module):
    """xavier_uniform weights, U(0,1) biases for every (transposed) convolution -- the initialisation
    every FlowNet2 sub-network applies to itself (e.g. FlowNetS.py:50-59)."""
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                nn.init.uniform_(m.bias)
            nn.init.xavier_uniform_(m.weight)


class tofp16(nn.Module):
    def forward(self, input):
        return input.half()


class tofp32(nn.Module):
    def forward(self, input):
        return input.float()


def add_layers(net, batchNorm, table):
    """table rows: (attribute name, in, out, kernel, stride) -> net.<name> = conv(...)."""
    for name, cin, cout, k, s in table:
        setattr(net, name, conv(batchNorm, cin, cout, kernel_size=k, stride=s))


FUSE_FLOW_UPSAMPLER = True


def _flow_upsampler_ok(conv, flow):
    return (isinstance(conv, nn.ConvTranspose2d) and conv.in_channels == 2 and conv.out_channels == 2 and conv.groups == 1
            and tuple(conv.kernel_size) == (4, 4) and tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (1, 1)
            and tuple(conv.output_padding) == (0, 0) and tuple(conv.dilation) == (1, 1)
            and flow.dtype == torch.float32 and flow.is_cuda and flow.shape[1] == 2
            and flow.is_contiguous(memory_format=torch.channels_last) and not conv.__dict__.get("_flowops_stock", False))


def _dense_weight(conv):
    """conv.weight as a plain contiguous [Cin, Cout, kh, kw] tensor (a channels_last module keeps a permuted one), cached."""
    w = conv.weight
    if w.is_contiguous():
        return w
    key = (w.data_ptr(), w._version)
    cache = conv.__dict__.get("_flowops_wdense")
    if cache is None or cache[0] != key:
        cache = (key, w.detach().contiguous())
        conv.__dict__["_flowops_wdense"] = cache
    return cache[1]


def refine(net, skips, top, levels, inter=False, skip_bufs=None):
    """Shared coarse-to-fine decoder of FlowNetC / FlowNetS / FlowNetSD (e.g. FlowNetS.py:70-90):
    at each level predict a flow, upsample it and the features, concatenate with the skip tensor.
    Returns the flows from finest to coarsest.

    Inference on a channels_last body: the concat buffer is allocated first (channel count rounded up to a
    multiple of PAD_CHANNELS, pad channels zero), the skip tensor and the upsampled flow are copied in, and the
    deconvolution's bias + LeakyReLU epilogue writes its slice directly."""
    feat = top
    flows = [apply_conv(getattr(net, "predict_flow%d" % (levels[0] + 1)), top)]
    for lv in levels:
        upconv = getattr(net, "upsampled_flow%d_to_%d" % (lv + 1, lv))
        deconv_lv = getattr(net, "deconv%d" % lv)
        skip = skips[lv]
        # the 2-channel flow upsampler as one libflowops kernel writing its slice of the concat buffer (inference,
        # channels_last, dense 2-channel flow); otherwise the cuDNN transposed convolution + copy
        buf = skip_bufs[lv].buf if skip_bufs is not None and skip_bufs.get(lv) is not None else None
        skip_ok = buf is not None or _F._cat_fast((skip,))      # already in its slice (possibly returned as a view of it)
        direct_up = (FUSE_FLOW_UPSAMPLER and deconv_lv.fusable(feat) and _F._is_nhwc(feat) and _flow_upsampler_ok(upconv, flows[0])
                     and skip_ok)
        up = None if direct_up else apply_conv(upconv, flows[0])
        if deconv_lv.fusable(feat) and _F._is_nhwc(feat) and skip_ok and (direct_up or _F._cat_fast((up,))):
            c_dec = deconv_lv[0].out_channels
            if buf is not None:                 # the encoder already wrote the skip tensor into its slice
                off = skip.shape[1]
            else:
                buf = _new_buffer(deconv_lv, "level", skip, skip.shape[1] + c_dec + 2)
                off = buf.copy_in(skip, 0)
            deconv_lv(feat, into=(buf, off))
            if direct_up and (off + c_dec) % 2 == 0:
                buf.flow_deconv_in(flows[0], _dense_weight(upconv), upconv.bias, off + c_dec)
            else:
                buf.copy_in(up if up is not None else apply_conv(upconv, flows[0]), off + c_dec)
            feat = buf.tensor
        else:
            feat = _F.cat_channels((skip, deconv_lv(feat), up if up is not None else apply_conv(upconv, flows[0])))
        head_in = apply_conv(getattr(net, "inter_conv%d" % lv), feat) if inter else feat
        flows.insert(0, apply_conv(getattr(net, "predict_flow%d" % lv), head_in))
    return flows
