"""Layer factories of the FlowNet2 family (counterpart of the reference's networks/submodules.py:7-38).

The conv body of FlowNet2 is stock cuDNN work and is NOT part of the hand-written hot path; it is
restated here only because BASELINE config 4 (FlowNet2 frame pairs/s) needs the network around the
three custom operators.  nn.Sequential indices are kept (``conv1.0.weight`` ...) so that a FlowNet2
checkpoint of the reference loads unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import functional as _F

LEAK = 0.1
FUSE_EPILOGUE = True      # inference only: conv -> one libflowops pass for bias + LeakyReLU


class ConvAct(nn.Sequential):
    """(Conv2d | ConvTranspose2d) + LeakyReLU with the reference's nn.Sequential indexing (so `conv1.0.weight`
    keeps its name).  Under no_grad on CUDA fp32 the bias add and the activation run as one in-place
    libflowops kernel after a bias-free convolution -- bit-identical to conv(+bias) -> LeakyReLU."""

    def forward(self, x):
        conv = self[0]
        if (FUSE_EPILOGUE and len(self) == 2 and conv.bias is not None and not torch.is_grad_enabled()
                and x.is_cuda and x.dtype == torch.float32):
            if isinstance(conv, nn.ConvTranspose2d):
                y = F.conv_transpose2d(x, conv.weight, None, conv.stride, conv.padding, conv.output_padding,
                                       conv.groups, conv.dilation)
            else:
                y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
            return _F.bias_lrelu_(y, conv.bias, self[1].negative_slope)
        return super().forward(x)


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


def refine(net, skips, top, levels, inter=False):
    """Shared coarse-to-fine decoder of FlowNetC / FlowNetS / FlowNetSD (e.g. FlowNetS.py:70-90):
    at each level predict a flow, upsample it and the features, concatenate with the skip tensor.
    Returns the flows from finest to coarsest."""
    feat = top
    flows = [getattr(net, "predict_flow%d" % (levels[0] + 1))(top)]
    for lv in levels:
        up = getattr(net, "upsampled_flow%d_to_%d" % (lv + 1, lv))(flows[0])
        dec = getattr(net, "deconv%d" % lv)(feat)
        feat = _F.cat_channels((skips[lv], dec, up))
        head_in = getattr(net, "inter_conv%d" % lv)(feat) if inter else feat
        flows.insert(0, getattr(net, "predict_flow%d" % lv)(head_in))
    return flows
