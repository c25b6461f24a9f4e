"""FlowNetC (reference networks/FlowNetC.py:13-131; 39,175,298 parameters), table-driven.

The one caller of the Correlation operator on the hot path (FlowNetC.py:31,89): both frames go
through a shared conv1-3 tower, the 256-channel conv3 features are correlated (pad 20, k 1, md 20,
s1 1, s2 2 -> 441 channels), and the cost volume joins a 32-channel redirect of frame-1 features.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import functional as _F
from . import submodules as _sm
from .correlation_package.correlation import Correlation
from .submodules import Skip, add_layers, deconv, flow_upsampler, predict_flow, refine, reference_init

TOWER = [("conv1", 3, 64, 7, 2), ("conv2", 64, 128, 5, 2), ("conv3", 128, 256, 5, 2), ("conv_redir", 256, 32, 1, 1)]
TRUNK = [("conv3_1", 473, 256, 3, 1), ("conv4", 256, 512, 3, 2), ("conv4_1", 512, 512, 3, 1), ("conv5", 512, 512, 3, 2),
         ("conv5_1", 512, 512, 3, 1), ("conv6", 512, 1024, 3, 2), ("conv6_1", 1024, 1024, 3, 1)]
DECODER = {5: (1024, 512), 4: (1026, 256), 3: (770, 128), 2: (386, 64)}
HEADS = {6: 1024, 5: 1026, 4: 770, 3: 386, 2: 194}


class FlowNetC(nn.Module):
    def __init__(self, args, batchNorm=True, div_flow=20):
        super(FlowNetC, self).__init__()
        self.fp16 = args.fp16
        self.batchNorm = batchNorm
        self.div_flow = div_flow
        add_layers(self, batchNorm, TOWER)
        self.corr = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
        self.corr_activation = nn.LeakyReLU(0.1, inplace=True)
        add_layers(self, batchNorm, TRUNK)
        for lv, (cin, cout) in DECODER.items():
            setattr(self, "deconv%d" % lv, deconv(cin, cout))
        for lv, cin in HEADS.items():
            setattr(self, "predict_flow%d" % lv, predict_flow(cin))
        for lv in (5, 4, 3, 2):
            setattr(self, "upsampled_flow%d_to_%d" % (lv + 1, lv), flow_upsampler(bias=True))
        reference_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')

    def conv1_s2d(self):
        """conv1 (3 -> 64, 7x7, stride 2, FlowNetC.py:18) as a 4x4 stride-1 convolution over the space-to-depth frame
        (functional.flownet2_prep_s2d): W'[co][(py*2+px)*4 + c][t][u] = W[co][c][2t+py-1][2u+px-1].  A ConvAct that shares
        conv1's bias, kept outside the module tree (the state dict is the reference's), rebuilt when conv1's weight changes."""
        w = self.conv1[0].weight
        key = (w.data_ptr(), w._version)
        cache = self.__dict__.get("_flowops_conv1_s2d")
        if cache is None or cache[0] != key:
            w7 = w.detach().contiguous()                                     # [64, 3, 7, 7]
            w8 = torch.zeros(w7.shape[0], 3, 8, 8, device=w.device, dtype=w.dtype)
            w8[:, :, 1:, 1:] = w7                                            # index k + 1: the tap "-1" is the zero at 0
            # w8[co, c, 2t + py, 2u + px] -> [co, py, px, c, t, u]
            w6 = w8.view(w7.shape[0], 3, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4)
            ws = torch.zeros(w7.shape[0], 2, 2, 4, 4, 4, device=w.device, dtype=w.dtype)
            ws[:, :, :, :3] = w6
            conv = nn.Conv2d(16, w7.shape[0], 4, 1, 1, bias=True).to(w.device)
            conv.weight = nn.Parameter(ws.view(w7.shape[0], 16, 4, 4).contiguous(memory_format=torch.channels_last), requires_grad=False)
            conv.bias = self.conv1[0].bias
            cache = (key, _sm.ConvAct(conv, nn.LeakyReLU(_sm.LEAK, inplace=True)))
            self.__dict__["_flowops_conv1_s2d"] = cache
        return cache[1]

    def tower(self, frame):
        c2 = self.conv2(self.conv1(frame))
        return c2, self.conv3(c2)

    def _fused_features_and_cost(self, x, frames=None):
        """Inference fast path on a channels_last conv body: conv3's bias + LeakyReLU epilogue writes the
        correlation's input planes directly (flowops_corr_planes_from_conv), so neither a separate activation
        pass nor the correlation's own layout pre-pass runs.  Same values as the plain path, bit for bit."""
        conv3, act3 = self.conv3[0], self.conv3[1]
        fa, fb = frames[:2] if frames is not None else (x[:, 0:3], x[:, 3:])      # frames: channels-last, zero-padded to 4 channels
        conv1 = self.conv1_s2d() if frames is not None and len(frames) > 2 and frames[2] == "s2d" else self.conv1
        c2a = self.conv2(conv1(fa))
        c2b = self.conv2(conv1(fb))
        y3a = F.conv2d(c2a, conv3.weight, None, conv3.stride, conv3.padding)
        y3b = F.conv2d(c2b, conv3.weight, None, conv3.stride, conv3.padding)
        if not (_F._is_nhwc(y3a) and _F._is_nhwc(y3b)):
            return None
        planes = _F.CorrelationPlanes(y3a.shape, y3a.device)
        c3a = planes.fill_from_conv_(y3a, conv3.bias, act3.negative_slope, 0, write_act=True)    # also feeds conv_redir
        planes.fill_from_conv_(y3b, conv3.bias, act3.negative_slope, 1, write_act=False)         # only the correlation reads it
        return c2a, c3a, planes

    def forward(self, x, frames=None):
        fused = None
        if (_sm.FUSE_EPILOGUE and not torch.is_grad_enabled() and not self.fp16 and not self.batchNorm and x.is_cuda
                and x.dtype == torch.float32 and self.conv3[0].weight.is_contiguous(memory_format=torch.channels_last)
                and not self.conv3[0].weight.is_contiguous()):
            fused = self._fused_features_and_cost(x, frames)
        if fused is not None:
            # the concat of FlowNetC.py:94 is allocated first (473 -> 480 channels, zero pad); conv_redir's epilogue and
            # the correlation (with corr_activation folded into its store) each write their channel slice
            c2a, c3a, planes = fused
            buf = _sm._new_buffer(self.conv_redir, "corr", c3a, self.conv_redir[0].out_channels + 441)
            off = self.conv_redir[0].out_channels
            self.conv_redir(c3a, into=(buf, 0))
            _F.correlation_planes_forward_into(planes, buf, off, self.corr_activation.negative_slope)
            cat = buf.tensor
        else:
            c2a, c3a = self.tower(x[:, 0:3])
            _, c3b = self.tower(x[:, 3:])
            if self.fp16 and c3a.dtype == torch.float16 and c3b.dtype == torch.float16:
                cost = self.corr(c3a, c3b)      # corr(a.float(), b.float()).half() (FlowNetC.py:86-87) as one operator call
            elif self.fp16:
                cost = self.corr(c3a.float(), c3b.float()).half()
            else:
                cost = self.corr(c3a, c3b)
            cat = torch.cat((self.conv_redir(c3a), self.corr_activation(cost)), 1)
        sk = {lv: Skip(self, lv) for lv in (5, 4, 3)}        # c2a is produced inside the tower: copied by refine()
        c3 = self.conv3_1(cat, skip=sk[3])
        c4 = self.conv4_1(self.conv4(c3), skip=sk[4])
        c5 = self.conv5_1(self.conv5(c4), skip=sk[5])
        c6 = self.conv6_1(self.conv6(c5))
        flows = refine(self, {5: c5, 4: c4, 3: c3, 2: c2a}, c6, (5, 4, 3, 2), skip_bufs=sk)
        return tuple(flows) if self.training else (flows[0],)
