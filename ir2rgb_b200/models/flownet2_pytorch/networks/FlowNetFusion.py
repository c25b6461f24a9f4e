"""FlowNetFusion (reference networks/FlowNetFusion.py:11-67; 581,226 parameters), table-driven."""
import torch
import torch.nn as nn

from .... import functional as _F
from . import submodules as _sm
from .submodules import add_layers, apply_conv, deconv, flow_upsampler, i_conv, predict_flow, reference_init

ENCODER = [("conv0", 11, 64, 3, 1), ("conv1", 64, 64, 3, 2), ("conv1_1", 64, 128, 3, 1), ("conv2", 128, 128, 3, 2),
           ("conv2_1", 128, 128, 3, 1)]


class FlowNetFusion(nn.Module):
    def __init__(self, args, batchNorm=True):
        super(FlowNetFusion, self).__init__()
        self.batchNorm = batchNorm
        add_layers(self, batchNorm, ENCODER)
        self.deconv1, self.deconv0 = deconv(128, 32), deconv(162, 16)
        self.inter_conv1, self.inter_conv0 = i_conv(batchNorm, 162, 32), i_conv(batchNorm, 82, 16)
        self.predict_flow2, self.predict_flow1, self.predict_flow0 = predict_flow(128), predict_flow(32), predict_flow(16)
        self.upsampled_flow2_to_1, self.upsampled_flow1_to_0 = flow_upsampler(), flow_upsampler()
        reference_init(self)

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1_1(self.conv1(c0))
        c2 = self.conv2_1(self.conv2(c1))
        flow2 = self.predict_flow2(c2)
        # inference: the concat buffers carry zero pad channels up to a multiple of 8 (162 -> 168, 82 -> 88)
        cat1 = _F.cat_channels((c1, self.deconv1(c2), self.upsampled_flow2_to_1(flow2)), pad_to=_sm.PAD_CHANNELS)
        flow1 = self.predict_flow1(apply_conv(self.inter_conv1, cat1))
        cat0 = _F.cat_channels((c0, apply_conv(self.deconv0, cat1), self.upsampled_flow1_to_0(flow1)), pad_to=_sm.PAD_CHANNELS)
        return self.predict_flow0(apply_conv(self.inter_conv0, cat0))
