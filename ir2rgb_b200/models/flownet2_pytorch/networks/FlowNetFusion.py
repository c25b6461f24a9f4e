"""FlowNetFusion (reference networks/FlowNetFusion.py:11-67; 581,226 parameters), table-driven."""
import torch
import torch.nn as nn

from .... import functional as _F
from . import submodules as _sm
from .submodules import Skip, add_layers, apply_conv, deconv, flow_upsampler, i_conv, predict_flow, reference_init

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

    def _concat(self, sk, skip, deconv_lv, feat, upconv, flow):
        """torch.cat((skip, deconv(feat), up), 1) (FlowNetFusion.py:54,60).  Inference on a channels_last body: the encoder
        layer already wrote `skip` into the level's concat buffer (sk.buf, channels rounded up to a multiple of 8:
        162 -> 168, 82 -> 88), the deconvolution's epilogue writes its slice, only the 2-channel flow is copied."""
        if sk.buf is not None and deconv_lv.fusable(feat) and _F._is_nhwc(feat):
            off = skip.shape[1]
            c_up = off + deconv_lv[0].out_channels
            fuse_up = _sm.FUSE_FLOW_UPSAMPLER and c_up % 2 == 0 and _sm._flow_upsampler_ok(upconv, flow)
            fu = {"flow": flow, "weight": _sm._dense_weight(upconv), "bias": upconv.bias, "done": False} if fuse_up else None
            deconv_lv(feat, into=(sk.buf, off), flow_up=fu)     # the depth-to-space epilogue takes the flow upsampler along
            if fuse_up and fu["done"]:
                pass
            elif fuse_up:
                sk.buf.flow_deconv_in(flow, _sm._dense_weight(upconv), upconv.bias, c_up)     # one kernel, straight into its slice
            else:
                sk.buf.copy_in(apply_conv(upconv, flow), c_up)
            return sk.buf.tensor
        return _F.cat_channels((skip, apply_conv(deconv_lv, feat), apply_conv(upconv, flow)), pad_to=_sm.PAD_CHANNELS)

    def forward(self, x):
        sk0, sk1 = Skip(self, 0), Skip(self, 1)
        c0 = self.conv0(x, skip=sk0)
        c1 = self.conv1_1(self.conv1(c0), skip=sk1)
        c2 = self.conv2_1(self.conv2(c1))
        flow2 = apply_conv(self.predict_flow2, c2)
        cat1 = self._concat(sk1, c1, self.deconv1, c2, self.upsampled_flow2_to_1, flow2)
        flow1 = apply_conv(self.predict_flow1, apply_conv(self.inter_conv1, cat1))
        cat0 = self._concat(sk0, c0, self.deconv0, cat1, self.upsampled_flow1_to_0, flow1)
        return apply_conv(self.predict_flow0, apply_conv(self.inter_conv0, cat0))
