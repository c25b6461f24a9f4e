"""FlowNetSD (reference networks/FlowNetSD.py:11-106; 45,371,666 parameters), table-driven."""
import torch.nn as nn

from .submodules import Skip, add_layers, deconv, flow_upsampler, i_conv, predict_flow, refine, reference_init

ENCODER = [("conv0", 6, 64, 3, 1), ("conv1", 64, 64, 3, 2), ("conv1_1", 64, 128, 3, 1), ("conv2", 128, 128, 3, 2),
           ("conv2_1", 128, 128, 3, 1), ("conv3", 128, 256, 3, 2), ("conv3_1", 256, 256, 3, 1), ("conv4", 256, 512, 3, 2),
           ("conv4_1", 512, 512, 3, 1), ("conv5", 512, 512, 3, 2), ("conv5_1", 512, 512, 3, 1), ("conv6", 512, 1024, 3, 2),
           ("conv6_1", 1024, 1024, 3, 1)]
DECODER = {5: (1024, 512), 4: (1026, 256), 3: (770, 128), 2: (386, 64)}
INTER = {5: (1026, 512), 4: (770, 256), 3: (386, 128), 2: (194, 64)}
HEADS = {6: 1024, 5: 512, 4: 256, 3: 128, 2: 64}


class FlowNetSD(nn.Module):
    def __init__(self, args, batchNorm=True):
        super(FlowNetSD, self).__init__()
        self.batchNorm = batchNorm
        add_layers(self, batchNorm, ENCODER)
        for lv, (cin, cout) in DECODER.items():
            setattr(self, "deconv%d" % lv, deconv(cin, cout))
        for lv, (cin, cout) in INTER.items():
            setattr(self, "inter_conv%d" % lv, i_conv(batchNorm, cin, cout))
        for lv, cin in HEADS.items():
            setattr(self, "predict_flow%d" % lv, predict_flow(cin))
        for lv in (5, 4, 3, 2):
            setattr(self, "upsampled_flow%d_to_%d" % (lv + 1, lv), flow_upsampler())
        reference_init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')

    def forward(self, x):
        c1 = self.conv1_1(self.conv1(self.conv0(x)))
        sk = {lv: Skip(self, lv) for lv in (5, 4, 3, 2)}
        c2 = self.conv2_1(self.conv2(c1), skip=sk[2])
        c3 = self.conv3_1(self.conv3(c2), skip=sk[3])
        c4 = self.conv4_1(self.conv4(c3), skip=sk[4])
        c5 = self.conv5_1(self.conv5(c4), skip=sk[5])
        c6 = self.conv6_1(self.conv6(c5))
        flows = refine(self, {5: c5, 4: c4, 3: c3, 2: c2}, c6, (5, 4, 3, 2), inter=True, skip_bufs=sk)
        return tuple(flows) if self.training else (flows[0],)
