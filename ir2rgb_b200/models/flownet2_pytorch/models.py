"""FlowNet2 (reference models/flownet2_pytorch/models.py:30-161; 162,518,834 parameters): the caller
that fixes how often the hot-path operators run per frame pair -- 1 Correlation (inside FlowNetC),
4 Resample2d, 6 ChannelNorm (models.py:105-156).

Same constructor (``FlowNet2(args=None, batchNorm=False, div_flow=20., fp16=False)``), same attribute
names (a reference checkpoint loads with ``load_state_dict``), same arithmetic.  The conv/deconv body
is stock cuDNN; what is new is the glue between the sub-networks:

  * ``warp -> img0 - warped -> ChannelNorm`` (models.py:109-111,121-123,133-137,146-150) is one
    libflowops kernel (``flowops_warp_diff_norm_fwd``) that reads frame 0, frame 1 and the flow once
    and writes the warped frame and the brightness-error magnitude -- instead of three kernels, a
    ``.contiguous()`` copy of the ``x[:,3:]`` slice (resample2d.py:45) and two intermediate tensors.
    It is used when no gradient is required (how vid2vid runs FlowNet2, flownet.py:21); with
    autograd the three separate drop-in operators run, exactly as in the reference.
"""
import torch
import torch.nn as nn

from ... import functional as _F
from .networks import FlowNetC, FlowNetFusion, FlowNetS, FlowNetSD
from .networks.channelnorm_package.channelnorm import ChannelNorm
from .networks.resample2d_package.resample2d import Resample2d
from .networks import submodules as _sm
from .networks.submodules import reference_init

'Parameter count = 162,518,834'


class MyDict(dict):
    pass


_SD_STREAMS = {}      # device index -> the stream FlowNetSD runs on beside FlowNetC -> S -> S (FlowNet2._forward_fused)


class fp16_resample2d(nn.Module):
    def __init__(self):
        super(fp16_resample2d, self).__init__()
        self.resample = Resample2d()

    def forward(self, input1, input2):
        if (input1.dtype == torch.float16 and input2.dtype == torch.float16 and input1.is_cuda and input1.dim() == 4
                and input1.shape[1] <= 3):
            # the same values from one kernel on the fp16 tensors (flowops_warp_fwd_16), no widening / narrowing copies
            return self.resample(input1, input2)
        return self.resample(input1.float(), input2.float()).half()


class FlowNet2(nn.Module):

    def __init__(self, args=None, batchNorm=False, div_flow=20., fp16=False):
        super(FlowNet2, self).__init__()
        if args is None:
            args = MyDict()
            args.rgb_max = 1
            args.fp16 = fp16
            args.grads = {}
        self.fp16 = fp16
        self.batchNorm = batchNorm
        self.div_flow = div_flow
        self.rgb_max = args.rgb_max
        self.args = args
        self.fuse_glue = True          # use the fused warp/diff/norm kernel when autograd is off
        self.fuse_fusion_input = True  # channels_last body: concat3 (models.py:129-152) from one kernel
        self.fuse_upsample = True      # channels_last body: the x4 bilinear upsamplings folded into the concat kernels
        self.s2d_conv1 = True          # channels_last body, TF32 convolutions: FlowNetC.conv1 on space-to-depth frames
        self.sd_pad16 = True           # channels_last body, TF32 convolutions: FlowNetSD.conv0 reads a 16-channel tensor
        self.overlap_sd = False        # channels_last body: FlowNetSD (models.py:141-142) on a second stream beside C -> S -> S
                                       # (optional: measured, no gain on one B200 -- 639 pairs/s either way)
        self._sd_warm = set()          # (device, shape) pairs whose first (serial, plan-building) forward has run

        self.channelnorm = ChannelNorm()
        self.flownetc = FlowNetC.FlowNetC(args, batchNorm=self.batchNorm)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')
        self.flownets_1 = FlowNetS.FlowNetS(args, batchNorm=self.batchNorm)
        self.upsample2 = nn.Upsample(scale_factor=4, mode='bilinear')
        self.flownets_2 = FlowNetS.FlowNetS(args, batchNorm=self.batchNorm)
        self.flownets_d = FlowNetSD.FlowNetSD(args, batchNorm=self.batchNorm)
        self.upsample3 = nn.Upsample(scale_factor=4, mode='nearest')
        self.upsample4 = nn.Upsample(scale_factor=4, mode='nearest')
        self.resample = Resample2d() if not args.fp16 else fp16_resample2d()
        self.flownetfusion = FlowNetFusion.FlowNetFusion(args, batchNorm=self.batchNorm)
        reference_init(self)

    # -- glue ------------------------------------------------------------------------------------
    def warp_error(self, x, flow):
        """(warped frame 1, |frame 0 - warped frame 1|) for the 6-channel stack x = [frame0, frame1]."""
        fusable = (self.fuse_glue and not torch.is_grad_enabled() and x.dtype == torch.float32
                   and flow.dtype == torch.float32 and x.is_cuda)
        if fusable:
            return _F.warp_diff_norm_forward(x, flow)
        warped = self.resample(x[:, 3:, :, :], flow)
        return warped, self.channelnorm(x[:, :3, :, :] - warped)

    def _channels_last_body(self):
        w = self.flownets_1.conv1[0].weight
        return w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()

    def _forward_fused(self, inputs):
        """Inference on a channels_last conv body: the same computation as forward(), with the
        tensors between the sub-networks produced directly in the layout and channel count their consumers read --
        no torch.cat, no slice copies, no NCHW<->NHWC conversions, no cuDNN channel re-padding."""
        rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1)
        # FlowNetSD.conv0 (3x3, 6 -> 64 at full resolution): cuDNN's fused conv + bias + LeakyReLU engine takes 967 us per 16
        # pairs on the frame stack padded to 8 channels and 678 us on the same stack padded to 16 (tools/conv_pad_probe.py);
        # the extra zero channels add exact zeros.  Only with TF32 convolutions, where the fused engines run at all.
        sd_c = 16 if self.sd_pad16 and torch.backends.cudnn.allow_tf32 else 8
        if self.s2d_conv1 and torch.backends.cudnn.allow_tf32 and inputs.shape[3] % 2 == 0 and inputs.shape[4] % 2 == 0:
            # FlowNetC's first layer on the space-to-depth frames: a 16-channel 4x4 convolution instead of a 3(4)-channel
            # 7x7 stride-2 one, which cuDNN runs on a pre-Blackwell kernel without shared-memory staging (4.8 % of the
            # step).  Same sums in another order: on when TF32 convolutions are (the bit-exact path otherwise).
            x, xa, xb, x8 = _F.flownet2_prep_s2d(inputs, rgb_mean, float(self.rgb_max), sd_c)
            frames = (xa, xb, "s2d")
        else:
            x, xa, xb, x8 = _F.flownet2_prep(inputs, rgb_mean, float(self.rgb_max), sd_c)
            frames = (xa, xb)

        # FlowNetSD reads only the frame stack (models.py:141-142), so it can run beside FlowNetC -> S1 -> S2: forked onto a
        # second stream here, joined in front of the fusion-network input.  Inside a CUDA graph the two chains become
        # parallel branches, and the sub-networks' low-resolution layers (grids of 32 .. 256 CTAs on 148 SMs) can fill each
        # other's idle SMs.  Off by default (no measurable gain: the step is power- and bandwidth-bound, not occupancy-bound).
        # The first forward of a shape stays serial: cuDNN autotuning and the fused-plan timings of
        # networks/submodules.py run there and must not be disturbed by a concurrent stream.
        key = (x.device.index, tuple(inputs.shape), bool(torch.backends.cudnn.allow_tf32), bool(torch.backends.cudnn.benchmark))
        fork = self.overlap_sd and key in self._sd_warm
        self._sd_warm.add(key)
        main = torch.cuda.current_stream(x.device)
        if fork:
            side = _SD_STREAMS.get(x.device.index)
            if side is None:
                side = _SD_STREAMS[x.device.index] = torch.cuda.Stream(device=x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                flow2_sd = self.flownets_d(x8)[0]
            x8.record_stream(side)

        flow2_c = self.flownetc(x, frames=frames)[0]
        if self.fuse_upsample and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0:
            # `upsample(flow2 * div_flow)` (models.py:106,118) folded into the concat kernel's flow read
            concat1 = _F.warp_diff_norm_concat_up4(x, flow2_c, self.div_flow, self.div_flow)
            concat2 = _F.warp_diff_norm_concat_up4(x, self.flownets_1(concat1)[0], self.div_flow, self.div_flow)
        else:
            concat1 = _F.warp_diff_norm_concat(x, self.upsample1(flow2_c * self.div_flow), self.div_flow)
            flownets1_flow = self.upsample2(self.flownets_1(concat1)[0] * self.div_flow)
            concat2 = _F.warp_diff_norm_concat(x, flownets1_flow, self.div_flow)
        flow2_s2 = self.flownets_2(concat2)[0]
        if fork:
            main.wait_stream(side)
            flow2_sd.record_stream(main)
        else:
            flow2_sd = self.flownets_d(x8)[0]
        if self.fuse_fusion_input and _sm.PAD_CHANNELS > 1 and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0:
            # scalings, nearest upsamplings, both ChannelNorms, both warp-error chains and the concat of models.py:129-152
            # as one kernel writing the 16-channel channels-last tensor conv0 of the fusion network reads
            return self.flownetfusion(_F.flownet2_fusion_input(x, flow2_s2, flow2_sd, self.div_flow))
        flownets2_flow = self.upsample4(flow2_s2 * self.div_flow)
        norm_flownets2_flow = self.channelnorm(flownets2_flow)
        _, diff_flownets2_img1 = self.warp_error(x, flownets2_flow)

        flownetsd_flow = self.upsample3(flow2_sd / self.div_flow)
        norm_flownetsd_flow = self.channelnorm(flownetsd_flow)
        _, diff_flownetsd_img1 = self.warp_error(x, flownetsd_flow)

        concat3 = torch.cat((x[:, :3, :, :], flownetsd_flow, flownets2_flow, norm_flownetsd_flow, norm_flownets2_flow,
                             diff_flownetsd_img1, diff_flownets2_img1), dim=1)
        return self.flownetfusion(concat3)

    def forward(self, inputs):
        if (self.fuse_glue and not torch.is_grad_enabled() and inputs.is_cuda and inputs.dtype == torch.float32
                and not self.fp16 and not self.batchNorm and inputs.dim() == 5 and self._channels_last_body()):
            return self._forward_fused(inputs)
        rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1,))
        x = (inputs - rgb_mean) / self.rgb_max
        x = torch.cat((x[:, :, 0, :, :], x[:, :, 1, :, :]), dim=1)

        # FlowNetC -> warp -> FlowNetS1 -> warp -> FlowNetS2 (models.py:104-137)
        flownetc_flow = self.upsample1(self.flownetc(x)[0] * self.div_flow)
        resampled_img1, norm_diff_img0 = self.warp_error(x, flownetc_flow)
        concat1 = torch.cat((x, resampled_img1, flownetc_flow / self.div_flow, norm_diff_img0), dim=1)

        flownets1_flow = self.upsample2(self.flownets_1(concat1)[0] * self.div_flow)
        resampled_img1, norm_diff_img0 = self.warp_error(x, flownets1_flow)
        concat2 = torch.cat((x, resampled_img1, flownets1_flow / self.div_flow, norm_diff_img0), dim=1)

        flownets2_flow = self.upsample4(self.flownets_2(concat2)[0] * self.div_flow)
        norm_flownets2_flow = self.channelnorm(flownets2_flow)
        _, diff_flownets2_img1 = self.warp_error(x, flownets2_flow)

        # FlowNetSD branch (models.py:141-150; note the division by div_flow, reproduced as is)
        flownetsd_flow = self.upsample3(self.flownets_d(x)[0] / self.div_flow)
        norm_flownetsd_flow = self.channelnorm(flownetsd_flow)
        _, diff_flownetsd_img1 = self.warp_error(x, flownetsd_flow)

        concat3 = torch.cat((x[:, :3, :, :], flownetsd_flow, flownets2_flow, norm_flownetsd_flow, norm_flownets2_flow,
                             diff_flownetsd_img1, diff_flownets2_img1), dim=1)
        return self.flownetfusion(concat3)


class FlowNet2C(FlowNetC.FlowNetC):
    def __init__(self, args, batchNorm=False, div_flow=20):
        super(FlowNet2C, self).__init__(args, batchNorm=batchNorm, div_flow=20)
