"""The part of the reference's models/base_model.py that lies on the flow hot path: the ``Model`` base
(``opt`` dict, ``Tensor``, ``save_dir``: base_model.py:9-15) with ``grid_sample`` / ``resample``
(base_model.py:123-136).  Checkpoint loading, pyramids and LR schedules are orchestration outside the
path."""
import os
from abc import ABC, abstractmethod

import torch

from . import networks


class Model(torch.nn.Module, ABC):

    def __init__(self, **opt):
        super(Model, self).__init__()
        self.opt = opt
        self.Tensor = torch.cuda.FloatTensor if opt.get('gpu_ids') else torch.LongTensor
        self.save_dir = os.path.join(opt.get('checkpoints_dir', '.'), opt.get('name', 'ir2rgb'))

    @abstractmethod
    def save(self, label):
        pass

    def grid_sample(self, input1, input2):
        if self.opt.get('fp16'):
            return torch.nn.functional.grid_sample(input1.float(), input2.float(), mode='bilinear',
                                                   padding_mode='border').half()
        return torch.nn.functional.grid_sample(input1, input2, mode='bilinear', padding_mode='border')

    def resample(self, image, flow):
        if (self.opt.get('fp16') and image.dtype == torch.float16 and flow.dtype == torch.float16 and image.is_cuda
                and image.dim() == 4 and image.shape[1] <= 3 and image.shape[2] > 1 and image.shape[3] > 1):
            if not (torch.is_grad_enabled() and (image.requires_grad or flow.requires_grad)):
                # base_model.py:123-136 as run on fp16 tensors -- grid arithmetic in the flow's dtype, fp32 sampling,
                # `.half()` -- in one kernel (flowops_warp_fwd_16, mode GRIDSAMPLE)
                return networks._F.warp_forward(image, flow, networks._F.WARP_GRIDSAMPLE)
            return self._resample_fp16_autograd(image, flow)
        return networks.resample(image, flow)

    def _resample_fp16_autograd(self, image, flow):
        """The reference's own op chain (base_model.py:129-136), for the rare case that an fp16 warp needs gradients:
        its 16-bit grid arithmetic has no fp32 backward kernel to borrow."""
        b, c, h, w = image.size()
        grid = networks.get_grid(b, h, w, device=flow.get_device(), dtype=flow.dtype)
        flow = torch.cat([flow[:, 0:1, :, :] / ((w - 1.0) / 2.0), flow[:, 1:2, :, :] / ((h - 1.0) / 2.0)], dim=1)
        final_grid = (grid + flow).permute(0, 2, 3, 1).cuda(image.get_device())
        return self.grid_sample(image, final_grid)
