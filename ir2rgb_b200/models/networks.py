"""Drop-in for the flow-warp part of the reference's models/networks.py: ``get_grid``
(networks.py:15-28) and ``BaseCompositeGeneratorModule.{grid_sample, resample}`` (networks.py:85-100;
the identical ``Model.resample`` of models/base_model.py:123-136 lives in base_model.py here).

The reference builds a normalised sampling grid tensor with four elementwise kernels (two scalar
divides, a cat, an add), permutes it and calls ``F.grid_sample(bilinear, border)`` with the default
``align_corners=False``.  ``resample`` below produces the same values with ONE libflowops kernel that
reads the pixel-unit flow directly and never materialises the grid; its backward (flow and/or image
gradient, whichever autograd asks for) is one more kernel.

Everything else in the reference's networks.py (generator / discriminator nn.Modules) is stock
PyTorch layers outside the flow hot path and is not rebuilt here.
"""
from abc import ABC

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import functional as _F


def get_grid(batch_size, rows, cols, device='cuda:0', dtype=torch.float32):
    """Same signature and values as the reference's get_grid.  Kept for callers that want the grid
    tensor; ``resample`` does not need it."""
    hor = torch.linspace(-1.0, 1.0, cols).view(1, 1, 1, cols).expand(batch_size, 1, rows, cols)
    ver = torch.linspace(-1.0, 1.0, rows).view(1, 1, rows, 1).expand(batch_size, 1, rows, cols)
    return torch.cat([hor, ver], 1).to(dtype).to(device)


class GridSampleWarpFunction(Function):
    """image [b,c,h,w], flow [b,2,h,w] (pixels) -> warped image; vid2vid's resample as run."""

    @staticmethod
    def forward(ctx, image, flow):
        ctx.save_for_backward(image, flow)
        return _F.warp_forward(image, flow, _F.WARP_GRIDSAMPLE)

    @staticmethod
    def backward(ctx, grad_output):
        image, flow = ctx.saved_tensors
        grad_image, grad_flow = _F.warp_backward(
            image, flow, grad_output, need_img=ctx.needs_input_grad[0], need_flow=ctx.needs_input_grad[1],
            mode=_F.WARP_GRIDSAMPLE)
        return grad_image, grad_flow


def resample(image, flow):
    """Functional form of ``self.resample(image, flow)`` (networks.py:93-100).  fp16 inputs are
    computed in fp32 and cast back, as base_model.py:124-125 does."""
    if image.dtype != torch.float32 or flow.dtype != torch.float32:
        return GridSampleWarpFunction.apply(image.float(), flow.float()).to(image.dtype)
    if flow.device != image.device:
        flow = flow.cuda(image.get_device())        # the reference moves the grid to the image's GPU
    return GridSampleWarpFunction.apply(image, flow)


class BaseCompositeGeneratorModule(nn.Module, ABC):
    """Mixin base of the composite generators: only the warp members are reproduced."""

    def __init__(self):
        super(BaseCompositeGeneratorModule, self).__init__()

    @staticmethod
    def grid_sample(input1, input2):
        # stock ATen sampler on an explicit grid: not on the fast path, kept for API compatibility
        return torch.nn.functional.grid_sample(input1, input2, mode='bilinear', padding_mode='border')

    def resample(self, image, flow):
        return resample(image, flow)
