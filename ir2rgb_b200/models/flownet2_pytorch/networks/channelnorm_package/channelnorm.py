"""Drop-in for the reference's channelnorm_package/channelnorm.py (channelnorm.py:5-38), backed by
libflowops.so instead of the `channelnorm_cuda` extension.

``norm_deg`` is accepted and, as in the reference kernel, ignored (always the L2 norm).  The
reference's ``backward`` calls an undefined name (channelnorm.py:25); this one works.
"""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import functional as _F


class ChannelNormFunction(Function):

    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        assert input1.is_contiguous()
        output = _F.channelnorm_forward(input1)
        ctx.save_for_backward(input1, output)
        ctx.norm_deg = norm_deg
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        grad_input1 = _F.channelnorm_backward(input1, output, grad_output)
        return grad_input1, None


class ChannelNorm(Module):

    def __init__(self, norm_deg=2):
        super(ChannelNorm, self).__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        # .contiguous() is free for NCHW tensors and lets channels_last activations through
        return ChannelNormFunction.apply(input1.contiguous(), self.norm_deg)
