"""Drop-in for the reference's resample2d_package/resample2d.py (resample2d.py:5-46), backed by
libflowops.so instead of the `resample2d_cuda` extension.

Same contract: ``input1`` (image) and ``input2`` (flow, pixels, ch0 = dx) must be contiguous in
``Resample2dFunction`` (the Module makes the image contiguous first, resample2d.py:45).  Only
``kernel_size == 1`` exists: larger kernels read out of bounds in the reference and are never used.
The gradient outputs are allocated by the library call (no separate zero-fill pass as in
resample2d.py:17,29-30), and a gradient that is not needed is not computed.
"""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import functional as _F


class Resample2dFunction(Function):

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        if kernel_size != 1:
            raise NotImplementedError("Resample2d: only kernel_size=1 is defined (reference reads out of "
                                      "bounds for larger kernels, resample2d_kernel.cu:53-58)")
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        return _F.warp_forward(input1, input2, _F.WARP_RESAMPLE2D)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_input1, grad_input2 = _F.warp_backward(
            input1, input2, grad_output.contiguous(),
            need_img=ctx.needs_input_grad[0], need_flow=ctx.needs_input_grad[1], mode=_F.WARP_RESAMPLE2D)
        return grad_input1, grad_input2, None


class Resample2d(Module):

    def __init__(self, kernel_size=1):
        super(Resample2d, self).__init__()
        self.kernel_size = kernel_size

    def forward(self, input1, input2):
        input1_c = input1.contiguous()
        return Resample2dFunction.apply(input1_c, input2.contiguous(), self.kernel_size)
