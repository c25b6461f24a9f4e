"""Drop-in for the reference's resample2d_package/resample2d.py (resample2d.py:5-46), backed by
libflowops.so instead of the `resample2d_cuda` extension.

Same contract: ``input1`` (image) and ``input2`` (flow, pixels, ch0 = dx) must be contiguous in
``Resample2dFunction`` (the Module makes the image contiguous first, resample2d.py:45).  Only
``kernel_size == 1`` exists: larger kernels read out of bounds in the reference and are never used.
The gradient outputs are allocated by the library call (no separate zero-fill pass as in
resample2d.py:17,29-30), and a gradient that is not needed is not computed.

fp16 / bf16 tensors (image and flow of the same dtype, CUDA) are accepted as well: the forward is then the
16-bit-storage kernel (flowops_warp_fwd_16), which equals ``Resample2d()(a.float(), b.float()).to(dtype)`` --
the reference's fp16_resample2d (models.py:22-28) -- bit for bit, and the backward is what autograd makes of
that chain (casts around the fp32 backward).
"""
import torch
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
        dtype = input1.dtype
        if dtype == torch.float64 or (dtype in (torch.float16, torch.bfloat16) and input1.shape[1] > 3):
            # the reference kernel is dispatched for Half / double and any channel count (resample2d_kernel.cu:213);
            # the one-pass 16-bit kernel covers the 1..3-channel frames of the hot path, everything else takes the
            # fp32 kernel between two casts -- the arithmetic the reference's fp16 mode runs anyway (models.py:22-28)
            return _F.warp_forward(input1.float(), input2.float(), _F.WARP_RESAMPLE2D).to(dtype)
        return _F.warp_forward(input1, input2, _F.WARP_RESAMPLE2D)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        dtype = input1.dtype
        grad_input1, grad_input2 = _F.warp_backward(
            input1.float(), input2.float(), grad_output.float().contiguous(),      # no-ops for fp32 tensors
            need_img=ctx.needs_input_grad[0], need_flow=ctx.needs_input_grad[1], mode=_F.WARP_RESAMPLE2D)
        if dtype != torch.float32:                                                 # 16-bit storage: cast back
            grad_input1 = grad_input1.to(dtype) if grad_input1 is not None else None
            grad_input2 = grad_input2.to(dtype) if grad_input2 is not None else None
        return grad_input1, grad_input2, None


class Resample2d(Module):

    def __init__(self, kernel_size=1):
        super(Resample2d, self).__init__()
        self.kernel_size = kernel_size

    def forward(self, input1, input2):
        input1_c = input1.contiguous()
        return Resample2dFunction.apply(input1_c, input2.contiguous(), self.kernel_size)
