"""Drop-in for the reference's correlation_package/correlation.py (same names, argument order and
defaults: correlation.py:7-55), backed by libflowops.so instead of the `correlation_cuda` extension.

Differences from the reference wrapper are bug fixes only:
  * the integer hyper-parameters are kept on ``ctx`` instead of being passed to
    ``save_for_backward`` (correlation.py:13 raises TypeError on any modern torch when a gradient is
    required);
  * ``backward`` returns one gradient per ``forward`` argument (correlation.py:39 returns 2 of 8).
``corr_multiply`` is accepted and, as in the reference kernels, ignored.

fp16 / bf16 inputs (both of the same dtype) run ``corr(a.float(), b.float()).to(dtype)`` -- what FlowNetC does
in fp16 mode (FlowNetC.py:86-87) -- as one operator call in the FlowNetC configuration (flowops_corr_fwd_16);
the backward is that chain's: casts around the fp32 backward.
"""
import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import functional as _F


class CorrelationFunction(Function):

    @staticmethod
    def forward(ctx, input1, input2,
                pad_size=3, kernel_size=3,
                max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        ctx.save_for_backward(input1, input2)
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2)
        if input1.dtype != torch.float32 and not _F.correlation_has_16bit_path(input1, *ctx.params):
            # 16-bit storage outside the FlowNetC configuration: the literal cast chain around the generic kernel
            return _F.correlation_forward(input1.float(), input2.float(), *ctx.params).to(input1.dtype)
        return _F.correlation_forward(input1, input2, *ctx.params)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        dtype = input1.dtype
        grad_input1, grad_input2 = _F.correlation_backward(
            input1.float(), input2.float(), grad_output.float(), *ctx.params,      # no-ops for fp32 tensors
            need1=ctx.needs_input_grad[0], need2=ctx.needs_input_grad[1])
        if dtype != torch.float32:
            grad_input1 = grad_input1.to(dtype) if grad_input1 is not None else None
            grad_input2 = grad_input2.to(dtype) if grad_input2 is not None else None
        return grad_input1, grad_input2, None, None, None, None, None, None


class Correlation(Module):
    """nn.Module front end; hyper-parameters are constructor constants (FlowNetC.py:31)."""
    _FIELDS = ("pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply")

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        for name, value in zip(self._FIELDS, (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)):
            setattr(self, name, value)

    def extra_repr(self):
        return ", ".join("%s=%s" % (f, getattr(self, f)) for f in self._FIELDS)

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, *(getattr(self, f) for f in self._FIELDS))
