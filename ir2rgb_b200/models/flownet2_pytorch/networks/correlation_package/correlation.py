"""Drop-in for the reference's correlation_package/correlation.py (same names, argument order and
defaults: correlation.py:7-55), backed by libflowops.so instead of the `correlation_cuda` extension.

Differences from the reference wrapper are bug fixes only:
  * the integer hyper-parameters are kept on ``ctx`` instead of being passed to
    ``save_for_backward`` (correlation.py:13 raises TypeError on any modern torch when a gradient is
    required);
  * ``backward`` returns one gradient per ``forward`` argument (correlation.py:39 returns 2 of 8).
``corr_multiply`` is accepted and, as in the reference kernels, ignored.
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
        return _F.correlation_forward(input1, input2, *ctx.params)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_input1, grad_input2 = _F.correlation_backward(
            input1, input2, grad_output, *ctx.params,
            need1=ctx.needs_input_grad[0], need2=ctx.needs_input_grad[1])
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
