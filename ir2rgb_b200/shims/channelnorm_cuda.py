"""`channelnorm_cuda` with the signature of the reference's pybind11 module (channelnorm_cuda.cc:9-31).  Dispatches on
the scalar type the way AT_DISPATCH_FLOATING_TYPES_AND_HALF does in channelnorm_kernel.cu:111,152."""
import torch

from .. import _lib


def _check(rc):
    if rc:
        raise RuntimeError(_lib.load().flowops_last_error().decode("utf-8", "replace"))


_DT16 = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def forward(input1, output, norm_deg):
    B, C, H, W = input1.shape
    st = torch.cuda.current_stream().cuda_stream
    lib = _lib.load()
    if input1.dtype == torch.float32:
        _check(lib.flowops_cnorm_fwd(input1.data_ptr(), output.data_ptr(), B, C, H, W, st))
    else:
        _check(lib.flowops_cnorm_fwd_16(input1.data_ptr(), output.data_ptr(), B, C, H, W, _DT16[input1.dtype], st))
    return 1


def backward(input1, output, grad_output, grad_input1, norm_deg):
    B, C, H, W = input1.shape
    st = torch.cuda.current_stream().cuda_stream
    lib = _lib.load()
    if input1.dtype == torch.float32:
        _check(lib.flowops_cnorm_bwd(input1.data_ptr(), output.data_ptr(), grad_output.data_ptr(),
                                     grad_input1.data_ptr(), B, C, H, W, st))
    else:
        _check(lib.flowops_cnorm_bwd_16(input1.data_ptr(), output.data_ptr(), grad_output.data_ptr(),
                                        grad_input1.data_ptr(), B, C, H, W, _DT16[input1.dtype], st))
    return 1
