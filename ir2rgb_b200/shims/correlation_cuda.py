"""`correlation_cuda` with the signature of the reference's pybind11 module (correlation_cuda.cc:36-167,169-172),
forwarding to libflowops.so through ctypes.  Tensors are fp32, contiguous, on the current device -- what the
reference's kernels silently assume (correlation_cuda_kernel.cu:58-68 ignores the strides it is passed)."""
import ctypes

import torch

from .. import _lib

_i = ctypes.c_int


def _check(rc):
    if rc:
        # AT_ERROR("CUDA call failed") -> RuntimeError, correlation_cuda.cc:81-83
        raise RuntimeError(_lib.load().flowops_last_error().decode("utf-8", "replace"))


def _workspace(rbot1, nbytes):
    # rbot1 (the reference's padded NHWC scratch, correlation_cuda.cc:39-42) doubles as the library's workspace;
    # rbot2 stays empty.  256-byte alignment comes from the caching allocator.
    rbot1.resize_((nbytes + 3) // 4 + 64)
    off = (-rbot1.data_ptr()) % 256
    return ctypes.c_void_p(rbot1.data_ptr() + off), nbytes


def forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2,
            corr_multiply):
    lib = _lib.load()
    B, C, H, W = input1.shape
    p = (int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
    oc, oh, ow = _i(), _i(), _i()
    _check(lib.flowops_corr_out_shape(H, W, *p, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)))
    output.resize_(B, oc.value, oh.value, ow.value)                     # correlation_cuda.cc:36-38; no fill_ needed
    ws, n = _workspace(rbot1, lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *p))
    _check(lib.flowops_corr_fwd(input1.data_ptr(), input2.data_ptr(), output.data_ptr(), B, C, H, W, *p,
                                0,                                      # FLOWOPS_LAYOUT_NCHW
                                ws, n, torch.cuda.current_stream().cuda_stream))
    return 1


def backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size, kernel_size,
             max_displacement, stride1, stride2, corr_multiply):
    lib = _lib.load()
    B, C, H, W = input1.shape
    p = (int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
    grad_input1.resize_(B, C, H, W)                                     # correlation_cuda.cc:112-113
    grad_input2.resize_(B, C, H, W)
    ws, n = _workspace(rbot1, lib.flowops_corr_bwd_workspace_bytes(B, C, H, W, *p))
    _check(lib.flowops_corr_bwd(input1.data_ptr(), input2.data_ptr(), grad_output.data_ptr(),
                                grad_input1.data_ptr(), grad_input2.data_ptr(), B, C, H, W, *p,
                                ws, n, torch.cuda.current_stream().cuda_stream))
    return 1
