"""ctypes binding of libflowops.so (include/flowops.h).  No torch C++ extension, no dispatch layer.

The library is built in-tree by ``python -m ir2rgb_b200.build`` (``__graft_entry__.build()`` does it).
If it is missing, loading fails loudly: there is no CPU or PyTorch fallback for the hot path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOWOPS_LIB") or os.path.join(_HERE, "libflowops.so")   # env override: development builds

WARP_RESAMPLE2D = 0
WARP_GRIDSAMPLE = 1
DTYPE_F16 = 1
DTYPE_BF16 = 2

_c_float_p = ctypes.c_void_p   # device pointers travel as plain addresses
_int = ctypes.c_int
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t

# name -> (restype, argtypes); every symbol include/flowops.h declares
SIGNATURES = {
    "flowops_version": (_int, []),
    "flowops_last_error": (ctypes.c_char_p, []),
    "flowops_cnorm_fwd": (_int, [_vp, _vp, _int, _int, _int, _int, _vp]),
    "flowops_cnorm_bwd": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "flowops_warp_fwd": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "flowops_warp_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "flowops_corr_out_shape": (_int, [_int] * 7 + [ctypes.POINTER(_int)] * 3),
    "flowops_corr_fwd_workspace_bytes": (_sz, [_int] * 9),
    "flowops_corr_bwd_workspace_bytes": (_sz, [_int] * 9),
    "flowops_corr_fwd": (_int, [_vp, _vp, _vp] + [_int] * 10 + [_vp, _sz, _vp]),
    "flowops_corr_planes_from_conv": (_int, [_vp, _vp, ctypes.c_float, _vp, _int] + [_int] * 9 + [_vp, _sz, _vp]),
    "flowops_corr_fwd_planes": (_int, [_vp] + [_int] * 9 + [_vp, _sz, _vp]),
    "flowops_corr_fwd_planes_nhwc": (_int, [_vp, _int, _int, ctypes.c_float] + [_int] * 9 + [_vp, _sz, _vp]),
    "flowops_warp_set_impl": (_int, [_int]),
    "flowops_warp_get_impl": (_int, []),
    "flowops_corr_set_impl": (_int, [_int]),
    "flowops_corr_get_impl": (_int, []),
    "flowops_corr_tc_trace": (_int, [_vp]),
    "flowops_corr_bwd": (_int, [_vp, _vp, _vp, _vp, _vp] + [_int] * 9 + [_vp, _sz, _vp]),
    "flowops_cnorm_fwd_16": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "flowops_cnorm_bwd_16": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "flowops_warp_fwd_16": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _vp, _int, _vp]),
    "flowops_corr_fwd_16": (_int, [_vp, _vp, _vp] + [_int] * 10 + [_vp, _sz, _vp]),
    "flowops_warp_diff_norm_fwd": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _sz, _int, _int, _int, _int, _vp]),
    "flowops_warp_conf_fwd": (_int, [_vp, _vp, _vp, _vp, ctypes.c_float, _int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "flowops_warp_diff_norm_concat_nhwc": (_int, [_vp, _vp, ctypes.c_float, _vp, _int, _int, _int, _int, _vp]),
    "flowops_warp_diff_norm_concat_up4_nhwc": (_int, [_vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _int, _int, _int, _int, _vp]),
    "flowops_flownet2_fusion_input_nhwc": (_int, [_vp, _vp, _vp, ctypes.c_float, _vp, _int, _int, _int, _int, _vp]),
    "flowops_flownet2_prep": (_int, [_vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _int, _int, _int, _vp]),
    "flowops_flownet2_prep_s2d": (_int, [_vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _int, _int, _int, _vp]),
    "flowops_flownet2_prep_pitched": (_int, [_vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "flowops_bias_lrelu": (_int, [_vp, _vp, _int, _int, _int, _int, ctypes.c_float, _vp]),
    "flowops_bias_lrelu_nhwc_to": (_int, [_vp, _vp, _vp, _sz, _int, _int, _int, ctypes.c_float, _vp, _vp]),
    "flowops_fill_channels_nhwc": (_int, [_vp, _sz, _int, _int, _int, ctypes.c_float, _vp]),
    "flowops_bias_lrelu_d2s_nhwc_to": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, ctypes.c_float, _vp]),
    "flowops_bias_lrelu_d2s_flowup_nhwc_to": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, ctypes.c_float, _vp, _vp, _vp, _int, _vp]),
    "flowops_flow_head_nhwc": (_int, [_vp, _int, _int, _vp, _vp, _vp, _int, _int, _int, _vp]),
    "flowops_flow_deconv_nhwc_to": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _vp]),
    "flowops_concat_nhwc": (_int, [_vp, _vp, _sz, _int, _int, _int, _vp]),
    "flowops_bench_ffma": (_int, [_vp, _int, ctypes.POINTER(ctypes.c_double), _vp]),
}

_lib = None


class FlowopsError(RuntimeError):
    pass


def load():
    """Load libflowops.so once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlowopsError(
            "libflowops.so not found at %s -- build it with `python -m ir2rgb_b200.build` "
            "(needs nvcc; there is no CPU fallback for the flow hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.flowops_version() != 1:
        raise FlowopsError("libflowops.so version mismatch")
    _lib = lib
    return lib


# kernels (and memsets) each library call launches, keyed by the name the wrappers pass to check();
# bench.py's `gpu_launches` is counted from this table through `launch_hook`
KERNELS_PER_CALL = {"corr_fwd": 2, "corr_fwd_16": 2, "corr_bwd": 6, "warp_bwd": 2}
launch_hook = None          # callable(what, n_kernels) or None


def kernels_per_call(what):
    n = KERNELS_PER_CALL.get(what, 1)
    if what == "corr_fwd" and _lib is not None and (_lib.flowops_corr_get_impl() & 1):
        n = 3          # tensor-core path of an eligible shape: layout pass, UMMA kernel, NCHW store pass
    return n


def check(rc, what):
    """Turn a non-zero return into a RuntimeError, like the reference's AT_ERROR
    (correlation_cuda.cc:81-83)."""
    if launch_hook is not None and rc == 0:
        launch_hook(what, kernels_per_call(what))
    if rc != 0:
        msg = load().flowops_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError("%s: %s" % (what, msg))
        raise FlowopsError("%s failed (%d): %s" % (what, rc, msg))
