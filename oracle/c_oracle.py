"""ctypes binding of oracle/flowops_oracle.c (numpy float32 in / out).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "flowops_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "liboracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: the only fused multiply-adds are the explicit fmaf() calls."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           _SRC, "-o", _SO, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_abi_version.restype = ctypes.c_int
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


_I = ctypes.c_int


def cnorm_fwd(x):
    x = _f32(x)
    B, C, H, W = x.shape
    y = np.empty((B, 1, H, W), np.float32)
    lib().oracle_cnorm_fwd(_p(x), _p(y), _I(B), _I(C), _I(H), _I(W))
    return y


def cnorm_bwd(x, y, gy):
    x, y, gy = _f32(x), _f32(y), _f32(gy)
    B, C, H, W = x.shape
    gx = np.empty_like(x)
    lib().oracle_cnorm_bwd(_p(x), _p(y), _p(gy), _p(gx), _I(B), _I(C), _I(H), _I(W))
    return gx


def resample2d_fwd(img, flow):
    img, flow = _f32(img), _f32(flow)
    B, C, H, W = img.shape
    assert flow.shape == (B, 2, H, W)
    out = np.empty_like(img)
    lib().oracle_resample2d_fwd(_p(img), _p(flow), _p(out), _I(B), _I(C), _I(H), _I(W))
    return out


def resample2d_bwd(img, flow, gout):
    img, flow, gout = _f32(img), _f32(flow), _f32(gout)
    B, C, H, W = img.shape
    gimg = np.zeros_like(img)
    gflow = np.empty_like(flow)
    lib().oracle_resample2d_bwd_img(_p(flow), _p(gout), _p(gimg), _I(B), _I(C), _I(H), _I(W))
    lib().oracle_resample2d_bwd_flow(_p(img), _p(flow), _p(gout), _p(gflow), _I(B), _I(C), _I(H), _I(W))
    return gimg, gflow


def corr_shape(H, W, pad, k, md, s1, s2):
    oc, oh, ow = _I(), _I(), _I()
    lib().oracle_corr_shape(_I(H), _I(W), _I(pad), _I(k), _I(md), _I(s1), _I(s2),
                            ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
    return oc.value, oh.value, ow.value


def corr_fwd(a, b, pad=20, k=1, md=20, s1=1, s2=2):
    a, b = _f32(a), _f32(b)
    B, C, H, W = a.shape
    oc, oh, ow = corr_shape(H, W, pad, k, md, s1, s2)
    out = np.empty((B, oc, oh, ow), np.float32)
    lib().oracle_corr_fwd(_p(a), _p(b), _p(out), _I(B), _I(C), _I(H), _I(W),
                          _I(pad), _I(k), _I(md), _I(s1), _I(s2))
    return out


def corr_bwd(a, b, gout, pad=20, k=1, md=20, s1=1, s2=2):
    a, b, gout = _f32(a), _f32(b), _f32(gout)
    B, C, H, W = a.shape
    ga, gb = np.empty_like(a), np.empty_like(b)
    lib().oracle_corr_bwd(_p(a), _p(b), _p(gout), _p(ga), _p(gb), _I(B), _I(C), _I(H), _I(W),
                          _I(pad), _I(k), _I(md), _I(s1), _I(s2))
    return ga, gb


def gridwarp_fwd(img, flow, lin_x, lin_y, inv_mode=1, fma_mode=1):
    img, flow, lin_x, lin_y = _f32(img), _f32(flow), _f32(lin_x), _f32(lin_y)
    B, C, H, W = img.shape
    out = np.empty_like(img)
    lib().oracle_gridwarp_fwd(_p(img), _p(flow), _p(out), _p(lin_x), _p(lin_y),
                              _I(B), _I(C), _I(H), _I(W), _I(inv_mode), _I(fma_mode))
    return out


# ---- 16-bit storage (values travel as float32 arrays holding representable values) ----
DTYPE_F16, DTYPE_BF16 = 1, 2


def round16(a, dtype):
    a = _f32(a)
    out = np.empty_like(a)
    lib().oracle_round16(_p(a), _p(out), ctypes.c_longlong(a.size), _I(dtype))
    return out


def cnorm_fwd_16(x, dtype):
    x = _f32(x)
    B, C, H, W = x.shape
    y = np.empty((B, 1, H, W), np.float32)
    lib().oracle_cnorm_fwd_16(_p(x), _p(y), _I(B), _I(C), _I(H), _I(W), _I(dtype))
    return y


def cnorm_bwd_16(x, y, gy, dtype):
    x, y, gy = _f32(x), _f32(y), _f32(gy)
    B, C, H, W = x.shape
    gx = np.empty_like(x)
    lib().oracle_cnorm_bwd_16(_p(x), _p(y), _p(gy), _p(gx), _I(B), _I(C), _I(H), _I(W), _I(dtype))
    return gx


def gridwarp_fwd_16(img, flow, lin_x, lin_y, dtype, inv_mode=1, fma_mode=1):
    img, flow, lin_x, lin_y = _f32(img), _f32(flow), _f32(lin_x), _f32(lin_y)
    B, C, H, W = img.shape
    out = np.empty_like(img)
    lib().oracle_gridwarp_fwd_16(_p(img), _p(flow), _p(out), _p(lin_x), _p(lin_y),
                                 _I(B), _I(C), _I(H), _I(W), _I(inv_mode), _I(fma_mode), _I(dtype))
    return out


def quotient_model(prod, d, delta):
    """(got, want) of oracle_quotient_model: the product kernel's reciprocal-and-correct quotient vs the reference's fp64 divide."""
    prod = _f32(prod)
    d = np.ascontiguousarray(d, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    got, want = np.empty_like(prod), np.empty_like(prod)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().oracle_quotient_model(_p(prod), d.ctypes.data_as(dp), delta.ctypes.data_as(dp), _p(got), _p(want),
                                ctypes.c_longlong(prod.size))
    return got, want
