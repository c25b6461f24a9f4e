"""Loader for the reference's own CUDA extensions rebuilt for sm_100 (oracle/_ref/, see build_ref.py).

Test infrastructure only (GPU-side oracle and the reference arm of bench.py).  The compiled
``*_cuda.forward/backward`` entry points are called directly because two of the three Python
wrappers in the reference cannot run their backward on a modern torch.
Tensors are allocated exactly as the reference wrappers do (correlation.py:16-18,30-34;
resample2d.py:17,29-30; channelnorm.py:11,23).
"""
import importlib.util
import os

import torch

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available():
    return all(os.path.exists(os.path.join(_DIR, n + ".so"))
               for n in ("correlation_cuda", "resample2d_cuda", "channelnorm_cuda"))


def _load(name):
    if name not in _mods:
        spec = importlib.util.spec_from_file_location(name, os.path.join(_DIR, name + ".so"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def correlation_forward(in1, in2, pad=20, k=1, md=20, s1=1, s2=2):
    m = _load("correlation_cuda")
    with torch.cuda.device_of(in1):
        rbot1, rbot2, out = in1.new(), in2.new(), in1.new()
        m.forward(in1, in2, rbot1, rbot2, out, pad, k, md, s1, s2, 1)
    return out


def correlation_backward(in1, in2, gout, pad=20, k=1, md=20, s1=1, s2=2):
    m = _load("correlation_cuda")
    with torch.cuda.device_of(in1):
        rbot1, rbot2, g1, g2 = in1.new(), in2.new(), in1.new(), in2.new()
        m.backward(in1, in2, rbot1, rbot2, gout, g1, g2, pad, k, md, s1, s2, 1)
    return g1, g2


def resample2d_forward(img, flow):
    m = _load("resample2d_cuda")
    out = torch.zeros_like(img)
    m.forward(img, flow, out, 1)
    return out


def resample2d_backward(img, flow, gout):
    m = _load("resample2d_cuda")
    gimg, gflow = torch.zeros_like(img), torch.zeros_like(flow)
    m.backward(img, flow, gout, gimg, gflow, 1)
    return gimg, gflow


def channelnorm_forward(x):
    m = _load("channelnorm_cuda")
    out = x.new_zeros((x.shape[0], 1, x.shape[2], x.shape[3]))
    m.forward(x, out, 2)
    return out


def channelnorm_backward(x, out, gout):
    m = _load("channelnorm_cuda")
    gx = torch.zeros_like(x)
    m.backward(x, out, gout, gx, 2)
    return gx
