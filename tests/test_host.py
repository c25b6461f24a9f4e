"""CPU tests of the host-side mirror of the reference interface: same names, import paths, argument
order and defaults; loud failure (never a silent CPU path) when asked to run off-GPU."""
import inspect

import pytest
import torch

from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm, ChannelNormFunction
from ir2rgb_b200.models.flownet2_pytorch.networks.correlation_package.correlation import Correlation, CorrelationFunction
from ir2rgb_b200.models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d, Resample2dFunction
from ir2rgb_b200.models import networks


def defaults(fn):
    sig = inspect.signature(fn)
    return [(n, p.default) for n, p in sig.parameters.items() if n not in ("self", "ctx")]


def test_constructor_signatures_match_reference():
    # reference correlation.py:43, resample2d.py:40, channelnorm.py:33
    assert defaults(Correlation.__init__) == [("pad_size", 0), ("kernel_size", 0), ("max_displacement", 0),
                                              ("stride1", 1), ("stride2", 2), ("corr_multiply", 1)]
    assert defaults(Resample2d.__init__) == [("kernel_size", 1)]
    assert defaults(ChannelNorm.__init__) == [("norm_deg", 2)]


def test_function_signatures_match_reference():
    # reference correlation.py:10-12, resample2d.py:8, channelnorm.py:8
    assert defaults(CorrelationFunction.forward) == [("input1", inspect._empty), ("input2", inspect._empty),
                                                     ("pad_size", 3), ("kernel_size", 3), ("max_displacement", 20),
                                                     ("stride1", 1), ("stride2", 2), ("corr_multiply", 1)]
    assert defaults(Resample2dFunction.forward) == [("input1", inspect._empty), ("input2", inspect._empty), ("kernel_size", 1)]
    assert defaults(ChannelNormFunction.forward) == [("input1", inspect._empty), ("norm_deg", 2)]
    assert defaults(networks.get_grid) == [("batch_size", inspect._empty), ("rows", inspect._empty), ("cols", inspect._empty),
                                           ("device", "cuda:0"), ("dtype", torch.float32)]


def test_modules_hold_no_parameters_and_keep_attribute_names():
    c = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
    assert (c.pad_size, c.kernel_size, c.max_displacement, c.stride1, c.stride2, c.corr_multiply) == (20, 1, 20, 1, 2, 1)
    for m in (c, Resample2d(), ChannelNorm()):
        assert list(m.state_dict().keys()) == []        # checkpoints of the reference load unchanged


def test_get_grid_values_match_reference_formula():
    g = networks.get_grid(2, 5, 7, device="cpu")
    assert g.shape == (2, 2, 5, 7)
    assert torch.equal(g[0, 0, 0], torch.linspace(-1.0, 1.0, 7))
    assert torch.equal(g[1, 1, :, 3], torch.linspace(-1.0, 1.0, 5))


@pytest.mark.parametrize("call", [
    lambda: ChannelNorm()(torch.zeros(1, 3, 4, 4)),
    lambda: Resample2d()(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)),
    lambda: Correlation(20, 1, 20, 1, 2, 1)(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8)),
    lambda: networks.resample(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)),
])
def test_no_cpu_fallback(call):
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        call()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ir2rgb_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libflowops.so"))
    with pytest.raises(_lib.FlowopsError, match="not found"):
        _lib.load()


def test_product_never_imports_the_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ir2rgb_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_flownet_resample_resolves_to_the_method_like_the_reference():
    """In the reference, `self.resample` inside FlowNet is Model.resample (the grid_sample warp), because the class
    method shadows the Resample2d submodule assigned in __init__ (flownet.py:17,50; base_model.py:129).  The drop-in
    keeps both the attribute layout and the lookup result."""
    from ir2rgb_b200.models.base_model import Model
    from ir2rgb_b200.models.flownet import FlowNet
    assert FlowNet.resample is Model.resample

    class Probe(Model):
        def __init__(self):
            super().__init__(gpu_ids=[], checkpoints_dir=".", name="probe")
            self.resample = torch.nn.Identity()

        def save(self, label):
            pass

    p = Probe()
    assert isinstance(p._modules["resample"], torch.nn.Identity)
    assert getattr(p.resample, "__func__", None) is Model.resample


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_16bit_tensors_have_no_cpu_fallback_either(dt):
    """fp16 / bf16 tensors select the *_16 entry points; off-GPU they fail as loudly as fp32 ones."""
    from ir2rgb_b200 import functional as F
    for call in (lambda: F.channelnorm_forward(torch.zeros(1, 3, 4, 4, dtype=dt)),
                 lambda: F.warp_forward(torch.zeros(1, 3, 4, 4, dtype=dt), torch.zeros(1, 2, 4, 4, dtype=dt)),
                 lambda: F.correlation_forward(torch.zeros(1, 4, 8, 8, dtype=dt), torch.zeros(1, 4, 8, 8, dtype=dt), 20, 1, 20, 1, 2)):
        with pytest.raises(RuntimeError, match="no CPU implementation"):
            call()


def test_fp16_modules_keep_the_reference_layout():
    """models.py:22-28,63: FlowNet2(fp16=True) swaps in fp16_resample2d, which wraps a Resample2d submodule named
    `resample`; neither holds parameters, so reference checkpoints load unchanged in fp16 mode too."""
    from ir2rgb_b200.models.flownet2_pytorch import models as fn2
    m = fn2.fp16_resample2d()
    assert isinstance(m.resample, Resample2d) and list(m.state_dict().keys()) == []
    assert defaults(fn2.fp16_resample2d.forward) == [("input1", inspect._empty), ("input2", inspect._empty)]
    assert defaults(fn2.FlowNet2.__init__) == [("args", None), ("batchNorm", False), ("div_flow", 20.), ("fp16", False)]


def test_model_resample_fp16_option_leaves_other_dtypes_alone(monkeypatch):
    """Model.resample takes the one-pass 16-bit kernel only for fp16 tensors under opt['fp16'] (base_model.py:123-136);
    everything else goes where it went before."""
    from ir2rgb_b200.models import base_model
    seen = []
    monkeypatch.setattr(base_model.networks, "resample", lambda image, flow: seen.append((image.dtype, flow.dtype)) or image)

    class M(base_model.Model):
        def save(self, label):
            pass
    m = M(fp16=True, gpu_ids=[], checkpoints_dir=".", name="t")
    img, flow = torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4)
    m.resample(img, flow)                       # fp32 under the fp16 option
    m.resample(img.half(), flow.half())         # fp16 but on the CPU: not the kernel's business
    m.resample(img.bfloat16(), flow.bfloat16())
    assert seen == [(torch.float32, torch.float32), (torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)]


def test_magic_number_floor_and_int_conversions_model():
    """csrc/warp_rows.cuh replaces floorf / (int) / (float) -- XU-pipe instructions on sm_100 -- by FP32-pipe arithmetic with
    the constant 1.5 * 2^23 and claims exactness for |v| < 2^22.  numpy float32 model of the same operation sequence against
    np.floor / integer casts on random values, every integer and half-integer neighbourhood near the range ends, and zeros."""
    import numpy as np
    M = np.float32(12582912.0)
    rng = np.random.default_rng(9)
    v = np.concatenate([
        (rng.uniform(-1, 1, 3_000_000) * 2.0 ** rng.integers(-30, 22, 3_000_000)).astype(np.float32),
        np.arange(-4096, 4096, dtype=np.float32) / np.float32(8.0),
        np.nextafter(np.arange(-2048, 2048, dtype=np.float32), np.float32(np.inf)),
        np.nextafter(np.arange(-2048, 2048, dtype=np.float32), np.float32(-np.inf)),
        np.float32(4194304.0) - np.arange(1, 4096, dtype=np.float32) / np.float32(2.0),
        -np.float32(4194304.0) + np.arange(1, 4096, dtype=np.float32) / np.float32(2.0),
        np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -1.5, -2.5], np.float32)])
    v = v[np.abs(v) < 4194304.0]
    f = (v + M) - M                                   # round to nearest integer (ulp is 1 on [2^23, 2^24))
    f = np.where(f > v, f - np.float32(1.0), f)       # step down if that rounded up
    assert f.dtype == np.float32 and np.array_equal(f, np.floor(v))
    # small_float_as_int: integer-valued float in [0, 2^22) -> int through the mantissa of v + 1.5 * 2^23
    iv = np.floor(np.abs(v)).astype(np.float32)
    as_int = (iv + M).view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(as_int, iv.astype(np.int32))
    # small_int_as_float: i in [0, 2^23) -> float by planting i in the mantissa of 2^23
    i = rng.integers(0, 1 << 23, 1_000_000).astype(np.int32)
    as_float = (np.int32(0x4B000000) | i).view(np.float32) - np.float32(8388608.0)
    assert np.array_equal(as_float, i.astype(np.float32))


def test_ctypes_shims_have_the_pybind_signatures():
    """ir2rgb_b200/shims mirrors the reference's three pybind11 modules (correlation_cuda.cc:169-172,
    resample2d_cuda.cc:25-28, channelnorm_cuda.cc:28-31): same function names and positional parameters."""
    import inspect
    import sys
    import ir2rgb_b200.shims as shims
    want = {
        "correlation_cuda": {"forward": 11, "backward": 13},
        "resample2d_cuda": {"forward": 4, "backward": 6},
        "channelnorm_cuda": {"forward": 3, "backward": 5},
    }
    for mod, fns in want.items():
        for fn, nargs in fns.items():
            assert len(inspect.signature(getattr(getattr(shims, mod), fn)).parameters) == nargs, (mod, fn)
    saved = {n: sys.modules.get(n) for n in shims.NAMES}
    try:
        shims.install()
        import correlation_cuda
        assert correlation_cuda is shims.correlation_cuda
    finally:
        shims.uninstall()
        for n, m in saved.items():
            if m is not None:
                sys.modules[n] = m


def test_deconv_as_conv3_weight_reproduces_conv_transpose():
    """submodules.deconv_as_conv3_weight: a ConvTranspose2d(k4, s2, p1) equals a 3x3 convolution with 4*C output channels at
    the input resolution followed by depth-to-space (fp64 on the CPU; the GPU test covers the epilogue kernel)."""
    import torch
    import torch.nn.functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks.submodules import deconv_as_conv3_weight
    torch.manual_seed(0)
    conv = torch.nn.ConvTranspose2d(10, 8, 4, 2, 1).double()
    x = torch.randn(2, 10, 5, 7, dtype=torch.float64)
    want = conv(x)
    y4 = F.conv2d(x, deconv_as_conv3_weight(conv, conv.weight), None, 1, 1)
    got = torch.zeros_like(want)
    for py in (0, 1):
        for px in (0, 1):
            g = (py * 2 + px) * 8
            got[:, :, py::2, px::2] = y4[:, g:g + 8] + conv.bias.view(1, -1, 1, 1)
    assert (got - want).abs().max().item() <= 1e-12


def test_space_to_depth_first_layer_weights():
    """FlowNetC.conv1_s2d: the 7x7 stride-2 first layer as a 4x4 stride-1 convolution over the space-to-depth frame
    (layout of flowops_flownet2_prep_s2d restated with torch indexing), fp64 on the CPU."""
    import torch
    import torch.nn.functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks import FlowNetC
    from ir2rgb_b200.models.flownet2_pytorch.models import MyDict
    torch.manual_seed(1)
    args = MyDict()
    args.rgb_max, args.fp16, args.grads = 1, False, {}
    net = FlowNetC.FlowNetC(args, batchNorm=False).double()
    B, H, W = 2, 12, 16
    x = torch.randn(B, 3, H, W, dtype=torch.float64)
    s2d = torch.zeros(B, 16, H // 2 + 1, W // 2 + 1, dtype=torch.float64)
    for py in (0, 1):
        for px in (0, 1):
            s2d[:, (py * 2 + px) * 4:(py * 2 + px) * 4 + 3, 1:, 1:] = x[:, :, py::2, px::2]
    with torch.no_grad():
        want = net.conv1(x)
        got = net.conv1_s2d()(s2d)
    assert got.shape == want.shape and (got - want).abs().max().item() <= 1e-12
    assert "_flowops_conv1_s2d" not in dict(net.named_modules()) and len(net.state_dict()) == len(FlowNetC.FlowNetC(args, batchNorm=False).state_dict())
