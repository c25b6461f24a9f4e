"""GPU parity tests of the 16-bit storage entry points (flowops_*_16), run with -m gpu.

What each operator must equal is what the reference's fp16 mode computes on the same tensors
(include/flowops.h, "16-bit storage variants"):
  * ChannelNorm: the reference kernels instantiated for at::Half -- checked against the C restatement
    (oracle_cnorm_*_16), the rebuilt reference extension run live on half tensors, and a torch emulation;
  * Resample2d: fp16_resample2d's cast chain around the fp32 operator (bit-identical);
  * Correlation: FlowNetC's cast chain around the fp32 operator (<= 1 ulp of the storage type);
  * Model.resample with opt['fp16']: the reference's own op chain run by torch on the GPU, and the C restatement.
Everything goes through the public drop-in modules -> ctypes -> the C ABI.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]


def code_of(c_oracle, dt):
    return c_oracle.DTYPE_F16 if dt == torch.float16 else c_oracle.DTYPE_BF16


def f32np(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def ops(flowops_lib):
    assert torch.cuda.is_available()
    from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    from ir2rgb_b200.models.flownet2_pytorch.networks.correlation_package.correlation import Correlation
    from ir2rgb_b200.models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d
    from ir2rgb_b200.models import base_model, networks
    from ir2rgb_b200.models.flownet2_pytorch import models as fn2

    class Ops:
        pass
    o = Ops()
    o.ChannelNorm, o.Correlation, o.Resample2d = ChannelNorm, Correlation, Resample2d
    o.networks, o.base_model, o.fn2 = networks, base_model, fn2
    return o


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_ext
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    return ref_ext


# =============================================================================================
# ChannelNorm
# =============================================================================================
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (1, 2, 13, 19), (1, 5, 8, 8), (3, 3, 64, 128), (1, 1, 1, 1), (2, 2, 31, 33),
                                   (1, 12, 20, 36)])
def test_cnorm16_vs_oracle_bitexact(ops, c_oracle, dt, shape):
    torch.manual_seed(11)
    x = (3 * torch.randn(shape)).to(dt)
    x[0, :, 0, 0] = 0
    if dt == torch.float16 and x.numel() > 4:
        x.view(-1)[3] = 300.0                                   # the square overflows fp16: inf, as in the reference kernel
    xt = x.cuda().requires_grad_()
    y = ops.ChannelNorm()(xt)
    assert y.dtype == dt and tuple(y.shape) == (shape[0], 1) + shape[2:]
    y_ref = c_oracle.cnorm_fwd_16(f32np(x), code_of(c_oracle, dt))
    assert np.array_equal(f32np(y), y_ref)
    gy = torch.randn(y.shape).to(dt)
    y.backward(gy.cuda())
    gx_ref = c_oracle.cnorm_bwd_16(f32np(x), y_ref, f32np(gy), code_of(c_oracle, dt))
    assert xt.grad.dtype == dt
    assert np.array_equal(f32np(xt.grad), gx_ref, equal_nan=True)


@pytest.mark.parametrize("dt", DTYPES)
def test_cnorm16_special_values(ops, c_oracle, dt):
    """Overflowed squares (y = inf: every gradient of the pixel is 0, not NaN), zeros of either sign, inf and NaN."""
    torch.manual_seed(14)
    x = torch.randn(1, 3, 8, 16).to(dt)
    x[0, 0, 0, 0] = 3e4 if dt == torch.float16 else 3e38
    x[0, :, 0, 1] = 0.0
    x[0, 1, 0, 2] = float("inf")
    x[0, 2, 0, 3] = float("nan")
    gy = torch.randn(1, 1, 8, 16).to(dt)
    gy[0, 0, 0, 4] = -0.0
    gy[0, 0, 0, 1] = -1.0
    xt = x.cuda().requires_grad_()
    y = ops.ChannelNorm()(xt)
    y_ref = c_oracle.cnorm_fwd_16(f32np(x), code_of(c_oracle, dt))
    assert np.isinf(y_ref[0, 0, 0, 0]) and np.array_equal(f32np(y), y_ref, equal_nan=True)
    y.backward(gy.cuda())
    gx, gx_ref = f32np(xt.grad), c_oracle.cnorm_bwd_16(f32np(x), y_ref, f32np(gy), code_of(c_oracle, dt))
    assert np.array_equal(gx, gx_ref, equal_nan=True)
    assert (gx[0, :, 0, 0] == 0).all()
    ok = ~np.isnan(gx_ref)
    assert np.array_equal(np.signbit(gx[ok]), np.signbit(gx_ref[ok]))


def test_cnorm16_is_not_the_cast_chain(ops):
    """The half kernel squares in half precision (channelnorm_kernel.cu:55): it differs from cast -> fp32 op -> cast,
    which is why it has kernels of its own."""
    torch.manual_seed(5)
    x = torch.randn(2, 3, 64, 64).half().cuda()
    y16 = ops.ChannelNorm()(x)
    ychain = ops.ChannelNorm()(x.float()).half()
    assert (y16 != ychain).any()
    assert ((y16.float() - ychain.float()).abs() <= 2e-3 * ychain.float().abs() + 1e-6).all()


@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (1, 2, 13, 19), (4, 3, 128, 256)])
def test_cnorm16_vs_reference_ext_live(ops, ref, shape):
    """The reference's own extension dispatches at::Half (channelnorm_kernel.cu:111,152): run it on the same tensors."""
    torch.manual_seed(12)
    x = (2 * torch.randn(shape)).half().cuda()
    y = ops.ChannelNorm()(x)
    y_ref = ref.channelnorm_forward(x)
    assert y_ref.dtype == torch.float16
    assert torch.equal(y, y_ref)
    gy = torch.randn_like(y)
    from ir2rgb_b200 import functional as F
    gx = F.channelnorm_backward(x, y, gy)
    gx_ref = ref.channelnorm_backward(x, y_ref, gy)
    assert torch.equal(gx, gx_ref)


def test_cnorm16_full_size_properties(ops, c_oracle):
    """Config-3 shape: a strided sample of pixels against the C restatement (the operator is pixel-wise), sign
    invariance, and agreement with the fp32 operator to 16-bit accuracy."""
    torch.manual_seed(13)
    x = (2 * torch.rand(16, 3, 512, 1024) - 1).half().cuda()
    y = ops.ChannelNorm()(x)
    assert torch.equal(ops.ChannelNorm()(-x), y)
    xs = x[:, :, ::37, ::41].permute(1, 0, 2, 3).reshape(1, 3, 1, -1)          # [1, 3, 1, N] pixels
    ys = y[:, :, ::37, ::41].reshape(1, 1, 1, -1)
    assert np.array_equal(f32np(ys), c_oracle.cnorm_fwd_16(f32np(xs), c_oracle.DTYPE_F16))
    y32 = ops.ChannelNorm()(x.float())
    assert ((y.float() - y32).abs() <= 2e-3 * y32 + 1e-4).all()


# =============================================================================================
# Resample2d (fp16_resample2d, models.py:22-28)
# =============================================================================================
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,sigma", [((2, 3, 16, 24), 3.0), ((1, 2, 13, 19), 6.0), ((1, 3, 12, 20), 40.0), ((1, 1, 5, 7), 1.0),
                                         ((2, 3, 70, 130), 6.0), ((1, 3, 256, 512), 5.0)])
def test_resample2d16_equals_cast_chain_bitexact(ops, dt, shape, sigma):
    torch.manual_seed(21)
    B, C, H, W = shape
    img = torch.randn(shape).to(dt).cuda()
    flow = (sigma * torch.randn(B, 2, H, W)).to(dt).cuda()
    out = ops.Resample2d()(img, flow)
    assert out.dtype == dt
    chain = ops.Resample2d()(img.float(), flow.float()).to(dt)
    assert torch.equal(out, chain)


def test_resample2d16_vs_c_oracle_and_reference_ext(ops, c_oracle, ref):
    torch.manual_seed(22)
    img = torch.randn(2, 3, 33, 65).half()
    flow = (4 * torch.randn(2, 2, 33, 65)).half()
    out = ops.Resample2d()(img.cuda(), flow.cuda())
    want = c_oracle.round16(c_oracle.resample2d_fwd(f32np(img), f32np(flow)), c_oracle.DTYPE_F16)
    assert np.array_equal(f32np(out), want)
    want_ref = ref.resample2d_forward(img.cuda().float(), flow.cuda().float()).half()       # the reference's chain, live
    assert torch.equal(out, want_ref)


def test_fp16_resample2d_module_uses_the_16bit_kernel(ops, monkeypatch):
    """The drop-in's fp16_resample2d (models.py:22-28) takes the one-pass path for fp16 CUDA tensors and returns the
    reference chain's values; mixed dtypes take the literal chain."""
    from ir2rgb_b200 import _lib
    calls = []
    monkeypatch.setattr(_lib, "launch_hook", lambda what, n: calls.append(what))
    torch.manual_seed(23)
    img, flow = torch.randn(1, 3, 32, 48).half().cuda(), (3 * torch.randn(1, 2, 32, 48)).half().cuda()
    m = ops.fn2.fp16_resample2d()
    out = m(img, flow)
    assert calls == ["warp_fwd_16"] and out.dtype == torch.float16
    assert torch.equal(out, ops.Resample2d()(img.float(), flow.float()).half())
    calls.clear()
    out2 = m(img.float(), flow)                                # fp32 frame, fp16 flow (AMP): literal chain
    assert calls == ["warp_fwd"] and out2.dtype == torch.float16


def test_resample2d16_backward_is_the_cast_chain(ops):
    torch.manual_seed(24)
    img = torch.randn(1, 3, 24, 40).half().cuda().requires_grad_()
    flow = (3 * torch.randn(1, 2, 24, 40)).half().cuda().requires_grad_()
    gout = torch.randn(1, 3, 24, 40).half().cuda()
    ops.Resample2d()(img, flow).backward(gout)
    i2, f2 = img.detach().clone().requires_grad_(), flow.detach().clone().requires_grad_()
    ops.Resample2d()(i2.float(), f2.float()).half().backward(gout)
    assert img.grad.dtype == torch.float16 and flow.grad.dtype == torch.float16
    assert torch.equal(flow.grad, f2.grad)                      # gather: deterministic
    # image gradient: fp32 atomics in either chain, then one rounding to fp16
    assert ((img.grad.float() - i2.grad.float()).abs() <= 2e-3 * i2.grad.float().abs() + 1e-5).all()


# =============================================================================================
# Model.resample with opt['fp16'] (models/base_model.py:123-136)
# =============================================================================================
def reference_chain_fp16(image, flow):
    """base_model.py:129-136 with opt['fp16'], literally, on the GPU (torch ops)."""
    b, c, h, w = image.size()
    hor = torch.linspace(-1.0, 1.0, w).view(1, 1, 1, w).expand(b, 1, h, w)
    ver = torch.linspace(-1.0, 1.0, h).view(1, 1, h, 1).expand(b, 1, h, w)
    grid = torch.cat([hor, ver], 1).to(flow.dtype).to(flow.device)
    flow = torch.cat([flow[:, 0:1, :, :] / ((w - 1.0) / 2.0), flow[:, 1:2, :, :] / ((h - 1.0) / 2.0)], dim=1)
    final_grid = (grid + flow).permute(0, 2, 3, 1)
    return torch.nn.functional.grid_sample(image.float(), final_grid.float(), mode='bilinear', padding_mode='border',
                                           align_corners=False).to(image.dtype)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,sigma", [((2, 3, 16, 24), 3.0), ((1, 2, 13, 19), 6.0), ((1, 3, 12, 20), 40.0),
                                         ((1, 3, 256, 512), 5.0), ((2, 3, 64, 300), 2.0)])
def test_gridsample16_vs_reference_chain_on_gpu(ops, c_oracle, dt, shape, sigma):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(31)
    B, C, H, W = shape
    img = torch.randn(shape).to(dt).cuda()
    flow = (sigma * torch.randn(B, 2, H, W)).to(dt).cuda()
    out = F.warp_forward(img, flow, F.WARP_GRIDSAMPLE)
    assert out.dtype == dt
    # (1) the C restatement, CUDA form (reciprocal multiply + fma): bit for bit
    lx = torch.linspace(-1, 1, W).to(dt).float().numpy()
    ly = torch.linspace(-1, 1, H).to(dt).float().numpy()
    want = c_oracle.gridwarp_fwd_16(f32np(img), f32np(flow), lx, ly, code_of(c_oracle, dt), inv_mode=1, fma_mode=1)
    assert np.array_equal(f32np(out), want)
    # (2) the reference's own op chain run by torch on this GPU.  The 16-bit grid is reproduced exactly, so the two can
    # differ only by ATen's fp32 rounding inside grid_sample before the final rounding to 16 bits: a last-place flip on a
    # small fraction of elements at most.
    chain = reference_chain_fp16(img, flow)
    diff = (out.float() - chain.float()).abs()
    ulp = 2.0 ** (-10 if dt == torch.float16 else -7)
    assert (diff <= ulp * chain.float().abs() + 1e-6).all()
    assert (out != chain).float().mean().item() < 1e-3


def test_model_resample_fp16_dispatch(ops, monkeypatch):
    from ir2rgb_b200 import _lib

    class M(ops.base_model.Model):
        def save(self, label):
            pass
    calls = []
    monkeypatch.setattr(_lib, "launch_hook", lambda what, n: calls.append(what))
    m = M(fp16=True, gpu_ids=[0])
    torch.manual_seed(32)
    img, flow = torch.randn(1, 3, 32, 48).half().cuda(), (3 * torch.randn(1, 2, 32, 48)).half().cuda()
    with torch.no_grad():
        out = m.resample(img, flow)
    assert calls == ["warp_fwd_16"] and out.dtype == torch.float16
    chain = reference_chain_fp16(img, flow)
    assert ((out.float() - chain.float()).abs() <= 2.0 ** -10 * chain.float().abs() + 1e-6).all()
    # with gradients required the reference's literal op chain runs (and is differentiable)
    calls.clear()
    f2 = flow.clone().requires_grad_()
    out2 = m.resample(img, f2)
    assert calls == [] and torch.equal(out2, chain)
    out2.float().sum().backward()
    assert f2.grad is not None and f2.grad.dtype == torch.float16
    # fp32 tensors are untouched by the fp16 option
    calls.clear()
    m.resample(img.float(), flow.float())
    assert calls == ["warp_fwd"]


# =============================================================================================
# Correlation (FlowNetC.py:86-87)
# =============================================================================================
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 8, 12), (1, 40, 6, 40), (1, 9, 7, 5), (2, 64, 24, 32), (1, 256, 64, 128)])
def test_correlation16_vs_cast_chain(ops, dt, shape):
    torch.manual_seed(41)
    a, b = torch.randn(shape).to(dt).cuda(), torch.randn(shape).to(dt).cuda()
    corr = ops.Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
    out = corr(a, b)
    assert out.dtype == dt and out.shape[1] == 441
    # the 16-bit entry point runs the FP32-FMA kernel: its result is that kernel's fp32 result rounded, bit for bit
    from ir2rgb_b200 import _lib
    lib = _lib.load()
    prev = lib.flowops_corr_get_impl()
    try:
        lib.flowops_corr_set_impl(0)
        chain32 = corr(a.float(), b.float())
    finally:
        lib.flowops_corr_set_impl(prev)
    assert torch.equal(out, chain32.to(dt))
    # and agrees with the default fp32 path (the tensor-core kernel where the shape allows it) to 16-bit rounding
    chain_default = corr(a.float(), b.float())
    tol = 2.0 ** -10 if dt == torch.float16 else 2.0 ** -7
    assert (out.float() - chain_default).abs().max().item() <= tol * chain_default.abs().max().item()


def test_correlation16_vs_c_oracle_and_reference_ext(ops, c_oracle, ref):
    torch.manual_seed(42)
    a, b = torch.randn(1, 40, 6, 40).half(), torch.randn(1, 40, 6, 40).half()
    corr = ops.Correlation(20, 1, 20, 1, 2, 1)
    out = corr(a.cuda(), b.cuda())
    want32 = c_oracle.corr_fwd(f32np(a), f32np(b), 20, 1, 20, 1, 2)
    want = c_oracle.round16(want32, c_oracle.DTYPE_F16)
    diff = np.abs(f32np(out) - want)
    assert (diff <= 2.0 ** -10 * np.abs(want) + 1e-7).all()           # <= 1 fp16 ulp (fp32 summation order differs)
    assert (f32np(out) != want).mean() < 2e-3
    want_ref = ref.correlation_forward(a.cuda().float(), b.cuda().float(), 20, 1, 20, 1, 2).half()   # FlowNetC.py:86-87, live
    assert ((out.float() - want_ref.float()).abs() <= 2.0 ** -10 * want_ref.float().abs() + 1e-7).all()


def test_correlation16_other_parameters_and_backward(ops):
    """Outside the FlowNetC configuration 16-bit tensors take the literal cast chain; the backward is the chain's."""
    torch.manual_seed(43)
    a, b = torch.randn(1, 8, 9, 11).half().cuda(), torch.randn(1, 8, 9, 11).half().cuda()
    corr = ops.Correlation(4, 1, 4, 1, 1, 1)
    out = corr(a, b)
    assert out.dtype == torch.float16 and torch.equal(out, corr(a.float(), b.float()).half())
    a1, b1 = torch.randn(1, 16, 8, 12).half().cuda().requires_grad_(), torch.randn(1, 16, 8, 12).half().cuda().requires_grad_()
    c2 = ops.Correlation(20, 1, 20, 1, 2, 1)
    out = c2(a1, b1)
    g = torch.randn_like(out)
    out.backward(g)
    a2, b2 = a1.detach().clone().requires_grad_(), b1.detach().clone().requires_grad_()
    c2(a2.float(), b2.float()).half().backward(g)
    assert a1.grad.dtype == torch.float16
    assert torch.equal(a1.grad, a2.grad) and torch.equal(b1.grad, b2.grad)       # deterministic kernels, same casts


def test_flownetc_fp16_branch_uses_the_16bit_operator(ops, monkeypatch):
    from ir2rgb_b200 import _lib
    from ir2rgb_b200.models.flownet2_pytorch.networks import FlowNetC
    args = ops.fn2.MyDict()
    args.fp16, args.rgb_max, args.grads = True, 1, {}
    torch.manual_seed(44)
    net = FlowNetC.FlowNetC(args).cuda().half().eval()
    x = torch.randn(1, 6, 64, 128).half().cuda()
    calls = []
    monkeypatch.setattr(_lib, "launch_hook", lambda what, n: calls.append(what))
    with torch.no_grad():
        flow = net(x)[0]
    assert "corr_fwd_16" in calls and "corr_fwd" not in calls
    assert flow.dtype == torch.float16 and tuple(flow.shape) == (1, 2, 16, 32) and torch.isfinite(flow.float()).all()


# =============================================================================================
# boundary behaviour
# =============================================================================================
def test_16bit_argument_errors(flowops_lib):
    from ir2rgb_b200 import functional as F
    img, flow = torch.randn(1, 3, 8, 8).half().cuda(), torch.randn(1, 2, 8, 8).cuda()
    with pytest.raises(TypeError):
        F.warp_forward(img, flow)                                  # mixed dtypes are the caller's to cast
    with pytest.raises(TypeError):
        F.channelnorm_backward(img, img[:, :1].float(), img[:, :1])
    with pytest.raises(NotImplementedError):
        F.warp_forward(torch.randn(1, 5, 8, 8).half().cuda(), flow.half())      # > 3 channels: cast and use the fp32 operator
    assert flowops_lib.flowops_cnorm_fwd_16(None, None, 1, 1, 1, 1, 1, None) == -1
    assert flowops_lib.flowops_cnorm_fwd_16(img.data_ptr(), img.data_ptr(), 1, 3, 8, 8, 7, None) == -1
    empty = torch.empty(0, 3, 8, 8, dtype=torch.float16, device="cuda")
    assert F.channelnorm_forward(empty).shape == (0, 1, 8, 8)


def test_16bit_ops_are_graph_capturable(ops):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(51)
    img, flow = torch.randn(2, 3, 64, 96).half().cuda(), (3 * torch.randn(2, 2, 64, 96)).half().cuda()
    a, b = torch.randn(1, 32, 16, 24).half().cuda(), torch.randn(1, 32, 16, 24).half().cuda()
    want = (F.warp_forward(img, flow), F.warp_forward(img, flow, F.WARP_GRIDSAMPLE), F.channelnorm_forward(img),
            F.correlation_forward(a, b, 20, 1, 20, 1, 2))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = (F.warp_forward(img, flow), F.warp_forward(img, flow, F.WARP_GRIDSAMPLE), F.channelnorm_forward(img),
                   F.correlation_forward(a, b, 20, 1, 20, 1, 2))
        for t in got:
            t.zero_()
        g.replay()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for w, t in zip(want, got):
        assert torch.equal(w, t)
