"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the public drop-in
modules -> ctypes -> the C ABI of libflowops.so; the oracle (C restatement, PyTorch fp64 closed
forms, the reference's rebuilt extensions when oracle/_ref is present) is only the checker.

Tolerances (BASELINE.json north_star): fp32 forward max-relative error <= 1e-5, backward <= 1e-4,
where max-relative = max|a-b| / max|b|.  Where the arithmetic can be reproduced exactly the tests
ask for bit-identity instead.
"""
import numpy as np
import pytest
import torch

from oracle import torch_ref as tr

pytestmark = pytest.mark.gpu

FWD_TOL, BWD_TOL = 1e-5, 1e-4


def maxrel(a, b):
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.from_numpy(np.asarray(b)).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops(flowops_lib):
    assert torch.cuda.is_available()
    from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    from ir2rgb_b200.models.flownet2_pytorch.networks.correlation_package.correlation import Correlation
    from ir2rgb_b200.models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d
    from ir2rgb_b200.models import networks

    class Ops:
        pass
    o = Ops()
    o.ChannelNorm, o.Correlation, o.Resample2d, o.networks = ChannelNorm, Correlation, Resample2d, networks
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
@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (1, 2, 13, 19), (1, 5, 8, 8), (3, 3, 64, 128), (1, 1, 1, 1),
                                   (2, 2, 31, 33), (1, 12, 20, 36)])
def test_cnorm_vs_oracle_bitexact(ops, c_oracle, shape):
    rng = np.random.default_rng(10)
    x = rng.standard_normal(shape).astype(np.float32)
    x[0, :, 0, 0] = 0.0
    xt = cu(x).requires_grad_()
    y = ops.ChannelNorm()(xt)
    y_ref = c_oracle.cnorm_fwd(x)
    assert np.array_equal(y.detach().cpu().numpy(), y_ref)
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    y.backward(cu(gy))
    gx_ref = c_oracle.cnorm_bwd(x, y_ref, gy)
    gx = xt.grad.cpu().numpy()
    assert maxrel(gx, gx_ref) <= 1e-7
    assert (gx == gx_ref).mean() >= 0.9999       # reciprocal + correction reproduces the fp64 divide


def test_cnorm_special_values_bit_exact(ops, c_oracle):
    """Zeros of either sign, squares that overflow to inf, inf and NaN inputs: forward and backward equal the
    restated reference kernels bit for bit, sign of zero included ((-0) / d is -0; finite / inf is 0, not NaN)."""
    rng = np.random.default_rng(17)
    x = rng.standard_normal((2, 3, 8, 16)).astype(np.float32)
    x[0, :, 0, 0] = 0.0
    x[0, 1, 0, 1] = -0.0
    x[0, 0, 0, 2] = 1e30            # square overflows: y = inf, every gradient of the pixel is 0
    x[0, 2, 0, 3] = np.inf
    x[0, 1, 0, 4] = np.nan
    x[1, :, 1, :] = 0.0
    gy = rng.standard_normal((2, 1, 8, 16)).astype(np.float32)
    gy[0, 0, 0, 5] = 0.0
    gy[0, 0, 0, 6] = -0.0
    gy[0, 0, 0, 7] = 3e38           # product overflows: the quotient is inf
    x[0, 0, 0, 7] = 7.0
    xt = cu(x).requires_grad_()
    y = ops.ChannelNorm()(xt)
    y_ref = c_oracle.cnorm_fwd(x)
    assert np.array_equal(y.detach().cpu().numpy(), y_ref, equal_nan=True)
    y.backward(cu(gy))
    gx, gx_ref = xt.grad.cpu().numpy(), c_oracle.cnorm_bwd(x, y_ref, gy)
    assert np.array_equal(gx, gx_ref, equal_nan=True)
    ok = ~np.isnan(gx_ref)            # the sign of a NaN is not defined (x86 and the GPU generate different ones)
    assert np.array_equal(np.signbit(gx[ok]), np.signbit(gx_ref[ok]))


def test_cnorm_golden(ops, golden_native):
    g = golden_native
    for name in ["cn3", "cn2", "cn5"]:
        xt = cu(g[name + "_x"]).requires_grad_()
        y = ops.ChannelNorm()(xt)
        assert np.array_equal(y.detach().cpu().numpy(), g[name + "_y"])
        y.backward(cu(g[name + "_gy"]))
        assert maxrel(xt.grad, g[name + "_gx"]) <= 1e-7


def test_cnorm_vs_reference_ext_full_size(ops, ref):
    torch.manual_seed(0)
    x = (2 * torch.rand(16, 3, 512, 1024, device="cuda") - 1)          # BASELINE config 3 at its stated size
    y = ops.ChannelNorm()(x)
    assert torch.equal(y, ref.channelnorm_forward(x))
    gy = torch.randn_like(y)
    from ir2rgb_b200 import functional as F
    gx = F.channelnorm_backward(x, y, gy)
    gx_ref = ref.channelnorm_backward(x, y, gy)
    assert maxrel(gx, gx_ref) <= 1e-7
    # properties at full size: norm of a scaled input scales, gradient is parallel to x
    assert torch.allclose(ops.ChannelNorm()(2 * x), 2 * y, rtol=1e-6)


def test_cnorm_empty_and_errors(ops):
    y = ops.ChannelNorm()(torch.empty(0, 3, 4, 4, device="cuda"))
    assert y.shape == (0, 1, 4, 4)
    with pytest.raises(RuntimeError):
        ops.ChannelNorm()(torch.zeros(1, 3, 4, 4))              # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ops.ChannelNorm()(torch.zeros(1, 3, 4, 4, device="cuda", dtype=torch.float64))    # fp32, fp16 and bf16 only


# =============================================================================================
# Resample2d
# =============================================================================================
@pytest.mark.parametrize("shape,sigma", [((2, 3, 16, 24), 3.0), ((1, 2, 13, 19), 6.0), ((1, 3, 12, 20), 40.0),
                                         ((1, 1, 5, 7), 1.0), ((2, 4, 33, 65), 2.0), ((1, 3, 64, 128), 0.0)])
def test_resample2d_vs_oracle(ops, c_oracle, shape, sigma):
    rng = np.random.default_rng(11)
    B, C, H, W = shape
    img = rng.standard_normal(shape).astype(np.float32)
    flow = (sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    it, ft = cu(img).requires_grad_(), cu(flow).requires_grad_()
    out = ops.Resample2d()(it, ft)
    out_ref = c_oracle.resample2d_fwd(img, flow)
    assert np.array_equal(out.detach().cpu().numpy(), out_ref), "forward is bit-identical to the reference arithmetic"
    if sigma == 0.0:
        assert torch.equal(out.detach(), it.detach())
    go = rng.standard_normal(shape).astype(np.float32)
    out.backward(cu(go))
    gi_ref, gf_ref = c_oracle.resample2d_bwd(img, flow, go)
    assert maxrel(it.grad, gi_ref) <= BWD_TOL and maxrel(it.grad, gi_ref) <= 1e-5
    assert maxrel(ft.grad, gf_ref) <= 1e-6


@pytest.mark.parametrize("shape,sigma", [((2, 3, 33, 65), 2.0), ((1, 3, 70, 130), 6.0), ((1, 2, 40, 100), 30.0), ((1, 1, 5, 7), 1.0),
                                         ((1, 3, 64, 64), 0.0)])
def test_warp_backward_ragged_frames_vs_oracle(ops, c_oracle, shape, sigma):
    """The row-walking backward (lane hand-over, vertical pairing, zero skipping) on ragged frames, flows from zero to far
    outside the frame, both coordinate conventions."""
    from oracle import torch_ref as tr
    rng = np.random.default_rng(31)
    B, C, H, W = shape
    img = rng.standard_normal(shape).astype(np.float32)
    flow = (sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    go = rng.standard_normal(shape).astype(np.float32)
    it, ft = cu(img).requires_grad_(), cu(flow).requires_grad_()
    ops.Resample2d()(it, ft).backward(cu(go))
    gi_ref, gf_ref = c_oracle.resample2d_bwd(img, flow, go)
    assert maxrel(it.grad, gi_ref) <= 1e-5 and maxrel(ft.grad, gf_ref) <= 1e-6
    assert abs(it.grad.double().sum().item() - float(np.float64(go).sum())) <= 1e-3 * np.abs(go).sum() ** 0.5 + 1e-2 or sigma > 0
    # grid_sample convention against autograd through the reference code path in fp64
    i2, f2 = cu(img).requires_grad_(), cu(flow).requires_grad_()
    ops.networks.resample(i2, f2).backward(cu(go))
    i64, f64 = cu(img).double().requires_grad_(), cu(flow).double().requires_grad_()
    tr.networks_resample(i64, f64).backward(cu(go).double())
    assert maxrel(i2.grad, i64.grad) <= 1e-4


def test_resample2d_golden(ops, golden_native):
    g = golden_native
    for name in ["res_small", "res_odd", "res_far"]:
        it, ft = cu(g[name + "_img"]).requires_grad_(), cu(g[name + "_flow"]).requires_grad_()
        out = ops.Resample2d()(it, ft)
        assert np.array_equal(out.detach().cpu().numpy(), g[name + "_out"])
        out.backward(cu(g[name + "_gout"]))
        assert maxrel(it.grad, g[name + "_gimg"]) <= 1e-5
        assert maxrel(ft.grad, g[name + "_gflow"]) <= 1e-6


def test_resample2d_integer_flow_is_shift(ops):
    img = torch.randn(1, 3, 20, 30, device="cuda")
    flow = torch.zeros(1, 2, 20, 30, device="cuda")
    flow[:, 0] = 3.0
    flow[:, 1] = -2.0
    out = ops.Resample2d()(img, flow)
    assert torch.equal(out[:, :, 2:, :27], img[:, :, :18, 3:])
    # border clamping: rows that would read above the image repeat row 0
    assert torch.equal(out[:, :, 0, :27], img[:, :, 0, 3:])


def test_resample2d_grad_flags(ops):
    img = torch.randn(1, 3, 8, 8, device="cuda", requires_grad=True)
    flow = torch.randn(1, 2, 8, 8, device="cuda")
    ops.Resample2d()(img, flow).sum().backward()
    assert img.grad is not None
    img2 = torch.randn(1, 3, 8, 8, device="cuda")
    flow2 = torch.randn(1, 2, 8, 8, device="cuda", requires_grad=True)
    ops.Resample2d()(img2, flow2).sum().backward()
    assert flow2.grad is not None
    with pytest.raises(NotImplementedError):
        ops.Resample2d(kernel_size=3)(img2, flow2)


def _special_flow(H, W, rng):
    """Flow values that exercise every branch of the coordinate arithmetic: exact integers, half-integers, the
    2^22 boundary of the FP32-pipe floor, huge / infinite / NaN displacements, signed zeros, every border."""
    specials = np.array([0.0, -0.0, 0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 1e-20, -1e-20, 0.99999994, -0.99999994,
                         4194303.5, -4194303.5, 4194304.0, -4194304.0, 4194304.5, 8388607.5, -8388607.5,
                         1e9, -1e9, 3e38, -3e38, np.inf, -np.inf, np.nan, W - 1.0, -(W - 1.0), W - 1.5, H - 1.0,
                         -(H - 1.0), 2.0 ** -126, 1e-40], dtype=np.float32)
    flow = (3.0 * rng.standard_normal((1, 2, H, W))).astype(np.float32)
    idx = rng.integers(0, specials.size, size=(1, 2, H, W))
    mask = rng.random((1, 2, H, W)) < 0.5
    flow[mask] = specials[idx[mask]]
    # make the flow cancel the pixel index at some places so that x + dx is exactly 0 / W-1 / -1
    flow[0, 0, 0, :] = -np.arange(W, dtype=np.float32)
    flow[0, 0, 1, :] = (W - 1) - np.arange(W, dtype=np.float32)
    flow[0, 0, 2, :] = -1.0 - np.arange(W, dtype=np.float32)
    return flow


def test_resample2d_special_values_bit_exact(ops, c_oracle, ref):
    """The forward replaces floorf / int conversions by FP32-pipe arithmetic (warp_rows.cuh); every special value
    must still give the reference's bits (C restatement of resample2d_kernel.cu:40-62 and, live, the reference's
    own kernel)."""
    rng = np.random.default_rng(5)
    H, W = 24, 40
    img = rng.standard_normal((1, 3, H, W)).astype(np.float32)
    flow = _special_flow(H, W, rng)
    out = ops.Resample2d()(cu(img), cu(flow)).cpu().numpy()
    want = c_oracle.resample2d_fwd(img, flow)
    def same_bits(a, b):      # NaN payloads differ between x86 and the GPU; everything else must match bit for bit
        nan = np.isnan(a)
        return np.array_equal(nan, np.isnan(b)) and np.array_equal(a.view(np.uint32)[~nan], b.view(np.uint32)[~nan])
    assert same_bits(out, want)
    live = ref.resample2d_forward(cu(img), cu(flow)).cpu().numpy()
    assert same_bits(out, live)
    assert np.isnan(out).mean() < 0.5


def test_networks_resample_special_values_match_aten(ops):
    """Same for the grid_sample mode: ATen's own CUDA kernel is the reference for these inputs."""
    from oracle import torch_ref as tr
    rng = np.random.default_rng(6)
    H, W = 24, 40
    img = cu(rng.standard_normal((1, 3, H, W)).astype(np.float32))
    flow = cu(_special_flow(H, W, rng))
    out = ops.networks.resample(img, flow)
    want = tr.networks_resample(img, flow)
    both_nan = torch.isnan(out) & torch.isnan(want)
    assert torch.equal(torch.isnan(out), torch.isnan(want))
    assert maxrel(torch.where(both_nan, torch.zeros_like(out), out), torch.where(both_nan, torch.zeros_like(out), want)) <= 1e-6


@pytest.mark.parametrize("flavour", ["randn", "bilinear_up", "nearest_up"])
def test_resample2d_vs_reference_ext_full_size(ops, ref, flavour):
    """SURVEY 8d C3 at its stated size (16 x 3 x 512 x 1024), the three flow flavours."""
    torch.manual_seed(0)
    B, H, W = 16, 512, 1024
    img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
    if flavour == "randn":
        flow = 4 * torch.randn(B, 2, H, W, device="cuda")
    else:
        low = 20 * torch.randn(B, 2, H // 4, W // 4, device="cuda")
        flow = torch.nn.functional.interpolate(low, scale_factor=4, mode="bilinear" if flavour == "bilinear_up" else "nearest")
    flow = flow.contiguous()
    it, ft = img.clone().requires_grad_(), flow.clone().requires_grad_()
    out = ops.Resample2d()(it, ft)
    assert torch.equal(out.detach(), ref.resample2d_forward(img, flow))
    go = torch.randn_like(out)
    out.backward(go)
    gi_ref, gf_ref = ref.resample2d_backward(img, flow, go)
    assert maxrel(ft.grad, gf_ref) <= 1e-6
    assert maxrel(it.grad, gi_ref) <= BWD_TOL
    # conservation: the scatter distributes exactly the incoming gradient mass
    assert abs(it.grad.double().sum().item() - go.double().sum().item()) <= 1e-3 * go.abs().double().sum().item() ** 0.5 + 1e-2


# =============================================================================================
# networks.resample (grid_sample path)
# =============================================================================================
@pytest.mark.parametrize("name", ["small", "odd", "border"])
def test_networks_resample_golden(ops, golden_resample, name):
    """Against outputs of the reference's own Python code path (CPU)."""
    g = golden_resample
    it, ft = cu(g[name + "_img"]).requires_grad_(), cu(g[name + "_flow"]).requires_grad_()
    out = ops.networks.resample(it, ft)
    # fp32 coordinate normalisation costs ~3e-5 px; CPU and CUDA ATen differ from each other by that
    assert maxrel(out, g[name + "_out"]) <= 2e-5
    out.backward(cu(g[name + "_gout"]))
    assert maxrel(it.grad, g[name + "_gimg"]) <= BWD_TOL
    assert maxrel(ft.grad, g[name + "_gflow"]) <= 2e-4


@pytest.mark.parametrize("shape,sigma", [((1, 3, 256, 512), 5.0), ((2, 3, 64, 96), 40.0), ((1, 3, 17, 23), 3.0)])
def test_networks_resample_vs_torch_cuda(ops, shape, sigma):
    """Against the reference code path run on this GPU (ATen grid_sampler_2d) and against fp64 truth."""
    torch.manual_seed(0)
    B, C, H, W = shape
    img = torch.randn(shape, device="cuda")
    flow = sigma * torch.randn(B, 2, H, W, device="cuda")
    it, ft = img.clone().requires_grad_(), flow.clone().requires_grad_()
    out = ops.networks.resample(it, ft)
    ir, fr = img.clone().requires_grad_(), flow.clone().requires_grad_()
    out_ref = tr.networks_resample(ir, fr)
    truth = tr.networks_resample(img.double(), flow.double())
    err_new, err_ref = maxrel(out, truth), maxrel(out_ref, truth)
    frac_equal = (out == out_ref).float().mean().item()
    print("resample %s: new-vs-fp64 %.2e  aten-vs-fp64 %.2e  new-vs-aten %.2e  bit-equal %.4f"
          % (shape, err_new, err_ref, maxrel(out, out_ref), frac_equal))
    assert err_new <= max(1.5 * err_ref, FWD_TOL)        # no less accurate than the reference path
    go = torch.randn_like(out)
    out.backward(go)
    out_ref.backward(go)
    i64, f64 = img.double().requires_grad_(), flow.double().requires_grad_()
    tr.networks_resample(i64, f64).backward(go.double())
    assert maxrel(it.grad, i64.grad) <= max(1.5 * maxrel(ir.grad, i64.grad), BWD_TOL)
    assert maxrel(ft.grad, f64.grad) <= max(1.5 * maxrel(fr.grad, f64.grad), BWD_TOL)


def test_networks_resample_smooth_frame_meets_1e5(ops):
    """On band-limited frames (what the generator warps) the 1e-5 forward target is met outright."""
    torch.manual_seed(1)
    H, W = 256, 512
    low = torch.randn(1, 3, H // 16, W // 16, device="cuda")
    img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
    flow = 5 * torch.randn(1, 2, H, W, device="cuda")
    out = ops.networks.resample(img, flow)
    truth = tr.networks_resample(img.double(), flow.double())
    assert maxrel(out, truth) <= FWD_TOL


# =============================================================================================
# Correlation
# =============================================================================================
FLOWNETC = (20, 1, 20, 1, 2)


@pytest.mark.parametrize("shape,params", [
    ((2, 16, 8, 12), FLOWNETC), ((1, 40, 6, 40), FLOWNETC), ((1, 8, 9, 11), (4, 1, 4, 1, 1)),
    ((1, 3, 5, 70), FLOWNETC), ((1, 9, 7, 5), FLOWNETC), ((2, 64, 24, 32), FLOWNETC),
    ((1, 8, 6, 6), (6, 1, 6, 1, 2)), ((1, 4, 6, 7), (3, 3, 2, 1, 1)),
])
def test_correlation_vs_oracle(ops, c_oracle, shape, params):
    rng = np.random.default_rng(12)
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    at, bt = cu(a).requires_grad_(), cu(b).requires_grad_()
    out = ops.Correlation(*params, 1)(at, bt)
    out_ref = c_oracle.corr_fwd(a, b, *params)
    assert tuple(out.shape) == out_ref.shape
    assert maxrel(out, out_ref) <= FWD_TOL
    go = rng.standard_normal(out_ref.shape).astype(np.float32)
    out.backward(cu(go))
    ga_ref, gb_ref = c_oracle.corr_bwd(a, b, go, *params)
    assert maxrel(at.grad, ga_ref) <= BWD_TOL
    assert maxrel(bt.grad, gb_ref) <= BWD_TOL


def test_correlation_stride1_forward_only(ops, c_oracle):
    rng = np.random.default_rng(13)
    a = rng.standard_normal((1, 8, 10, 10)).astype(np.float32)
    b = rng.standard_normal((1, 8, 10, 10)).astype(np.float32)
    out = ops.Correlation(4, 1, 4, 2, 2, 1)(cu(a), cu(b))
    assert maxrel(out, c_oracle.corr_fwd(a, b, 4, 1, 4, 2, 2)) <= FWD_TOL
    at = cu(a).requires_grad_()
    with pytest.raises(NotImplementedError):
        ops.Correlation(4, 1, 4, 2, 2, 1)(at, cu(b)).sum().backward()


def test_correlation_golden(ops, golden_native):
    g = golden_native
    for name in ["corr_c", "corr_wide", "corr_s1", "corr_s2"]:
        p = [int(v) for v in g[name + "_params"]]
        at, bt = cu(g[name + "_a"]).requires_grad_(), cu(g[name + "_b"]).requires_grad_()
        out = ops.Correlation(*p, 1)(at, bt)
        assert maxrel(out, g[name + "_out"]) <= FWD_TOL
        if name + "_gout" in g:
            out.backward(cu(g[name + "_gout"]))
            assert maxrel(at.grad, g[name + "_ga"]) <= BWD_TOL
            assert maxrel(bt.grad, g[name + "_gb"]) <= BWD_TOL


def test_correlation_config2_vs_reference_ext_and_fp64(ops, ref):
    """BASELINE config 2: 8x256x48x64, FlowNetC parameters, fwd + bwd."""
    torch.manual_seed(0)
    a = torch.randn(8, 256, 48, 64, device="cuda")
    b = torch.randn(8, 256, 48, 64, device="cuda")
    at, bt = a.clone().requires_grad_(), b.clone().requires_grad_()
    out = ops.Correlation(*FLOWNETC, 1)(at, bt)
    out_ref = ref.correlation_forward(a, b, *FLOWNETC)
    assert maxrel(out, out_ref) <= FWD_TOL
    # fp64 truth on one batch item: the FP32-FMA kernel is no further from it than the reference kernel; the tensor-core
    # kernel (3xTF32, fp32 accumulation inside the tensor core over C / 8 x 3 UMMA steps) stays within half the tolerance
    truth = tr.correlation(a[:1].double(), b[:1].double(), *FLOWNETC)
    from ir2rgb_b200 import _lib
    tc = _lib.load().flowops_corr_get_impl() & 1
    assert maxrel(out[:1], truth) <= (5e-6 if tc else max(1.5 * maxrel(out_ref[:1], truth), 1e-6))
    go = torch.randn_like(out)
    out.backward(go)
    ga_ref, gb_ref = ref.correlation_backward(a, b, go, *FLOWNETC)
    assert maxrel(at.grad, ga_ref) <= BWD_TOL
    assert maxrel(bt.grad, gb_ref) <= BWD_TOL


def test_correlation_flownet2_shape_vs_reference_ext(ops, ref):
    """The FlowNet2 operating point of BASELINE config 4 (conv3 features of 512x1024 frames, 16 pairs): forward
    against the reference's own kernel at the full size, backward on a batch slice of 2 (the reference backward
    launches 2*B single-warp-block kernels and takes seconds per item)."""
    torch.manual_seed(2)
    a = torch.randn(16, 256, 64, 128, device="cuda")
    b = torch.randn(16, 256, 64, 128, device="cuda")
    out = ops.Correlation(*FLOWNETC, 1)(a, b)
    out_ref = ref.correlation_forward(a, b, *FLOWNETC)
    assert out.shape == out_ref.shape == (16, 441, 64, 128)
    assert maxrel(out, out_ref) <= FWD_TOL
    del out_ref
    at, bt = a[:2].clone().requires_grad_(), b[:2].clone().requires_grad_()
    o2 = ops.Correlation(*FLOWNETC, 1)(at, bt)
    assert torch.equal(o2.detach(), out[:2])                  # batch rows are independent: a slice gives the same bits
    go = torch.randn_like(o2)
    o2.backward(go)
    ga_ref, gb_ref = ref.correlation_backward(a[:2].contiguous(), b[:2].contiguous(), go, *FLOWNETC)
    assert maxrel(at.grad, ga_ref) <= BWD_TOL
    assert maxrel(bt.grad, gb_ref) <= BWD_TOL


def test_correlation_properties_full_size(ops):
    """Size-independent properties at the FlowNet2 512x1024 feature size (B x 256 x 64 x 128)."""
    torch.manual_seed(1)
    a = torch.randn(2, 256, 64, 128, device="cuda")
    b = torch.randn(2, 256, 64, 128, device="cuda")
    corr = ops.Correlation(*FLOWNETC, 1)
    out = corr(a, b)
    assert out.shape == (2, 441, 64, 128)
    # centre channel (tj = ti = 0) is the per-pixel mean of a*b
    assert maxrel(out[:, 220], (a * b).mean(1)) <= FWD_TOL
    # bilinearity
    assert maxrel(corr(2 * a, b), 2 * out) <= FWD_TOL
    assert maxrel(corr(a, b + a), out + corr(a, a)) <= 2e-5
    # displacement structure: channel (tj, ti) equals mean_c a * shift(b); spot-check three taps incl. padding
    for tj, ti in [(-10, -10), (3, -7), (10, 10)]:
        sh = torch.zeros_like(b)
        dy, dx = 2 * tj, 2 * ti
        ys, ye = max(0, -dy), min(64, 64 - dy)
        xs, xe = max(0, -dx), min(128, 128 - dx)
        sh[:, :, ys:ye, xs:xe] = b[:, :, ys + dy:ye + dy, xs + dx:xe + dx]
        assert maxrel(out[:, (tj + 10) * 21 + ti + 10], (a * sh).mean(1)) <= FWD_TOL


def test_correlation_errors(ops):
    a = torch.randn(1, 4, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        ops.Correlation(*FLOWNETC, 1)(a, torch.randn(1, 4, 8, 9, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.Correlation(*FLOWNETC, 1)(a.cpu(), a.cpu())
    out = ops.Correlation(*FLOWNETC, 1)(torch.empty(0, 4, 8, 8, device="cuda"), torch.empty(0, 4, 8, 8, device="cuda"))
    assert out.shape == (0, 441, 8, 8)


def test_ops_are_graph_capturable_and_stream_ordered(ops):
    """No allocation-free guarantee is needed from torch here: the library calls themselves must not
    synchronise or touch the default stream."""
    a = torch.randn(1, 16, 16, 16, device="cuda")
    b = torch.randn(1, 16, 16, 16, device="cuda")
    corr = ops.Correlation(*FLOWNETC, 1)
    expect = corr(a, b)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            corr(a, b)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = corr(a, b)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, expect)


@pytest.mark.parametrize("shape", [(2, 64, 24, 32), (1, 40, 6, 70), (1, 9, 7, 5), (2, 256, 64, 128)])
def test_correlation_channels_last_inputs(ops, shape):
    """A channels_last conv body hands its features over without a layout copy: same planes, same result."""
    torch.manual_seed(14)
    a, b = torch.randn(shape, device="cuda"), torch.randn(shape, device="cuda")
    corr = ops.Correlation(*FLOWNETC, 1)
    ref_out = corr(a, b)
    a_cl, b_cl = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    out = corr(a_cl, b_cl)
    assert out.is_contiguous() and torch.equal(out, ref_out)
    # mixed layouts fall back to the NCHW path
    assert torch.equal(corr(a_cl, b), ref_out)


# =============================================================================================
# inference glue for a channels_last FlowNet2 body: concat buffers filled slice by slice
# =============================================================================================
@pytest.mark.parametrize("shape", [(2, 64, 24, 40), (1, 256, 64, 128), (1, 32, 13, 21)])
def test_correlation_nhwc_output_with_lrelu_matches_nchw_path(ops, shape):
    """flowops_corr_fwd_planes_nhwc: the cost volume written channels-last into a slice of the (padded) concat
    buffer with corr_activation folded in == correlation -> LeakyReLU -> cat, bit for bit (FlowNetC.py:89-94)."""
    from ir2rgb_b200 import functional as F
    torch.manual_seed(21)
    B, C, H, W = shape
    a = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    b = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    want = torch.nn.functional.leaky_relu(F.correlation_forward(a, b, 20, 1, 20, 1, 2), 0.1)
    planes = F.CorrelationPlanes(a.shape, a.device)
    zero_bias = torch.zeros(C, device="cuda")
    planes.fill_from_conv_(a.clone(memory_format=torch.channels_last), zero_bias, 1.0, 0, write_act=False)   # slope 1: identity
    planes.fill_from_conv_(b.clone(memory_format=torch.channels_last), zero_bias, 1.0, 1, write_act=False)
    buf = F.ConcatBuffer(a, 32 + 441, 8)
    assert buf.c_pad == 480 and buf.tensor.is_contiguous(memory_format=torch.channels_last)
    buf.tensor[:, :32] = 7.0
    end = F.correlation_planes_forward_into(planes, buf, 32, 0.1)
    assert end == 473
    assert torch.equal(buf.tensor[:, 32:473], want)
    assert (buf.tensor[:, :32] == 7.0).all() and (buf.tensor[:, 473:] == 0).all()


@pytest.mark.parametrize("c,c_total,c_off", [(64, 130, 64), (30, 45, 7), (512, 1026, 512)])
def test_epilogue_into_concat_buffer(ops, c, c_total, c_off):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(22)
    y = torch.randn(2, c, 12, 20, device="cuda").contiguous(memory_format=torch.channels_last)
    bias = torch.randn(c, device="cuda")
    buf = F.ConcatBuffer(y, c_total, 8)
    assert buf.c_pad % 8 == 0 and buf.c_pad >= c_total
    buf.tensor[:, :c_total] = -3.0
    buf.bias_lrelu_in(y, bias, 0.1, c_off)
    want = torch.nn.functional.leaky_relu(y + bias.view(1, -1, 1, 1), 0.1)
    assert torch.equal(buf.tensor[:, c_off:c_off + c], want)
    untouched = torch.ones(buf.c_pad, dtype=torch.bool, device="cuda")
    untouched[c_off:c_off + c] = False
    untouched[c_total:] = False
    assert (buf.tensor[:, untouched] == -3.0).all() and (buf.tensor[:, c_total:] == 0).all()
    # padded concat == torch.cat plus zero channels
    parts = (torch.randn(2, 5, 12, 20, device="cuda").contiguous(memory_format=torch.channels_last), y)
    with torch.no_grad():
        cat = F.cat_channels(parts, pad_to=8)
    assert cat.shape[1] == -(-(c + 5) // 8) * 8 and torch.equal(cat[:, :c + 5], torch.cat(parts, 1)) and (cat[:, c + 5:] == 0).all()


# =============================================================================================
# warp backward with owned per-warp accumulation windows (flowops_warp_set_impl bit 0)
# =============================================================================================
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", [(2, 3, 17, 28, 3.0), (1, 1, 40, 64, 10.0), (2, 2, 33, 100, 40.0), (1, 3, 64, 96, 0.0),
                                  (3, 3, 70, 132, 0.5), (1, 3, 256, 512, 5.0)])
def test_warp_backward_window_kernel_matches_direct_reductions(flowops_lib, c_oracle, case, mode):
    """The shared-memory-window image gradient (warp_win_bwd.cuh) against the direct-reduction kernel and, on the small
    cases, the C oracle: same sums up to the order of the additions."""
    from ir2rgb_b200 import functional as F
    B, C, H, W, amp = case
    torch.manual_seed(21)
    img = torch.randn(B, C, H, W, device="cuda")
    flow = (amp * torch.randn(B, 2, H, W, device="cuda")).contiguous()
    go = torch.randn(B, C, H, W, device="cuda")
    prev = flowops_lib.flowops_warp_get_impl()
    try:
        flowops_lib.flowops_warp_set_impl(0)
        gi_d, gf_d = F.warp_backward(img, flow, go, True, True, mode)
        flowops_lib.flowops_warp_set_impl(1)
        gi_w, gf_w = F.warp_backward(img, flow, go, True, True, mode)
        gi_w2, _ = F.warp_backward(img, flow, go, True, False, mode)
    finally:
        flowops_lib.flowops_warp_set_impl(prev)
    assert torch.equal(gf_w, gf_d)                       # the flow gradient is the same gather
    assert maxrel(gi_w, gi_d) <= 1e-5 and maxrel(gi_w2, gi_d) <= 1e-5
    if mode == 0 and H * W <= 4096:
        gi_o, _ = c_oracle.resample2d_bwd(img.cpu().numpy(), flow.cpu().numpy(), go.cpu().numpy())
        assert maxrel(gi_w, gi_o) <= BWD_TOL


@pytest.mark.parametrize("case", [(2, 3, 40, 56, 3.0), (1, 2, 33, 70, 0.0), (2, 3, 64, 96, 40.0), (1, 1, 7, 5, 2.0), (2, 3, 128, 160, 1.0),
                                  (1, 3, 96, 64, "nan"), (1, 3, 64, 64, "huge"), (1, 3, 64, 64, "tiny")])
@pytest.mark.parametrize("mode", [0, 1])
def test_warp_backward_fixed_point_kernel_matches_direct_reductions(flowops_lib, c_oracle, case, mode):
    """The fixed-point shared-memory image gradient (warp_fx_bwd.cuh, flowops_warp_set_impl(4)) against the direct-reduction
    kernel and, on the small cases, the C oracle: tiles that fit the window (small flows), tiles that fall back (40-pixel
    noise), ragged edges, NaN flows and Inf gradients (must propagate exactly as in the direct kernel), gradients of 1e30 and
    1e-25 (the scaling is relative to the tile's largest gradient)."""
    from ir2rgb_b200 import functional as F
    B, C, H, W, amp = case
    torch.manual_seed(23)
    img = torch.randn(B, C, H, W, device="cuda")
    go = torch.randn(B, C, H, W, device="cuda")
    if isinstance(amp, str):
        flow = (2.0 * torch.randn(B, 2, H, W, device="cuda")).contiguous()
        if amp == "nan":
            flow[0, 0, 10, 12] = float("nan")
            flow[0, 1, 50, 3] = float("inf")
            go[0, 1, 70, 40] = float("inf")
        elif amp == "huge":
            go *= 1e30
        else:
            go *= 1e-25
    else:
        flow = (amp * torch.randn(B, 2, H, W, device="cuda")).contiguous()
    prev = flowops_lib.flowops_warp_get_impl()
    try:
        flowops_lib.flowops_warp_set_impl(0)
        gi_d, gf_d = F.warp_backward(img, flow, go, True, True, mode)
        flowops_lib.flowops_warp_set_impl(4)
        gi_f, gf_f = F.warp_backward(img, flow, go, True, True, mode)
        gi_f2, _ = F.warp_backward(img, flow, go, True, False, mode)
    finally:
        flowops_lib.flowops_warp_set_impl(prev)
    assert torch.equal(gf_f, gf_d, ) or torch.equal(torch.nan_to_num(gf_f, nan=7.0), torch.nan_to_num(gf_d, nan=7.0))     # same gather
    if amp == "nan":
        assert torch.equal(torch.isnan(gi_f), torch.isnan(gi_d)) and torch.equal(torch.isinf(gi_f), torch.isinf(gi_d))
        ok = torch.isfinite(gi_d)
        assert maxrel(gi_f[ok], gi_d[ok]) <= 1e-5
    else:
        assert maxrel(gi_f, gi_d) <= 1e-5 and maxrel(gi_f2, gi_d) <= 1e-5
    if mode == 0 and H * W <= 4096 and not isinstance(amp, str):
        gi_o, _ = c_oracle.resample2d_bwd(img.cpu().numpy(), flow.cpu().numpy(), go.cpu().numpy())
        assert maxrel(gi_f, gi_o) <= BWD_TOL


@pytest.mark.parametrize("flavour", ["randn", "smooth", "border"])
def test_resample2d_tolerance_mode_vs_reference_ext(ops, ref, flavour):
    """fp32-weight blend (functional.warp_tolerance_mode, what FlowNet runs): within 1e-6 of the reference's kernel --
    ten times inside the 1e-5 tolerance -- and the default path stays bit-identical to it."""
    from ir2rgb_b200 import functional as F
    torch.manual_seed(31)
    B, H, W = 4, 256, 512
    img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
    if flavour == "randn":
        flow = 4 * torch.randn(B, 2, H, W, device="cuda")
    elif flavour == "smooth":
        flow = torch.nn.functional.interpolate(20 * torch.randn(B, 2, 4, 8, device="cuda"), size=(H, W), mode="bicubic").contiguous()
    else:
        flow = 80 * (torch.rand(B, 2, H, W, device="cuda") - 0.5)
    want = ref.resample2d_forward(img, flow)
    with F.warp_tolerance_mode(True):
        fast = ops.Resample2d()(img, flow)
        x6 = torch.cat([img, img.flip(0)], 1).contiguous()
        warped_f, norm_f = F.warp_diff_norm_forward(x6, flow)
    exact = ops.Resample2d()(img, flow)
    assert torch.equal(exact, want)
    assert maxrel(fast, want) <= 1e-6
    warped_e, norm_e = F.warp_diff_norm_forward(x6, flow)
    assert maxrel(warped_f, warped_e) <= 1e-6 and maxrel(norm_f, norm_e) <= 1e-6


@pytest.mark.parametrize("C", [1, 2, 3])
def test_gridsample_forward_two_rows_in_flight_kernel_is_bit_identical(flowops_lib, C):
    """Large frames take warp_rows_mlp_kernel in GRIDSAMPLE mode (two rows of corner gathers in flight per thread); flag bit 3
    forces the row-walking kernel, which the goldens and the ATen comparisons above pin.  Ragged frame (neither dimension a
    multiple of the block shape, bottom strip shorter than 4 rows), NaN / Inf / huge flow values."""
    from ir2rgb_b200 import functional as F
    B, H, W = 3, 1021, 1531                                   # > 3.03 M pixels: 4-row strips
    torch.manual_seed(31)
    img = 2 * torch.rand(B, C, H, W, device="cuda") - 1
    flow = torch.nn.functional.interpolate(30 * torch.randn(B, 2, 16, 24, device="cuda"), size=(H, W), mode="bicubic",
                                           align_corners=False).contiguous()
    flow[0, 0, 5, 7] = float("nan"); flow[1, 1, 100, 200] = float("inf"); flow[2, 0, 0, 0] = -3e38
    flow[2, 1, H - 1, W - 1] = 5e6; flow[1, :, H - 2:, :] = 4 * torch.randn(2, 2, W, device="cuda")
    prev = flowops_lib.flowops_warp_get_impl()
    try:
        flowops_lib.flowops_warp_set_impl(prev & ~8)
        new = F.warp_forward(img, flow, F.WARP_GRIDSAMPLE)
        flowops_lib.flowops_warp_set_impl(prev | 8)
        old = F.warp_forward(img, flow, F.WARP_GRIDSAMPLE)
    finally:
        flowops_lib.flowops_warp_set_impl(prev)
    assert torch.equal(new.view(torch.int32), old.view(torch.int32))
