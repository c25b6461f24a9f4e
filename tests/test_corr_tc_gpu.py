"""The tcgen05 (3xTF32) Correlation forward, csrc/corr_tc.cu: parity against the reference's own CUDA extension at both
BASELINE shapes, against the FP32-FMA kernel of the same library, through every entry point that reaches it, and the
edge cases of its tiling (partial tiles, single tile, zero features, large dynamic range)."""
import numpy as np
import pytest
import torch

from oracle import torch_ref as tr

pytestmark = pytest.mark.gpu

P = (20, 1, 20, 1, 2)
FWD_TOL = 1e-5


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture()
def lib(flowops_lib):
    prev = flowops_lib.flowops_corr_get_impl()
    flowops_lib.flowops_corr_set_impl(1)
    yield flowops_lib
    flowops_lib.flowops_corr_set_impl(prev)


def both(lib, a, b):
    from ir2rgb_b200 import functional as F
    lib.flowops_corr_set_impl(0)
    ffma = F.correlation_forward(a, b, *P)
    lib.flowops_corr_set_impl(1)
    tc = F.correlation_forward(a, b, *P)
    return tc, ffma


@pytest.mark.parametrize("shape", [(8, 256, 48, 64), (16, 256, 64, 128)])
def test_tc_vs_reference_extension_at_baseline_shapes(lib, shape):
    from ir2rgb_b200 import functional as F
    from oracle import ref_ext
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    torch.manual_seed(0)
    a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
    out = F.correlation_forward(a, b, *P)
    ref = ref_ext.correlation_forward(a, b, *P)
    assert maxrel(out, ref) <= FWD_TOL
    truth = tr.correlation(a[:1].double(), b[:1].double(), *P)
    assert maxrel(out[:1], truth) <= 5e-6          # half the tolerance against fp64 truth


@pytest.mark.parametrize("shape", [(1, 32, 2, 2), (1, 32, 32, 16), (2, 64, 34, 18), (1, 96, 6, 70), (3, 32, 66, 34), (1, 256, 20, 12)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_tc_matches_fp32_kernel_on_awkward_tilings(lib, shape, layout):
    torch.manual_seed(1)
    a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
    if layout == "nhwc":
        a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    tc, ffma = both(lib, a, b)
    assert tc.shape == ffma.shape == (shape[0], 441, shape[2], shape[3])
    assert maxrel(tc, ffma) <= FWD_TOL


def test_tc_unsupported_shapes_fall_back_to_the_fp32_kernel(lib):
    """C % 32 != 0 or odd H / W: the FP32-FMA kernel runs (bit-identical to it with the switch off)."""
    torch.manual_seed(2)
    for shape in [(1, 40, 8, 12), (1, 32, 9, 12), (1, 32, 8, 11)]:
        a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
        tc, ffma = both(lib, a, b)
        assert torch.equal(tc, ffma)


def test_tc_special_inputs(lib):
    torch.manual_seed(3)
    a = torch.randn(1, 64, 16, 24, device="cuda")
    z = torch.zeros_like(a)
    tc, ffma = both(lib, a, z)
    assert torch.count_nonzero(tc) == 0 and torch.count_nonzero(ffma) == 0
    # scale invariance: the hi/lo split is relative, so tiny and huge features keep the same relative accuracy
    for s in (1e-20, 1e15):
        tc, ffma = both(lib, a * s, a.flip(3) * s)
        assert maxrel(tc, ffma) <= FWD_TOL
    # centre channel = per-pixel mean of a*b; bilinearity in the first argument
    b = torch.randn_like(a)
    tc, _ = both(lib, a, b)
    assert maxrel(tc[:, 220], (a * b).mean(1)) <= FWD_TOL
    tc2, _ = both(lib, 3 * a, b)
    assert maxrel(tc2, 3 * tc) <= FWD_TOL


def test_tc_split_entry_points_and_channels_last_store(lib):
    """What FlowNetC runs: planes written by the conv3 epilogue (bias + LeakyReLU), cost volume stored channels-last with
    LeakyReLU into the conv3_1 concat buffer -- against the same calls with the FP32-FMA kernel."""
    from ir2rgb_b200 import functional as F
    torch.manual_seed(4)
    shape = (2, 256, 32, 48)
    ya = torch.randn(*shape, device="cuda").contiguous(memory_format=torch.channels_last)
    yb = torch.randn(*shape, device="cuda").contiguous(memory_format=torch.channels_last)
    bias = torch.randn(256, device="cuda")
    outs = []
    for impl in (0, 1):
        lib.flowops_corr_set_impl(impl)
        planes = F.CorrelationPlanes(ya.shape, ya.device)
        a_act = ya.clone()
        planes.fill_from_conv_(a_act, bias, 0.1, 0, write_act=True)
        planes.fill_from_conv_(yb.clone(), bias, 0.1, 1, write_act=False)
        buf = F.ConcatBuffer(ya, 473, 8)
        F.correlation_planes_forward_into(planes, buf, 32, 0.1)
        nchw = F.correlation_planes_forward(planes)
        outs.append((a_act, buf.tensor[:, 32:473].clone(), nchw))
    lib.flowops_corr_set_impl(1)
    assert torch.equal(outs[0][0], outs[1][0])                     # activated frame-0 features written back: same bits
    assert maxrel(outs[1][1], outs[0][1]) <= FWD_TOL
    assert maxrel(outs[1][2], outs[0][2]) <= FWD_TOL
    want = torch.nn.functional.leaky_relu(outs[1][2], 0.1)
    assert maxrel(outs[1][1], want) <= 1e-7


def test_tc_is_graph_capturable_and_deterministic(lib):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(5)
    a, b = torch.randn(2, 64, 32, 32, device="cuda"), torch.randn(2, 64, 32, 32, device="cuda")
    eager = F.correlation_forward(a, b, *P)
    assert torch.equal(eager, F.correlation_forward(a, b, *P))      # no atomics anywhere: bit-reproducible
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        F.correlation_forward(a, b, *P)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = F.correlation_forward(a, b, *P)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


# ---------------------------------------------------------------------------------------------------------------
# backward (csrc/corr_tc_bwd.cu)
# ---------------------------------------------------------------------------------------------------------------
BWD_TOL = 1e-4


def both_bwd(lib, a, b, go, **kw):
    from ir2rgb_b200 import functional as F
    lib.flowops_corr_set_impl(0)
    ffma = F.correlation_backward(a, b, go, *P, **kw)
    lib.flowops_corr_set_impl(1)
    tc = F.correlation_backward(a, b, go, *P, **kw)
    return tc, ffma


@pytest.mark.parametrize("shape", [(8, 256, 48, 64), (4, 256, 64, 128)])
def test_tc_backward_vs_reference_extension_at_baseline_shapes(lib, shape):
    """Config 2 at its stated size and the FlowNet2 feature shape, against the reference's own backward kernels."""
    from ir2rgb_b200 import functional as F
    from oracle import ref_ext
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    torch.manual_seed(0)
    a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
    go = torch.randn(shape[0], 441, shape[2], shape[3], device="cuda")
    g1, g2 = F.correlation_backward(a, b, go, *P)
    r1, r2 = ref_ext.correlation_backward(a, b, go, *P)
    assert maxrel(g1, r1) <= BWD_TOL and maxrel(g2, r2) <= BWD_TOL
    # fp64 truth by autograd of the closed form on one pair
    a1, b1 = a[:1].double().requires_grad_(), b[:1].double().requires_grad_()
    tr.correlation(a1, b1, *P).backward(go[:1].double())
    assert maxrel(g1[:1], a1.grad) <= 1e-5 and maxrel(g2[:1], b1.grad) <= 1e-5


@pytest.mark.parametrize("shape", [(1, 32, 2, 2), (1, 32, 32, 16), (2, 64, 34, 18), (1, 96, 6, 70), (3, 32, 66, 34), (1, 256, 20, 12),
                                   (1, 512, 8, 8)])
def test_tc_backward_matches_fp32_kernel_on_awkward_tilings(lib, shape):
    """Partial tiles, both tile shapes (16 x 8 and 8 x 16 pixels), single tiles, two passes of 256 channels."""
    torch.manual_seed(1)
    a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
    go = torch.randn(shape[0], 441, shape[2], shape[3], device="cuda")
    (t1, t2), (f1, f2) = both_bwd(lib, a, b, go)
    assert maxrel(t1, f1) <= 1e-5 and maxrel(t2, f2) <= 1e-5
    # one gradient at a time: the same bits as when both are computed
    (o1, none2), _ = both_bwd(lib, a, b, go, need1=True, need2=False)
    (none1, o2), _ = both_bwd(lib, a, b, go, need1=False, need2=True)
    assert none1 is None and none2 is None and torch.equal(o1, t1) and torch.equal(o2, t2)


def test_tc_backward_variants_and_fallbacks(lib):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(2)
    a, b = torch.randn(2, 64, 20, 36, device="cuda"), torch.randn(2, 64, 20, 36, device="cuda")
    go = torch.randn(2, 441, 20, 36, device="cuda")
    lib.flowops_corr_set_impl(1)
    t1, t2 = F.correlation_backward(a, b, go, *P)
    lib.flowops_corr_set_impl(3)                                   # A tiles in shared memory instead of TMEM: same UMMAs, same bits
    s1, s2 = F.correlation_backward(a, b, go, *P)
    lib.flowops_corr_set_impl(1)
    assert torch.equal(s1, t1) and torch.equal(s2, t2)
    assert torch.equal(F.correlation_backward(a, b, go, *P)[0], t1)           # no atomics anywhere: bit-reproducible
    # shapes the tensor-core kernel does not take run the FP32-FMA kernel
    for shape in [(1, 40, 8, 12), (1, 32, 9, 12), (1, 320, 8, 8)]:
        a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
        go = torch.randn(shape[0], 441, shape[2], shape[3], device="cuda")
        (t1, t2), (f1, f2) = both_bwd(lib, a, b, go)
        assert torch.equal(t1, f1) and torch.equal(t2, f2)


def test_tc_backward_special_inputs_and_linearity(lib):
    torch.manual_seed(3)
    a, b = torch.randn(1, 64, 16, 24, device="cuda"), torch.randn(1, 64, 16, 24, device="cuda")
    go = torch.randn(1, 441, 16, 24, device="cuda")
    (t1, t2), _ = both_bwd(lib, a, b, torch.zeros_like(go))
    assert torch.count_nonzero(t1) == 0 and torch.count_nonzero(t2) == 0
    # the centre displacement alone: gI1 = gO_centre * f2 / C, gI2 = gO_centre * f1 / C
    gc = torch.zeros_like(go)
    gc[:, 220] = go[:, 220]
    (t1, t2), _ = both_bwd(lib, a, b, gc)
    assert maxrel(t1, go[:, 220:221] * b / 64) <= 1e-5 and maxrel(t2, go[:, 220:221] * a / 64) <= 1e-5
    # linear in gO, and scale-invariant accuracy (the hi / lo split is relative)
    (u1, u2), (f1, f2) = both_bwd(lib, a * 1e-12, b * 1e9, go * 1e6)
    assert maxrel(u1, f1) <= 1e-5 and maxrel(u2, f2) <= 1e-5


def test_tc_backward_through_autograd_and_graph_capture(lib):
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks.correlation_package.correlation import Correlation
    torch.manual_seed(4)
    a = torch.randn(2, 32, 16, 16, device="cuda", requires_grad=True)
    b = torch.randn(2, 32, 16, 16, device="cuda", requires_grad=True)
    out = Correlation(20, 1, 20, 1, 2, 1)(a, b)
    go = torch.randn_like(out)
    out.backward(go)
    a64, b64 = a.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    tr.correlation(a64, b64, *P).backward(go.double())
    assert maxrel(a.grad, a64.grad) <= 1e-5 and maxrel(b.grad, b64.grad) <= 1e-5
    ad, bd = a.detach(), b.detach()
    eager = F.correlation_backward(ad, bd, go, *P)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        F.correlation_backward(ad, bd, go, *P)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cap = F.correlation_backward(ad, bd, go, *P)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(cap[0], eager[0]) and torch.equal(cap[1], eager[1])
