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
