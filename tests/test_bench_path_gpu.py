"""The configuration bench.py actually times -- channels_last conv body, every fusion on, 8-channel-padded concat
buffers, the whole FlowNet.forward replayed from a CUDA graph -- compared DIRECTLY with the reference network at the
benchmark's frame size (512 x 1024): the reference's own `FlowNet2` class with its own wrappers and CUDA extensions
(oracle/_ref, kind "refclass") and, as a second witness, the restated architecture with the reference's extensions
("ref").  Weights are shared through the state dict.

Tolerances: the operators agree to <= 1e-5 (tests/test_ops_gpu.py); ~100 convolution layers amplify that, so the
flow is held to max-relative <= 1e-3 with fp32 convolutions; the confidence mask is a hard threshold and is
compared by flip fraction (<= 1 %).  With cuDNN's TF32 convolutions (bench.py's setting, torch's default, the same in
both arms) the two networks run different cuDNN kernels (NHWC vs NCHW) whose TF32 roundings differ: that number
is recorded (gpurun_out/bench_path_parity.json) and only bounded loosely.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 512, 1024


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _record(key, value):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "bench_path_parity.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = value
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


@pytest.fixture(scope="module")
def bench_net(flowops_lib):
    """Exactly bench.build_native('channels_last') wrapped in GraphedFlowNet."""
    import bench
    from ir2rgb_b200.runtime import GraphedFlowNet
    torch.manual_seed(0)
    eager = bench.build_native(torch.device("cuda", 0), "channels_last")
    return eager, GraphedFlowNet(eager)


def _frames(B, seed):
    torch.manual_seed(seed)
    im1 = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
    im2 = (im1.roll(shifts=(2, -3), dims=(2, 3)) + 0.05 * torch.randn_like(im1)).clamp(-1, 1)     # a real displacement
    return im1, im2


@pytest.mark.parametrize("kind", ["refclass", "ref"])
def test_benchmarked_path_matches_reference_network_fp32_convs(bench_net, kind):
    from oracle import harness, ref_ext
    if kind == "refclass" and not harness.reference_flownet2_available():
        pytest.skip("oracle/_ref/refpy not built (needs /root/reference at build time)")
    if kind == "ref" and not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    eager, graphed = bench_net
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = False, False
    try:
        graphed.reset()                                   # graphs captured under another conv-math setting are stale
        im1, im2 = _frames(2, 11)
        flow_new, conf_new = graphed(im1, im2)
        other = harness.OracleFlowNet(kind, "cuda", state_dict=eager.flowNet.state_dict())
        flow_ref, conf_ref = other(im1, im2)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
        graphed.reset()
    assert flow_new.shape == flow_ref.shape == (2, 2, H, W)
    rel = maxrel(flow_new, flow_ref)
    flips = (conf_new != conf_ref).float().mean().item()
    _record("fp32_convs_vs_" + kind, {"flow_maxrel": rel, "conf_flip_fraction": flips, "frame": [H, W], "batch": 2})
    assert rel <= 1e-3, rel
    assert flips <= 0.01, flips


def test_benchmarked_path_with_bench_conv_math_is_recorded(bench_net):
    """bench.py's own setting: cuDNN TF32 convolutions + autotuning in both arms."""
    from oracle import harness
    if not harness.reference_flownet2_available():
        pytest.skip("oracle/_ref/refpy not built")
    eager, graphed = bench_net
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = True, True
    try:
        graphed.reset()
        im1, im2 = _frames(2, 12)
        flow_new, conf_new = graphed(im1, im2)
        other = harness.OracleFlowNet("refclass", "cuda", state_dict=eager.flowNet.state_dict())
        flow_ref, conf_ref = other(im1, im2)
        # how far TF32 alone moves the REFERENCE network: its fp32-conv result is the yardstick
        torch.backends.cudnn.allow_tf32 = False
        flow_ref32, _ = other(im1, im2)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
        graphed.reset()
    rel = maxrel(flow_new, flow_ref)
    ref_tf32_noise = maxrel(flow_ref, flow_ref32)
    flips = (conf_new != conf_ref).float().mean().item()
    _record("tf32_convs_vs_refclass", {"flow_maxrel": rel, "conf_flip_fraction": flips,
                                       "reference_tf32_vs_its_own_fp32_maxrel": ref_tf32_noise})
    # the new path is no further from the reference than TF32 moves the reference from itself (x4 head-room)
    assert rel <= max(4 * ref_tf32_noise, 1e-3), (rel, ref_tf32_noise)
    assert flips <= 0.05, flips


def test_reference_class_arm_makes_no_libflowops_call():
    from ir2rgb_b200 import _lib
    from oracle import harness
    if not harness.reference_flownet2_available():
        pytest.skip("oracle/_ref/refpy not built")
    calls = []
    prev = _lib.launch_hook
    _lib.launch_hook = lambda what, n: calls.append(what)
    try:
        torch.manual_seed(1)
        net = harness.OracleFlowNet("refclass", "cuda")
        assert type(net.flowNet).__module__ == "flownet2_pytorch.models"          # the reference's class, not the restatement
        assert type(net.flowNet.flownetc.corr).__module__.startswith("flownet2_pytorch.networks.correlation_package")
        im = 2 * torch.rand(1, 3, 64, 128, device="cuda") - 1
        flow, conf = net(im, im.flip(3))
        assert flow.shape == (1, 2, 64, 128) and torch.isfinite(flow).all()
    finally:
        _lib.launch_hook = prev
    assert calls == [], calls


def test_sd_branch_on_second_stream_changes_nothing(bench_net):
    """FlowNetSD forked onto a second stream beside FlowNetC -> S -> S (FlowNet2.overlap_sd; reference models.py:141-142
    is independent of :104-137): the same kernels in another schedule, so flow and confidence must be bit-identical to
    the in-line order -- launched eagerly and replayed from the CUDA graph.  (Only the captured graphs are dropped
    between the two settings: the per-layer plan choices made by timing must stay the same on both sides.)"""
    eager, graphed = bench_net
    fn = eager.flowNet
    im1, im2 = _frames(2, 13)
    prev = fn.overlap_sd
    try:
        graphed.reset()
        fn.overlap_sd = False
        with torch.no_grad():
            eager(im1, im2)                               # builds every plan of this shape
            f_serial, c_serial = eager(im1, im2)
            f_serial_g, c_serial_g = graphed(im1, im2)
        fn.overlap_sd = True
        graphed._graphs.clear()
        with torch.no_grad():
            assert any(k[:2] == (im1.device.index, (2, 3, 2, H, W)) for k in fn._sd_warm)     # so the next forward forks
            f_fork, c_fork = eager(im1, im2)
            f_fork_g, c_fork_g = graphed(im1, im2)        # parallel branches inside the graph
            im1b, im2b = _frames(2, 14)
            f_fork_g2, c_fork_g2 = graphed(im1b, im2b)    # replay with other contents
            f_fork_2, c_fork_2 = eager(im1b, im2b)
        torch.cuda.synchronize()
    finally:
        fn.overlap_sd = prev
        graphed.reset()
    assert torch.isfinite(f_fork).all()
    assert torch.equal(f_serial, f_fork) and torch.equal(c_serial, c_fork)
    assert torch.equal(f_serial_g, f_fork_g) and torch.equal(c_serial_g, c_fork_g)
    assert torch.equal(f_fork_g2, f_fork_2) and torch.equal(c_fork_g2, c_fork_2)
