"""GPU integration tests: FlowNet2 / FlowNet wrapper with the new operators against the same network with
(a) the reference's rebuilt CUDA extensions and (b) the pure-PyTorch oracle, same random weights."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def nets(flowops_lib):
    from ir2rgb_b200.models.flownet import FlowNet
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False          # remove conv-algorithm noise from the comparison
    new = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
    yield new
    torch.backends.cudnn.allow_tf32 = True


def test_state_dict_matches_reference_architecture(nets):
    arch = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "flownet2_arch.json")))
    sd = nets.flowNet.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == arch["state_dict"]
    assert sum(v.numel() for v in sd.values()) == arch["n_params"] == 162518834


@pytest.mark.parametrize("kind", ["ref", "torch"])
def test_flownet_matches_oracle_network(nets, kind):
    from oracle import ref_ext
    from oracle.harness import OracleFlowNet
    if kind == "ref" and not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    torch.manual_seed(3)
    im1 = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
    im2 = (im1 + 0.05 * torch.randn_like(im1)).clamp(-1, 1)
    other = OracleFlowNet(kind, "cuda", state_dict=nets.flowNet.state_dict())
    flow_new, conf_new = nets(im1, im2)
    flow_ref, conf_ref = other(im1, im2)
    assert flow_new.shape == (2, 2, 128, 192) and conf_new.shape == (2, 1, 128, 192)
    # operators agree to ~1e-7; a 100+-layer conv stack amplifies that somewhat
    assert maxrel(flow_new, flow_ref) <= 1e-3
    assert (conf_new != conf_ref).float().mean().item() <= 0.01      # hard threshold: compare by flip fraction
    assert set(conf_new.unique().tolist()) <= {0.0, 1.0}


def test_oracle_network_never_calls_libflowops(nets):
    """The comparison network (and bench.py's reference arm) must run the reference's operators and stock torch layers
    only: not one libflowops call may happen on that path."""
    from ir2rgb_b200 import _lib
    from oracle import ref_ext
    from oracle.harness import OracleFlowNet
    calls = []
    prev = _lib.launch_hook
    _lib.launch_hook = lambda what, n: calls.append(what)
    try:
        im = 2 * torch.rand(1, 3, 64, 64, device="cuda") - 1
        for kind in ["torch"] + (["ref"] if ref_ext.available() else []):
            OracleFlowNet(kind, "cuda", state_dict=nets.flowNet.state_dict())(im, im.flip(3))
        assert calls == [], calls
        nets(im, im.flip(3))
        assert len(calls) > 0          # the hook does see the product path
    finally:
        _lib.launch_hook = prev


def test_fused_glue_is_bit_identical_to_operator_chain(nets, c_oracle):
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    from ir2rgb_b200.models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d
    torch.manual_seed(4)
    x = 2 * torch.rand(2, 6, 64, 96, device="cuda") - 1
    flow = 6 * torch.randn(2, 2, 64, 96, device="cuda")
    warped, norm = F.warp_diff_norm_forward(x, flow)
    w_ref = Resample2d()(x[:, 3:], flow)
    n_ref = ChannelNorm()((x[:, :3] - w_ref).contiguous())
    assert torch.equal(warped, w_ref) and torch.equal(norm, n_ref)
    # ... and directly against the C restatement of the reference kernels (resample2d_kernel.cu:16-64, channelnorm_kernel.cu:19-60)
    xn, fn = x.cpu().numpy(), flow.cpu().numpy()
    w_o = c_oracle.resample2d_fwd(np.ascontiguousarray(xn[:, 3:]), fn)
    n_o = c_oracle.cnorm_fwd(np.ascontiguousarray(xn[:, :3] - w_o))
    assert np.array_equal(warped.cpu().numpy(), w_o) and np.array_equal(norm.cpu().numpy(), n_o)
    # writing straight into a concat buffer
    buf = torch.zeros(2, 12, 64, 96, device="cuda")
    F.warp_diff_norm_forward(x, flow, out=(buf, 6, 11))
    assert torch.equal(buf[:, 6:9], w_ref) and torch.equal(buf[:, 11:12], n_ref) and buf[:, :6].abs().sum() == 0
    _, norm_only = F.warp_diff_norm_forward(x, flow, need_warped=False)
    assert torch.equal(norm_only, n_ref)
    # confidence mask
    conf = F.warp_conf_forward(x[:, :3].contiguous(), x[:, 3:].contiguous(), flow, 0.02, F.WARP_RESAMPLE2D)
    t = x[:, :3] - w_ref
    conf_ref = (torch.sum(t * t, dim=1, keepdim=True) < 0.02).float()
    assert (conf != conf_ref).float().mean().item() <= 1e-4
    # as-run mode of the reference's FlowNet: the grid_sample warp (Model.resample shadows the Resample2d submodule)
    from ir2rgb_b200.models import networks
    conf_gs = F.warp_conf_forward(x[:, :3].contiguous(), x[:, 3:].contiguous(), flow, 0.02, F.WARP_GRIDSAMPLE)
    t = x[:, :3] - networks.resample(x[:, 3:].contiguous(), flow)
    assert (conf_gs != (torch.sum(t * t, dim=1, keepdim=True) < 0.02).float()).float().mean().item() <= 1e-4


def test_input_prep_and_concat_assembly_are_bit_identical(nets):
    """flowops_flownet2_prep and flowops_warp_diff_norm_concat_nhwc against the torch expressions of models.py:97-114."""
    from ir2rgb_b200 import functional as F
    torch.manual_seed(8)
    B, H, W = 2, 64, 96
    inputs = 2 * torch.rand(B, 3, 2, H, W, device="cuda") - 1
    rgb_mean = inputs.contiguous().view(B, 3, -1).mean(dim=-1)
    x_ref = (inputs - rgb_mean.view(B, 3, 1, 1, 1)) / 1.0
    x_ref = torch.cat((x_ref[:, :, 0], x_ref[:, :, 1]), dim=1)
    for rgb_max in (1.0, 255.0):
        x, xa, xb, x8 = F.flownet2_prep(inputs, rgb_mean, rgb_max)
        want = torch.cat((((inputs - rgb_mean.view(B, 3, 1, 1, 1)) / rgb_max)[:, :, 0],
                          ((inputs - rgb_mean.view(B, 3, 1, 1, 1)) / rgb_max)[:, :, 1]), dim=1)
        assert torch.equal(x, want)
        assert torch.equal(xa[:, :3], want[:, :3]) and torch.equal(xb[:, :3], want[:, 3:]) and torch.equal(x8[:, :6], want)
        assert (xa[:, 3:] == 0).all() and (xb[:, 3:] == 0).all() and (x8[:, 6:] == 0).all()
        assert xa.is_contiguous(memory_format=torch.channels_last) and x8.is_contiguous(memory_format=torch.channels_last)
    x = x_ref.contiguous()
    flow = 6 * torch.randn(B, 2, H, W, device="cuda")
    cat = F.warp_diff_norm_concat(x, flow, 20.0)
    warped, norm = F.warp_diff_norm_forward(x, flow)
    want = torch.cat((x, warped, flow / 20.0, norm), dim=1)
    assert cat.shape == (B, 16, H, W) and cat.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(cat[:, :12], want) and (cat[:, 12:] == 0).all()


@pytest.mark.parametrize("s2d", [False, True])
def test_input_prep_at_a_16_channel_pitch(nets, s2d):
    """flowops_flownet2_prep_pitched: the both-frames tensor FlowNetSD.conv0 reads, as the first 8 channels of a cached,
    pre-zeroed 16-channel channels_last tensor (the rest is never written) -- same values as the dense 8-channel form,
    also on the second call, which reuses the tensor."""
    from ir2rgb_b200 import functional as F
    prep = F.flownet2_prep_s2d if s2d else F.flownet2_prep
    B, H, W = 2, 64, 96
    ptr = None
    for seed in (1, 2):
        torch.manual_seed(seed)
        inputs = 2 * torch.rand(B, 3, 2, H, W, device="cuda") - 1
        mean = inputs.view(B, 3, -1).mean(dim=-1)
        x, xa, xb, x8 = prep(inputs, mean, 255.0, 8)
        x2, xa2, xb2, x16 = prep(inputs, mean, 255.0, 16)
        assert x16.shape == (B, 16, H, W) and x16.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(x, x2) and torch.equal(xa, xa2) and torch.equal(xb, xb2)
        assert torch.equal(x16[:, :8], x8) and (x16[:, 8:] == 0).all()
        assert ptr in (None, x16.data_ptr())              # one tensor per shape and device
        ptr = x16.data_ptr()
    with pytest.raises(Exception):
        prep(inputs, mean, 255.0, 6)                      # a multiple of 4, >= 8


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 128, 320), (3, 8, 12)])
def test_fusion_input_is_bit_identical_to_operator_chain(nets, shape):
    """flowops_flownet2_fusion_input_nhwc against models.py:129-152 spelled out with the separate operators."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    torch.manual_seed(18)
    B, H, W = shape
    x = (2 * torch.rand(B, 6, H, W, device="cuda") - 1).contiguous()
    lo_s2 = 0.3 * torch.randn(B, 2, H // 4, W // 4, device="cuda")          # network units: * div_flow -> pixels
    lo_sd = 60.0 * torch.randn(B, 2, H // 4, W // 4, device="cuda")         # / div_flow -> pixels
    lo_s2[0, :, 0, 0] = 0.0
    lo_sd[0, 0, 0, 1] = 1e4                                                 # far outside the frame: clamped corners
    up = torch.nn.Upsample(scale_factor=4, mode='nearest')
    flow_s2, flow_sd = up(lo_s2 * 20.0), up(lo_sd / 20.0)
    _, diff_s2 = F.warp_diff_norm_forward(x, flow_s2)
    _, diff_sd = F.warp_diff_norm_forward(x, flow_sd)
    want = torch.cat((x[:, :3], flow_sd, flow_s2, ChannelNorm()(flow_sd), ChannelNorm()(flow_s2), diff_sd, diff_s2), dim=1)
    got = F.flownet2_fusion_input(x, lo_s2, lo_sd, 20.0)
    assert got.shape == (B, 16, H, W) and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got[:, :11], want) and (got[:, 11:] == 0).all()
    got12 = F.flownet2_fusion_input(x, lo_s2, lo_sd, 20.0, c_pad=12)
    assert torch.equal(got12[:, :11], want) and (got12[:, 11:] == 0).all()
    with pytest.raises(ValueError):
        F.flownet2_fusion_input(x, lo_s2[:, :, :-1], lo_sd, 20.0)


def test_fused_fusion_input_does_not_change_the_network_output(flowops_lib):
    """channels_last FlowNet2 with and without the one-pass concat3: same flow up to cuDNN's summation order on the
    padded channel count of the fusion network's first convolution."""
    from ir2rgb_b200.models.flownet import FlowNet
    torch.manual_seed(19)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, True, False
    try:
        net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
        net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
        a = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
        b = (a + 0.05 * torch.randn_like(a)).clamp(-1, 1)
        from ir2rgb_b200 import _lib
        calls = []
        prev_hook = _lib.launch_hook
        _lib.launch_hook = lambda what, n: calls.append(what)
        try:
            flow_fused, conf_fused = net(a, b)
        finally:
            _lib.launch_hook = prev_hook
        assert "flownet2_fusion_input_nhwc" in calls
        net.flowNet.fuse_fusion_input = False
        flow_plain, conf_plain = net(a, b)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = prev
    rel = ((flow_fused - flow_plain).abs().max() / flow_plain.abs().max()).item()
    assert rel <= 1e-4, rel
    assert (conf_fused != conf_plain).float().mean().item() <= 1e-3


def test_flownet_wrapper_shapes_and_resize_path(nets):
    # 5-D input (b, n, c, h, w) and a height that is not a multiple of 64 (flownet.py:27-33,41-47,51-53)
    a = 2 * torch.rand(1, 2, 3, 96, 128, device="cuda") - 1
    b = 2 * torch.rand(1, 2, 3, 96, 128, device="cuda") - 1
    flow, conf = nets(a, b)
    assert flow.shape == (1, 2, 2, 96, 128) and conf.shape == (1, 2, 1, 96, 128)
    assert torch.isfinite(flow).all()


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("kind,cin,cout,k,s", [("conv", 12, 64, 7, 2), ("conv", 64, 66, 3, 1), ("deconv", 34, 16, 4, 2)])
def test_fused_bias_lrelu_is_bit_identical(flowops_lib, channels_last, kind, cin, cout, k, s):
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(5)
    mod = (sm.conv(False, cin, cout, kernel_size=k, stride=s) if kind == "conv" else sm.deconv(cin, cout)).cuda()
    x = torch.randn(2, cin, 24, 40, device="cuda")
    if channels_last:
        mod = mod.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    prev, prev_det = torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True       # cuDNN's transposed-conv algorithms may otherwise use atomics
    try:
        with torch.no_grad():
            fused = mod(x)
            sm.FUSE_EPILOGUE = False
            try:
                plain = mod(x)
            finally:
                sm.FUSE_EPILOGUE = True
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic = prev, prev_det
    assert torch.equal(fused, plain)
    # with autograd enabled the stock path runs (and is differentiable)
    y = mod(x.requires_grad_())
    y.sum().backward()
    assert x.grad is not None


def test_graphed_flownet_matches_eager(nets):
    from ir2rgb_b200.runtime import GraphedFlowNet
    torch.manual_seed(6)
    a = 2 * torch.rand(2, 3, 64, 128, device="cuda") - 1
    b = 2 * torch.rand(2, 3, 64, 128, device="cuda") - 1
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        flow_e, conf_e = nets(a, b)
        g = GraphedFlowNet(nets)
        flow_g, conf_g = g(a, b)
        flow_g2, _ = g(a.cpu().pin_memory(), b.cpu().pin_memory())     # pinned host inputs, same graph
        a2 = a.flip(0).contiguous()
        flow_g3, _ = g(a2, b)                                          # new data through the replayed graph
        flow_e3, _ = nets(a2, b)
    finally:
        torch.backends.cudnn.deterministic = prev
    assert torch.equal(flow_g, flow_e) and torch.equal(conf_g, conf_e)
    assert torch.equal(flow_g2, flow_e)
    assert torch.equal(flow_g3, flow_e3)
    g.reset()                                                          # what a caller does after loading new weights
    assert not g._graphs
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        flow_g4, _ = g(a, b)
    finally:
        torch.backends.cudnn.deterministic = prev
    assert torch.equal(flow_g4, flow_e)


def test_host_pipeline_matches_direct_calls(nets):
    from ir2rgb_b200.runtime import HostPipeline
    torch.manual_seed(7)
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        h1 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory()
        h2 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory()
        hflow, hconf = torch.empty(5, 2, 64, 128).pin_memory(), torch.empty(5, 1, 64, 128).pin_memory()
        HostPipeline(nets, torch.device("cuda", 0))(h1, h2, 2, hflow, hconf)       # micro-batches of 2, 2, 1
        torch.cuda.synchronize()
        for s in (0, 2, 4):
            f, c = nets(h1[s:s + 2].cuda(), h2[s:s + 2].cuda())
            assert torch.equal(hflow[s:s + 2], f.cpu()) and torch.equal(hconf[s:s + 2], c.cpu())
        # a stream of batches: the next batch's first copy-in is issued by the previous call, the last copy-out is not
        # waited for; a batch that was NOT announced (g1) must still be computed from its own frames
        g1 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory()
        g2 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory()
        outs = [(torch.empty(5, 2, 64, 128).pin_memory(), torch.empty(5, 1, 64, 128).pin_memory()) for _ in range(3)]
        pipe = HostPipeline(nets, torch.device("cuda", 0))
        pipe(h1, h2, 2, *outs[0], next_inputs=(h1, h2), wait=False)
        pipe(h1, h2, 2, *outs[1], next_inputs=(h1, h2), wait=False)      # prefetched
        pipe(g1, g2, 2, *outs[2], wait=False)                            # announced h1, got g1: prefetch discarded
        pipe.synchronize()
        assert torch.equal(outs[0][0], hflow) and torch.equal(outs[1][0], hflow) and torch.equal(outs[1][1], hconf)
        f, c = nets(g1[:2].cuda(), g2[:2].cuda())
        assert torch.equal(outs[2][0][:2], f.cpu()) and torch.equal(outs[2][1][:2], c.cpu())
    finally:
        torch.backends.cudnn.deterministic = prev


@pytest.mark.parametrize("chans", [(512, 512, 2), (64, 64), (128, 32, 2), (5, 3)])
def test_cat_channels_matches_torch_cat(flowops_lib, chans):
    from ir2rgb_b200 import functional as F
    torch.manual_seed(8)
    parts = [torch.randn(2, c, 9, 14, device="cuda").contiguous(memory_format=torch.channels_last) for c in chans]
    with torch.no_grad():
        out = F.cat_channels(parts)
    ref = torch.cat(parts, 1)
    assert torch.equal(out, ref) and out.is_contiguous(memory_format=torch.channels_last)
    # NCHW inputs and autograd fall back to torch.cat
    nchw = [p.contiguous() for p in parts]
    assert torch.equal(F.cat_channels(nchw), ref)
    req = [p.clone().requires_grad_() for p in parts]
    F.cat_channels(req).sum().backward()
    assert all(r.grad is not None for r in req)


def test_channels_last_flownet_with_fused_conv3_epilogue_matches_plain_path(flowops_lib):
    """channels_last conv body: conv3's bias+LeakyReLU epilogue writes the correlation's planes directly.
    Same network with every fusion switched off must give the same flow, bit for bit."""
    from ir2rgb_b200.models.flownet import FlowNet
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(9)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = False, True, False
    try:
        net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
        net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
        a = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
        b = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
        pad = sm.PAD_CHANNELS
        sm.PAD_CHANNELS = 1         # channel padding changes cuDNN's algorithm choice: tested separately below
        # two round-2 fusions are equal to the layers they replace only to a few ulp (their own tests say how close):
        # the x4 bilinear upsampling folded into the concat kernel and the one-kernel flow upsampler.  Off for the
        # bit-for-bit comparison, on for the tolerance comparison below.
        sm.FUSE_FLOW_UPSAMPLER = False
        net.flowNet.fuse_upsample = False
        try:
            flow_fused, conf_fused = net(a, b)
            sm.FUSE_EPILOGUE = False
            net.flowNet.fuse_glue = False
            net.fuse_conf = False
            try:
                flow_plain, conf_plain = net(a, b)
            finally:
                sm.FUSE_EPILOGUE = True
                net.flowNet.fuse_glue = True
                net.fuse_conf = True
        finally:
            sm.PAD_CHANNELS = pad
            sm.FUSE_FLOW_UPSAMPLER = True
            net.flowNet.fuse_upsample = True
        flow_padded, conf_padded = net(a, b)          # everything on: padded concat buffers, folded upsampling, flow upsampler kernel
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = prev
    assert torch.equal(flow_fused, flow_plain)
    assert (conf_fused != conf_plain).float().mean().item() <= 1e-4
    # zero channels add exact zeros to every sum; only cuDNN's summation order may differ with the channel count
    rel = ((flow_padded - flow_plain).abs().max() / flow_plain.abs().max()).item()
    assert rel <= 1e-4, rel
    assert (conf_padded != conf_plain).float().mean().item() <= 1e-3


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 128, 192), (2, 512, 1024)])
def test_concat_with_folded_bilinear_upsampling(flowops_lib, shape):
    """`upsample_bilinear_x4(flow2 * div_flow)` (models.py:106,118) formed inside the concat kernel: the flow channels
    agree with nn.Upsample to 1e-6 max-relative (the blend may contract its FMAs differently from ATen's kernel: a few
    ulp), the channels computed FROM the flow (warped frame, error magnitude) to the operator tolerance 1e-5."""
    from ir2rgb_b200 import functional as F
    B, H, W = shape
    torch.manual_seed(23)
    x = 2 * torch.rand(B, 6, H, W, device="cuda") - 1
    lo = torch.randn(B, 2, H // 4, W // 4, device="cuda")
    up = torch.nn.Upsample(scale_factor=4, mode="bilinear")
    want = F.warp_diff_norm_concat(x, up(lo * 20.0), 20.0)
    got = F.warp_diff_norm_concat_up4(x, lo, 20.0, 20.0)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got[:, :6], want[:, :6]) and torch.equal(got[:, 12:], want[:, 12:])     # frames, zero padding
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    assert rel(got[:, 9:11], want[:, 9:11]) <= 1e-6          # flow / div_flow
    # a flow that moves by an ulp can move a pixel's floor across an integer: compare the warped frame statistically
    d = (got[:, 6:9] - want[:, 6:9]).abs()
    assert (d > 1e-4).float().mean().item() <= 1e-4 and rel(got[:, 11:12], want[:, 11:12]) <= 2.0
    assert d.median().item() <= 1e-6


def test_flow_upsampler_kernel_matches_conv_transpose(flowops_lib):
    """flowops_flow_deconv_nhwc_to (the decoders' 2-channel ConvTranspose2d(k4, s2, p1)) against torch, written into a slice
    of a padded channels-last concat buffer; with and without bias."""
    from ir2rgb_b200 import functional as F
    torch.manual_seed(29)
    for (B, h, w, bias) in [(2, 5, 7, True), (1, 16, 32, False), (3, 1, 1, True), (2, 64, 128, True)]:
        conv = torch.nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=bias).cuda()
        flow = torch.randn(B, 2, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
        like = torch.empty(B, 1, 2 * h, 2 * w, device="cuda")
        buf = F.ConcatBuffer(like, 10, 8)
        buf.tensor.fill_(7.0)
        with torch.no_grad():
            want = conv(flow)
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            try:
                want = conv(flow)
            finally:
                torch.backends.cudnn.allow_tf32 = prev
            buf.flow_deconv_in(flow, conv.weight.detach().contiguous(), conv.bias, 6)
        got = buf.tensor[:, 6:8]
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-6
        assert (buf.tensor[:, :6] == 7.0).all() and (buf.tensor[:, 8:] == 7.0).all()      # neighbours untouched


def test_flownetc_first_layer_on_space_to_depth_frames(flowops_lib):
    """FlowNetC.conv1 (3 -> 64, 7x7, stride 2) as a 4x4 stride-1 convolution over the space-to-depth frame written by
    flowops_flownet2_prep_s2d: the same sums in another order (fp32 convolutions: <= 1e-5)."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.models import FlowNet2
    torch.manual_seed(33)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = False, False
    try:
        net = FlowNet2().cuda().eval().to(memory_format=torch.channels_last)
        inputs = 2 * torch.rand(2, 3, 2, 64, 96, device="cuda") - 1
        mean = inputs.contiguous().view(2, 3, -1).mean(dim=-1)
        with torch.no_grad():
            x, xa, xb, x8 = F.flownet2_prep(inputs, mean, 1.0)
            x2, sa, sb, x82 = F.flownet2_prep_s2d(inputs, mean, 1.0)
            assert torch.equal(x, x2) and torch.equal(x8, x82)
            assert sa.shape == (2, 16, 33, 49) and (sa[:, :, 0] == 0).all() and (sa[:, :, :, 0] == 0).all()
            for frame, s2d in ((xa, sa), (xb, sb)):
                want = net.flownetc.conv1(frame)
                got = net.flownetc.conv1_s2d()(s2d)
                assert got.shape == want.shape
                assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5
            # a second call reuses the cached s2d tensors (border still zero, interior rewritten)
            inputs2 = inputs.flip(0).contiguous()
            _, sa2, _, _ = F.flownet2_prep_s2d(inputs2, inputs2.view(2, 3, -1).mean(dim=-1), 1.0)
            assert sa2.data_ptr() == sa.data_ptr() and (sa2[:, :, 0] == 0).all()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev


@pytest.mark.parametrize("cin,cout,h,w", [(168, 16, 32, 48), (24, 8, 9, 7), (40, 16, 1, 1)])
def test_narrow_deconv_as_conv3_with_depth_to_space_epilogue(flowops_lib, cin, cout, h, w):
    """ConvTranspose2d(k4, s2, p1) + bias + LeakyReLU computed as a 3x3 convolution with 4*C output channels and the
    depth-to-space epilogue kernel (flowops_bias_lrelu_d2s_nhwc_to), against torch; written into a concat slice."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(37)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        conv = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1).cuda().to(memory_format=torch.channels_last)
        x = torch.randn(2, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
        buf = F.ConcatBuffer(x, cout + 12, 8, shape=(2, 2 * h, 2 * w))
        buf.tensor.fill_(5.0)
        with torch.no_grad():
            want = torch.nn.functional.leaky_relu(conv(x), 0.1)
            w3 = sm.deconv_as_conv3_weight(conv, conv.weight)
            buf.bias_lrelu_d2s_in(torch.nn.functional.conv2d(x, w3, None, 1, 1), conv.bias, 0.1, 8)
        got = buf.tensor[:, 8:8 + cout]
        assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5
        assert (buf.tensor[:, :8] == 5.0).all() and (buf.tensor[:, 8 + cout:] == 5.0).all()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("cin,cout,h,w", [(168, 16, 32, 48), (24, 8, 9, 7), (40, 32, 1, 1)])
def test_depth_to_space_epilogue_with_the_flow_upsampler_folded_in(flowops_lib, cin, cout, h, w):
    """flowops_bias_lrelu_d2s_flowup_nhwc_to: the deconvolution slice equals the plain depth-to-space epilogue and the two flow
    channels behind it equal flowops_flow_deconv_nhwc_to, bit for bit; the rest of the concat buffer is untouched."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(41)
    conv = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1).cuda().to(memory_format=torch.channels_last)
    up = torch.nn.ConvTranspose2d(2, 2, 4, 2, 1).cuda()
    x = torch.randn(2, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    flow = (3 * torch.randn(2, 2, h, w, device="cuda")).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y4 = torch.nn.functional.conv2d(x, sm.deconv_as_conv3_weight(conv, conv.weight), None, 1, 1)
        fw = up.weight.detach().contiguous()
        sep = F.ConcatBuffer(x, cout + 14, 8, shape=(2, 2 * h, 2 * w))
        sep.tensor.fill_(5.0)
        sep.bias_lrelu_d2s_in(y4, conv.bias, 0.1, 8)
        sep.flow_deconv_in(flow, fw, up.bias, 8 + cout)
        one = F.ConcatBuffer(x, cout + 14, 8, shape=(2, 2 * h, 2 * w))
        one.tensor.fill_(5.0)
        assert one.bias_lrelu_d2s_in(y4, conv.bias, 0.1, 8, (flow, fw, up.bias)) == 8 + cout + 2
        want_up = up(flow)
    assert torch.equal(one.tensor, sep.tensor)
    assert ((one.tensor[:, 8 + cout:10 + cout] - want_up).abs().max() / want_up.abs().max()).item() <= 1e-6
    assert (one.tensor[:, :8] == 5.0).all() and (one.tensor[:, 10 + cout:] == 5.0).all()


@pytest.mark.parametrize("cout,tail", [(16, 6), (20, 2), (8, 6)])
def test_flow_slice_that_ends_the_record_rewrites_the_pad_channels(flowops_lib, cout, tail):
    """A 2-channel flow slice that ends a concat buffer's real channels is written together with the buffer's zero pad
    channels (tail_zero of flowops_flow_deconv_nhwc_to / flowops_bias_lrelu_d2s_flowup_nhwc_to: whole 16-byte stores, whole
    sectors): every real channel is bit-identical to the 8-byte-store form, the pad channels are (still) zero, and the
    channels in front of the slice are untouched."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(43)
    h, w, cin = 12, 20, 24
    conv = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1).cuda().to(memory_format=torch.channels_last)
    up = torch.nn.ConvTranspose2d(2, 2, 4, 2, 1).cuda()
    x = torch.randn(2, cin, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    flow = (3 * torch.randn(2, 2, h, w, device="cuda")).contiguous(memory_format=torch.channels_last)
    c_total = 8 + cout + 2
    results = {}
    prev = F.D2S_WRITE_PAD
    try:
        with torch.no_grad():
            y4 = torch.nn.functional.conv2d(x, sm.deconv_as_conv3_weight(conv, conv.weight), None, 1, 1)
            fw = up.weight.detach().contiguous()
            for mode in (False, True):
                F.D2S_WRITE_PAD = mode
                one = F.ConcatBuffer(x, c_total, 8, shape=(2, 2 * h, 2 * w))
                assert one.c_pad - one.c_total == tail and one._pad_tail(c_total, 8 + cout) == (tail if mode else 0)
                one.tensor[:, :8] = 5.0
                one.bias_lrelu_d2s_in(y4, conv.bias, 0.1, 8, (flow, fw, up.bias))
                two = F.ConcatBuffer(x, c_total, 8, shape=(2, 2 * h, 2 * w))
                two.tensor[:, :8] = 5.0
                two.bias_lrelu_d2s_in(y4, conv.bias, 0.1, 8)
                two.flow_deconv_in(flow, fw, up.bias, 8 + cout)
                results[mode] = (one.tensor.clone(), two.tensor.clone())
    finally:
        F.D2S_WRITE_PAD = prev
    ref = results[False][0]
    assert (ref[:, :8] == 5.0).all() and (ref[:, c_total:] == 0).all()
    for t in (results[False][1], results[True][0], results[True][1]):
        assert torch.equal(t, ref)


@pytest.mark.parametrize("B,c_real,c_pad,H,W,bias", [(2, 16, 16, 33, 70, True), (1, 194, 200, 17, 45, True), (3, 22, 24, 5, 3, False),
                                                     (2, 1026, 1032, 4, 9, True), (1, 32, 32, 64, 129, True), (2, 8, 8, 2, 3, True), (1, 16, 16, 70, 40, True)])
def test_flow_head_kernel_matches_fp32_convolution(flowops_lib, B, c_real, c_pad, H, W, bias):
    """flowops_flow_head_nhwc (predict_flow: nn.Conv2d(C, 2, 3, 1, 1), networks/submodules.py:40-41) against cuDNN's fp32
    convolution: FP32 sums in another order (<= 1e-5 of the largest output); zero pad channels, ragged widths, one-pixel
    frames, a channel-slice view as the input, with and without bias."""
    from ir2rgb_b200 import functional as F
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    torch.manual_seed(c_real + W)
    conv = torch.nn.Conv2d(c_real, 2, 3, 1, 1, bias=bias).cuda()
    x = torch.zeros(B, c_pad, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    x[:, :c_real] = torch.randn(B, c_real, H, W, device="cuda")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            want = conv(x[:, :c_real])
            got = F.flow_head(x, F.pack_flow_head_weight(conv.weight, c_pad), conv.bias)
            assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
            assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-5
            # NaN / Inf in a pad channel must not leak (the kernel never reads beyond the layer's channel quads ...)
            if c_pad - c_real >= 4:
                x2 = x.clone()
                x2[:, -4:] = float("nan")
                got2 = F.flow_head(x2[:, :c_pad - 4], F.pack_flow_head_weight(conv.weight, c_pad - 4), conv.bias)    # a channel-slice view
                assert torch.equal(got2, got) or ((got2 - want).abs().max() / want.abs().max()).item() <= 1e-5
            # ... and through the layer dispatch of the conv body (TF32 convolutions allowed: the kernel runs)
            torch.backends.cudnn.allow_tf32 = True
            assert sm._flow_head_ok(conv, x) == (c_pad <= sm.FLOW_HEAD_MAX_CHANNELS)
            if bias and sm._flow_head_ok(conv, x):
                assert torch.equal(sm.apply_conv(conv, x), got)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
