"""Needs at least two GPUs in one box (`gpurun --gpus 2`); skipped otherwise.

SURVEY.md section 4 / 8e: frame pairs are independent, the batch is split into contiguous chunks, and the sharded
forward must equal the single-GPU forward BIT FOR BIT for every operator.  Also the single-process multi-GPU use the
reference itself makes of these operators (generator.py:113-180 places frames on several devices from one process):
the per-device kernel attributes (opt-in shared memory) must be set on every device, not once per process.
"""
import pytest
import torch

from ir2rgb_b200.sharding import shard_bounds

pytestmark = pytest.mark.gpu

FLOWNETC = (20, 1, 20, 1, 2)


@pytest.fixture(scope="module")
def two_gpus(flowops_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    return torch.device("cuda", 0), torch.device("cuda", 1)


def _ops():
    from ir2rgb_b200.models import networks
    from ir2rgb_b200.models.flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    from ir2rgb_b200.models.flownet2_pytorch.networks.correlation_package.correlation import Correlation
    from ir2rgb_b200.models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d
    return Correlation(*FLOWNETC, 1), Resample2d(), ChannelNorm(), networks.resample


def test_correlation_on_second_device_after_first(two_gpus):
    d0, d1 = two_gpus
    corr = _ops()[0]
    torch.manual_seed(0)
    a, b = torch.randn(2, 64, 16, 40), torch.randn(2, 64, 16, 40)
    outs, grads = [], []
    for dev in (d0, d1, d0):
        at, bt = a.to(dev).requires_grad_(), b.to(dev).requires_grad_()
        out = corr(at, bt)                              # > 48 KB of dynamic shared memory: a per-device attribute
        out.backward(torch.ones_like(out))
        torch.cuda.synchronize(dev)
        outs.append(out.detach().cpu())
        grads.append((at.grad.cpu(), bt.grad.cpu()))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


def test_sharded_forward_equals_single_gpu_forward_bit_for_bit(two_gpus):
    devs = two_gpus
    corr, resample2d, cnorm, grid_resample = _ops()
    torch.manual_seed(1)
    B = 6
    fa, fb = torch.randn(B, 256, 32, 64), torch.randn(B, 256, 32, 64)
    img = 2 * torch.rand(B, 3, 256, 512) - 1
    flow = 4 * torch.randn(B, 2, 256, 512)

    def run(dev, sl):
        with torch.no_grad():
            a, b = fa[sl].to(dev), fb[sl].to(dev)
            i, f = img[sl].to(dev), flow[sl].to(dev)
            warped = resample2d(i, f)
            res = (corr(a, b), warped, cnorm(i - warped), grid_resample(i, f))
            torch.cuda.synchronize(dev)
            return [r.cpu() for r in res]

    full = run(devs[0], slice(0, B))
    parts = [run(devs[r], slice(*shard_bounds(B, r, 2))) for r in range(2)]
    for k, name in enumerate(["correlation", "resample2d", "channelnorm", "networks.resample"]):
        assert torch.equal(torch.cat([p[k] for p in parts]), full[k]), name


def test_sharded_flownet_wrapper_matches_single_gpu(two_gpus):
    """The whole FlowNet wrapper on a 2-way split: cuDNN picks its algorithms per batch size, so this one is held to
    the network tolerance instead of bit-identity."""
    from ir2rgb_b200.models.flownet import FlowNet
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = False, False
    try:
        torch.manual_seed(2)
        nets = []
        sd = None
        for dev in two_gpus:
            net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[dev.index], checkpoints_dir=".", name="t").eval()
            if sd is None:
                sd = {k: v.cpu() for k, v in net.flowNet.state_dict().items()}
            else:
                net.flowNet.load_state_dict(sd)
            net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
            nets.append(net)
        im1 = 2 * torch.rand(4, 3, 128, 256) - 1
        im2 = (im1.roll(2, 3) + 0.05 * torch.randn_like(im1)).clamp(-1, 1)
        flow_full, conf_full = nets[0](im1.to(two_gpus[0]), im2.to(two_gpus[0]))
        parts = []
        for r, dev in enumerate(two_gpus):
            s, e = shard_bounds(4, r, 2)
            with torch.cuda.device(dev):
                f, c = nets[r](im1[s:e].to(dev), im2[s:e].to(dev))
            parts.append((f.cpu(), c.cpu()))
        flow_sh = torch.cat([p[0] for p in parts])
        conf_sh = torch.cat([p[1] for p in parts])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
    rel = ((flow_sh.double() - flow_full.cpu().double()).abs().max() / flow_full.abs().max().double()).item()
    assert rel <= 1e-3, rel
    assert (conf_sh != conf_full.cpu()).float().mean().item() <= 0.01
