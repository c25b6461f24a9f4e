"""GPU: the vid2vid training iteration (BASELINE configs[4]) with the flow hot path on libflowops against the same
iteration with the pure-PyTorch / ATen flow path, same weights and frames."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_training_iteration_matches_aten_flow_path(flowops_lib):
    from ir2rgb_b200.models.flownet import FlowNet
    from ir2rgb_b200.train.vid2vid_step import Vid2VidStep
    from oracle import torch_ref as tr
    dev = torch.device("cuda", 0)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32 = False, True, False
    try:
        torch.manual_seed(0)
        flow_net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
        cache = {}

        def frozen_flow(a, b):            # both arms get the SAME reference flow: the comparison is about the warps
            key = (a.data_ptr(), tuple(a.shape))
            if key not in cache:
                cache[key] = flow_net(a, b)
            return cache[key]
        torch.manual_seed(1)
        new = Vid2VidStep(frozen_flow, dev, ngf=8, ndf=8, n_blocks=2)                               # libflowops resample
        old = Vid2VidStep(frozen_flow, dev, ngf=8, ndf=8, n_blocks=2, resample=tr.networks_resample)  # ATen chain
        old.netG.load_state_dict(new.netG.state_dict())
        old.netD.load_state_dict(new.netD.state_dict())
        for a, b in zip(old.netD_T, new.netD_T):
            a.load_state_dict(b.state_dict())
        torch.manual_seed(2)
        frames = [(2 * torch.rand(1, 3, 3, 64, 128, device=dev) - 1, 2 * torch.rand(1, 3, 3, 64, 128, device=dev) - 1) for _ in range(3)]
        for A, B in frames:
            out_new, out_old = new.step(A, B), old.step(A, B)
            for k in ("G", "D", "F_Flow", "F_Warp", "G_Warp"):
                assert torch.allclose(out_new[k], out_old[k], rtol=2e-4, atol=1e-5), (k, out_new[k].item(), out_old[k].item())
        assert out_new["temporal_scales_active"] == 1
        # after three optimizer steps the generators are still the same network.  Adam's first steps move every weight by
        # ~lr * sign(grad), so a gradient that is zero up to rounding may step either way: bound the mean, not the max
        diff = torch.cat([(p - q).abs().reshape(-1) for p, q in zip(new.netG.parameters(), old.netG.parameters())])
        assert diff.mean().item() <= 2e-5 and diff.max().item() <= 3 * 2 * 2e-4 + 1e-6, (diff.mean().item(), diff.max().item())
        assert set(new.timer_ms()) >= {"flownet", "flow_losses", "generator_fwd", "generator_bwd"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32 = prev


def test_generator_warp_gradients_reach_flow_and_previous_frame(flowops_lib):
    from ir2rgb_b200.train import vid2vid_nets as N
    torch.manual_seed(3)
    g = N.CompositeGenerator(9, 3, 6, 8, 2, 2).cuda()
    labels = torch.randn(1, 9, 32, 64, device="cuda")
    prev = torch.randn(1, 6, 32, 64, device="cuda", requires_grad=True)
    final, flow, weight, raw = g(labels, prev)
    assert final.shape == raw.shape == (1, 3, 32, 64) and flow.shape == (1, 2, 32, 64) and weight.shape == (1, 1, 32, 64)
    final.sum().backward()
    assert prev.grad is not None and prev.grad[:, 3:].abs().sum() > 0          # through the warp of the last previous frame
    assert g.model_final_flow[1].weight.grad.abs().sum() > 0                   # through the warp into the flow head
