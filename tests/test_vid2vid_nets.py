"""The vid2vid training-step harness (BASELINE configs[4]): restated generator / discriminators and the loss plumbing
against the reference's own classes (imported from /root/reference when it is there -- the build container -- and
against the committed architecture fixture everywhere)."""
import json
import os
import sys

import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def test_architecture_matches_reference_fixture():
    from ir2rgb_b200.train import vid2vid_nets as N
    arch = json.load(open(os.path.join(HERE, "golden", "vid2vid_arch.json")))
    with torch.device("meta"):
        nets = {"G": N.CompositeGenerator(9, 3, 6, 128, 3, 9), "D": N.MultiScaleDiscriminator(6, 64, 3, "batch", 2),
                "D_T": N.MultiScaleDiscriminator(13, 64, 3, "batch", 2)}
    for name, net in nets.items():
        assert {k: list(v.shape) for k, v in net.state_dict().items()} == arch[name], name
        assert sum(p.numel() for p in net.parameters()) == arch["n_params"][name]
    assert arch["n_params"]["G"] == 364770438


@pytest.fixture(scope="module")
def ref_modules():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box): covered by the architecture fixture")
    sys.path.insert(0, REF)
    try:
        from models import discriminator, networks
    finally:
        sys.path.remove(REF)
    return networks, discriminator


def test_generator_and_discriminator_forward_match_reference(ref_modules):
    ref, _ = ref_modules
    from ir2rgb_b200.train import vid2vid_nets as N
    torch.manual_seed(0)
    kw = dict(gen_blocks=4, n_local_enhancers=1, feat_num=3, n_blocks_local=3, fg=False, no_flow=False)
    g_ref = ref.build_generator_module(9, 3, 6, 8, "composite", 2, "batch", 0, **kw)
    g_new = N.CompositeGenerator(9, 3, 6, 8, 2, 4)
    assert list(g_ref.state_dict().keys()) == list(g_new.state_dict().keys())
    g_new.load_state_dict(g_ref.state_dict())
    labels, prev = torch.randn(1, 9, 32, 48), torch.randn(1, 6, 32, 48)
    # use_raw_only: every conv stack of the generator without the warp (which is CUDA-only on both sides)
    want = g_ref(labels, prev, None, None, None, None, True)
    got = g_new(labels, prev, use_raw_only=True)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]) and torch.equal(got[3], want[3])

    d_ref = ref.build_discriminator_module(6, 64, 3, "batch", 2, True)
    d_new = N.MultiScaleDiscriminator(6, 64, 3, "batch", 2)
    d_new.load_state_dict(d_ref.state_dict())
    x = torch.randn(2, 6, 64, 96)
    for a, b in zip(d_ref(x), d_new(x)):
        assert len(a) == len(b) == 5 and all(torch.equal(p, q) for p, q in zip(a, b))


def test_loss_plumbing_matches_reference_discriminator_model(ref_modules):
    """Vid2VidModelD.forward (discriminator.py:90-151) and the temporal path (:104-118, 169-184) against Vid2VidStep's
    loss code on the same weights and tensors, CPU, with the pure-torch warp on both sides."""
    _, RD = ref_modules
    from ir2rgb_b200.train.vid2vid_step import Vid2VidStep, skipped_frames
    from ir2rgb_b200.train import vid2vid_nets as N
    from oracle import torch_ref as tr
    torch.manual_seed(1)
    opt = dict(gpu_ids=[0], gen_gpus=1, batch_size=1, debug=True, n_frames_D=3, output_nc=3, label_nc=0, input_nc=3,
               use_instance=False, first_layer_dis_filters=8, n_layers_D=3, norm="batch", num_D=2, no_ganFeat=False,
               n_scales_temporal=2, continue_train=False, load_pretrained=False, gan_mode="ls", no_vgg=True, lr=2e-4, TTUR=False,
               beta1=0.5, lambda_feat=10.0, lambda_F=10.0, lambda_T=10.0, n_scales_spatial=1, no_first_img=False,
               checkpoints_dir=".", name="t", fp16=False)
    ref_d = RD.Vid2VidModelD(**opt)
    ref_d.resample = lambda image, flow: tr.networks_resample(image, flow)       # the reference's own method calls .cuda()
    ref_d.compute_loss_D.__func__.__globals__["print"] = lambda *a, **k: None     # discriminator.py:154 prints shapes
    st = Vid2VidStep(None, torch.device("cpu"), ngf=8, ndf=8, n_blocks=2, resample=tr.networks_resample)
    st.netD.load_state_dict(ref_d.netD.state_dict())
    for s in range(2):
        st.netD_T[s].load_state_dict(getattr(ref_d, "netD_T%d" % s).state_dict())
    h, w = 32, 48
    rB, fB, fBraw, rA, rBp, fBp = (torch.randn(1, 3, h, w) for _ in range(6))
    flow, flow_ref = 3 * torch.randn(1, 2, h, w), 3 * torch.randn(1, 2, h, w)
    weight, conf = torch.rand(1, 1, h, w), (torch.rand(1, 1, h, w) > 0.4).float()
    names = ref_d.loss_names
    ref_losses = dict(zip(names, [x.mean() for x in ref_d(0, [rB, fB, fBraw, rA, rBp, fBp, flow, weight, flow_ref, conf])]))
    # the same quantities from the step's pieces
    d_real, d_fake, g_gan, g_fm = st._loss_D(st.netD, torch.cat((rA, rB), 1), torch.cat((rA, fB), 1))
    r2, f2, gg2, gf2 = st._loss_D(st.netD, torch.cat((rA, rB), 1), torch.cat((rA, fBraw), 1))
    mine = {"G_GAN": g_gan + gg2, "G_GAN_Feat": g_fm + gf2, "D_real": d_real + r2, "D_fake": d_fake + f2,
            "F_Flow": N.masked_l1(flow, flow_ref, conf) * 10.0,
            "F_Warp": N.masked_l1(tr.networks_resample(rBp, flow), rB, conf) * 10.0,
            "G_Warp": N.masked_l1(fB, tr.networks_resample(fBp, flow_ref), conf) * 10.0}
    for k, v in mine.items():
        assert torch.allclose(v, ref_losses[k], rtol=1e-6, atol=1e-7), (k, v.item(), ref_losses[k].item())
    assert ref_losses["G_VGG"].abs().max() == 0 and ref_losses["W"].abs().max() == 0
    # temporal scale 0: three consecutive frames plus the two reference flows between them
    real5, fake5 = torch.randn(1, 3, 3, h, w), torch.randn(1, 3, 3, h, w)
    flow5, conf5 = 3 * torch.randn(1, 2, 2, h, w), torch.rand(1, 2, 1, h, w)
    tl = dict(zip(ref_d.loss_names_T, [x.mean() for x in ref_d(1, [real5, fake5, flow5, conf5])]))
    rb, fb = real5.view(-1, 9, h, w), fake5.view(-1, 9, h, w)
    fr = (flow5 / 20).view(-1, 4, h, w)
    t_real, t_fake, t_gan, t_fm = st._loss_D(st.netD_T[0], torch.cat([rb, fr], 1), torch.cat([fb, fr], 1))
    for k, v in {"G_T_GAN": t_gan, "G_T_GAN_Feat": t_fm, "D_T_real": t_real, "D_T_fake": t_fake}.items():
        assert torch.allclose(v, tl[k], rtol=1e-6, atol=1e-7), (k, v.item(), tl[k].item())
    # frame bookkeeping of the temporal scales
    all_r = all_n = None
    for _ in range(9):
        B = torch.randn(1, 1, 3, 4, 4)
        all_r, sk_r = RD.get_skipped_frames(all_r, B, 2, 3)
        all_n, sk_n = skipped_frames(all_n, B, 2, 3)
        assert torch.equal(all_r, all_n)
        assert all((a is None and b is None) or torch.equal(a, b) for a, b in zip(sk_r, sk_n))


def test_step_runs_a_sequence_on_cpu_with_stand_in_flow():
    """Eight iterations of one sequence: temporal scale 0 switches on at the 3rd frame, scale 1 at the 7th."""
    from ir2rgb_b200.train.vid2vid_step import Vid2VidStep
    from oracle import torch_ref as tr
    torch.manual_seed(0)

    def flow_net(a, b):
        n, t, _, h, w = a.shape
        return 0.5 * torch.randn(n, t, 2, h, w), (torch.rand(n, t, 1, h, w) > 0.3).float()
    st = Vid2VidStep(flow_net, torch.device("cpu"), ngf=8, ndf=8, n_blocks=2, resample=tr.networks_resample)
    active = []
    for _ in range(8):
        out = st.step(2 * torch.rand(1, 3, 3, 32, 64) - 1, 2 * torch.rand(1, 3, 3, 32, 64) - 1)
        assert torch.isfinite(out["G"]) and torch.isfinite(out["D"])
        active.append(out["temporal_scales_active"])
    assert active == [0, 0, 1, 1, 1, 1, 2, 2]


# ---- data-parallel path: world_size-2 gloo run on CPU -----------------------------------------------------------
def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    from ir2rgb_b200.train.vid2vid_step import Vid2VidStep
    from oracle import torch_ref as tr
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(0)                              # same initial weights on every rank

    def flow_net(a, b):
        n, t, _, h, w = a.shape
        return torch.zeros(n, t, 2, h, w), torch.ones(n, t, 1, h, w)
    st = Vid2VidStep(flow_net, torch.device("cpu"), ngf=4, ndf=4, n_blocks=2, resample=tr.networks_resample, world_size=world)
    torch.manual_seed(100 + rank)                     # different frames on every rank
    for _ in range(3):
        st.step(2 * torch.rand(1, 3, 3, 16, 32) - 1, 2 * torch.rand(1, 3, 3, 16, 32) - 1)
    flat = torch.cat([p.detach().reshape(-1) for net in [st.netG, st.netD] + st.netD_T for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        ret["same"] = bool(all(torch.equal(gathered[0], g) for g in gathered[1:]))
        ret["moved"] = bool((flat - flat.mean()).abs().sum() > 0)
    dist.destroy_process_group()


def test_two_rank_gloo_training_keeps_replicas_identical():
    """Every rank sees different frames; after the per-network gradient all-reduce the replicas must stay bit-identical."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret.get("same") is True and ret.get("moved") is True
