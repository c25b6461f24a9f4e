"""One vid2vid training iteration as the reference's train loop runs it (train_vid2vid.py:54-111), for
BASELINE configs[4]: single spatial scale, composite generator, two temporal discriminator scales, no VGG loss,
one frame generated per iteration (max_frames_per_gpu = 1), batch 1 per GPU.

What is on the flow hot path inside it (all through libflowops):
  * FlowNet(real_B, real_B_prev) under no_grad                   train_vid2vid.py:65  (+ flownet calls for the skipped
    frames of the temporal scales >= 1, discriminator.py:274-284),
  * the generator's warp resample(img_prev, flow), fwd + bwd       networks.py:207,
  * loss_F_Warp resample(real_B_prev, flow), fwd + bwd             discriminator.py:120,
  * loss_G_Warp resample(fake_B_prev, flow_ref), fwd only          discriminator.py:137.

Multi-GPU: one process per GPU, every rank trains on its own frames, gradients are averaged with one NCCL
all-reduce per network over NVLink after each backward (the data-parallel replacement of the reference's
replicate / frame-placement scheme, generator.py:113-180, discriminator.py:21-25) -- there is no collective on
the flow path itself.
"""
import torch
import torch.distributed as dist

from ..models import networks as _warp
from . import vid2vid_nets as N


def _reshape(t):
    """train_vid2vid.py:172-178: fold the frame axis into the batch."""
    if t is None:
        return None
    _, _, ch, h, w = t.size()
    return t.contiguous().view(-1, ch, h, w)


def skipped_frames(B_all, B, t_scales, tD):
    """discriminator.py:253-270: temporally subsampled frame groups for every temporal scale."""
    B_all = torch.cat([B_all.detach(), B], dim=1) if B_all is not None else B
    B_skipped = [None] * t_scales
    for s in range(t_scales):
        tDs = tD ** s
        span = tDs * (tD - 1)
        n_groups = min(B_all.size(1) - span, B.size(1))
        if n_groups > 0:
            for t in range(0, n_groups, tD):
                skip = B_all[:, (-span - t - 1):-t:tDs].contiguous() if t != 0 else B_all[:, -span - 1::tDs].contiguous()
                B_skipped[s] = torch.cat([B_skipped[s], skip]) if B_skipped[s] is not None else skip
    max_prev = tD ** (t_scales - 1) * (tD - 1)
    if B_all.size(1) > max_prev:
        B_all = B_all[:, -max_prev:]
    return B_all, B_skipped


class Vid2VidStep:
    def __init__(self, flow_net, device, ngf=128, ndf=64, n_downsampling=3, n_blocks=9, n_layers_D=3, num_D=2, tG=3, tD=3,
                 t_scales=2, norm="batch", lr=2e-4, beta1=0.5, lambda_feat=10.0, lambda_F=10.0, lambda_T=10.0,
                 resample=None, world_size=1, channels_last=False):
        self.dev, self.tG, self.tD, self.t_scales, self.world = device, tG, tD, t_scales, world_size
        self.n_layers_D, self.num_D = n_layers_D, num_D
        self.lambda_feat, self.lambda_F, self.lambda_T = lambda_feat, lambda_F, lambda_T
        self.flow_net = flow_net
        self.resample = resample or _warp.resample
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self.netG = N.CompositeGenerator(3 * tG, 3, 3 * (tG - 1), ngf, n_downsampling, n_blocks, norm).to(device, memory_format=fmt)
        if resample is not None:
            self.netG.resample = resample
        self.netD = N.MultiScaleDiscriminator(3 + 3, ndf, n_layers_D, norm, num_D).to(device, memory_format=fmt)
        self.netD_T = [N.MultiScaleDiscriminator(3 * tD + 2 * (tD - 1), ndf, n_layers_D, norm, num_D).to(device, memory_format=fmt)
                       for _ in range(t_scales)]
        adam = lambda net: torch.optim.Adam(net.parameters(), lr=lr, betas=(beta1, 0.999))
        self.opt_G, self.opt_D, self.opt_D_T = adam(self.netG), adam(self.netD), [adam(n) for n in self.netD_T]
        self.gan = N.GANLoss()
        self.timers = {}                    # name -> list of (start, end) CUDA events of the last iteration
        self.reset_sequence()

    # -- sequence state (train_vid2vid.py:44-52) ------------------------------------------------------------
    def reset_sequence(self):
        self.fake_B_prev_last = None
        self.frames_all = [None] * 4        # real_B_all, fake_B_all, flow_ref_all, conf_ref_all
        self.i = 0

    def _timed(self, name):
        step = self

        class _T:
            def __enter__(self):
                self.e0 = torch.cuda.Event(enable_timing=True) if step.dev.type == "cuda" else None
                if self.e0 is not None:
                    self.e0.record()

            def __exit__(self, *a):
                if self.e0 is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    step.timers.setdefault(name, []).append((self.e0, e1))
        return _T()

    # -- pieces ---------------------------------------------------------------------------------------------
    def _gan_and_fm(self, pred_real, pred_fake):
        """discriminator.py:186-201."""
        loss_gan = self.gan(pred_fake, True)
        loss_fm = torch.zeros_like(loss_gan)
        feat_w, d_w = 4.0 / (self.n_layers_D + 1), 1.0 / self.num_D
        for i in range(min(len(pred_fake), self.num_D)):
            for j in range(len(pred_fake[i]) - 1):
                loss_fm = loss_fm + d_w * feat_w * torch.nn.functional.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * self.lambda_feat
        return loss_gan, loss_fm

    def _loss_D(self, netD, real, fake):
        """discriminator.py:153-184 for already concatenated inputs: (D_real, D_fake, G_GAN, G_GAN_Feat)."""
        pred_real = netD(real)
        pred_fake = netD(fake.detach())
        loss_real, loss_fake = self.gan(pred_real, True), self.gan(pred_fake, False)
        g_gan, g_fm = self._gan_and_fm(pred_real, netD(fake))
        return loss_real, loss_fake, g_gan, g_fm

    def _allreduce(self, net, name):
        if self.world > 1:
            with self._timed("allreduce_" + name):
                grads = [p.grad for p in net.parameters() if p.grad is not None]
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat)
                flat.div_(self.world)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()

    def _backward(self, loss, net, opt, name):
        """train_vid2vid.py `loss_backward`: zero_grad, backward, step -- with the data-parallel gradient average."""
        opt.zero_grad()
        loss.backward()
        self._allreduce(net, name)
        opt.step()

    # -- one iteration ----------------------------------------------------------------------------------------
    def step(self, input_A, input_B):
        """input_A, input_B: [1, tG, 3, h, w] (labels / real frames of this window).  Returns a dict of losses."""
        tG, tD = self.tG, self.tD
        self.timers = {}
        bs, _, _, h, w = input_A.shape
        first = self.fake_B_prev_last is None
        # ---- generator (generator.py:99-182, one scale, one frame) ----
        with self._timed("generator_fwd"):
            prev = input_B[:, :tG - 1] if first else self.fake_B_prev_last
            fake_B, flow, weight, fake_B_raw = self.netG(input_A[:, :tG].reshape(bs, -1, h, w), prev.detach().reshape(bs, -1, h, w))
            seq = torch.cat([prev, fake_B.unsqueeze(1)], dim=1)
            fake_B_last = seq[:, -tG + 1:].detach()
            fake_B5, flow5, weight5, raw5 = fake_B.unsqueeze(1), flow.unsqueeze(1), weight.unsqueeze(1), fake_B_raw.unsqueeze(1)
            real_A, real_Bp = input_A[:, tG - 1:], input_B[:, tG - 2:]
        real_B_prev, real_B = real_Bp[:, :-1], real_Bp[:, 1:]
        # ---- reference flow (train_vid2vid.py:65) ----
        with self._timed("flownet"):
            flow_ref, conf_ref = self.flow_net(real_B, real_B_prev)
        fake_B_prev = real_B_prev[:, 0:1] if first else self.fake_B_prev_last[:, -1:]     # generator.py:286-290
        self.fake_B_prev_last = fake_B_last

        # ---- frame discriminator and flow / warp losses (discriminator.py:104-151) ----
        rB, fB, fBraw, rA = _reshape(real_B), _reshape(fake_B5), _reshape(raw5), _reshape(real_A)
        rBprev, fBprev, fl, wt, flr, cfr = (_reshape(t) for t in (real_B_prev, fake_B_prev, flow5, weight5, flow_ref, conf_ref))
        with self._timed("flow_losses"):
            loss_F_Flow = N.masked_l1(fl, flr, cfr) * self.lambda_F
            loss_F_Warp = N.masked_l1(self.resample(rBprev, fl), rB, cfr) * self.lambda_T
            loss_W = torch.zeros_like(wt)
            loss_G_Warp = N.masked_l1(fB, self.resample(fBprev, flr).detach(), cfr) * self.lambda_T
        with self._timed("discriminator_fwd"):
            d_real, d_fake, g_gan, g_fm = self._loss_D(self.netD, torch.cat((rA, rB), 1), torch.cat((rA, fB), 1))
            r2, f2, gg2, gf2 = self._loss_D(self.netD, torch.cat((rA, rB), 1), torch.cat((rA, fBraw), 1))
            d_real, d_fake, g_gan, g_fm = d_real + r2, d_fake + f2, g_gan + gg2, g_fm + gf2
        loss_G = g_gan + g_fm + loss_G_Warp + loss_F_Flow + loss_F_Warp + loss_W.mean()
        loss_D = (d_fake + d_real) * 0.5

        # ---- temporal discriminators (discriminator.py:219-234, 274-284, 109-118) ----
        loss_D_T = []
        with self._timed("temporal_fwd"):
            self.frames_all[0], real_skipped = skipped_frames(self.frames_all[0], real_B, self.t_scales, tD)
            self.frames_all[1], fake_skipped = skipped_frames(self.frames_all[1], fake_B5, self.t_scales, tD)
            self.frames_all[2], fl_s = skipped_frames(self.frames_all[2], flow_ref, 1, tD)
            self.frames_all[3], cf_s = skipped_frames(self.frames_all[3], conf_ref, 1, tD)
            flow_skipped, conf_skipped = [None] * self.t_scales, [None] * self.t_scales
            if fl_s[0] is not None:
                flow_skipped[0], conf_skipped[0] = fl_s[0][:, 1:], cf_s[0][:, 1:]
            for s in range(1, self.t_scales):
                if real_skipped[s] is not None and real_skipped[s].size(1) == tD:
                    with self._timed("flownet"):
                        flow_skipped[s], conf_skipped[s] = self.flow_net(real_skipped[s][:, 1:], real_skipped[s][:, :-1])
            for s in range(self.t_scales):
                if real_skipped[s] is None:
                    continue
                rb = real_skipped[s].reshape(-1, 3 * tD, h, w)
                fb = fake_skipped[s].reshape(-1, 3 * tD, h, w)
                if flow_skipped[s] is not None:
                    fr = (flow_skipped[s] / 20).reshape(-1, 2 * (tD - 1), h, w)
                    rb, fb = torch.cat([rb, fr], dim=1), torch.cat([fb, fr], dim=1)
                t_real, t_fake, t_gan, t_fm = self._loss_D(self.netD_T[s], rb, fb)
                loss_G = loss_G + t_gan + t_fm
                loss_D_T.append((s, (t_fake + t_real) * 0.5))

        # ---- backward passes (train_vid2vid.py:104-111) ----
        with self._timed("generator_bwd"):
            self._backward(loss_G, self.netG, self.opt_G, "G")
        with self._timed("discriminator_bwd"):
            self._backward(loss_D, self.netD, self.opt_D, "D")
            for s, l in loss_D_T:
                self._backward(l, self.netD_T[s], self.opt_D_T[s], "D_T%d" % s)
        self.i += 1
        return {"G": loss_G.detach(), "D": loss_D.detach(), "D_T": [l.detach() for _, l in loss_D_T],
                "F_Flow": loss_F_Flow.detach(), "F_Warp": loss_F_Warp.detach(), "G_Warp": loss_G_Warp.detach(),
                "temporal_scales_active": len(loss_D_T)}

    def timer_ms(self):
        torch.cuda.synchronize(self.dev)
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.timers.items()}
