"""The callers of `networks.resample` in a vid2vid training step: the composite generator and the multi-scale
PatchGAN discriminators (reference models/networks.py:103-220, 547-586, 627-719), restated as the harness for
BASELINE configs[4].  These are stock cuDNN layers -- nothing here is a kernel target; the only
hot-path call is `resample` (networks.py:207), which goes to libflowops through ir2rgb_b200.models.networks.

Attribute names and nn.Sequential indices are those of the reference (`model_down_seg.1.weight`,
`scale0_layer2.0.weight`, ...), so a reference checkpoint loads unchanged and parity can be checked weight for
weight (tests/test_vid2vid_nets.py).  Only what the configuration needs is built: single spatial scale, no
foreground model, flow branch on.
"""
import copy
import functools

import torch
import torch.nn as nn
from torch.nn import init

from ..models import networks as _warp


def weights_init(m):
    """networks.py:31-38."""
    if isinstance(m, (nn.Conv2d, nn.Conv3d)):
        init.normal_(m.weight, 0.0, 0.02)
    if isinstance(m, nn.BatchNorm2d):
        init.normal_(m.weight, 1.0, 0.02)
        init.zeros_(m.bias)
    if isinstance(m, nn.InstanceNorm2d) and m.weight is not None:
        init.normal_(m.weight, 1.0, 0.02)


def norm_layer_of(kind):
    """networks.py:41-48."""
    if kind == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True)
    if kind == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=True)
    raise NotImplementedError("normalization layer %s is not found" % kind)


class ResnetBlock(nn.Module):
    """x + conv-norm-relu-conv-norm with reflection padding (networks.py:547-586)."""

    def __init__(self, dim, norm_layer):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), norm_layer(dim), nn.ReLU(True),
                                        nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), norm_layer(dim))

    def forward(self, x):
        return x + self.conv_block(x)


def _stem(cin, ngf, norm_layer):
    return [nn.ReflectionPad2d(3), nn.Conv2d(cin, ngf, 7), norm_layer(ngf), nn.ReLU(True)]


def _head(ngf, cout, act=None):
    layers = [nn.ReflectionPad2d(3), nn.Conv2d(ngf, cout, 7)]
    return layers + ([act] if act is not None else [])


class CompositeGenerator(nn.Module):
    """networks.py:103-220 with use_fg_model=False, no_flow=False.

    forward(label_frames [b, tG*nc, h, w], previous_frames [b, (tG-1)*3, h, w]) ->
        (final image, flow, weight, raw image): final = raw * w + resample(last previous frame, flow) * (1 - w)."""

    def __init__(self, input_nc, output_nc, prev_output_nc, ngf, n_downsampling, n_blocks, norm="batch"):
        super().__init__()
        nl = norm_layer_of(norm)
        down = []
        for i in range(n_downsampling):
            c = ngf * 2 ** i
            down += [nn.Conv2d(c, 2 * c, 3, stride=2, padding=1), nl(2 * c), nn.ReLU(True)]
        top = ngf * 2 ** n_downsampling
        trunk = [ResnetBlock(top, nl) for _ in range(n_blocks - n_blocks // 2)]
        self.model_down_seg = nn.Sequential(*(_stem(input_nc, ngf, nl) + down + trunk))
        self.model_down_img = nn.Sequential(*(_stem(prev_output_nc, ngf, nl) + copy.deepcopy(down + trunk)))
        self.model_res_img = nn.Sequential(*[ResnetBlock(top, nl) for _ in range(n_blocks // 2)])
        up = []
        for i in range(n_downsampling):
            c = ngf * 2 ** (n_downsampling - i)
            up += [nn.ConvTranspose2d(c, c // 2, 3, stride=2, padding=1, output_padding=1), nl(c // 2), nn.ReLU(True)]
        self.model_up_img = nn.Sequential(*up)
        self.model_final_img = nn.Sequential(*_head(ngf, output_nc, nn.Tanh()))
        self.model_res_flow = copy.deepcopy(self.model_res_img)
        self.model_up_flow = copy.deepcopy(self.model_up_img)
        self.model_final_flow = nn.Sequential(*_head(ngf, 2))
        self.model_final_w = nn.Sequential(*_head(ngf, 1, nn.Sigmoid()))
        self.apply(weights_init)

    def resample(self, image, flow):
        return _warp.resample(image, flow)            # libflowops: the hot-path call of the generator (networks.py:207)

    def forward(self, labels, img_prev, use_raw_only=False):
        downsample = self.model_down_seg(labels) + self.model_down_img(img_prev)
        img_raw = self.model_final_img(self.model_up_img(self.model_res_img(downsample)))
        flow_feat = self.model_up_flow(self.model_res_flow(downsample))
        flow = self.model_final_flow(flow_feat) * 20
        weight = self.model_final_w(flow_feat)
        if use_raw_only:
            return img_raw, flow, weight, img_raw
        img_warp = self.resample(img_prev[:, -3:, ...], flow)
        w = weight.expand_as(img_raw)
        return img_raw * w + img_warp * (1 - w), flow, weight, img_raw


class MultiScaleDiscriminator(nn.Module):
    """num_D PatchGAN discriminators on an average-pooled pyramid, intermediate features exposed for the feature-matching
    loss (networks.py:627-719, getIntermFeat=True)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm="batch", num_D=2):
        super().__init__()
        nl = norm_layer_of(norm)
        self.num_D, self.n_layers = num_D, n_layers
        for i in range(num_D):
            width = min(64, ndf * 2 ** (num_D - 1 - i))
            for j, layer in enumerate(self._patch_layers(input_nc, width, n_layers, nl)):
                setattr(self, "scale%d_layer%d" % (i, j), layer)
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
        self.apply(weights_init)

    @staticmethod
    def _patch_layers(input_nc, ndf, n_layers, nl):
        kw, pad = 4, 2
        layers = [nn.Sequential(nn.Conv2d(input_nc, ndf, kw, stride=2, padding=pad), nn.LeakyReLU(0.2, True))]
        nf = ndf
        for n in range(1, n_layers + 1):
            nf_prev, nf = nf, min(nf * 2, 512)
            stride = 2 if n < n_layers else 1
            layers.append(nn.Sequential(nn.Conv2d(nf_prev, nf, kw, stride=stride, padding=pad), nl(nf), nn.LeakyReLU(0.2, True)))
        layers.append(nn.Sequential(nn.Conv2d(nf, 1, kw, stride=1, padding=pad)))
        return layers

    def forward(self, x):
        result = []
        for i in range(self.num_D):
            feats, h = [], x
            for j in range(self.n_layers + 2):
                h = getattr(self, "scale%d_layer%d" % (self.num_D - 1 - i, j))(h)
                feats.append(h)
            result.append(feats)
            if i != self.num_D - 1:
                x = self.downsample(x)
        return result


class GANLoss(nn.Module):
    """Least-squares GAN loss over the last feature map of every scale (models/loss.py:8-42 with use_lsgan truthy,
    which is what gan_mode='ls' gives at discriminator.py:61)."""

    def forward(self, preds, target_is_real):
        loss = 0
        for scale in preds:
            p = scale[-1]
            loss = loss + torch.nn.functional.mse_loss(p, torch.full_like(p, 1.0 if target_is_real else 0.0))
        return loss


def masked_l1(a, b, mask):
    """models/loss.py:105-113."""
    mask = mask.expand(-1, a.size(1), -1, -1)
    return torch.nn.functional.l1_loss(a * mask, b * mask)
