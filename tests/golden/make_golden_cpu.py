"""Generate golden vectors for networks.resample from the reference's OWN Python code path.

Runs in the build container only (needs /root/reference; torch CPU).  The reference's
``get_grid`` (models/networks.py:15-28) and ``BaseCompositeGeneratorModule.grid_sample``
(networks.py:89-91) are imported and called; the two lines between them (flow normalisation and
grid add/permute, networks.py:97-98) are restated verbatim because ``resample`` itself ends in
``.cuda(image.get_device())`` and cannot run without a GPU.

Output: tests/golden/resample_cpu.npz  (inputs + outputs, small shapes) and
        tests/golden/resample_c1_digest.npz (BASELINE config 1, 1x3x256x512: seeds + sampled outputs).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("IR2RGB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_networks():
    spec = importlib.util.spec_from_file_location("ref_networks", os.path.join(REF, "models", "networks.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_resample(net, image, flow):
    b, c, h, w = image.size()
    grid = net.get_grid(b, h, w, device='cpu', dtype=flow.dtype)                     # networks.py:96
    nflow = torch.cat([flow[:, 0:1, :, :] / ((w - 1.0) / 2.0),
                       flow[:, 1:2, :, :] / ((h - 1.0) / 2.0)], dim=1)                # networks.py:97
    final_grid = (grid + nflow).permute(0, 2, 3, 1)                                  # networks.py:98
    return net.BaseCompositeGeneratorModule.grid_sample(image, final_grid)           # networks.py:99


def main():
    torch.set_num_threads(1)
    net = load_reference_networks()
    cases = {}
    rng = np.random.default_rng(1234)
    for name, (B, C, H, W, sigma) in {
        "small": (2, 3, 16, 24, 3.0),
        "odd": (1, 2, 13, 19, 6.0),
        "border": (1, 3, 12, 20, 40.0),
    }.items():
        img = rng.standard_normal((B, C, H, W)).astype(np.float32)
        flow = (sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
        it = torch.from_numpy(img).requires_grad_()
        ft = torch.from_numpy(flow).requires_grad_()
        out = reference_resample(net, it, ft)
        gout = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(torch.from_numpy(gout))
        cases.update({name + "_img": img, name + "_flow": flow, name + "_out": out.detach().numpy(),
                      name + "_gout": gout, name + "_gimg": it.grad.numpy(), name + "_gflow": ft.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "resample_cpu.npz"), **cases)

    # BASELINE config 1: torch.manual_seed(0); img = randn(1,3,256,512); flow = 5*randn
    torch.manual_seed(0)
    img = torch.randn(1, 3, 256, 512)
    flow = 5 * torch.randn(1, 2, 256, 512)
    out = reference_resample(net, img, flow)
    idx = np.random.default_rng(7).integers(0, out.numel(), size=4096)
    np.savez_compressed(os.path.join(HERE, "resample_c1_digest.npz"),
                        idx=idx, samples=out.flatten().numpy()[idx],
                        sum=np.float64(out.double().sum().item()), sumsq=np.float64((out.double() ** 2).sum().item()),
                        img_sum=np.float64(img.double().sum().item()), flow_sum=np.float64(flow.double().sum().item()))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    sys.exit(main())
