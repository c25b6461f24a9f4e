"""A/B of the warp backward kernels (flowops_warp_set_impl bit 0: owned accumulation in per-warp windows) on a B200:
parity of the two against each other on awkward shapes, and time per call at BASELINE config 3 for the benchmark's
flow flavours."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()


def maxrel(x, y):
    return ((x.double() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-30)).item()


def both(img, flow, go, mode, need_flow=True):
    lib.flowops_warp_set_impl(0)
    r = F.warp_backward(img, flow, go, True, need_flow, mode)
    lib.flowops_warp_set_impl(1)
    n = F.warp_backward(img, flow, go, True, need_flow, mode)
    torch.cuda.synchronize()
    return r, n


torch.manual_seed(0)
out = {"parity": [], "timing": []}
for (B, C, H, W, amp, mode) in [(2, 3, 17, 28, 3.0, 0), (1, 1, 40, 64, 10.0, 0), (2, 2, 33, 100, 40.0, 0), (1, 3, 64, 96, 0.0, 0),
                                (2, 3, 48, 68, 3.0, 1), (1, 3, 30, 36, 40.0, 1), (3, 3, 70, 132, 0.5, 0), (1, 3, 256, 512, 5.0, 1)]:
    img = torch.randn(B, C, H, W, device="cuda")
    flow = (amp * torch.randn(B, 2, H, W, device="cuda")).contiguous()
    go = torch.randn(B, C, H, W, device="cuda")
    (gi_r, gf_r), (gi_n, gf_n) = both(img, flow, go, mode)
    out["parity"].append({"case": [B, C, H, W, amp, mode], "gimg": maxrel(gi_n, gi_r), "gflow": maxrel(gf_n, gf_r),
                          "mass": abs(gi_n.double().sum().item() - gi_r.double().sum().item())})
print(json.dumps(out["parity"]))

B, H, W = 16, 512, 1024
img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
go = torch.randn(B, 3, H, W, device="cuda")
coarse = 20 * torch.randn(B, 2, H // 64, W // 64, device="cuda")
low = 20 * torch.randn(B, 2, H // 4, W // 4, device="cuda")
flows = {"smooth": torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False).contiguous(),
         "nearest_up": torch.nn.functional.interpolate(low, scale_factor=4, mode="nearest").contiguous(),
         "bilinear_up": torch.nn.functional.interpolate(low, scale_factor=4, mode="bilinear").contiguous(),
         "randn4": (4 * torch.randn(B, 2, H, W, device="cuda")).contiguous(),
         "zero": torch.zeros(B, 2, H, W, device="cuda")}
imgs = [img.clone() for _ in range(3)]            # rotate inputs: > L2
for name, flow in flows.items():
    for mode in (0, 1):
        (gi_r, gf_r), (gi_n, gf_n) = both(img, flow, go, mode)
        rec = {"flow": name, "mode": mode, "gimg_maxrel": maxrel(gi_n, gi_r), "gflow_maxrel": maxrel(gf_n, gf_r)}
        for impl in (0, 1):
            lib.flowops_warp_set_impl(impl)
            for need_flow in (True, False):
                for _ in range(2):
                    F.warp_backward(img, flow, go, True, need_flow, mode)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for i in range(12):
                    F.warp_backward(imgs[i % 3], flow, go, True, need_flow, mode)
                e1.record(); torch.cuda.synchronize()
                rec["us_impl%d_%s" % (impl, "both" if need_flow else "img")] = e0.elapsed_time(e1) / 12 * 1e3
        out["timing"].append(rec)
        print(json.dumps(rec), flush=True)
lib.flowops_warp_set_impl(0)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "warp_probe.json"), "w"), indent=1)
