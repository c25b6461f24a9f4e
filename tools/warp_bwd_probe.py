"""Warp backward, config-3 shape: fixed-point shared-memory accumulation (flowops_warp_set_impl(4)) vs the default direct
reductions on flows of increasing roughness."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
B, H, W = 16, 512, 1024
torch.manual_seed(0)
img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
gout = torch.randn(B, 3, H, W, device="cuda")
up = torch.nn.functional.interpolate


def smooth(amp, gh, gw):
    return up(amp * torch.randn(B, 2, gh, gw, device="cuda"), size=(H, W), mode="bicubic", align_corners=False).contiguous()


flows = {
    "zero": torch.zeros(B, 2, H, W, device="cuda"),
    "gentle (20 px on a 2x4 grid, ~0.1 px/px)": smooth(20, 2, 4),
    "medium (20 px on a 4x8 grid, ~0.2 px/px)": smooth(20, 4, 8),
    "smooth (20 px on an 8x16 grid: opbench)": smooth(20, 8, 16),
    "nearest x4 of a gentle quarter-res flow": up(smooth(20, 2, 4)[:, :, ::4, ::4].contiguous(), scale_factor=4, mode="nearest").contiguous(),
    "4 randn per pixel": 4 * torch.randn(B, 2, H, W, device="cuda"),
}
res = {}
for name, flow in flows.items():
    row = {}
    for tag, flags in (("fixed_point", 4), ("direct", 0)):
        lib.flowops_warp_set_impl(flags)
        for mode, mname in ((F.WARP_RESAMPLE2D, "r2d"),):
            for _ in range(3):
                F.warp_backward(img, flow, gout, True, True, mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                F.warp_backward(img, flow, gout, True, True, mode)
            e1.record()
            e1.synchronize()
            row[tag + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    lib.flowops_warp_set_impl(4)
    a = F.warp_backward(img, flow, gout, True, True, F.WARP_RESAMPLE2D)
    lib.flowops_warp_set_impl(0)
    b = F.warp_backward(img, flow, gout, True, True, F.WARP_RESAMPLE2D)
    row["gimg_maxrel"] = ((a[0].double() - b[0].double()).abs().max() / b[0].double().abs().max()).item()
    row["gflow_equal"] = bool(torch.equal(a[1], b[1]))
    res[name] = row
    print(name, json.dumps(row), flush=True)
lib.flowops_warp_set_impl(0)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
