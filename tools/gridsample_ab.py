"""GRIDSAMPLE forward through the default path (warp_rows_mlp_kernel on large frames) against the row-walking kernel (flag bit 3)
at the config-3 shape and at the 1-channel shape of the vid2vid step.    python tools/gridsample_ab.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
up = torch.nn.functional.interpolate
torch.manual_seed(0)
res = {}
for C in (3, 1):
    B, H, W = 16, 512, 1024
    img = 2 * torch.rand(B, C, H, W, device="cuda") - 1
    flows = {
        "zero": torch.zeros(B, 2, H, W, device="cuda"),
        "smooth": up(20 * torch.randn(B, 2, 8, 16, device="cuda"), size=(H, W), mode="bicubic", align_corners=False).contiguous(),
        "nearest": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="nearest").contiguous(),
        "bilinear": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="bilinear", align_corners=False).contiguous(),
        "randn": 4 * torch.randn(B, 2, H, W, device="cuda"),
    }
    for name, flow in flows.items():
        row = {}
        for tag, flag in (("two_rows_in_flight_us", 0), ("row_walking_us", 8), ("two_rows_in_flight_again_us", 0)):
            lib.flowops_warp_set_impl(flag)
            for _ in range(3):
                F.warp_forward(img, flow, F.WARP_GRIDSAMPLE)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                F.warp_forward(img, flow, F.WARP_GRIDSAMPLE)
            e1.record()
            e1.synchronize()
            row[tag] = round(e0.elapsed_time(e1) / 30 * 1e3, 1)
        res["C%d/%s" % (C, name)] = row
        print(C, name, json.dumps(row), flush=True)
lib.flowops_warp_set_impl(0)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
