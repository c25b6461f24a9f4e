"""Resample2d forward / backward (default kernels) at the config-3 shape on flows of increasing roughness, from the flows video
frames have (zero, a constant translation, a few pixels of slowly varying motion) to the three synthetic flows of SURVEY 8d.
Shows how much of the distance to the HBM roofline is the access pattern of the benchmark flow and how much is the kernel.
    python tools/warp_flow_sweep.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
B, H, W = 16, 512, 1024
FWD_BYTES = B * (3 + 2 + 3) * H * W * 4                 # image + flow in, image out
BWD_BYTES = B * (3 + 2 + 3 + 3 + 2) * H * W * 4         # image, flow, gout in; gimg, gflow out (memset of gimg not counted)
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6439.5) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6439.5
torch.manual_seed(0)
img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
gout = torch.randn(B, 3, H, W, device="cuda")
up = torch.nn.functional.interpolate


def smooth(amp, gh, gw):
    return up(amp * torch.randn(B, 2, gh, gw, device="cuda"), size=(H, W), mode="bicubic", align_corners=False).contiguous()


const = torch.empty(B, 2, H, W, device="cuda")
const[:, 0] = 3.3
const[:, 1] = -1.7
flows = {
    "zero": torch.zeros(B, 2, H, W, device="cuda"),
    "constant translation (3.3, -1.7) px": const,
    "3 px on a 2x4 grid (~0.015 px/px)": smooth(3, 2, 4),
    "20 px on a 2x4 grid (~0.1 px/px)": smooth(20, 2, 4),
    "20 px on a 4x8 grid (~0.2 px/px)": smooth(20, 4, 8),
    "20 px on an 8x16 grid (~0.34 px/px: the 'smooth' row of bench.py)": smooth(20, 8, 16),
    "nearest x4 of 20 randn (SURVEY 8d)": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="nearest").contiguous(),
    "bilinear x4 of 20 randn (SURVEY 8d)": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="bilinear",
                                              align_corners=False).contiguous(),
    "4 randn per pixel (SURVEY 8d)": 4 * torch.randn(B, 2, H, W, device="cuda"),
}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = {}
for name, flow in flows.items():
    row = {}
    us = timed(lambda: F.warp_forward(img, flow, F.WARP_RESAMPLE2D))
    row["fwd_us"], row["fwd_frac_hbm"] = round(us, 1), round(FWD_BYTES / us * 1e-3 / PEAK, 3)
    us = timed(lambda: F.warp_backward(img, flow, gout, True, True, F.WARP_RESAMPLE2D))
    row["bwd_us"], row["bwd_frac_hbm"] = round(us, 1), round(BWD_BYTES / us * 1e-3 / PEAK, 3)
    us = timed(lambda: F.warp_backward(img, flow, gout, False, True, F.WARP_RESAMPLE2D))
    row["bwd_flow_only_us"] = round(us, 1)
    res[name] = row
    print(name, json.dumps(row), flush=True)
res["_note"] = ("16 x 3 x 512 x 1024, default kernels, 20 back-to-back launches after 3 warm-ups; working set 268 / 436 MB > L2; "
                "frac = algorithmic bytes / time / %.1f GB/s" % PEAK)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
