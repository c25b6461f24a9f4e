"""Forward warp at the config-3 shape: the row-walking kernel (variant 0) against the variants of warp_rows_mlp_kernel that keep
more loads in flight per thread (flowops_warp_set_impl bits 3..5), in all three arithmetic modes, on flows of increasing
roughness; bit-identity of every variant with the row-walking kernel, also on a ragged frame.
Needs a variant build with every variant compiled in:
    python -m ir2rgb_b200.build --out tools/_exp/libflowops_mlp.so -DFLOWOPS_TUNE_WARP_MLP
    FLOWOPS_LIB=tools/_exp/libflowops_mlp.so python tools/warp_mlp_probe.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
B, H, W = 16, 512, 1024
VARIANTS = [int(v) for v in os.environ.get("PROBE_VARIANTS", "0,1,2,3,4,5,6,7").split(",")]
torch.manual_seed(0)
img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
up = torch.nn.functional.interpolate


def smooth(amp, gh, gw, b=B, h=H, w=W):
    return up(amp * torch.randn(b, 2, gh, gw, device="cuda"), size=(h, w), mode="bicubic", align_corners=False).contiguous()


flows = {
    "zero": torch.zeros(B, 2, H, W, device="cuda"),
    "gentle": smooth(20, 2, 4),
    "smooth": smooth(20, 8, 16),
    "nearest": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="nearest").contiguous(),
    "bilinear": up(20 * torch.randn(B, 2, H // 4, W // 4, device="cuda"), scale_factor=4, mode="bilinear", align_corners=False).contiguous(),
    "randn": 4 * torch.randn(B, 2, H, W, device="cuda"),
}
modes = {"bit_exact": (F.WARP_RESAMPLE2D, 0), "fp32_blend": (F.WARP_RESAMPLE2D, 2), "gridsample": (F.WARP_GRIDSAMPLE, 0)}


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
    return round(e0.elapsed_time(e1) / n * 1e3, 1)


res = {"us": {}, "identical": {}}
# ragged frame (neither dimension a multiple of the block shape), special values in the flow
rimg = 2 * torch.rand(4, 3, 509, 1531, device="cuda") - 1
rflow = smooth(30, 8, 16, 4, 509, 1531)
rflow[0, 0, 5, 7] = float("nan"); rflow[1, 1, 100, 200] = float("inf"); rflow[2, 0, 0, 0] = -3e38; rflow[3, 1, 508, 1530] = 5e6
for mname, (mode, bit) in modes.items():
    ref = {}
    for v in VARIANTS:
        lib.flowops_warp_set_impl(bit | (v << 3))
        row = {k: timed(lambda: F.warp_forward(img, f, mode)) for k, f in flows.items()}
        res["us"]["%s/v%d" % (mname, v)] = row
        outs = [F.warp_forward(img, flows["smooth"], mode), F.warp_forward(img, flows["randn"], mode), F.warp_forward(rimg, rflow, mode)]
        if v == 0:
            ref = outs
        else:
            res["identical"]["%s/v%d" % (mname, v)] = [bool(torch.equal(a.view(torch.int32), b.view(torch.int32))) for a, b in zip(outs, ref)]
        print(mname, v, json.dumps(row), res["identical"].get("%s/v%d" % (mname, v)), flush=True)
lib.flowops_warp_set_impl(0)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
