"""A/B timing of flowops_flow_head_nhwc's rows-per-strip knob (FLOWOPS_TUNE_HEAD_ROWS) at the flow-head shapes of FlowNet2
(16 pairs, 512 x 1024)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ir2rgb_b200 import functional as F  # noqa: E402


def time_us(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B = 16
    res = []
    for c_real, hh, ww in ((16, 512, 1024), (32, 256, 512), (194, 128, 256)):
        c_pad = -(-c_real // 8) * 8
        torch.manual_seed(c_real)
        x = torch.randn(B, c_pad, hh, ww, device="cuda").contiguous(memory_format=torch.channels_last)
        w = torch.randn(2, c_real, 3, 3, device="cuda")
        bias = torch.randn(2, device="cuda")
        wp = F.pack_flow_head_weight(w, c_pad)
        rec = {"cin": c_real, "hw": [hh, ww]}
        for rows in ("auto", "4", "8", "16", "32", "64"):
            if rows == "auto":
                os.environ.pop("FLOWOPS_TUNE_HEAD_ROWS", None)
            else:
                os.environ["FLOWOPS_TUNE_HEAD_ROWS"] = rows
            rec["rows_" + rows] = round(time_us(lambda: F.flow_head(x, wp, bias)), 1)
        print(json.dumps(rec), flush=True)
        res.append(rec)
    os.environ.pop("FLOWOPS_TUNE_HEAD_ROWS", None)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "head_tune.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
