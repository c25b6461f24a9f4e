"""Does the input-channel padding of the full-resolution first layers change what cuDNN makes of them?
FlowNetSD.conv0 reads the 6-channel frame stack padded to 8 channels (977 us per 16 pairs in the launch list), the fusion
network's conv0 reads 11 channels padded to 16 (698 us): the same 3x3 / 64-filter layer, more input, less time.
Times conv -> bias -> LeakyReLU through ir2rgb_b200.cudnn_fused (best fused plan) and the two-kernel path for both paddings."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import cudnn_fused  # noqa: E402
from ir2rgb_b200 import functional as F  # noqa: E402

torch.backends.cudnn.benchmark = True


def time_it(fn, n=10):
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
    B = int(os.environ.get("PROBE_B", "16"))
    out = []
    for name, cin_real, k, s, co in (("SD.conv0", 6, 3, 1, 64), ("S.conv1", 12, 7, 2, 64), ("fusion.conv0", 11, 3, 1, 64)):
        for cpad in (8, 16, 32):
            if cpad < cin_real:
                continue
            torch.manual_seed(0)
            x = torch.zeros(B, cpad, 512, 1024, device="cuda").contiguous(memory_format=torch.channels_last)
            x[:, :cin_real] = torch.randn(B, cin_real, 512, 1024, device="cuda")
            conv = torch.nn.Conv2d(cin_real, co, k, s, (k - 1) // 2).cuda()
            w = torch.zeros(co, cpad, k, k, device="cuda")
            w[:, :cin_real] = conv.weight.detach()
            w = w.contiguous(memory_format=torch.channels_last)
            pad = (k - 1) // 2

            def plain():
                return F.bias_lrelu_(torch.nn.functional.conv2d(x, w, None, s, pad), conv.bias.detach(), 0.1)
            rec = {"layer": name, "cin": cin_real, "cpad": cpad, "B": B}
            with torch.no_grad():
                rec["us_conv_only"] = time_it(lambda: torch.nn.functional.conv2d(x, w, None, s, pad))
                rec["us_two_kernels"] = time_it(plain)
                y = cudnn_fused.conv_bias_lrelu(conv, x, w, 0.1, unfused=plain)
                if y is NotImplemented:
                    rec["fused"] = "not kept / unavailable"
                else:
                    rec["us_fused"] = time_it(lambda: cudnn_fused.conv_bias_lrelu(conv, x, w, 0.1))
                    want = plain()
                    rec["maxrel"] = ((y - want).abs().max() / want.abs().max()).item()
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del x, w
            torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/conv_pad_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
