"""What the FlowNet2 step is made of, layer by layer, without ncu: (1) one forward of bench.py's network at 16 pairs, then the
record cuDNN-fusion keeps per layer (fused plan vs conv + epilogue pass, which one runs); (2) event timings of the
depth-to-space / flow-slice epilogues at the shapes of the fusion network, with and without the whole-sector tail write."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ir2rgb_b200 import cudnn_fused  # noqa: E402
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
    B = int(os.environ.get("PROBE_B", "16"))
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    out = {"B": B}
    if os.environ.get("PROBE_NET", "1") != "0":
        torch.manual_seed(0)
        net = bench.build_native(dev)
        im1 = 2 * torch.rand(B, 3, 512, 1024, device=dev) - 1
        im2 = 2 * torch.rand(B, 3, 512, 1024, device=dev) - 1
        with torch.no_grad():
            net(im1, im2)
            net(im1, im2)
            out["forward_us"] = time_us(lambda: net(im1, im2), n=5)
        rep = cudnn_fused.report()
        out["layers"] = rep
        for r in rep:
            print(json.dumps(r))
        del net
        torch.cuda.empty_cache()

    # depth-to-space epilogue of FlowNetFusion.deconv0 (162 -> 16, output 512 x 1024, concat record 64 + 16 + 2 (+6) channels)
    h, w, C = 256, 512, 16
    y4 = torch.randn(B, 4 * C, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    flow = torch.randn(B, 2, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    bias, fw, fb = torch.randn(C, device=dev), torch.randn(2, 2, 4, 4, device=dev), torch.randn(2, device=dev)
    buf = F.ConcatBuffer(y4, 64 + C + 2, 8, shape=(B, 2 * h, 2 * w))
    prev = F.D2S_WRITE_PAD
    for mode in (False, True):
        F.D2S_WRITE_PAD = mode
        tag = "whole_sectors" if mode else "8_byte_flow_store"
        out["d2s_flowup_" + tag] = time_us(lambda: buf.bias_lrelu_d2s_in(y4, bias, 0.1, 64, (flow, fw, fb)))
        out["flow_deconv_fullres_" + tag] = time_us(lambda: buf.flow_deconv_in(flow, fw, fb, 64 + C))
    out["d2s_plain"] = time_us(lambda: buf.bias_lrelu_d2s_in(y4, bias, 0.1, 64))
    gb = (y4.numel() * 4 * 2 + B * 4 * h * w * 8) / 1e9
    out["d2s_flowup_alg_GB"] = gb
    # the flow slice of a FlowNetS decoder level (level 2: 128 + 64 + 2 (+6) channels at 128 x 256)
    h2, w2 = 64, 128
    flow2 = torch.randn(B, 2, h2, w2, device=dev).contiguous(memory_format=torch.channels_last)
    buf2 = F.ConcatBuffer(flow2, 194, 8, shape=(B, 2 * h2, 2 * w2))
    for mode in (False, True):
        F.D2S_WRITE_PAD = mode
        out["flow_deconv_level2_" + ("whole_sectors" if mode else "8_byte_flow_store")] = time_us(lambda: buf2.flow_deconv_in(flow2, fw, fb, 192))
    F.D2S_WRITE_PAD = prev

    # the flow heads (predict_flow: C -> 2, 3x3) at every decoder level: libflowops' direct FP32 kernel vs the cuDNN path
    from ir2rgb_b200.models.flownet2_pytorch.networks import submodules as sm
    heads = []
    for c_real, hh, ww in ((1024, 8, 16), (1026, 16, 32), (770, 32, 64), (386, 64, 128), (194, 128, 256), (32, 256, 512), (16, 512, 1024)):
        torch.manual_seed(c_real)
        conv = sm.predict_flow(c_real).cuda()
        c_pad = -(-c_real // 8) * 8
        xh = torch.zeros(B, c_pad, hh, ww, device=dev).contiguous(memory_format=torch.channels_last)
        xh[:, :c_real] = torch.randn(B, c_real, hh, ww, device=dev)
        rec = {"cin": c_real, "c_pad": c_pad, "hw": [hh, ww]}
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = False
            want = torch.nn.functional.conv2d(xh[:, :c_real], conv.weight, conv.bias, 1, 1)
            torch.backends.cudnn.allow_tf32 = True
            for on in (True, False):
                sm.FLOW_HEAD_KERNEL = on
                got = sm.apply_conv(conv, xh)
                rec["maxrel_vs_fp32_" + ("kernel" if on else "cudnn_tf32")] = ((got - want).abs().max() / want.abs().max()).item()
                rec["us_" + ("kernel" if on else "cudnn_path")] = time_us(lambda: sm.apply_conv(conv, xh))
            sm.FLOW_HEAD_KERNEL = True
        heads.append(rec)
        print(json.dumps(rec))
    out["flow_heads"] = heads
    print(json.dumps({k: v for k, v in out.items() if k not in ("layers", "flow_heads")}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "step_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
