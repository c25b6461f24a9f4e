"""Tiny launcher for ncu: runs ONE operator a few times at its BASELINE config.
    python tools/profile_target.py <corr_fwd|corr_fwd_c4|corr_fwd_c4b16|corr_fwd_nhwc_b16|corr_bwd|warp_fwd|warp_bwd|cnorm_fwd|cnorm_bwd|fused|fusion_input|cnorm_fwd_f16|cnorm_bwd_f16|warp_fwd_f16> [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import functional as F  # noqa: E402

op = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
P = (20, 1, 20, 1, 2)
if op in ("corr_fwd", "corr_bwd"):
    a, b = torch.randn(8, 256, 48, 64, device="cuda"), torch.randn(8, 256, 48, 64, device="cuda")
    go = torch.randn(8, 441, 48, 64, device="cuda")
    fn = (lambda: F.correlation_forward(a, b, *P)) if op == "corr_fwd" else (lambda: F.correlation_backward(a, b, go, *P))
elif op in ("corr_fwd_nhwc_b16", "corr_fwd_nhwc_b32"):    # what the FlowNet2 step launches: planes from the conv3 epilogue, channels-last store + LeakyReLU
    nb = int(op[-2:])
    a = torch.randn(nb, 256, 64, 128, device="cuda").contiguous(memory_format=torch.channels_last)
    b = torch.randn(nb, 256, 64, 128, device="cuda").contiguous(memory_format=torch.channels_last)
    planes = F.CorrelationPlanes(a.shape, a.device)
    zb = torch.zeros(256, device="cuda")
    planes.fill_from_conv_(a, zb, 1.0, 0, write_act=False)
    planes.fill_from_conv_(b, zb, 1.0, 1, write_act=False)
    buf = F.ConcatBuffer(a, 473, 8)
    fn = lambda: F.correlation_planes_forward_into(planes, buf, 32, 0.1)
elif op in ("flow_head_c16", "flow_head_c194"):       # predict_flow of the fusion network (full resolution) / of a decoder level 2
    c_real, hh, ww = (16, 512, 1024) if op == "flow_head_c16" else (194, 128, 256)
    c_pad = -(-c_real // 8) * 8
    xh = torch.randn(16, c_pad, hh, ww, device="cuda").contiguous(memory_format=torch.channels_last)
    wp = F.pack_flow_head_weight(torch.randn(2, c_real, 3, 3, device="cuda"), c_pad)
    hb = torch.randn(2, device="cuda")
    fn = lambda: F.flow_head(xh, wp, hb)
elif op == "corr_fwd_c4b16":       # the shape bench.py's FlowNet2 step runs per micro-batch of 16 pairs
    a, b = torch.randn(16, 256, 64, 128, device="cuda"), torch.randn(16, 256, 64, 128, device="cuda")
    fn = lambda: F.correlation_forward(a, b, *P)
elif op == "corr_fwd_c4":
    a, b = torch.randn(8, 256, 64, 128, device="cuda"), torch.randn(8, 256, 64, 128, device="cuda")
    fn = lambda: F.correlation_forward(a, b, *P)
else:
    B, H, W = 16, 512, 1024
    img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
    coarse = 20 * torch.randn(B, 2, H // 64, W // 64, device="cuda")      # smooth, realistic flow field
    flow = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False).contiguous()
    gout = torch.randn(B, 3, H, W, device="cuda")
    if op == "warp_fwd":
        fn = lambda: F.warp_forward(img, flow, F.WARP_RESAMPLE2D)
    elif op == "warp_bwd":
        fn = lambda: F.warp_backward(img, flow, gout, True, True, F.WARP_RESAMPLE2D)
    elif op == "cnorm_fwd":
        fn = lambda: F.channelnorm_forward(img)
    elif op == "cnorm_bwd":
        y = F.channelnorm_forward(img)
        gy = torch.randn_like(y)
        fn = lambda: F.channelnorm_backward(img, y, gy)
    elif op == "fusion_input":       # concat3 of the FlowNet2 step, one micro-batch of 16 pairs
        x = torch.cat([img, img.flip(0)], 1).contiguous()
        lo_s2 = flow[:, :, ::4, ::4].contiguous() / 20.0
        lo_sd = flow[:, :, ::4, ::4].flip(0).contiguous() * 20.0
        fn = lambda: F.flownet2_fusion_input(x, lo_s2, lo_sd, 20.0)
    elif op == "cnorm_fwd_f16":
        img16 = img.half()
        fn = lambda: F.channelnorm_forward(img16)
    elif op == "cnorm_bwd_f16":
        img16 = img.half()
        y16 = F.channelnorm_forward(img16)
        gy16 = torch.randn_like(y16)
        fn = lambda: F.channelnorm_backward(img16, y16, gy16)
    elif op == "warp_fwd_f16":
        img16, flow16 = img.half(), flow.half()
        fn = lambda: F.warp_forward(img16, flow16, F.WARP_RESAMPLE2D)
    elif op == "fused":
        x = torch.cat([img, img.flip(0)], 1).contiguous()
        fn = lambda: F.warp_diff_norm_forward(x, flow)
    else:
        raise SystemExit("unknown op " + op)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok", op)
