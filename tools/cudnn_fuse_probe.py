"""Does cuDNN (frontend 1.18 / backend 9.22) have an sm_100 engine for  conv -> +bias -> LeakyReLU  with fp32 I/O and TF32
math on channels-last tensors, and is it faster than cuDNN's plain convolution followed by libflowops' bias_lrelu pass?
Probed on the layers of FlowNet2 whose epilogue pass costs most at 512x1024 (16 pairs)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudnn  # noqa: E402
from ir2rgb_b200 import functional as F  # noqa: E402

torch.backends.cudnn.benchmark = True
handle = cudnn.create_handle()


def build(x, w, b, y, stride, pad, slope):
    cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)
    g = cudnn.pygraph(handle=handle, io_data_type=cudnn.data_type.FLOAT, intermediate_data_type=cudnn.data_type.FLOAT,
                      compute_data_type=cudnn.data_type.FLOAT)
    X = g.tensor(name="X", dim=list(x.shape), stride=list(x.stride()), data_type=cudnn.data_type.FLOAT)
    W = g.tensor(name="W", dim=list(w.shape), stride=list(w.stride()), data_type=cudnn.data_type.FLOAT)
    Bt = g.tensor(name="B", dim=[1, b.numel(), 1, 1], stride=[b.numel(), 1, b.numel(), b.numel()], data_type=cudnn.data_type.FLOAT)
    c = g.conv_fprop(image=X, weight=W, padding=[pad, pad], stride=[stride, stride], dilation=[1, 1])
    t = g.bias(input=c, bias=Bt)
    o = g.leaky_relu(input=t, negative_slope=slope)
    o.set_output(True).set_dim(list(y.shape)).set_stride(list(y.stride())).set_data_type(cudnn.data_type.FLOAT)
    g.validate()
    g.build_operation_graph()
    g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    g.check_support()
    g.build_plans(cudnn.build_plan_policy.ALL)
    ws = torch.empty(max(g.get_workspace_size(), 1), device="cuda", dtype=torch.uint8)
    return g, {X: x, W: w, Bt: b, o: y}, ws


def time_it(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = []
# (name, B, Cin, H, W, Cout, k, stride)
layers = [("fusion.conv0", 16, 16, 512, 1024, 64, 3, 1), ("S.conv1", 16, 16, 512, 1024, 64, 7, 2), ("S.conv2", 16, 64, 256, 512, 128, 5, 2),
          ("fusion.conv1", 16, 64, 512, 1024, 64, 3, 2), ("S.conv3", 16, 128, 128, 256, 256, 5, 2), ("C.conv3_1", 16, 480, 64, 128, 256, 3, 1)]
for name, B, ci, H, W, co, k, s in layers:
    torch.manual_seed(0)
    x = torch.randn(B, ci, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = (0.05 * torch.randn(co, ci, k, k, device="cuda")).contiguous(memory_format=torch.channels_last)
    b = torch.randn(co, device="cuda")
    pad = (k - 1) // 2
    y_ref = torch.nn.functional.conv2d(x, w, None, s, pad)
    rec = {"layer": name, "in": [B, ci, H, W], "out": list(y_ref.shape), "k": k, "stride": s}

    def plain():
        y = torch.nn.functional.conv2d(x, w, None, s, pad)
        return F.bias_lrelu_(y, b, 0.1)
    rec["us_conv_only"] = time_it(lambda: torch.nn.functional.conv2d(x, w, None, s, pad))
    rec["us_conv_plus_epilogue"] = time_it(plain)
    want = plain()
    try:
        y = torch.empty_like(y_ref)
        t0 = time.time()
        g, pack, ws = build(x, w, b, y, s, pad, 0.1)
        rec["build_s"] = time.time() - t0
        rec["plans"] = g.get_execution_plan_count()
        g.execute(pack, ws, handle=handle)
        torch.cuda.synchronize()
        rec["maxrel_vs_plain"] = ((y - want).abs().max() / want.abs().max()).item()
        rec["us_fused"] = time_it(lambda: g.execute(pack, ws, handle=handle))
        best = None
        for i in range(g.get_execution_plan_count()):
            try:
                wsi = torch.empty(max(g.get_workspace_size_plan_at_index(i), 1), device="cuda", dtype=torch.uint8)
                us = time_it(lambda: g.execute_plan_at_index(pack, wsi, i, handle=handle), n=5)
                if best is None or us < best[0]:
                    best = (us, i, g.get_plan_name_at_index(i))
            except Exception:
                pass
        rec["best_plan"] = best
    except Exception as e:
        rec["error"] = repr(e)[:400]
    print(json.dumps(rec), flush=True)
    res.append(rec)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "cudnn_fuse_probe.json"), "w"), indent=1)
