"""Does cuDNN's graph path fuse  conv_dgrad -> +bias -> LeakyReLU  (a ConvTranspose2d(k4, s2, p1) with its epilogue) for fp32 /
TF32 channels-last tensors on sm_100, writing into a channel slice of a concat buffer -- and is it faster than torch's
conv_transpose2d followed by libflowops' epilogue pass?  FlowNet2 decoder shapes at 512x1024, 16 pairs."""
import json
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudnn  # noqa: E402
from ir2rgb_b200 import cudnn_fused as cf, functional as F  # noqa: E402

torch.backends.cudnn.benchmark = True
FL = cudnn.data_type.FLOAT
res = []
for name, B, ci, h, w, co in [("deconv5", 16, 1024, 8, 16, 512), ("deconv4", 16, 1032, 16, 32, 256), ("deconv3", 16, 776, 32, 64, 128),
                              ("deconv2", 16, 392, 64, 128, 64), ("fusion.deconv1", 16, 128, 128, 256, 32), ("fusion.deconv0", 16, 168, 256, 512, 16)]:
    torch.manual_seed(0)
    conv = torch.nn.ConvTranspose2d(ci, co, 4, 2, 1).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(B, ci, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
    buf = F.ConcatBuffer(x, co + 10, 8, shape=(B, 2 * h, 2 * w))
    dst = buf.tensor[:, 8:8 + co]
    rec = {"layer": name}
    with torch.no_grad():
        def unfused():
            buf.bias_lrelu_in(torch.nn.functional.conv_transpose2d(x, conv.weight, None, 2, 1), conv.bias, 0.1, 8)
        unfused()
        want = dst.clone()
        rec["us_unfused"] = cf._time3(unfused) / 3 * 1e3
        rec["us_dgrad_only"] = cf._time3(lambda: torch.nn.functional.conv_transpose2d(x, conv.weight, None, 2, 1)) / 3 * 1e3
        dst.zero_()
        try:
            handle = cf._handle(x.device)
            cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)
            g = cudnn.pygraph(handle=handle, io_data_type=FL, intermediate_data_type=FL, compute_data_type=FL)
            wt = conv.weight.detach()
            DY = g.tensor(name="DY", dim=list(x.shape), stride=list(x.stride()), data_type=FL)
            W = g.tensor(name="W", dim=list(wt.shape), stride=list(wt.stride()), data_type=FL)
            Bt = g.tensor(name="B", dim=[1, co, 1, 1], stride=[co, 1, co, co], data_type=FL)
            dx = g.conv_dgrad(loss=DY, filter=W, padding=[1, 1], stride=[2, 2], dilation=[1, 1])
            dx.set_dim([B, co, 2 * h, 2 * w])
            t = g.bias(input=dx, bias=Bt)
            o = g.leaky_relu(input=t, negative_slope=0.1)
            o.set_output(True).set_dim(list(dst.shape)).set_stride(list(dst.stride())).set_data_type(FL)
            g.validate(); g.build_operation_graph()
            g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
            g.check_support(); g.build_plans(cudnn.build_plan_policy.ALL)
            pack = {DY: x, W: wt, Bt: conv.bias, o: dst}
            best = None
            for i in range(g.get_execution_plan_count()):
                try:
                    ws = torch.empty(max(g.get_workspace_size_plan_at_index(i), 1), device="cuda", dtype=torch.uint8)
                    ms = cf._time3(lambda: g.execute_plan_at_index(pack, ws, i, handle=handle))
                    if best is None or ms < best[0]:
                        best = (ms, i, ws)
                except Exception:
                    pass
            rec["plans"] = g.get_execution_plan_count()
            dst.zero_()
            g.execute_plan_at_index(pack, best[2], best[1], handle=handle)
            torch.cuda.synchronize()
            rec["maxrel"] = ((dst - want).abs().max() / want.abs().max()).item()
            rec["us_fused"] = best[0] / 3 * 1e3
        except Exception:
            rec["error"] = traceback.format_exc()[-700:]
    print(json.dumps(rec), flush=True)
    res.append(rec)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "cudnn_dgrad_probe.json"), "w"), indent=1)
