"""Where the cycles of the tensor-core Correlation kernel go: per-role wait / work counters (flowops_corr_tc_trace)
averaged over the CTAs, for the FlowNet2 in-step launch (16 pairs, planes from NHWC features, channels-last store)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
a = torch.randn(B, 256, 64, 128, device="cuda").contiguous(memory_format=torch.channels_last)
b = torch.randn(B, 256, 64, 128, device="cuda").contiguous(memory_format=torch.channels_last)
planes = F.CorrelationPlanes(a.shape, a.device)
zb = torch.zeros(256, device="cuda")
planes.fill_from_conv_(a, zb, 1.0, 0, write_act=False)
planes.fill_from_conv_(b, zb, 1.0, 1, write_act=False)
buf = F.ConcatBuffer(a, 473, 8)
fn = lambda: F.correlation_planes_forward_into(planes, buf, 32, 0.1)
for _ in range(3):
    fn()
trace = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
lib.flowops_corr_tc_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record()
torch.cuda.synchronize()
lib.flowops_corr_tc_trace(None)
t = trace.view(148, 8).double().cpu()
names = ["producer_wait_empty", "mma_wait_split", "mma_wait_tmem_empty", "split_wait_full", "epi_wait_tmem_full",
         "epi_busy", "cta_lifetime", "split_busy"]
res = {n: round(t[:, i].mean().item()) for i, n in enumerate(names)}
res["us"] = e0.elapsed_time(e1) * 1e3
res["items_per_cta"] = B * 4 * 2 * 8 * 4 / 148       # planes x 16 tiles x 4 window quarters
print(json.dumps(res))
