"""Kernel-time breakdown of one FlowNet forward (micro-batch 8, 512x1024) with torch.profiler.
    python tools/profile_flownet.py [channels_last|contiguous] [out.txt]
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

mf = sys.argv[1] if len(sys.argv) > 1 else "channels_last"
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = bench.build_native(dev, mf)
im1 = 2 * torch.rand(8, 3, 512, 1024, device=dev) - 1
im2 = 2 * torch.rand(8, 3, 512, 1024, device=dev) - 1
for _ in range(3):
    net(im1, im2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        net(im1, im2)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
print(tab)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(tab)
