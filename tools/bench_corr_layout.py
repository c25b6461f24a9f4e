import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import functional as F
from torch.profiler import ProfilerActivity, profile
P = (20, 1, 20, 1, 2)
a = torch.randn(16, 256, 64, 128, device="cuda"); b = torch.randn(16, 256, 64, 128, device="cuda")
acl, bcl = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
for _ in range(3):
    F.correlation_forward(a, b, *P); F.correlation_forward(acl, bcl, *P)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        F.correlation_forward(a, b, *P); F.correlation_forward(acl, bcl, *P)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "flowops" in e.key:
        print("%-60s %8.1f us x %d" % (e.key[:60], e.device_time_total / e.count, e.count))
