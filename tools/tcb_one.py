import sys, os, torch
sys.path.insert(0, os.getcwd())
from ir2rgb_b200 import functional as F
P = (20, 1, 20, 1, 2)
torch.manual_seed(0)
a, b = torch.randn(8, 256, 48, 64, device="cuda"), torch.randn(8, 256, 48, 64, device="cuda")
go = torch.randn(8, 441, 48, 64, device="cuda")
for _ in range(3):
    F.correlation_backward(a, b, go, *P)
torch.cuda.synchronize()
