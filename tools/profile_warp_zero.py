"""ncu target: the forward warp at the config-3 shape on a ZERO flow (perfectly coalesced gathers).
    python tools/profile_warp_zero.py [impl_flags]       # 0 = bit-exact blend, 2 = fp32 blend"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
lib.flowops_warp_set_impl(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
img = 2 * torch.rand(16, 3, 512, 1024, device="cuda") - 1
flow = torch.zeros(16, 2, 512, 1024, device="cuda")
for _ in range(5):
    F.warp_forward(img, flow, F.WARP_RESAMPLE2D)
torch.cuda.synchronize()
