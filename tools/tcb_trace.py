"""Where the cycles of the tensor-core Correlation backward go: per-role wait counters (flowops_corr_tc_trace) averaged over
the CTAs, config 2 by default."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

lib = _lib.load()
P = (20, 1, 20, 1, 2)
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 256, 48, 64)
a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
go = torch.randn(shape[0], 441, shape[2], shape[3], device="cuda")
for flags in (1, 5, 3):
    lib.flowops_corr_set_impl(flags)
    for _ in range(3):
        F.correlation_backward(a, b, go, *P)
    trace = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    lib.flowops_corr_tc_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); F.correlation_backward(a, b, go, *P); e1.record()
    torch.cuda.synchronize()
    lib.flowops_corr_tc_trace(None)
    t = trace.view(148, 8).double().cpu()
    names = ["producer_wait_empty", "mma_wait_ready", "mma_wait_tmem_empty", "split_wait_full", "skew_wait_empty", "split_warp0_busy",
             "cta_lifetime", "skew_busy"]
    res = {n: round(t[:, i].mean().item()) for i, n in enumerate(names)}
    res["cta_lifetime_max"] = round(t[:, 6].max().item())
    res["us"] = e0.elapsed_time(e1) * 1e3
    res["variant"] = {1: "A in TMEM", 5: "A in TMEM, ONE UMMA per K block (timing experiment, wrong numerics)", 3: "A in shared memory"}[flags]
    print(json.dumps(res))
lib.flowops_corr_set_impl(1)
