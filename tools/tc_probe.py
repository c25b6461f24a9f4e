"""Bring-up / A-B probe of the tcgen05 Correlation forward (csrc/corr_tc.cu) on a B200.

    python tools/tc_probe.py [--out gpurun_out/tc_probe.jsonl]

Each case runs in its own interpreter (a trapped kernel kills the CUDA context, not the probe) under a time limit.
For every (shape, impl flags) it reports the max-relative error against the FP32-FMA kernel of the same library and
against an fp64 einsum on one batch item, and the time per call (CUDA events, 20 back-to-back launches, inputs > L2 at
the large shapes).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import json, sys, torch
sys.path.insert(0, %(root)r)
from ir2rgb_b200 import _lib, functional as F
from oracle import torch_ref as tr
shape, flags, layout = %(shape)r, %(flags)d, %(layout)r
lib = _lib.load()
P = (20, 1, 20, 1, 2)
torch.manual_seed(0)
a = torch.randn(*shape, device="cuda"); b = torch.randn(*shape, device="cuda")
if layout == "nhwc":
    a = a.contiguous(memory_format=torch.channels_last); b = b.contiguous(memory_format=torch.channels_last)
lib.flowops_corr_set_impl(0)
ref = F.correlation_forward(a, b, *P)
lib.flowops_corr_set_impl(flags)
out = F.correlation_forward(a, b, *P)
torch.cuda.synchronize()
def maxrel(x, y): return ((x.double() - y.double()).abs().max() / y.double().abs().max()).item()
res = {"shape": shape, "flags": flags, "layout": layout, "vs_ffma": maxrel(out, ref)}
truth = tr.correlation(a[:1].double().contiguous(), b[:1].double().contiguous(), *P)
res["vs_fp64"] = maxrel(out[:1], truth); res["ffma_vs_fp64"] = maxrel(ref[:1], truth)
bad = (out - ref).abs() > 1e-3 * ref.abs().max()
res["n_bad"] = int(bad.sum().item())
if res["n_bad"]:
    idx = bad.nonzero()[:8].tolist(); res["bad_idx"] = idx
def time_it(fl):
    lib.flowops_corr_set_impl(fl)
    for _ in range(3): F.correlation_forward(a, b, *P)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): F.correlation_forward(a, b, *P)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
res["us"] = time_it(flags); res["us_ffma"] = time_it(0)
print("RESULT " + json.dumps(res))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tc_probe.jsonl"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    cases = [((1, 32, 32, 16), 5, "nchw"), ((1, 32, 32, 16), 1, "nchw"), ((1, 32, 32, 16), 3, "nchw"),
             ((2, 64, 40, 24), 1, "nchw"), ((2, 64, 40, 24), 1, "nhwc"), ((1, 32, 34, 18), 1, "nchw"),
             ((8, 256, 48, 64), 1, "nchw"), ((8, 256, 48, 64), 3, "nchw"),
             ((16, 256, 64, 128), 1, "nchw"), ((16, 256, 64, 128), 1, "nhwc")]
    if args.quick:
        cases = cases[:4]
    with open(args.out, "a") as f:
        for shape, flags, layout in cases:
            code = CASE % {"root": ROOT, "shape": shape, "flags": flags, "layout": layout}
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[0][7:]) if line else {"shape": shape, "flags": flags, "layout": layout, "rc": r.returncode,
                                                            "stderr": r.stderr[-1500:]}
            except subprocess.TimeoutExpired:
                rec = {"shape": shape, "flags": flags, "layout": layout, "timeout": True}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
