"""Bring-up probe of the tensor-core Correlation backward (csrc/corr_tc_bwd.cu): max-relative error of both gradients
against the FP32-FMA backward of the same library on a list of shapes, and timings at the BASELINE shapes.

    python tools/tcb_probe.py [--json gpurun_out/tcb_probe.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ir2rgb_b200 import _lib, functional as F  # noqa: E402

P = (20, 1, 20, 1, 2)


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--flags", type=int, default=1)
    args = ap.parse_args()
    lib = _lib.load()
    out = []
    torch.manual_seed(0)
    for shape in [(1, 32, 2, 2), (1, 32, 32, 16), (1, 64, 16, 24), (2, 64, 34, 18), (1, 96, 6, 70), (1, 256, 20, 12), (1, 512, 8, 8),
                  (8, 256, 48, 64), (8, 256, 64, 128)]:
        a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
        go = torch.randn(shape[0], 441, shape[2], shape[3], device="cuda")
        lib.flowops_corr_set_impl(0)
        r1, r2 = F.correlation_backward(a, b, go, *P)
        lib.flowops_corr_set_impl(args.flags)
        try:
            g1, g2 = F.correlation_backward(a, b, go, *P)
            torch.cuda.synchronize()
            e = {"shape": shape, "g1": maxrel(g1, r1), "g2": maxrel(g2, r2)}
            h1, _ = F.correlation_backward(a, b, go, *P, need1=True, need2=False)
            _, h2 = F.correlation_backward(a, b, go, *P, need1=False, need2=True)
            e["single_output_calls_equal"] = bool(torch.equal(h1, g1) and torch.equal(h2, g2))
            lib.flowops_corr_set_impl(3)
            k1, k2 = F.correlation_backward(a, b, go, *P)
            e["smemA_g1"], e["smemA_g2"] = maxrel(k1, r1), maxrel(k2, r2)
            lib.flowops_corr_set_impl(args.flags)
        except Exception as ex:                      # noqa: BLE001
            e = {"shape": shape, "error": repr(ex)}
            print(json.dumps(e), flush=True)
            out.append(e)
            break
        if shape[0] == 8:
            for name, flags in (("tc", args.flags), ("tc_smemA", 3), ("ffma", 0)):
                lib.flowops_corr_set_impl(flags)
                for _ in range(3):
                    F.correlation_backward(a, b, go, *P)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    F.correlation_backward(a, b, go, *P)
                e1.record()
                e1.synchronize()
                e[name + "_us"] = e0.elapsed_time(e1) / 20 * 1e3
        print(json.dumps(e), flush=True)
        out.append(e)
    lib.flowops_corr_set_impl(1)
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
