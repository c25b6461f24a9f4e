"""A/B timing of libflowops builds in ONE process on ONE GPU, rounds interleaved.

Box-to-box and run-to-run differences on this pool are a few per cent -- as large as most kernel changes -- so two
builds must be compared inside one process: the ChannelNorm-backward work of round 1 only converged once four builds
were timed this way.  Every variant is a complete libflowops.so (build it from a scratch copy
of csrc/, e.g. into tools/_exp/, which is git-ignored but travels to the GPU box); the operators run through the
normal ctypes wrappers, which are pointed at one variant after the other.

    python tools/ab_ops.py --lib base=ir2rgb_b200/libflowops.so --lib new=tools/_exp/libflowops_new.so \
                           --only cnorm_bwd [--rounds 4] [--check]

--only is the row filter of tools/opbench.py (substring of the row name).  --check only loads the variants and verifies
that each exports every symbol of include/flowops.h (no GPU needed).
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ir2rgb_b200 import _lib  # noqa: E402


def load_variant(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    missing = []
    for name, (res, args) in _lib.SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype, fn.argtypes = res, args
    return lib, missing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", action="append", required=True, metavar="NAME=PATH")
    ap.add_argument("--only", default=None)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    variants = {}
    for spec in args.lib:
        name, _, path = spec.partition("=")
        lib, missing = load_variant(path)
        if missing:
            print("%s: missing symbols %s (rows that need them will fail)" % (name, ", ".join(missing)))
        variants[name] = lib
    if args.check:
        print("loaded", ", ".join(variants))
        return
    import opbench
    from ir2rgb_b200 import functional as F
    _lib._lib = next(iter(variants.values()))
    ffma = F.ffma_peak_tflops()
    table = {}
    for rnd in range(args.rounds):
        for name, lib in variants.items():
            _lib._lib = lib                       # functional.* fetch the library through _lib.load() on every call
            res = opbench.run(iters=5, skip_ref=True, quiet=True, ffma=ffma, only=args.only)
            for row, e in res["ops"].items():
                table.setdefault(row, {}).setdefault(name, []).append(e["us"])
    for row, by in table.items():
        print(row)
        for name, us in by.items():
            print("    %-12s %s   min %.1f us" % (name, " ".join("%.1f" % u for u in us), min(us)))
    print(json.dumps({row: {n: min(u) for n, u in by.items()} for row, by in table.items()}))


if __name__ == "__main__":
    main()
