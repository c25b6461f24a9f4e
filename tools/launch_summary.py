"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): kernel families by share, top launches of a step.
    python tools/launch_summary.py profiles/launches_r02_flownet2_b16.csv [steps_in_file]"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        out.append((r[ki], v))
    return out


def family(name):
    if "flowops" in name or name.startswith("tc::") or "tc::" in name:
        return "libflowops"
    if any(s in name for s in ("cudnn", "cutlass", "xmma", "implicit_gemm", "nhwcAddPadding", "convertTensor", "engines_precompiled")):
        return "cuDNN / CUTLASS"
    return "torch elementwise / copies"


if __name__ == "__main__":
    L = load(sys.argv[1])
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    tot = sum(v for _, v in L)
    agg, fam = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(float)
    for k, v in L:
        agg[k][0] += 1
        agg[k][1] += v
        fam[family(k)] += v
    print("launches %d (%d per step), %.1f us total, %.1f us per step" % (len(L), len(L) // steps, tot, tot / steps))
    print({k: "%.1f %%" % (100 * v / tot) for k, v in fam.items()})
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
        print("%6.2f%% %5d %9.1f  %s" % (100 * t / tot, n, t, k[:110]))
    n = len(L) // steps
    print("-- largest launches of the last step")
    for k, v in sorted(L[-n:], key=lambda kv: -kv[1])[:14]:
        print("%8.1f  %s" % (v, k[:110]))
