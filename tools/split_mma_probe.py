"""CPU feasibility probe for a tensor-core Correlation: which split-precision product schemes keep
the 441 x C dot products inside BASELINE.json's forward tolerance (max-relative error <= 1e-5)?

Emulates TF32 (10 explicit mantissa bits) and BF16 (7) operand rounding in numpy, forms the split products
hi*hi + hi*lo + lo*hi (+ ...), accumulates in fp64 (an upper bound on accuracy) and in fp32 in the worst order
(every partial product rounded into one running fp32 sum), and compares with the fp64 dot product.

    python tools/split_mma_probe.py [C] [N]
"""
import sys

import numpy as np


def round_mantissa(x, drop):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    half = (1 << (drop - 1)) - 1
    r = (u + half + ((u >> drop) & 1)) & ~np.uint64((1 << drop) - 1)
    return r.astype(np.uint32).view(np.float32)


def tf32(x):
    return round_mantissa(x, 13)


def bf16(x):
    return round_mantissa(x, 16)


def main(C=256, N=20000):
    rng = np.random.default_rng(0)
    for dist in ("randn", "leaky-relu(0.1) features"):
        a = rng.standard_normal((N, C)).astype(np.float32)
        b = rng.standard_normal((N, C)).astype(np.float32)
        if dist != "randn":
            a, b = np.maximum(a, 0.1 * a), np.maximum(b, 0.1 * b)
        ref = (a.astype(np.float64) * b.astype(np.float64)).sum(1) / C

        def err(x):
            return np.abs(x - ref).max() / np.abs(ref).max()

        def fp32_running(terms):
            acc = np.zeros(N, np.float32)
            for c in range(C):
                for p, q in terms:
                    acc = acc + (p[:, c] * q[:, c]).astype(np.float32)
            return acc / np.float32(C)

        out = {"fp32 FFMA chain": err((a * b).sum(1, dtype=np.float32) / np.float32(C))}
        ah, bh = tf32(a), tf32(b)
        al, bl = tf32(a - ah), tf32(b - bh)
        out["1xTF32"] = err((ah.astype(np.float64) * bh).sum(1) / C)
        out["3xTF32, fp64 acc"] = err(((ah.astype(np.float64) * bh) + (ah.astype(np.float64) * bl) + (al.astype(np.float64) * bh)).sum(1) / C)
        out["3xTF32, fp32 running acc"] = err(fp32_running([(ah, bh), (ah, bl), (al, bh)]))
        h, g = bf16(a), bf16(b)
        m, n = bf16(a - h), bf16(b - g)
        lo, ko = bf16(a - h - m), bf16(b - g - n)
        out["1xBF16"] = err((h.astype(np.float64) * g).sum(1) / C)
        out["3xBF16, fp64 acc"] = err(((h.astype(np.float64) * g) + (h.astype(np.float64) * n) + (m.astype(np.float64) * g)).sum(1) / C)
        out["3xBF16, fp32 running acc"] = err(fp32_running([(h, g), (h, n), (m, g)]))
        out["6xBF16, fp64 acc"] = err(((h.astype(np.float64) * g) + (h.astype(np.float64) * n) + (m.astype(np.float64) * g) +
                                       (m.astype(np.float64) * n) + (h.astype(np.float64) * ko) + (lo.astype(np.float64) * g)).sum(1) / C)
        print("%s, C = %d, %d dot products (tolerance 1e-5):" % (dist, C, N))
        for k, v in out.items():
            print("    %-28s %.1e  %s" % (k, v, "ok" if v <= 1e-5 else "OUT"))


if __name__ == "__main__":
    main(*(int(v) for v in sys.argv[1:3]))
