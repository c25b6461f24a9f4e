"""Generate golden vectors for Correlation / Resample2d / ChannelNorm from the reference's OWN CUDA
extensions (rebuilt for sm_100 by oracle/build_ref.py into oracle/_ref/).

Runs on a GPU box:   gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.  Inputs come from
numpy.random.default_rng(seed) so they are identical everywhere; they are stored with the outputs.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_ext  # noqa: E402


def t(a):
    return torch.from_numpy(a).cuda()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    assert ref_ext.available(), "oracle/_ref/*.so missing: run python oracle/build_ref.py first"
    rng = np.random.default_rng(2024)
    g = {}

    # Correlation, FlowNetC parameters (FlowNetC.py:31) and two off-config parameter sets
    for name, (B, C, H, W, params) in {
        "corr_c": (2, 16, 8, 12, (20, 1, 20, 1, 2)),
        "corr_wide": (1, 40, 6, 40, (20, 1, 20, 1, 2)),      # C not a multiple of 32 or 8; W > one tile
        "corr_s1": (1, 8, 9, 11, (4, 1, 4, 1, 1)),           # stride2 = 1, md = 4 -> 81 channels
        "corr_s2": (1, 8, 10, 10, (4, 1, 4, 2, 2)),          # stride1 = 2 (forward only)
    }.items():
        a = rng.standard_normal((B, C, H, W)).astype(np.float32)
        b = rng.standard_normal((B, C, H, W)).astype(np.float32)
        out = ref_ext.correlation_forward(t(a), t(b), *params)
        g[name + "_a"], g[name + "_b"], g[name + "_params"] = a, b, np.array(params)
        g[name + "_out"] = out.cpu().numpy()
        if params[3] == 1:
            go = rng.standard_normal(tuple(out.shape)).astype(np.float32)
            ga, gb = ref_ext.correlation_backward(t(a), t(b), t(go), *params)
            g[name + "_gout"], g[name + "_ga"], g[name + "_gb"] = go, ga.cpu().numpy(), gb.cpu().numpy()

    # Resample2d
    for name, (B, C, H, W, sigma) in {"res_small": (2, 3, 16, 24, 3.0), "res_odd": (1, 2, 13, 19, 6.0),
                                      "res_far": (1, 3, 12, 20, 40.0)}.items():
        img = rng.standard_normal((B, C, H, W)).astype(np.float32)
        flow = (sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
        out = ref_ext.resample2d_forward(t(img), t(flow))
        go = rng.standard_normal((B, C, H, W)).astype(np.float32)
        gi, gf = ref_ext.resample2d_backward(t(img), t(flow), t(go))
        g.update({name + "_img": img, name + "_flow": flow, name + "_out": out.cpu().numpy(),
                  name + "_gout": go, name + "_gimg": gi.cpu().numpy(), name + "_gflow": gf.cpu().numpy()})

    # ChannelNorm
    for name, (B, C, H, W) in {"cn3": (2, 3, 16, 24), "cn2": (1, 2, 13, 19), "cn5": (1, 5, 8, 8)}.items():
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        x[0, :, 0, 0] = 0.0                                   # exercises the 1e-9 in the backward
        y = ref_ext.channelnorm_forward(t(x))
        gy = rng.standard_normal((B, 1, H, W)).astype(np.float32)
        gx = ref_ext.channelnorm_backward(t(x), y, t(gy))
        g.update({name + "_x": x, name + "_y": y.cpu().numpy(), name + "_gy": gy, name + "_gx": gx.cpu().numpy()})

    torch.cuda.synchronize()
    path = os.path.join(out_dir, "native_ops_ref_sm100.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "on", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
