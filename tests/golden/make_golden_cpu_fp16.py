"""Generate golden vectors for the fp16 form of ``Model.resample`` from the reference's OWN Python code.

Runs in the build container only (needs /root/reference; torch CPU).  ``get_grid``
(models/networks.py:15-28) and ``Model.grid_sample`` (models/base_model.py:123-127, the
``opt['fp16']`` branch: ``grid_sample(input1.float(), input2.float(), ...).half()``) are imported and
called; the lines of ``Model.resample`` between them (base_model.py:131-134: grid in the flow's dtype,
flow normalisation, grid add / permute) are restated verbatim because the method ends in
``.cuda(image.get_device())`` and cannot run without a GPU.  On the CPU ``half / scalar`` is a true
divide (the CUDA kernel multiplies by the fp32 reciprocal): the C oracle restates both
(``inv_mode``), these vectors pin the CPU form.

Output: tests/golden/resample_fp16_cpu.npz (fp16 inputs and outputs stored as float32).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("IR2RGB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_base_model():
    pkg = types.ModuleType("ref_models")
    pkg.__path__ = [os.path.join(REF, "models")]          # namespace-style: no __init__ side effects
    sys.modules["ref_models"] = pkg
    return importlib.import_module("ref_models.base_model"), importlib.import_module("ref_models.networks")


def reference_resample_fp16(base_model, net, image, flow):
    stub = types.SimpleNamespace(opt={"fp16": True})
    b, c, h, w = image.size()
    grid = net.get_grid(b, h, w, device="cpu", dtype=flow.dtype)                                   # base_model.py:132
    flow = torch.cat([flow[:, 0:1, :, :] / ((w - 1.0) / 2.0), flow[:, 1:2, :, :] / ((h - 1.0) / 2.0)], dim=1)   # :133
    final_grid = (grid + flow).permute(0, 2, 3, 1)                                                 # :134
    return base_model.Model.grid_sample(stub, image, final_grid)                                   # :135, :123-127


def main():
    torch.set_num_threads(1)
    base_model, net = load_reference_base_model()
    rng = np.random.default_rng(4321)
    cases = {}
    for name, (B, C, H, W, sigma) in {"small": (2, 3, 16, 24, 3.0), "odd": (1, 2, 13, 19, 6.0),
                                      "border": (1, 3, 12, 20, 40.0), "wide": (1, 3, 8, 300, 2.0)}.items():
        img = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).half()
        flow = torch.from_numpy((sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)).half()
        out = reference_resample_fp16(base_model, net, img, flow)
        assert out.dtype == torch.float16
        cases.update({name + "_img": img.float().numpy(), name + "_flow": flow.float().numpy(),
                      name + "_out": out.float().numpy()})
    path = os.path.join(HERE, "resample_fp16_cpu.npz")
    np.savez_compressed(path, **cases)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
