"""INTEGRATION.md section 3 ("Option B") executed: the reference's UNMODIFIED Python wrappers -- correlation.py,
resample2d.py, channelnorm.py, byte-compiled from /root/reference into oracle/_ref/refpy by oracle/build_ref.py --
import `correlation_cuda` / `resample2d_cuda` / `channelnorm_cuda`, which here resolve to the ctypes shims of
ir2rgb_b200/shims over libflowops.so.  Results are compared with the same wrappers running on the reference's own
rebuilt extensions.

Each configuration runs in a fresh interpreter: the wrappers bind their `*_cuda` module at import time.
"""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, torch
    sys.path.insert(0, %(root)r)
    backend = sys.argv[1]
    from oracle import build_ref, ref_ext
    if backend == "shims":
        import ir2rgb_b200.shims as shims
        shims.install()
    else:
        for name in ("correlation_cuda", "resample2d_cuda", "channelnorm_cuda"):
            sys.modules[name] = ref_ext._load(name)
    build_ref.install_finder()
    from flownet2_pytorch.networks.correlation_package.correlation import Correlation
    from flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d
    from flownet2_pytorch.networks.channelnorm_package.channelnorm import ChannelNorm
    import correlation_cuda, channelnorm_cuda
    torch.manual_seed(0)
    out = {}
    a, b = torch.randn(2, 32, 16, 24, device="cuda"), torch.randn(2, 32, 16, 24, device="cuda")
    with torch.no_grad():                     # correlation.py:13 passes ints to save_for_backward: forward only under no_grad
        out["corr"] = Correlation(20, 1, 20, 1, 2, 1)(a, b)
    go = torch.randn_like(out["corr"])
    g1, g2 = a.new(), a.new()
    correlation_cuda.backward(a, b, a.new(), a.new(), go, g1, g2, 20, 1, 20, 1, 2, 1)      # what correlation.py:36 calls
    out["corr_g1"], out["corr_g2"] = g1, g2
    img = torch.randn(2, 3, 20, 28, device="cuda").requires_grad_()
    flow = (3 * torch.randn(2, 2, 20, 28, device="cuda")).requires_grad_()
    warped = Resample2d()(img, flow)          # the reference wrapper's forward AND backward work as shipped
    out["warp"] = warped.detach()
    gw = torch.randn_like(warped)
    warped.backward(gw)
    out["warp_gimg"], out["warp_gflow"] = img.grad, flow.grad
    with torch.no_grad():
        x = torch.randn(2, 3, 20, 28, device="cuda")
        out["cnorm"] = ChannelNorm()(x)
        x16 = x.half()
        out["cnorm16"] = ChannelNorm()(x16).float()
    gx = torch.zeros_like(x)
    channelnorm_cuda.backward(x, out["cnorm"], torch.ones_like(out["cnorm"]), gx, 2)       # channelnorm.py:25 (NameError as shipped)
    out["cnorm_gx"] = gx
    torch.cuda.synchronize()
    torch.save({k: v.cpu() for k, v in out.items()}, sys.argv[2])
''')


def _run(backend, path):
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, backend, path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]


def test_reference_wrappers_run_unmodified_on_the_ctypes_shims(flowops_lib, tmp_path):
    from oracle import harness
    if not harness.reference_flownet2_available():
        pytest.skip("oracle/_ref (extensions + refpy) not built")
    _run("shims", str(tmp_path / "shims.pt"))
    _run("reference", str(tmp_path / "ref.pt"))
    new, ref = torch.load(tmp_path / "shims.pt"), torch.load(tmp_path / "ref.pt")

    def maxrel(k):
        return ((new[k].double() - ref[k].double()).abs().max() / ref[k].double().abs().max().clamp_min(1e-30)).item()

    assert new["corr"].shape == ref["corr"].shape == (2, 441, 16, 24)
    assert maxrel("corr") <= 1e-5
    assert maxrel("corr_g1") <= 1e-4 and maxrel("corr_g2") <= 1e-4
    assert torch.equal(new["warp"], ref["warp"])
    assert maxrel("warp_gimg") <= 1e-4 and maxrel("warp_gflow") <= 1e-4
    assert torch.equal(new["cnorm"], ref["cnorm"]) and torch.equal(new["cnorm16"], ref["cnorm16"])
    assert maxrel("cnorm_gx") <= 1e-6
