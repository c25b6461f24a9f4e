"""CPU tests: the oracle against (a) the golden vectors produced by the reference itself and (b) the
pure-PyTorch fp64 closed forms.  No GPU, no libflowops compute."""
import numpy as np
import pytest
import torch

from oracle import torch_ref as tr


def maxrel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---- (a) golden vectors from the reference's own CUDA extensions (bit-exactness expected) --------
@pytest.mark.parametrize("name", ["cn3", "cn2", "cn5"])
def test_cnorm_oracle_vs_reference_golden(c_oracle, golden_native, name):
    g = golden_native
    y = c_oracle.cnorm_fwd(g[name + "_x"])
    assert np.array_equal(y, g[name + "_y"]), "ChannelNorm forward must be bit-identical to the reference kernel"
    gx = c_oracle.cnorm_bwd(g[name + "_x"], g[name + "_y"], g[name + "_gy"])
    assert np.array_equal(gx, g[name + "_gx"]), "ChannelNorm backward must be bit-identical to the reference kernel"


@pytest.mark.parametrize("name", ["res_small", "res_odd", "res_far"])
def test_resample2d_oracle_vs_reference_golden(c_oracle, golden_native, name):
    g = golden_native
    out = c_oracle.resample2d_fwd(g[name + "_img"], g[name + "_flow"])
    assert np.array_equal(out, g[name + "_out"]), "Resample2d forward must be bit-identical to the reference kernel"
    gimg, gflow = c_oracle.resample2d_bwd(g[name + "_img"], g[name + "_flow"], g[name + "_gout"])
    assert np.array_equal(gflow, g[name + "_gflow"])
    # the image gradient is an atomic scatter on the GPU: order differs, values agree to fp32 rounding
    assert maxrel(gimg, g[name + "_gimg"]) <= 1e-6


@pytest.mark.parametrize("name", ["corr_c", "corr_wide", "corr_s1", "corr_s2"])
def test_correlation_oracle_vs_reference_golden(c_oracle, golden_native, name):
    g = golden_native
    p = [int(v) for v in g[name + "_params"]]
    out = c_oracle.corr_fwd(g[name + "_a"], g[name + "_b"], *p)
    assert out.shape == g[name + "_out"].shape
    assert np.array_equal(out, g[name + "_out"]), "Correlation forward must be bit-identical to the reference kernel"
    if name + "_gout" in g:
        ga, gb = c_oracle.corr_bwd(g[name + "_a"], g[name + "_b"], g[name + "_gout"], *p)
        assert np.array_equal(ga, g[name + "_ga"])
        assert np.array_equal(gb, g[name + "_gb"])


# ---- (a') golden vectors from the reference's Python path for networks.resample ------------------
@pytest.mark.parametrize("name", ["small", "odd", "border"])
def test_gridwarp_oracle_vs_reference_python_golden(c_oracle, golden_resample, name):
    g = golden_resample
    img, flow = g[name + "_img"], g[name + "_flow"]
    H, W = img.shape[2:]
    lx, ly = torch.linspace(-1, 1, W).numpy(), torch.linspace(-1, 1, H).numpy()
    # CPU op chain (true divide; ATen's vectorised unnormalize is a fused multiply-add):
    # bit-identical to the CPU-generated golden
    out_cpu = c_oracle.gridwarp_fwd(img, flow, lx, ly, inv_mode=0, fma_mode=1)
    assert np.array_equal(out_cpu, g[name + "_out"])
    # CUDA op chain (reciprocal multiply + fma): same values to coordinate-rounding accuracy
    out_gpu_chain = c_oracle.gridwarp_fwd(img, flow, lx, ly, inv_mode=1, fma_mode=1)
    assert maxrel(out_gpu_chain, g[name + "_out"]) <= 2e-5


def test_torch_ref_networks_resample_matches_golden(golden_resample):
    g = golden_resample
    for name in ["small", "odd", "border"]:
        it = torch.from_numpy(g[name + "_img"]).requires_grad_()
        ft = torch.from_numpy(g[name + "_flow"]).requires_grad_()
        out = tr.networks_resample(it, ft)
        assert np.array_equal(out.detach().numpy(), g[name + "_out"])
        out.backward(torch.from_numpy(g[name + "_gout"]))
        assert maxrel(it.grad.numpy(), g[name + "_gimg"]) <= 1e-6
        assert maxrel(ft.grad.numpy(), g[name + "_gflow"]) <= 1e-6


def test_c1_digest_reproducible():
    """BASELINE config 1 (1x3x256x512) through the restated code path reproduces the digest the
    reference's own functions produced."""
    d = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "resample_c1_digest.npz"))
    torch.manual_seed(0)
    img = torch.randn(1, 3, 256, 512)
    flow = 5 * torch.randn(1, 2, 256, 512)
    assert abs(img.double().sum().item() - float(d["img_sum"])) < 1e-9
    out = tr.networks_resample(img, flow)
    assert np.array_equal(out.flatten().numpy()[d["idx"]], d["samples"])
    assert abs(out.double().sum().item() - float(d["sum"])) <= 1e-6 * abs(float(d["sum"])) + 1e-6


# ---- (b) oracle vs fp64 closed forms --------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 3, 7, 9), (1, 2, 16, 16), (1, 7, 5, 3)])
def test_cnorm_oracle_vs_fp64(c_oracle, shape):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape).astype(np.float32)
    y = c_oracle.cnorm_fwd(x)
    assert maxrel(y, tr.channelnorm(torch.from_numpy(x).double()).numpy()) <= 1e-6
    gy = rng.standard_normal(y.shape).astype(np.float32)
    gx = c_oracle.cnorm_bwd(x, y, gy)
    ref = tr.channelnorm_bwd(torch.from_numpy(x).double(), torch.from_numpy(y).double(), torch.from_numpy(gy).double())
    assert maxrel(gx, ref.numpy()) <= 1e-6


@pytest.mark.parametrize("sigma", [0.0, 0.7, 4.0, 60.0])
def test_resample2d_oracle_vs_fp64(c_oracle, sigma):
    rng = np.random.default_rng(1)
    img = rng.standard_normal((2, 3, 11, 13)).astype(np.float32)
    flow = (sigma * rng.standard_normal((2, 2, 11, 13))).astype(np.float32)
    it = torch.from_numpy(img).double().requires_grad_()
    ft = torch.from_numpy(flow).double().requires_grad_()
    ot = tr.resample2d(it, ft)
    out = c_oracle.resample2d_fwd(img, flow)
    assert maxrel(out, ot.detach().numpy()) <= 1e-6
    if sigma == 0.0:
        assert np.array_equal(out, img)          # zero flow is the identity
    go = rng.standard_normal(out.shape).astype(np.float32)
    ot.backward(torch.from_numpy(go).double())
    gi, gf = c_oracle.resample2d_bwd(img, flow, go)
    assert maxrel(gi, it.grad.numpy()) <= 2e-6
    assert maxrel(gf, ft.grad.numpy()) <= 2e-6


def test_resample2d_equals_grid_sample_align_corners_true(c_oracle):
    """Resample2d == grid_sample(bilinear, border, align_corners=True) on vid2vid's grid."""
    rng = np.random.default_rng(2)
    img = rng.standard_normal((1, 3, 9, 14)).astype(np.float32)
    flow = (3 * rng.standard_normal((1, 2, 9, 14))).astype(np.float32)
    it, ft = torch.from_numpy(img).double(), torch.from_numpy(flow).double()
    H, W = 9, 14
    grid = tr.vid2vid_grid(1, H, W, "cpu", torch.float64)
    lin = torch.cat([torch.linspace(-1, 1, W, dtype=torch.float64).view(1, 1, 1, W).expand(1, 1, H, W),
                     torch.linspace(-1, 1, H, dtype=torch.float64).view(1, 1, H, 1).expand(1, 1, H, W)], 1)
    nflow = torch.cat([ft[:, 0:1] / ((W - 1.0) / 2.0), ft[:, 1:2] / ((H - 1.0) / 2.0)], 1)
    ref = torch.nn.functional.grid_sample(it, (lin + nflow).permute(0, 2, 3, 1), mode="bilinear",
                                          padding_mode="border", align_corners=True)
    assert maxrel(c_oracle.resample2d_fwd(img, flow), ref.numpy()) <= 1e-6
    del grid


@pytest.mark.parametrize("params,shape", [((20, 1, 20, 1, 2), (2, 16, 12, 14)), ((4, 1, 4, 1, 1), (1, 5, 7, 9)),
                                          ((6, 1, 6, 1, 2), (1, 33, 6, 6))])
def test_correlation_oracle_vs_fp64(c_oracle, params, shape):
    rng = np.random.default_rng(3)
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    at = torch.from_numpy(a).double().requires_grad_()
    bt = torch.from_numpy(b).double().requires_grad_()
    ot = tr.correlation(at, bt, *params)
    out = c_oracle.corr_fwd(a, b, *params)
    assert out.shape == tuple(ot.shape)
    assert maxrel(out, ot.detach().numpy()) <= 1e-6
    go = rng.standard_normal(out.shape).astype(np.float32)
    ot.backward(torch.from_numpy(go).double())
    ga, gb = c_oracle.corr_bwd(a, b, go, *params)
    assert maxrel(ga, at.grad.numpy()) <= 2e-6
    assert maxrel(gb, bt.grad.numpy()) <= 2e-6


def test_correlation_channel_order_dy_outer(c_oracle):
    """tc = (tj + r) * D + (ti + r): the vertical displacement is the slow index
    (correlation_cuda_kernel.cu:107-110,139-140)."""
    a = np.zeros((1, 1, 8, 8), np.float32)
    b = np.zeros((1, 1, 8, 8), np.float32)
    a[0, 0, 4, 4] = 1.0
    b[0, 0, 4 + 2, 4 - 4] = 1.0      # dy = +2 (tj = +1), dx = -4 (ti = -2)
    out = c_oracle.corr_fwd(a, b, 20, 1, 20, 1, 2)
    tc = (1 + 10) * 21 + (-2 + 10)
    assert out[0, tc, 4, 4] == 1.0 and np.count_nonzero(out) == 1


# ---- (c) 16-bit storage: the reference's fp16 mode ------------------------------------------------
def test_round16_matches_numpy_float16_and_torch_bfloat16(c_oracle):
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2 ** 32, size=500_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    special = np.array([0.0, -0.0, 65504, 65519.99, 65520, 65536, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.000001, 3 * 2.0 ** -25,
                        2.0 ** -14, 6.1e-5, np.inf, -np.inf, np.nan], np.float32)
    x = np.concatenate([x, special, rng.standard_normal(200_000).astype(np.float32),
                        (1e-5 * rng.standard_normal(200_000)).astype(np.float32)])
    with np.errstate(all="ignore"):
        want = x.astype(np.float16).astype(np.float32)
    got = c_oracle.round16(x, c_oracle.DTYPE_F16)
    assert np.array_equal(got, want, equal_nan=True) and np.array_equal(np.signbit(got), np.signbit(want))
    want = torch.from_numpy(x.copy()).bfloat16().float().numpy()
    got = c_oracle.round16(x, c_oracle.DTYPE_BF16)
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("name", ["small", "odd", "border", "wide"])
def test_gridwarp16_oracle_vs_reference_python_fp16_golden(c_oracle, name):
    """Model.resample with opt['fp16'] on half tensors, run by the reference's own code on the CPU
    (tests/golden/make_golden_cpu_fp16.py): the oracle's CPU form (true divide) must reproduce it bit for bit."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "resample_fp16_cpu.npz"))
    img, flow = g[name + "_img"], g[name + "_flow"]
    H, W = img.shape[2:]
    lx = torch.linspace(-1, 1, W).half().float().numpy()
    ly = torch.linspace(-1, 1, H).half().float().numpy()
    out = c_oracle.gridwarp_fwd_16(img, flow, lx, ly, c_oracle.DTYPE_F16, inv_mode=0, fma_mode=1)
    assert np.array_equal(out, g[name + "_out"])
    # the CUDA form (reciprocal multiply) differs from it only where the 16-bit rounding of flow / scale flips
    out_cuda_form = c_oracle.gridwarp_fwd_16(img, flow, lx, ly, c_oracle.DTYPE_F16, inv_mode=1, fma_mode=1)
    assert np.mean(out_cuda_form != g[name + "_out"]) < 0.02


@pytest.mark.parametrize("dtype_name", ["float16", "bfloat16"])
def test_cnorm16_oracle_vs_torch_emulation(c_oracle, dtype_name):
    """channelnorm_kernel.cu:55-59 instantiated for a 16-bit type, emulated with torch CPU tensors of that type:
    square rounded to the type, fp32 sum in channel order, fp32 sqrt, rounded to the type."""
    dt = getattr(torch, dtype_name)
    code = c_oracle.DTYPE_F16 if dt == torch.float16 else c_oracle.DTYPE_BF16
    torch.manual_seed(3)
    x = (3 * torch.randn(2, 3, 9, 11)).to(dt)
    x[0, :, 0, 0] = 0
    x[0, 0, 0, 1] = 300.0                       # 300^2 overflows fp16: the reference kernel yields inf there
    acc = torch.zeros(2, 9, 11)
    for c in range(3):
        v = x[:, c].float()
        acc = acc + (v * v).to(dt).float()
    want = acc.sqrt().to(dt).float().numpy()[:, None]
    got = c_oracle.cnorm_fwd_16(x.float().numpy(), code)
    assert np.array_equal(got, want)
    if dt == torch.float16:
        assert np.isinf(got[0, 0, 0, 1])
    gy = torch.randn(2, 1, 9, 11).to(dt)
    y = torch.from_numpy(got)
    q = (gy.float() * x.float()).double() / (y.double() + 1e-9)
    want_gx = q.float().to(dt).float().numpy()
    got_gx = c_oracle.cnorm_bwd_16(x.float().numpy(), got, gy.float().numpy(), code)
    assert np.array_equal(got_gx, want_gx, equal_nan=True)


def test_cnorm_backward_quotient_model(c_oracle):
    """The product's ChannelNorm backward replaces the reference's fp64 divide by a reciprocal seed + two Newton steps +
    one residual correction (csrc/cnorm.cu).  CPU model of that arithmetic against (float)((double)prod / d) for seeds far
    worse than the hardware's: the fp32 results agree everywhere (the double quotient may differ in its last place, which
    survives the rounding to fp32 only once in ~2^29 cases)."""
    rng = np.random.default_rng(5)
    n = 4_000_000
    y = np.abs(rng.standard_normal(n)).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6, n).astype(np.float32)
    y[:1000] = 0.0                                                   # d = 1e-9
    d = y.astype(np.float64) + 1e-9
    prod = (rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)).astype(np.float32)
    prod[1000:2000] = 0.0
    prod[2000:2100] = -0.0
    delta = rng.uniform(-1.0, 1.0, n) * 2.0 ** -16
    got, want = c_oracle.quotient_model(prod, d, delta)
    assert np.array_equal(np.signbit(got), np.signbit(want))
    assert (got != want).sum() <= 2, (got != want).sum()
    assert np.abs(got.astype(np.float64) - want).max() <= np.abs(want).max() * 2.0 ** -23
