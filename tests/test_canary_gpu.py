"""Out-of-bounds write detection without compute-sanitizer (closed on this pool): every output and workspace
handed to the C ABI is a window inside a larger buffer pre-filled with a sentinel; the guard bands must come
back untouched and the window fully overwritten.  Odd / ragged shapes on purpose."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SENT = 123456.0
GUARD = 4096          # floats on each side


def guarded(numel):
    buf = torch.full((numel + 2 * GUARD,), SENT, device="cuda", dtype=torch.float32)
    return buf, buf[GUARD:GUARD + numel]


def check(buf, numel, name, must_fill=True):
    assert torch.all(buf[:GUARD] == SENT), name + ": wrote before the buffer"
    assert torch.all(buf[GUARD + numel:] == SENT), name + ": wrote past the buffer"
    if must_fill:
        assert not torch.any(buf[GUARD:GUARD + numel] == SENT), name + ": left part of the output unwritten"


def p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


@pytest.fixture(scope="module")
def lib(flowops_lib):
    return flowops_lib


@pytest.mark.parametrize("shape", [(2, 16, 8, 12), (1, 5, 7, 9), (1, 40, 6, 70), (1, 64, 13, 36), (2, 3, 24, 32)])
def test_correlation_fast_path_stays_in_bounds(lib, shape):
    B, C, H, W = shape
    torch.manual_seed(0)
    a, b = torch.randn(shape, device="cuda"), torch.randn(shape, device="cuda")
    P = (20, 1, 20, 1, 2)
    n_out = B * 441 * H * W
    obuf, out = guarded(n_out)
    ws_bytes = lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *P)
    wbuf, ws = guarded((ws_bytes + 3) // 4 + 64)
    ws_off = (-ws.data_ptr()) % 256 // 4                      # 256-byte aligned start inside the window
    ws = ws[ws_off:]
    rc = lib.flowops_corr_fwd(p(a), p(b), p(out), B, C, H, W, *P, 0, p(ws), ws_bytes, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(obuf, n_out, "corr_fwd out")
    check(wbuf, (ws_bytes + 3) // 4 + 64, "corr_fwd workspace", must_fill=False)

    go = torch.randn(B, 441, H, W, device="cuda")
    g1buf, g1 = guarded(a.numel())
    g2buf, g2 = guarded(a.numel())
    ws_bytes = lib.flowops_corr_bwd_workspace_bytes(B, C, H, W, *P)
    wbuf, ws = guarded((ws_bytes + 3) // 4 + 64)
    ws = ws[(-ws.data_ptr()) % 256 // 4:]
    rc = lib.flowops_corr_bwd(p(a), p(b), p(go), p(g1), p(g2), B, C, H, W, *P, p(ws), ws_bytes, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(g1buf, a.numel(), "corr_bwd grad1")
    check(g2buf, a.numel(), "corr_bwd grad2")
    check(wbuf, (ws_bytes + 3) // 4 + 64, "corr_bwd workspace", must_fill=False)


@pytest.mark.parametrize("params,shape", [((4, 1, 4, 1, 1), (1, 5, 9, 11)), ((3, 3, 2, 1, 1), (1, 4, 6, 7)), ((4, 1, 4, 2, 2), (1, 8, 10, 10))])
def test_correlation_generic_stays_in_bounds(lib, params, shape):
    B, C, H, W = shape
    a, b = torch.randn(shape, device="cuda"), torch.randn(shape, device="cuda")
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.flowops_corr_out_shape(H, W, *params, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
    n_out = B * oc.value * oh.value * ow.value
    obuf, out = guarded(n_out)
    assert lib.flowops_corr_fwd(p(a), p(b), p(out), B, C, H, W, *params, 0, None, 0, None) == 0
    torch.cuda.synchronize()
    check(obuf, n_out, "corr_fwd generic")
    if params[3] == 1:
        go = torch.randn(n_out, device="cuda")
        g1buf, g1 = guarded(a.numel())
        g2buf, g2 = guarded(a.numel())
        assert lib.flowops_corr_bwd(p(a), p(b), p(go), p(g1), p(g2), B, C, H, W, *params, None, 0, None) == 0
        torch.cuda.synchronize()
        check(g1buf, a.numel(), "corr_bwd generic grad1")
        check(g2buf, a.numel(), "corr_bwd generic grad2")


@pytest.mark.parametrize("shape", [(2, 3, 17, 23), (1, 2, 5, 3), (1, 3, 64, 128), (1, 7, 9, 33)])
@pytest.mark.parametrize("mode", [0, 1])
def test_warp_and_cnorm_stay_in_bounds(lib, shape, mode):
    B, C, H, W = shape
    torch.manual_seed(1)
    img = torch.randn(shape, device="cuda")
    flow = 30 * torch.randn(B, 2, H, W, device="cuda")        # far out-of-range samples: exercises the clamps
    lx, ly = torch.linspace(-1, 1, W).cuda(), torch.linspace(-1, 1, H).cuda()
    n = img.numel()
    obuf, out = guarded(n)
    assert lib.flowops_warp_fwd(p(img), p(flow), p(out), B, C, H, W, mode, p(lx), p(ly), None) == 0
    gout = torch.randn(shape, device="cuda")
    gibuf, gi = guarded(n)
    gfbuf, gf = guarded(flow.numel())
    assert lib.flowops_warp_bwd(p(img), p(flow), p(gout), p(gi), p(gf), B, C, H, W, mode, p(lx), p(ly), None) == 0
    ybuf, y = guarded(B * H * W)
    assert lib.flowops_cnorm_fwd(p(img), p(y), B, C, H, W, None) == 0
    gxbuf, gx = guarded(n)
    gy = torch.randn(B * H * W, device="cuda")
    assert lib.flowops_cnorm_bwd(p(img), p(y), p(gy), p(gx), B, C, H, W, None) == 0
    torch.cuda.synchronize()
    check(obuf, n, "warp_fwd")
    check(gibuf, n, "warp_bwd gimg")
    check(gfbuf, flow.numel(), "warp_bwd gflow")
    check(ybuf, B * H * W, "cnorm_fwd")
    check(gxbuf, n, "cnorm_bwd")


def test_fused_glue_stays_in_bounds(lib):
    B, H, W = 2, 19, 27
    x = torch.randn(B, 6, H, W, device="cuda")
    flow = 20 * torch.randn(B, 2, H, W, device="cuda")
    hw = H * W
    wbuf, warped = guarded(B * 3 * hw)
    nbuf, norm = guarded(B * hw)
    x0 = ctypes.c_void_p(x.data_ptr())
    x1 = ctypes.c_void_p(x.data_ptr() + 4 * 3 * hw)
    assert lib.flowops_warp_diff_norm_fwd(x0, x1, 6 * hw, p(flow), p(warped), 3 * hw, p(norm), hw, B, 3, H, W, None) == 0
    cbuf, conf = guarded(B * hw)
    a, b = x[:, :3].contiguous(), x[:, 3:].contiguous()
    assert lib.flowops_warp_conf_fwd(p(a), p(b), p(flow), p(conf), ctypes.c_float(0.02), B, 3, H, W, 0, None, None, None) == 0
    ybuf, y = guarded(B * 5 * hw)
    y.copy_(torch.randn(B * 5 * hw, device="cuda"))
    bias = torch.randn(5, device="cuda")
    assert lib.flowops_bias_lrelu(p(y), p(bias), B, 5, hw, 0, ctypes.c_float(0.1), None) == 0
    torch.cuda.synchronize()
    check(wbuf, B * 3 * hw, "fused warped")
    check(nbuf, B * hw, "fused norm")
    check(cbuf, B * hw, "conf")
    check(ybuf, B * 5 * hw, "bias_lrelu", must_fill=False)
