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


@pytest.mark.parametrize("shape", [(1, 32, 13, 21), (2, 16, 8, 70), (1, 8, 6, 34)])
def test_correlation_nhwc_store_stays_in_bounds(lib, shape):
    """flowops_corr_planes_from_conv + flowops_corr_fwd_planes_nhwc on ragged shapes: the channels-last destination
    (473 -> 480 channels) is written exactly in channels [32, 473)."""
    B, C, H, W = shape
    torch.manual_seed(3)
    P = (20, 1, 20, 1, 2)
    ya = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    yb = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, device="cuda")
    ws_bytes = lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *P)
    wbuf, ws = guarded((ws_bytes + 3) // 4 + 64)
    ws = ws[(-ws.data_ptr()) % 256 // 4:]
    for which, y in ((0, ya), (1, yb)):
        rc = lib.flowops_corr_planes_from_conv(p(y), p(bias), ctypes.c_float(0.1), None, which, B, C, H, W, *P, p(ws), ws_bytes, None)
        assert rc == 0, lib.flowops_last_error()
    c_dst = 480
    n_out = B * H * W * c_dst
    obuf, out = guarded(n_out)
    rc = lib.flowops_corr_fwd_planes_nhwc(p(out), c_dst, 32, ctypes.c_float(0.1), B, C, H, W, *P, p(ws), ws_bytes, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(obuf, n_out, "corr_fwd_planes_nhwc out", must_fill=False)
    check(wbuf, (ws_bytes + 3) // 4 + 64, "corr workspace", must_fill=False)
    o = out.view(B * H * W, c_dst)
    assert torch.all(o[:, :32] == SENT) and torch.all(o[:, 473:] == SENT), "wrote outside its channel slice"
    assert not torch.any(o[:, 32:473] == SENT), "left part of its channel slice unwritten"
    # values: LeakyReLU(correlation of the activated features)
    act = lambda y: torch.nn.functional.leaky_relu(y + bias.view(1, -1, 1, 1), 0.1)
    from ir2rgb_b200 import functional as F
    want = torch.nn.functional.leaky_relu(F.correlation_forward(act(ya).contiguous(), act(yb).contiguous(), *P), 0.1)
    got = o[:, 32:473].reshape(B, H, W, 441).permute(0, 3, 1, 2)
    assert torch.equal(got, want)


@pytest.mark.parametrize("B,H,W", [(1, 5, 7), (2, 33, 65), (1, 64, 300)])
def test_glue_kernels_stay_in_bounds(lib, B, H, W):
    """flowops_flownet2_prep, flowops_warp_diff_norm_concat_nhwc, flowops_bias_lrelu_nhwc_to, flowops_fill_channels_nhwc."""
    torch.manual_seed(4)
    hw = H * W
    inputs = torch.randn(B, 3, 2, H, W, device="cuda")
    mean = inputs.view(B, 3, -1).mean(-1).contiguous()
    bufs = [guarded(n) for n in (B * 6 * hw, B * 4 * hw, B * 4 * hw, B * 8 * hw)]
    rc = lib.flowops_flownet2_prep(p(inputs), p(mean), ctypes.c_float(255.0), *[p(v) for _, v in bufs], B, H, W, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    for (buf, v), name in zip(bufs, ("x", "xa", "xb", "x8")):
        check(buf, v.numel(), "flownet2_prep " + name)

    x = bufs[0][1].view(B, 6, H, W)
    flow = 5 * torch.randn(B, 2, H, W, device="cuda")
    cbuf, cat = guarded(B * hw * 16)
    rc = lib.flowops_warp_diff_norm_concat_nhwc(p(x), p(flow), ctypes.c_float(20.0), p(cat), 16, B, H, W, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(cbuf, cat.numel(), "warp_diff_norm_concat_nhwc")

    C, c_dst, c_off = 12, 40, 20
    y = torch.randn(B * hw, C, device="cuda")
    bias = torch.randn(C, device="cuda")
    dbuf, dst = guarded(B * hw * c_dst)
    rc = lib.flowops_bias_lrelu_nhwc_to(p(y), p(bias), p(dst), B * hw, C, c_dst, c_off, ctypes.c_float(0.1), p(y), None)
    assert rc == 0, lib.flowops_last_error()
    rc = lib.flowops_fill_channels_nhwc(p(dst), B * hw, c_dst, 35, 5, ctypes.c_float(0.0), None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(dbuf, dst.numel(), "bias_lrelu_nhwc_to / fill_channels_nhwc", must_fill=False)
    d = dst.view(B * hw, c_dst)
    assert torch.all(d[:, :c_off] == SENT) and torch.all(d[:, c_off + C:35] == SENT) and torch.all(d[:, 35:] == 0)
    assert torch.equal(d[:, c_off:c_off + C], y)              # `also` aliased y: y now holds the activated values too


# ---- 16-bit storage entry points and the fusion-network input: sentinel guard bands in the storage type ----
SENT16 = 12345.0          # representable in fp16 and bf16


def guarded16(numel, dt):
    buf = torch.full((numel + 2 * GUARD,), SENT16, device="cuda", dtype=dt)
    return buf, buf[GUARD:GUARD + numel]


def check16(buf, numel, name):
    assert torch.all(buf[:GUARD] == SENT16), name + ": wrote before the buffer"
    assert torch.all(buf[GUARD + numel:] == SENT16), name + ": wrote past the buffer"
    assert not torch.any(buf[GUARD:GUARD + numel] == SENT16), name + ": left part of the output unwritten"


@pytest.mark.parametrize("dt,code", [(torch.float16, 1), (torch.bfloat16, 2)])
@pytest.mark.parametrize("shape", [(2, 3, 17, 23), (1, 2, 8, 16), (1, 1, 5, 7), (3, 3, 33, 40)])
def test_16bit_entry_points_stay_in_bounds(lib, dt, code, shape):
    B, C, H, W = shape
    torch.manual_seed(1)
    x = torch.randn(shape, device="cuda").to(dt)
    flow = (3 * torch.randn(B, 2, H, W, device="cuda")).to(dt)
    ybuf, y = guarded16(B * H * W, dt)
    assert lib.flowops_cnorm_fwd_16(p(x), p(y), B, C, H, W, code, None) == 0, lib.flowops_last_error()
    gy = torch.randn(B, 1, H, W, device="cuda").to(dt)
    gbuf, gx = guarded16(x.numel(), dt)
    assert lib.flowops_cnorm_bwd_16(p(x), p(y), p(gy), p(gx), B, C, H, W, code, None) == 0, lib.flowops_last_error()
    lx = torch.linspace(-1, 1, W).to(dt).float().cuda()
    ly = torch.linspace(-1, 1, H).to(dt).float().cuda()
    for mode in (0, 1):
        obuf, out = guarded16(x.numel(), dt)
        assert lib.flowops_warp_fwd_16(p(x), p(flow), p(out), B, C, H, W, mode, p(lx), p(ly), code, None) == 0, lib.flowops_last_error()
        torch.cuda.synchronize()
        check16(obuf, x.numel(), "warp_fwd_16 mode %d" % mode)
    torch.cuda.synchronize()
    check16(ybuf, B * H * W, "cnorm_fwd_16")
    check16(gbuf, x.numel(), "cnorm_bwd_16")


@pytest.mark.parametrize("dt,code", [(torch.float16, 1), (torch.bfloat16, 2)])
@pytest.mark.parametrize("shape", [(2, 16, 8, 12), (1, 5, 7, 9), (1, 40, 6, 70)])
def test_correlation_16bit_stays_in_bounds(lib, dt, code, shape):
    B, C, H, W = shape
    torch.manual_seed(2)
    a, b = torch.randn(shape, device="cuda").to(dt), torch.randn(shape, device="cuda").to(dt)
    P = (20, 1, 20, 1, 2)
    n_out = B * 441 * H * W
    obuf, out = guarded16(n_out, dt)
    ws_bytes = lib.flowops_corr_fwd_workspace_bytes(B, C, H, W, *P)
    wbuf, ws = guarded((ws_bytes + 3) // 4 + 64)
    ws = ws[(-ws.data_ptr()) % 256 // 4:]
    rc = lib.flowops_corr_fwd_16(p(a), p(b), p(out), B, C, H, W, *P, code, p(ws), ws_bytes, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check16(obuf, n_out, "corr_fwd_16 out")
    check(wbuf, (ws_bytes + 3) // 4 + 64, "corr_fwd_16 workspace", must_fill=False)
    assert lib.flowops_corr_fwd_16(p(a), p(b), p(out), B, C, H, W, 4, 1, 4, 1, 1, code, None, 0, None) == -2     # FlowNetC configuration only


@pytest.mark.parametrize("B,H,W,c_dst", [(2, 8, 12, 16), (1, 36, 52, 12), (3, 4, 4, 16)])
def test_fusion_input_stays_in_bounds(lib, B, H, W, c_dst):
    torch.manual_seed(3)
    x = torch.randn(B, 6, H, W, device="cuda")
    lo1, lo2 = torch.randn(B, 2, H // 4, W // 4, device="cuda"), 50 * torch.randn(B, 2, H // 4, W // 4, device="cuda")
    n = B * H * W * c_dst
    obuf, out = guarded(n)
    off = (-out.data_ptr()) % 16 // 4
    assert off == 0, "guard size keeps the window 16-byte aligned"
    rc = lib.flowops_flownet2_fusion_input_nhwc(p(x), p(lo1), p(lo2), ctypes.c_float(20.0), p(out), c_dst, B, H, W, None)
    assert rc == 0, lib.flowops_last_error()
    torch.cuda.synchronize()
    check(obuf, n, "flownet2_fusion_input_nhwc", must_fill=False)
    v = out.view(B, H, W, c_dst)
    assert not torch.any(v[..., :11] == SENT) and torch.all(v[..., 11:] == 0)
    assert lib.flowops_flownet2_fusion_input_nhwc(p(x), p(lo1), p(lo2), ctypes.c_float(20.0), p(out), c_dst, B, H + 1, W, None) == -1
