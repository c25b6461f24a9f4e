"""Debugging aid for csrc/corr_tc_bwd.cu: calls the C ABI with its own workspace and inspects the intermediate buffers."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ir2rgb_b200 import _lib, functional as F  # noqa: E402
from ir2rgb_b200.functional import _p, _stream, check  # noqa: E402

P = (20, 1, 20, 1, 2)


def run(shape, a=None, b=None, go=None, flags=1):
    lib = _lib.load()
    B, C, H, W = shape
    a = torch.randn(*shape, device="cuda") if a is None else a
    b = torch.randn(*shape, device="cuda") if b is None else b
    go = torch.randn(B, 441, H, W, device="cuda") if go is None else go
    lib.flowops_corr_set_impl(0)
    r1, r2 = F.correlation_backward(a, b, go, *P)
    lib.flowops_corr_set_impl(flags)
    nbytes = lib.flowops_corr_bwd_workspace_bytes(B, C, H, W, *P)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    g1 = torch.full_like(a, 7.0)
    g2 = torch.full_like(a, 7.0)
    check(lib.flowops_corr_bwd(_p(a), _p(b), _p(go), _p(g1), _p(g2), B, C, H, W, *P, _p(ws), nbytes, _stream()), "corr_bwd")
    torch.cuda.synchronize()
    lib.flowops_corr_set_impl(1)
    plane_bytes = 4 * B * C * H * W
    rec_bytes = (4 * B * H * W * 588 + 255) // 256 * 256
    wsf = ws.view(torch.float32)
    PH, PW = H // 2, W // 2
    P1 = wsf[:plane_bytes // 4].view(B, 4, C // 8, PH, PW, 8)
    P2 = wsf[plane_bytes // 4: 2 * plane_bytes // 4].view(B, 4, C // 8, PH, PW, 8)
    G1 = wsf[2 * plane_bytes // 4: 2 * plane_bytes // 4 + B * H * W * 588].view(B, 4, PH, PW, 21, 28)
    o2 = 2 * plane_bytes // 4 + rec_bytes // 4
    G2 = wsf[o2: o2 + B * H * W * 588].view(B, 4, PH, PW, 21, 28)
    # expected planes
    def planes(x):
        return torch.stack([x[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], 1).view(B, 4, C // 8, 8, PH, PW).permute(0, 1, 2, 4, 5, 3)
    print("shape", shape, "flags", flags)
    print(" P1 ok", torch.equal(P1, planes(a)), " P2 ok", torch.equal(P2, planes(b)))
    gop = torch.stack([go[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], 1).view(B, 4, 21, 21, PH, PW)   # [B,4,tj,ti,Y,X]
    e1 = torch.zeros(B, 4, PH, PW, 21, 28, device="cuda")
    e2 = torch.zeros(B, 4, PH, PW, 21, 28, device="cuda")
    gpad = torch.nn.functional.pad(gop, (10, 10, 10, 10))          # plane coords + 10
    for X in range(PW):
        cx = X % 8
        e1[:, :, :, X, :, cx:cx + 21] = gop[:, :, :, :, :, X].permute(0, 1, 4, 2, 3)
        for tj in range(21):
            for ti in range(21):
                # gO[(20-tj, 20-ti)][Y + tj - 10][X + ti - 10]
                e2[:, :, :, X, tj, cx + ti] = gpad[:, :, 20 - tj, 20 - ti, tj:tj + PH, X + ti]
    print(" G1 ok", torch.equal(G1, e1), " G2 ok", torch.equal(G2, e2), " G1 nonzero", int(torch.count_nonzero(G1)), int(torch.count_nonzero(e1)))
    for name, g, r in (("g1", g1, r1), ("g2", g2, r2)):
        err = ((g.double() - r.double()).abs().max() / r.double().abs().max()).item()
        print(" %s: maxrel %.3e  zeros %d  sevens %d  nan %d  of %d   |g|max %.4f |r|max %.4f" % (
            name, err, int((g == 0).sum()), int((g == 7).sum()), int(torch.isnan(g).sum()), g.numel(), g.abs().max().item(), r.abs().max().item()))
    return g1, g2, r1, r2


def pattern_tests():
    lib = _lib.load()
    torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
    B, C, H, W = 1, 32, 16, 16
    a = torch.zeros(B, C, H, W, device="cuda")
    for name, tj, ti in (("centre", 10, 10), ("shift(+2,+3)", 12, 13)):
        go = torch.zeros(B, 441, H, W, device="cuda")
        go[:, tj * 21 + ti] = 1.0
        # test 1: channel pattern
        b = (torch.arange(C, device="cuda").float() + 1).view(1, C, 1, 1).expand(B, C, H, W).contiguous()
        lib.flowops_corr_set_impl(5)          # single TF32 product: small integers are exact
        g1, _ = F.correlation_backward(a, b, go, *P, need2=False)
        lib.flowops_corr_set_impl(0)
        r1, _ = F.correlation_backward(a, b, go, *P, need2=False)
        print("==", name, "channel pattern: g1[0,:,5,5]*C =", (g1[0, :, 5, 5] * C).tolist())
        print("   expected                              =", (r1[0, :, 5, 5] * C).tolist())
        print("   g1[0,3]*C grid\n", g1[0, 3] * C)
        # test 2: position pattern
        yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
        b = (100 * yy + xx).float().view(1, 1, H, W).expand(B, C, H, W).contiguous()
        lib.flowops_corr_set_impl(5)
        g1, _ = F.correlation_backward(a, b, go, *P, need2=False)
        lib.flowops_corr_set_impl(0)
        r1, _ = F.correlation_backward(a, b, go, *P, need2=False)
        print("== position pattern g1[0,0]*C\n", g1[0, 0] * C)
        print("   expected\n", r1[0, 0] * C)
        print("   channel 7 equal to channel 0:", torch.equal(g1[0, 7], g1[0, 0]))
    lib.flowops_corr_set_impl(1)


if __name__ == "__main__":
    torch.manual_seed(0)
    pattern_tests()
