"""Per-operator timing on one GPU: new kernels vs the reference's rebuilt extensions (oracle/_ref),
with roofline fractions.  CUDA events on the current stream after warm-up; L2-cold by rotating among
buffer sets larger than L2, plus single-launch-with-L2-flush and L2-warm figures.

    python tools/opbench.py [--iters 20] [--json gpurun_out/opbench.json] [--skip-ref]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ir2rgb_b200 import functional as F  # noqa: E402


def peaks():
    """HBM copy bandwidth and the dense TF32 tensor peak.  MEASURED_PEAKS.json holds a measured bf16 GEMM rate; TF32 runs
    at half the bf16 rate on the same pipe, so bf16 / 2 is the TF32 proxy (burst figure: kernels here are timed alone)."""
    p = {"hbm_gbs": 6650.0, "tf32_tflops": 1590.0 / 2, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        m = json.load(open(path))
        p = {"hbm_gbs": float(m["hbm_gbs"]), "tf32_tflops": float(m["bf16_tflops"]) / 2,
             "tf32_tflops_sustained": float(m.get("bf16_tflops_sustained", m["bf16_tflops"])) / 2, "source": "measured"}
    return p


class L2Flusher:
    def __init__(self):
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def __call__(self):
        self.buf.zero_()


def time_op(fn, iters, warmup=3, flush=None):
    """One launch per measurement, L2 flushed in between (includes ~2-3 us of event/launch gap)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    best = 1e30
    for _ in range(iters):
        if flush is not None:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        t = e0.elapsed_time(e1)
        total += t
        best = min(best, t)
    return total / iters * 1e3, best * 1e3      # microseconds (mean, best)


def time_rotating(fns, launches=100, warmup=1):
    """CUDA events around `launches` back-to-back launches that rotate among buffer sets whose
    combined footprint exceeds L2, so every launch finds its inputs in HBM and the launch gap is amortised."""
    n = len(fns)
    for _ in range(warmup):
        for f in fns:
            f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(launches):
        fns[i % n]()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / launches * 1e3


def run(iters=20, skip_ref=False, small=False, quiet=False, ffma=None, only=None):
    torch.manual_seed(0)
    pk = peaks()
    flush = L2Flusher()
    ffma = ffma or F.ffma_peak_tflops()
    res = {"peaks": dict(pk, ffma_tflops=ffma, ffma_nominal_tflops=148 * 128 * 2 * 1.965e9 / 1e12), "ops": {}}
    ref = None
    if not skip_ref:
        from oracle import ref_ext
        if ref_ext.available():
            ref = ref_ext

    def add(name, fn, args, ref_fn=None, alg_bytes=None, alg_flop=None, ref_iters=None, tensor_flop=None):
        """fn(*args) is the new operator, ref_fn(*args) the reference's.  Primary timing: back-to-back
        launches rotating among clones of `args` whose combined footprint exceeds 2x L2 (L2-cold, launch gap
        amortised); also reported: one launch per event pair with an L2 flush in between, and L2-warm."""
        if only and only not in name:
            return
        foot = sum(t.numel() * t.element_size() for t in args)
        nsets = int(min(64, max(2, -(-(300 << 20) // max(foot, 1)))))
        sets = [args] + [tuple(t.clone() for t in args) for _ in range(nsets - 1)]
        launches = 20 if small else 100
        mean = time_rotating([(lambda s=s_: fn(*s)) for s_ in sets], launches=launches)
        single, best = time_op(lambda: fn(*args), iters, flush=flush)
        warm = time_rotating([lambda: fn(*args)], launches=launches)
        del sets
        e = {"us": mean, "us_single_flushed": single, "us_l2_warm": warm, "rotating_sets": nsets, "launches": launches}
        if alg_bytes is not None:
            e.update(bound="hbm", alg_bytes=alg_bytes, achieved_gbs=alg_bytes / mean / 1e3,
                     frac=alg_bytes / mean / 1e3 / pk["hbm_gbs"], frac_of_8TBs=alg_bytes / mean / 1e3 / 8000.0)
        if alg_flop is not None and tensor_flop is None:
            e.update(bound="fp32", alg_flop=alg_flop, achieved_tflops=alg_flop / mean / 1e6,
                     frac=alg_flop / mean / 1e6 / ffma)
        if tensor_flop is not None:
            # tcgen05 kernel: `achieved` counts the ALGORITHMIC (useful) flops, 2*B*H*W*441*C; the tensor pipe executes
            # tensor_flop = 3 (3xTF32) x 1024/441 (dense 128 x 256 tiles around a banded contraction) times that
            e.update(bound="tensor", alg_flop=alg_flop, achieved_tflops=alg_flop / mean / 1e6,
                     frac=alg_flop / mean / 1e6 / pk["tf32_tflops"], executed_tflops=tensor_flop / mean / 1e6,
                     frac_executed=tensor_flop / mean / 1e6 / pk["tf32_tflops"],
                     x_fp32_pipe_peak=alg_flop / mean / 1e6 / ffma)
        if ref is not None and ref_fn is not None:
            rmean, rbest = time_op(lambda: ref_fn(*args), ref_iters or max(3, iters // 4), warmup=1, flush=flush)
            e.update(ref_us=rmean, speedup_vs_ref=rmean / single)
        res["ops"][name] = e
        if not quiet:
            print(name, json.dumps(e), flush=True)

    # ---- C2: Correlation 8x256x48x64 ----
    B, C, H, W = (2, 64, 24, 32) if small else (8, 256, 48, 64)
    P = (20, 1, 20, 1, 2)
    a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    go = torch.randn(B, 441, H, W, device="cuda")
    flop = 2.0 * B * H * W * 441 * C
    from ir2rgb_b200 import _lib
    lib = _lib.load()
    tc_on = bool(lib.flowops_corr_get_impl() & 1)
    tc_x = 3.0 * 1024.0 / 441.0 if tc_on else None       # executed / useful flops of the tensor-core kernel

    def with_impl(flags, f):
        def g(*x):
            prev = lib.flowops_corr_get_impl()
            lib.flowops_corr_set_impl(flags)
            try:
                return f(*x)
            finally:
                lib.flowops_corr_set_impl(prev)
        return g
    add("corr_fwd_c2", lambda a, b: F.correlation_forward(a, b, *P), (a, b),
        (lambda a, b: ref.correlation_forward(a, b, *P)) if ref else None, alg_flop=flop,
        tensor_flop=flop * tc_x if tc_on else None)
    if tc_on:
        add("corr_fwd_c2_fp32fma", with_impl(0, lambda a, b: F.correlation_forward(a, b, *P)), (a, b), None, alg_flop=flop)
    # backward on the tensor cores (csrc/corr_tc_bwd.cu): K = the 36 x 28 window (1008 positions, 441 of them inside the band)
    tcb_x = 3.0 * 1008.0 / 441.0 if tc_on else None
    add("corr_bwd_c2", lambda a, b, go: F.correlation_backward(a, b, go, *P), (a, b, go),
        (lambda a, b, go: ref.correlation_backward(a, b, go, *P)) if ref else None, alg_flop=2 * flop, ref_iters=2,
        tensor_flop=2 * flop * tcb_x if tc_on else None)
    if tc_on:
        add("corr_bwd_c2_fp32fma", with_impl(0, lambda a, b, go: F.correlation_backward(a, b, go, *P)), (a, b, go), None,
            alg_flop=2 * flop)
    # ---- C4-shaped correlation (FlowNet2 at 512x1024, per-GPU batch 8) ----
    if not small:
        a4, b4 = torch.randn(8, 256, 64, 128, device="cuda"), torch.randn(8, 256, 64, 128, device="cuda")
        add("corr_fwd_c4_b8", lambda a, b: F.correlation_forward(a, b, *P), (a4, b4),
            (lambda a, b: ref.correlation_forward(a, b, *P)) if ref else None,
            alg_flop=2.0 * 8 * 64 * 128 * 441 * 256, ref_iters=2,
            tensor_flop=2.0 * 8 * 64 * 128 * 441 * 256 * tc_x if tc_on else None)
        if tc_on:
            add("corr_fwd_c4_b8_fp32fma", with_impl(0, lambda a, b: F.correlation_forward(a, b, *P)), (a4, b4), None,
                alg_flop=2.0 * 8 * 64 * 128 * 441 * 256)
        go4 = torch.randn(8, 441, 64, 128, device="cuda")
        add("corr_bwd_c4_b8", lambda a, b, go: F.correlation_backward(a, b, go, *P), (a4, b4, go4), None,
            alg_flop=4.0 * 8 * 64 * 128 * 441 * 256, tensor_flop=4.0 * 8 * 64 * 128 * 441 * 256 * tcb_x if tc_on else None)
        if tc_on:
            add("corr_bwd_c4_b8_fp32fma", with_impl(0, lambda a, b, go: F.correlation_backward(a, b, go, *P)), (a4, b4, go4), None,
                alg_flop=4.0 * 8 * 64 * 128 * 441 * 256)
        del a4, b4, go4

    # ---- C3: Resample2d + ChannelNorm on 16x3x512x1024 ----
    B, H, W = (2, 128, 256) if small else (16, 512, 1024)
    plane = B * H * W * 4
    img = 2 * torch.rand(B, 3, H, W, device="cuda") - 1
    gout = torch.randn(B, 3, H, W, device="cuda")
    low = 20 * torch.randn(B, 2, H // 4, W // 4, device="cuda")
    coarse = 20 * torch.randn(B, 2, max(H // 64, 2), max(W // 64, 2), device="cuda")
    flows = {
        # realistic optical flow: smooth field, |flow| ~ 20 px, gradient well below 1 px/px
        "smooth": torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False).contiguous(),
        "randn": 4 * torch.randn(B, 2, H, W, device="cuda"),
        "bilinear": torch.nn.functional.interpolate(low, scale_factor=4, mode="bilinear").contiguous(),
        "nearest": torch.nn.functional.interpolate(low, scale_factor=4, mode="nearest").contiguous(),
    }
    R2D, GS = F.WARP_RESAMPLE2D, F.WARP_GRIDSAMPLE
    for fl, flow in flows.items():
        add("resample2d_fwd_" + fl, lambda i, f: F.warp_forward(i, f, R2D), (img, flow),
            (lambda i, f: ref.resample2d_forward(i, f)) if ref else None, alg_bytes=8 * plane)
        def _fast(i, f):
            with F.warp_tolerance_mode(True):
                return F.warp_forward(i, f, R2D)
        # tolerance mode (fp32 bilinear weights; what FlowNet runs): ~1e-7 from the reference kernel instead of bit-exact
        add("resample2d_fwd_fp32blend_" + fl, _fast, (img, flow), None, alg_bytes=8 * plane)
        add("resample2d_bwd_" + fl, lambda i, f, g: F.warp_backward(i, f, g, True, True, R2D), (img, flow, gout),
            (lambda i, f, g: ref.resample2d_backward(i, f, g)) if ref else None, alg_bytes=13 * plane)
        add("resample2d_bwd_flowonly_" + fl, lambda i, f, g: F.warp_backward(i, f, g, False, True, R2D), (img, flow, gout),
            None, alg_bytes=10 * plane)
    flow = flows["bilinear"]
    from oracle import torch_ref as _tr
    add("gridsample_fwd_bilinear", lambda i, f: F.warp_forward(i, f, GS), (img, flow),
        lambda i, f: _tr.networks_resample(i, f), alg_bytes=8 * plane)
    add("gridsample_fwd_smooth", lambda i, f: F.warp_forward(i, f, GS), (img, flows["smooth"]),
        lambda i, f: _tr.networks_resample(i, f), alg_bytes=8 * plane)
    add("gridsample_bwd_smooth", lambda i, f, g: F.warp_backward(i, f, g, True, True, GS), (img, flows["smooth"], gout),
        None, alg_bytes=13 * plane)
    # ---- C1: vid2vid generator flow-warp, 1x3x256x512 (BASELINE configs[0], the reference's CPU-runnable case) ----
    if not small:
        torch.manual_seed(0)
        img1 = torch.randn(1, 3, 256, 512)
        flow1 = 5 * torch.randn(1, 2, 256, 512)
        img1c, flow1c = img1.cuda(), flow1.cuda()
        gout1 = torch.randn(1, 3, 256, 512, device="cuda")
        add("gridsample_fwd_c1", lambda i, f: F.warp_forward(i, f, GS), (img1c, flow1c), None, alg_bytes=8 * 256 * 512 * 4)
        add("gridsample_bwd_c1", lambda i, f, g: F.warp_backward(i, f, g, True, True, GS), (img1c, flow1c, gout1), None,
            alg_bytes=13 * 256 * 512 * 4)
        import time as _time
        for _ in range(3):
            _tr.networks_resample(img1, flow1)
        t0 = _time.perf_counter()
        for _ in range(20):
            _tr.networks_resample(img1, flow1)
        if "gridsample_fwd_c1" in res["ops"]:
            res["ops"]["gridsample_fwd_c1"]["cpu_reference_us"] = (_time.perf_counter() - t0) / 20 * 1e6
            res["ops"]["gridsample_fwd_c1"]["cpu_threads"] = torch.get_num_threads()
    y = F.channelnorm_forward(img)
    gy = torch.randn_like(y)
    add("cnorm_fwd_c3", lambda x: F.channelnorm_forward(x), (img,),
        (lambda x: ref.channelnorm_forward(x)) if ref else None, alg_bytes=4 * plane)
    add("cnorm_bwd_c3", lambda x, y, gy: F.channelnorm_backward(x, y, gy), (img, y, gy),
        (lambda x, y, gy: ref.channelnorm_backward(x, y, gy)) if ref else None, alg_bytes=8 * plane)
    # ---- 16-bit storage variants: same shapes, half the algorithmic bytes; "ref" is the chain the
    # reference's fp16 mode runs on the same tensors (casts around its fp32 extension, or its at::Half kernels) ----
    img16, gout16 = img.half(), gout.half()
    y16 = F.channelnorm_forward(img16)
    gy16 = gy.half()
    add("cnorm_fwd_c3_f16", lambda x: F.channelnorm_forward(x), (img16,),
        (lambda x: ref.channelnorm_forward(x)) if ref else None, alg_bytes=2 * plane)
    add("cnorm_bwd_c3_f16", lambda x, y, g: F.channelnorm_backward(x, y, g), (img16, y16, gy16),
        (lambda x, y, g: ref.channelnorm_backward(x, y, g)) if ref else None, alg_bytes=4 * plane)
    for fl in ("smooth", "nearest"):
        flow16 = flows[fl].half()
        add("resample2d_fwd_%s_f16" % fl, lambda i, f: F.warp_forward(i, f, R2D), (img16, flow16),
            (lambda i, f: ref.resample2d_forward(i.float(), f.float()).half()) if ref else None, alg_bytes=4 * plane)

    def _chain16(i, f):      # models/base_model.py:123-136 with opt['fp16'], torch ops
        b, c, h, w = i.shape
        grid = torch.cat([torch.linspace(-1.0, 1.0, w, device=i.device).view(1, 1, 1, w).expand(b, 1, h, w),
                          torch.linspace(-1.0, 1.0, h, device=i.device).view(1, 1, h, 1).expand(b, 1, h, w)], 1).to(f.dtype)
        nf = torch.cat([f[:, 0:1] / ((w - 1.0) / 2.0), f[:, 1:2] / ((h - 1.0) / 2.0)], dim=1)
        return torch.nn.functional.grid_sample(i.float(), (grid + nf).permute(0, 2, 3, 1).float(), mode="bilinear",
                                               padding_mode="border", align_corners=False).half()
    add("gridsample_fwd_smooth_f16", lambda i, f: F.warp_forward(i, f, GS), (img16, flows["smooth"].half()), _chain16,
        alg_bytes=4 * plane)
    del img16, gout16, y16, gy16
    # ---- FlowNet2 glue of the fusion network at 16 pairs: the full-resolution flow head and the
    # depth-to-space epilogue of deconv0 with the flow upsampler folded in ----
    if not small:
        xh = torch.randn(16, 16, 512, 1024, device="cuda").contiguous(memory_format=torch.channels_last)
        wh = F.pack_flow_head_weight(torch.randn(2, 16, 3, 3, device="cuda"), 16)
        bh = torch.randn(2, device="cuda")
        add("flow_head_c16_fullres", lambda x, w, b: F.flow_head(x, w, b), (xh, wh, bh), None, alg_bytes=xh.numel() * 4 + 16 * 512 * 1024 * 8)
        del xh
        y4 = torch.randn(16, 64, 256, 512, device="cuda").contiguous(memory_format=torch.channels_last)
        fl = torch.randn(16, 2, 256, 512, device="cuda").contiguous(memory_format=torch.channels_last)
        cb = F.ConcatBuffer(y4, 82, 8, shape=(16, 512, 1024))
        b16c, fw, fb = torch.randn(16, device="cuda"), torch.randn(2, 2, 4, 4, device="cuda"), torch.randn(2, device="cuda")
        # bytes: y4 read + 16 + 2 (+ 6 pad) channels written per output pixel + the low-resolution flow
        add("d2s_flowup_deconv0", lambda y, f: cb.bias_lrelu_d2s_in(y, b16c, 0.1, 64, (f, fw, fb)), (y4, fl), None,
            alg_bytes=y4.numel() * 4 + 16 * 512 * 1024 * 24 * 4 + fl.numel() * 4)
        del y4, fl, cb
    B, C, H, W = (2, 64, 24, 32) if small else (8, 256, 48, 64)
    a16, b16 = torch.randn(B, C, H, W, device="cuda").half(), torch.randn(B, C, H, W, device="cuda").half()
    add("corr_fwd_c2_f16", lambda a, b: F.correlation_forward(a, b, *P), (a16, b16),
        (lambda a, b: ref.correlation_forward(a.float(), b.float(), *P).half()) if ref else None,
        alg_flop=2.0 * B * H * W * 441 * C)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--only", default=None, help="run only the rows whose name contains this substring")
    args = ap.parse_args()
    out = run(args.iters, args.skip_ref, args.small, only=args.only)
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)
