"""Convolution with its bias + LeakyReLU epilogue fused, through cuDNN's runtime-fusion engines (cuDNN frontend graph
conv_fprop -> bias -> leaky_relu, fp32 I/O, TF32 math, channels-last).  Library code, like the convolutions themselves:
it removes the separate HBM pass `flowops_bias_lrelu` makes after a bias-free cuDNN convolution
(reference networks/submodules.py:7-38 builds conv -> LeakyReLU for every layer).

Used by ConvAct (networks/submodules.py) for inference on a channels_last body when TF32 convolutions are allowed
(torch.backends.cudnn.allow_tf32 -- the fused engines are tensor-op engines; with TF32 off the unfused, bit-exact path
runs).  One graph per (layer, shapes, strides); all execution plans cuDNN offers are built and timed once, the fastest
is kept.  Measured on B200 for the layers of FlowNet2 at 512x1024, 16 pairs (tools/cudnn_fuse_probe.py): FlowNetFusion
conv0 683 us fused vs 546 + 669 us unfused, FlowNetS conv1 627 vs 963 + 180.
"""
import torch

try:
    import cudnn
except Exception:          # pragma: no cover - the image ships it; anything else keeps the unfused path
    cudnn = None

ENABLED = True
MIN_OUT_ELEMENTS = 1 << 22      # smaller layers: the epilogue pass is launch-latency sized and plan building is not worth it
_handles = {}
_cache = {}
_errors = {}         # key -> why no fused plan could be built
_rejected = {}       # key -> (ms fused, ms two kernels) of the layers where the fused plan lost


def available():
    return ENABLED and cudnn is not None


def _handle(device):
    h = _handles.get(device.index)
    if h is None:
        with torch.cuda.device(device):
            h = _handles[device.index] = cudnn.create_handle()
    return h


class _Plan:
    def __init__(self, x, w, bias, y, stride, padding, dilation, slope, handle, unfused=None, after=None):
        FL = cudnn.data_type.FLOAT
        g = cudnn.pygraph(handle=handle, io_data_type=FL, intermediate_data_type=FL, compute_data_type=FL)
        self.X = g.tensor(name="X", dim=list(x.shape), stride=list(x.stride()), data_type=FL)
        self.W = g.tensor(name="W", dim=list(w.shape), stride=list(w.stride()), data_type=FL)
        n = bias.numel()
        self.B = g.tensor(name="B", dim=[1, n, 1, 1], stride=[n, 1, n, n], data_type=FL)
        c = g.conv_fprop(image=self.X, weight=self.W, padding=list(padding), stride=list(stride), dilation=list(dilation))
        t = g.bias(input=c, bias=self.B)
        o = g.leaky_relu(input=t, negative_slope=float(slope))
        o.set_output(True).set_dim(list(y.shape)).set_stride(list(y.stride())).set_data_type(FL)
        self.Y = o
        g.validate()
        g.build_operation_graph()
        g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
        g.check_support()
        g.build_plans(cudnn.build_plan_policy.ALL)
        self.g = g
        self.index, self.ws, self.ms = self._pick(x, w, bias, y, handle)
        # trust, but verify: the chosen plan must reproduce the two-kernel result on this very layer (TF32 engines differ
        # from each other by ~1e-3 of the largest output; a layout or stride misunderstanding shows up as O(1))
        self.run(x, w, bias, y, handle)
        want = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, w, bias, tuple(stride), tuple(padding), tuple(dilation)),
                                              float(slope))
        err = ((y - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()
        if not err <= 1e-2:
            raise RuntimeError("fused plan disagrees with the unfused convolution (max-relative %.3g)" % err)
        # a fused engine is not always faster than cuDNN's best plain convolution followed by the epilogue pass (the
        # plain convolution has more algorithms to choose from): keep the fused plan only where it wins
        self.ms_unfused = _time3(unfused) if unfused is not None else None
        if after is not None:                          # work the fused path needs on top (copy into a concat buffer)
            self.ms = _time3(lambda: (self.run(x, w, bias, y, handle), after(y)))
        self.wins = self.ms_unfused is None or self.ms < 0.97 * self.ms_unfused

    def _pick(self, x, w, bias, y, handle):
        """Time every plan (3 launches each after one warm-up) and keep the fastest."""
        pack = {self.X: x, self.W: w, self.B: bias, self.Y: y}
        best = None
        for i in range(self.g.get_execution_plan_count()):
            try:
                ws = torch.empty(max(self.g.get_workspace_size_plan_at_index(i), 1), device=x.device, dtype=torch.uint8)
                ms = _time3(lambda: self.g.execute_plan_at_index(pack, ws, i, handle=handle))
            except Exception:
                continue
            if best is None or ms < best[0]:
                best = (ms, i, ws)
        if best is None:
            raise RuntimeError("no executable plan")
        return best[1], best[2], best[0]

    def run(self, x, w, bias, y, handle):
        self.g.execute_plan_at_index({self.X: x, self.W: w, self.B: bias, self.Y: y}, self.ws, self.index, handle=handle)


def _time3(fn):
    """Milliseconds for 3 launches of fn after two warm-up calls (the first may autotune), best of two rounds."""
    fn()
    fn()
    best = None
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def conv_bias_lrelu(conv, x, w, slope, out=None, unfused=None, after=None):
    """LeakyReLU(conv2d(x, w) + conv.bias) in one cuDNN launch, or NotImplemented (caller runs the unfused path).
    x: dense channels_last fp32; w: the (possibly zero-padded) weight the unfused path would use; out: an optional
    preallocated destination (may be a channel slice of a channels_last concat buffer); unfused: a callable running the
    two-kernel path for the same layer -- timed once against the best fused plan (plus `after(y)`, extra work only the
    fused path needs), the faster of the two is kept.  `after` itself is run by the caller."""
    if not available() or conv.groups != 1 or conv.padding_mode != "zeros" or isinstance(conv.padding, str):
        return NotImplemented
    B, _, H, W = x.shape
    kh, kw = conv.kernel_size
    oh = (H + 2 * conv.padding[0] - conv.dilation[0] * (kh - 1) - 1) // conv.stride[0] + 1
    ow = (W + 2 * conv.padding[1] - conv.dilation[1] * (kw - 1) - 1) // conv.stride[1] + 1
    if B * conv.out_channels * oh * ow < MIN_OUT_ELEMENTS:
        return NotImplemented
    y = out if out is not None else torch.empty((B, conv.out_channels, oh, ow), device=x.device, dtype=torch.float32,
                                                memory_format=torch.channels_last)
    if y.data_ptr() % 16 or x.data_ptr() % 16 or w.data_ptr() % 16:
        return NotImplemented
    key = (x.device.index, tuple(x.shape), tuple(x.stride()), tuple(w.shape), tuple(w.stride()), tuple(y.shape), tuple(y.stride()),
           tuple(conv.stride), tuple(conv.padding), tuple(conv.dilation), float(slope))
    handle = _handle(x.device)
    cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream(x.device).cuda_stream)
    plan = _cache.get(key)
    if plan is None:
        if torch.cuda.is_current_stream_capturing():
            return NotImplemented                      # plans are built (and timed) during the warm-up calls only
        try:
            plan = _Plan(x, w, conv.bias, y, conv.stride, conv.padding, conv.dilation, slope, handle, unfused, after)
            if not plan.wins:
                _rejected[key] = (plan.ms, plan.ms_unfused)
                plan = False
        except Exception as e:
            plan = False                               # no engine for this layer: remember, use the unfused path
            _errors[key] = repr(e)[:200]
        _cache[key] = plan
    if plan is False:
        return NotImplemented
    plan.run(x, w, conv.bias, y, handle)
    return y


def report():
    """One record per layer seen so far: shapes, the best fused plan's time, the two-kernel time it was measured against, and
    which of the two runs (tools/step_probe.py prints it)."""
    out = []
    for key, plan in _cache.items():
        rec = {"x": list(key[1]), "w": list(key[3]), "y": list(key[5]), "stride": list(key[7]), "slope": key[10],
               "fused": bool(plan)}
        if key in _rejected:
            rec.update(us_fused=round(_rejected[key][0] / 3 * 1e3, 1), us_two_kernels=round(_rejected[key][1] / 3 * 1e3, 1))
        if key in _errors:
            rec["error"] = _errors[key]
        if plan:
            rec.update(us_fused=round(plan.ms / 3 * 1e3, 1), us_two_kernels=None if plan.ms_unfused is None else round(plan.ms_unfused / 3 * 1e3, 1))
        out.append(rec)
    return out


def reset():
    _cache.clear()
    _rejected.clear()
    _errors.clear()

