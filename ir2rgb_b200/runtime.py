"""CUDA-graph execution of the per-frame flow path.

FlowNet2 inference launches ~450 kernels per micro-batch; at small per-GPU batches (8 pairs per rank when
the batch of 64 is sharded over 8 GPUs) the Python/launch overhead is as long as the GPU work.  Every
libflowops entry point is asynchronous, allocation-free and capturable, so the whole `FlowNet.forward`
can be captured once per input shape and replayed.
"""
import torch


class GraphedFlowNet(torch.nn.Module):
    """Wraps a `FlowNet` (models/flownet.py API): same call signature and results, replayed from a CUDA graph.

    One graph per (shape, dtype) of the inputs.  Outputs are copies, so they stay valid across calls.

    The graphs hold device pointers: weights updated IN PLACE (``load_state_dict``, optimizer steps) are picked up by the
    next replay, but tensors derived from the weights on the host side are not -- the zero-padded weight copies that the
    inference path of the conv body caches (networks/submodules.py ``padded_weight``).  Call ``reset()`` after loading
    or changing weights.
    """

    accepts_host_inputs = True      # pinned host frames are copied straight into the graph's static inputs

    def __init__(self, flownet, warmup=3):
        super().__init__()
        self.net = flownet
        self.warmup = warmup
        self._graphs = {}

    def _capture(self, a, b):
        dev = next(self.net.parameters()).device
        sa, sb = torch.empty_like(a, device=dev), torch.empty_like(b, device=dev)
        sa.copy_(a)
        sb.copy_(b)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):          # cuDNN autotuning and lazy initialisation happen here
                self.net(sa, sb)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.net(sa, sb)
        return g, sa, sb, out

    def reset(self):
        """Drop the captured graphs (and the padded-weight caches they point into); the next call captures again."""
        self._graphs.clear()
        for m in self.net.modules():
            for k in ("_flowops_wpad", "_flowops_wdense", "_flowops_cbuf", "_flowops_conv1_s2d", "_flowops_wconv3", "_flowops_whead"):
                m.__dict__.pop(k, None)
            if isinstance(m.__dict__.get("_sd_warm"), set):
                m._sd_warm.clear()                     # FlowNet2: the next forward builds its plans serially again

    @torch.no_grad()
    def forward(self, input_A, input_B):
        key = (tuple(input_A.shape), input_A.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._graphs[key] = self._capture(input_A, input_B)
        g, sa, sb, out = entry
        sa.copy_(input_A, non_blocking=True)       # device->device, or pinned host->device
        sb.copy_(input_B, non_blocking=True)
        g.replay()
        return tuple(o.clone() for o in out)


class HostPipeline:
    """Runs `net(a, b) -> (flow, conf)` over a batch that lives in PINNED HOST memory, micro-batch by micro-batch,
    with the host->device copy of micro-batch i+1 and the device->host copy of the results of micro-batch i-1
    overlapping the computation of micro-batch i (three streams, double-buffered staging).

    Across calls: `next_inputs=(host_a, host_b)` starts the copy of the NEXT call's first micro-batch behind this call's
    last computation, and `wait=False` lets this call's last device->host copy run under the next call's computation
    (the caller synchronises -- `synchronize()` -- before it reads the host outputs).  With both, a stream of batches
    keeps the GPU busy across the batch boundaries; at 8 pairs per rank (a batch of 64 over 8 GPUs) the un-hidden first
    copy-in and last copy-out were 11 % of the step."""

    def __init__(self, net, device):
        self.net, self.device = net, device
        self.copy_in = torch.cuda.Stream(device=device)
        self.copy_out = torch.cuda.Stream(device=device)
        self._staging = {}
        self._count = 0            # micro-batches issued so far: staging slot = count & 1
        self._prefetched = None    # (key of the host slices, slot) of a copy-in already issued for the next call

    def _buffers(self, shape, dtype):
        key = (tuple(shape), dtype)
        if key not in self._staging:
            mk = lambda: torch.empty(shape, dtype=dtype, device=self.device)
            self._staging[key] = ([mk(), mk()], [mk(), mk()],
                                  [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()])
            # the caching allocator may hand out blocks that earlier kernels on the compute stream still use:
            # order the copy stream after everything already queued there before it first writes the new buffers
            self.copy_in.wait_stream(torch.cuda.current_stream(self.device))
        return self._staging[key]

    @staticmethod
    def _key(host_a, host_b, s, n):
        return (host_a.data_ptr(), host_b.data_ptr(), s, n, tuple(host_a.shape[1:]), host_a.dtype)

    def _copy_in(self, host_a, host_b, s, n, k):
        sa, sb, ready, free = self._buffers((n,) + tuple(host_a.shape[1:]), host_a.dtype)
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(free[k])                   # staging slot k was consumed two micro-batches ago
            sa[k].copy_(host_a[s:s + n], non_blocking=True)
            sb[k].copy_(host_b[s:s + n], non_blocking=True)
            ready[k].record(self.copy_in)

    def synchronize(self):
        """Wait until every result copied out so far has landed in host memory."""
        self.copy_out.synchronize()

    @torch.no_grad()
    def __call__(self, host_a, host_b, micro_batch, out_flow, out_conf, next_inputs=None, wait=True):
        compute = torch.cuda.current_stream(self.device)
        B = host_a.shape[0]
        last = None
        for s in range(0, B, micro_batch):
            n = min(micro_batch, B - s)
            sa, sb, ready, free = self._buffers((n,) + tuple(host_a.shape[1:]), host_a.dtype)
            k = self._count & 1
            if self._prefetched == (self._key(host_a, host_b, s, n), k):
                self._prefetched = None                            # this copy-in was issued by the previous call
            else:
                self._prefetched = None
                self._copy_in(host_a, host_b, s, n, k)
            self._count += 1
            compute.wait_event(ready[k])
            flow, conf = self.net(sa[k], sb[k])
            free[k].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            if s + n >= B and next_inputs is not None:
                # the next call's first micro-batch goes to the other staging slot while this one is being computed
                na, nb = next_inputs
                nn_ = min(micro_batch, na.shape[0])
                self._copy_in(na, nb, 0, nn_, self._count & 1)
                self._prefetched = (self._key(na, nb, 0, nn_), self._count & 1)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(done)
                out_flow[s:s + n].copy_(flow, non_blocking=True)
                out_conf[s:s + n].copy_(conf, non_blocking=True)
            flow.record_stream(self.copy_out)
            conf.record_stream(self.copy_out)
            last = flow
        if wait:
            compute.wait_stream(self.copy_out)
        return last
