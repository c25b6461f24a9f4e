#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: FlowNet2 frame pairs/s at 512x1024, batch 64 sharded over
N B200s, with the hot-path operators' roofline fractions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the flow hot path over one batch of synthetic frame pairs: config[3] of
BASELINE.json -- FlowNet2 (C+S+SD+fusion) forward + confidence mask through the public `FlowNet` wrapper
(models/flownet.py API) on 2*rand-1 frames of 512x1024, random-init weights, global batch 64 split into
contiguous chunks of 64/N per rank (strong scaling, no collective on the data path), processed in
micro-batches.  `value` times the step with the frames already in HBM; `e2e` times the same call with
the frames in pinned host memory (H2D of both frame batches and D2H of flow + confidence inside the
timed region).  Rank 0 prints ONE JSON line.

--impl reference runs the same workload with the reference's own operators: the CUDA extensions of
/root/reference rebuilt for sm_100 (oracle/_ref, see oracle/build_ref.py) plugged into the same FlowNet2
conv body with the reference's unfused glue; if they cannot be loaded it falls back to the pure-PyTorch
CPU port on the host cores (a bounded sample).  Nothing of ir2rgb_b200's native code runs in that arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 512, 1024
GLOBAL_BATCH = 64
METRIC = "flownet2_frame_pairs_per_s"


# ---------------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--micro-batch", type=int, default=32,
                    help="frame pairs per CUDA-graph replay (32: +4 %% over 16 on one B200; capped by the per-GPU batch)")
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--no-ops", action="store_true", help="skip the per-operator roofline table")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vid2vid", action="store_true", help="skip the BASELINE configs[4] training-iteration timing (`extra`)")
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    ap.add_argument("--memory-format", default="channels_last", choices=["channels_last", "contiguous"],
                    help="layout of the FlowNet2 conv body in the native arm (the reference arm keeps the stock NCHW)")
    ap.add_argument("--no-cuda-graph", action="store_true", help="native arm: launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--sd-overlap", action="store_true", help="native arm: run FlowNetSD on a second stream beside FlowNetC -> S -> S (A/B; no gain measured)")
    ap.add_argument("--no-sd-pad16", action="store_true", help="native arm: FlowNetSD.conv0 reads the 8-channel frame stack instead of the 16-channel one (A/B)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="keep cuDNN autotuning launches out of ncu launch lists")
    return ap.parse_args()


def peaks():
    """HBM copy bandwidth and the dense-TF32 tensor peak.  The driver measures a bf16 GEMM; TF32 runs at half the bf16
    rate on the same pipe, so bf16 / 2 is the TF32 proxy -- the sustained figure for a kernel timed inside a long step."""
    p = {"hbm_gbs": 6650.0, "tf32_tflops": 1400.0 / 2, "tf32_tflops_burst": 1590.0 / 2, "source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        m = json.load(open(path))
        p = {"hbm_gbs": float(m["hbm_gbs"]), "tf32_tflops": float(m.get("bf16_tflops_sustained", m["bf16_tflops"])) / 2,
             "tf32_tflops_burst": float(m["bf16_tflops"]) / 2, "source": "measured (MEASURED_PEAKS.json; TF32 = bf16 / 2)"}
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# (timed entry point, (B, C, H, W), tensor-core kernel?) -> DRAM bytes per launch measured with `ncu --set full` (profiles/)
NCU_DRAM_BYTES = {("correlation_planes_forward_into", (16, 256, 64, 128), False): (281336576 + 196153088, "profiles/ncu_corr_fwd_nhwc_b16_r01.txt"),
                  ("correlation_planes_forward_into", (16, 256, 64, 128), True): (273361920 + 191170304, "profiles/ncu_corr_fwd_tc_b16_r02.txt"),
                  ("correlation_planes_forward_into", (32, 256, 64, 128), True): (549026560 + 431249664, "profiles/ncu_corr_fwd_tc_b32_r02.txt")}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12        # 74.5


# ---------------------------------------------------------------------------------------------------
class Launches:
    """Counts this repo's kernel launches and times the dominant operator (Correlation forward) with
    CUDA events on the launching stream, live inside the timed region."""

    TIMED = ("correlation_forward", "correlation_planes_forward", "correlation_planes_forward_into")
    NO_KERNEL = ("corr_out_shape",)

    def __init__(self):
        self.count = 0
        self.corr_events = []
        self.epi_events = []        # (e0, e1, bytes) of the in-place bias + LeakyReLU epilogue launches
        self.enabled = False

    def install(self):
        import torch
        from ir2rgb_b200 import _lib
        from ir2rgb_b200 import functional as F

        def hook(what, n):          # every libflowops call that launches kernels goes through _lib.check()
            if self.enabled and what not in self.NO_KERNEL:
                self.count += n
        _lib.launch_hook = hook
        for name in self.TIMED:
            orig = getattr(F, name)

            def wrapped(*a, _orig=orig, _name=name, **k):
                if not self.enabled:
                    return _orig(*a, **k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = _orig(*a, **k)
                e1.record()
                self.corr_events.append((e0, e1, tuple(a[0].shape), _name))
                return out
            setattr(F, name, wrapped)
        orig_epi = F.bias_lrelu_

        def epi(y, *a, **k):
            # the 2-channel flow heads (a few hundred KB, launch-latency sized) are not part of the bandwidth figure
            if not self.enabled or y.numel() * 8 < (8 << 20):
                return orig_epi(y, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_epi(y, *a, **k)
            e1.record()
            self.epi_events.append((e0, e1, 2 * y.numel() * 4))
            return out
        F.bias_lrelu_ = epi


def build_native(device, memory_format="channels_last", overlap_sd=False, sd_pad16=True):
    import torch
    from ir2rgb_b200.models.flownet import FlowNet
    net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[device.index], checkpoints_dir=".", name="bench")
    if memory_format == "channels_last":
        # conv body tuning: NHWC weights/activations let cuDNN run its tensor-op kernels
        # without the NCHW<->NHWC transposes that take ~30 % of the stock forward (profiles/launches_r01_*)
        net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
    net.flowNet.overlap_sd = overlap_sd
    net.flowNet.sd_pad16 = sd_pad16
    return net.eval()


def run_step(net, im1, im2, mb, host_out=None):
    """One pass over this rank's batch in micro-batches.  im1/im2 on the device, or pinned host tensors
    (then each micro-batch is copied H2D here and its flow/conf are copied back into host_out)."""
    import torch
    B = im1.shape[0]
    dev = next(net.parameters()).device
    last = None
    for s in range(0, B, mb):
        a, b = im1[s:s + mb], im2[s:s + mb]
        if not a.is_cuda and not getattr(net, "accepts_host_inputs", False):
            a = a.to(dev, non_blocking=True)
            b = b.to(dev, non_blocking=True)
        flow, conf = net(a, b)
        if host_out is not None:
            host_out[0][s:s + mb].copy_(flow, non_blocking=True)
            host_out[1][s:s + mb].copy_(conf, non_blocking=True)
        last = flow
    return last


def timed(fn, steps, warmup, dist, device):
    import torch
    for _ in range(warmup):
        fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("bench_timed")       # ncu --nvtx --nvtx-include "bench_timed/" lists exactly these launches
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    torch.cuda.nvtx.range_pop()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms / steps


# ---------------------------------------------------------------------------------------------------
def cpu_port_pairs_per_s(max_pairs=8, min_seconds=10.0, threads=None):
    """The pure-PyTorch CPU port (oracle) of the same workload on the host cores: FlowNet2 forward +
    confidence on `max_pairs` 512x1024 pairs (bounded sample)."""
    import torch
    from oracle.harness import OracleFlowNet
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = OracleFlowNet("torch", "cpu")
    im1, im2 = 2 * torch.rand(1, 3, H, W) - 1, 2 * torch.rand(1, 3, H, W) - 1
    t0 = time.perf_counter()
    n = 0
    while n < max_pairs and (n == 0 or time.perf_counter() - t0 < min_seconds):
        net(im1, im2)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": "%d frame pair(s) of 512x1024 through the pure-PyTorch CPU oracle (FlowNet2 + conf), %.1f s" % (n, dt)}


def ops_table(pk, ffma):
    """Per-operator roofline fractions at BASELINE configs 2 and 3 (outside the timed region)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import opbench
    res = opbench.run(iters=10, skip_ref=True, quiet=True, ffma=ffma)
    out = {}
    for k, v in res["ops"].items():
        out[k] = {"us": round(v["us"], 2), "us_single_flushed": round(v["us_single_flushed"], 2),
                  "bound": v["bound"], "frac": round(v["frac"], 4),
                  "achieved": round(v.get("achieved_gbs", v.get("achieved_tflops", 0.0)), 2),
                  "unit": "GB/s" if v["bound"] == "hbm" else "TFLOP/s"}
        for extra in ("frac_executed", "executed_tflops", "x_fp32_pipe_peak"):
            if extra in v:
                out[k][extra] = round(v[extra], 4)
        if "cpu_reference_us" in v:
            out[k]["cpu_reference_us"] = round(v["cpu_reference_us"], 1)
            out[k]["cpu_threads"] = v["cpu_threads"]
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)

    import torch

    use_cpu_ref = False
    if args.impl == "reference":
        from oracle import ref_ext
        use_cpu_ref = args.ref_device == "cpu" or (args.ref_device == "auto" and not (ref_ext.available() and torch.cuda.is_available()))
        if use_cpu_ref and rank != 0:
            return 0            # the CPU arm runs on rank 0 alone

    if args.impl == "reference" and use_cpu_ref:
        base = cpu_port_pairs_per_s()
        line = {"metric": METRIC, "value": base["value"], "unit": "pairs/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": "flownet2_fwd+conf_512x1024", "global_batch": args.global_batch, "device": "cpu"},
                "cpu_baseline": base, "gpu_launches": 0,
                "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the hot path)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # whatever NCCL does print goes to stderr
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark     # True: as the reference does (models/models.py:26)

    from ir2rgb_b200.sharding import shard_bounds
    lo, hi = shard_bounds(args.global_batch, rank, world)      # contiguous chunk of the global batch
    B_local = hi - lo
    mb = min(args.micro_batch, B_local)
    torch.manual_seed(1234 + rank)
    im1 = 2 * torch.rand(B_local, 3, H, W, device=device) - 1
    im2 = 2 * torch.rand(B_local, 3, H, W, device=device) - 1
    h1, h2 = im1.cpu().pin_memory(), im2.cpu().pin_memory()
    hflow = torch.empty(B_local, 2, H, W).pin_memory()
    hconf = torch.empty(B_local, 1, H, W).pin_memory()

    launches = Launches()
    ffma_idle = None
    if args.impl == "native":
        from ir2rgb_b200 import functional as F0
        ffma_idle = F0.ffma_peak_tflops()            # FP32-FMA pipe peak on the idle, cool GPU (before any step has run)
        launches.install()
        torch.manual_seed(0)
        net = build_native(device, args.memory_format, args.sd_overlap, not args.no_sd_pad16)
        if not args.no_cuda_graph:
            from ir2rgb_b200.runtime import GraphedFlowNet
            net = GraphedFlowNet(net)
    else:
        from oracle.harness import OracleFlowNet, reference_flownet2_available
        torch.manual_seed(0)
        # the reference's own FlowNet2 class + wrappers + rebuilt extensions when oracle/_ref holds them (refclass);
        # otherwise the restated architecture with the reference's extensions swapped in
        ref_kind = "refclass" if reference_flownet2_available() else "ref"
        net = OracleFlowNet(ref_kind, device)

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda: run_step(net, im1, im2, mb), args.steps, args.warmup, dist, device)
    clocks = sampler.stop() if rank == 0 else None
    if args.impl == "native":
        # public runtime helper: H2D of micro-batch i+1 and D2H of i-1 overlap the computation of micro-batch i
        from ir2rgb_b200.runtime import HostPipeline
        pipe = HostPipeline(net, device)
        mb_host = mb      # copies overlap computation ACROSS steps (next_inputs / wait=False below), so one micro-batch per step is fine
        # a stream of batches: the next batch's first copy-in and this batch's last copy-out run under computation
        # (every step still copies all of its frames in and all of its results out inside the timed region; timed()
        # synchronises the device, and with it both copy streams, before it stops the clock)
        e2e_step = lambda: pipe(h1, h2, mb_host, hflow, hconf, next_inputs=(h1, h2), wait=False)
    else:
        e2e_step = lambda: run_step(net, h1, h2, mb, host_out=(hflow, hconf))
    ms_e2e = timed(e2e_step, max(1, args.steps), 1, dist, device)

    # Launch count and live timing of the dominant hand-written kernel: one more step of the same workload,
    # launched eagerly (the same kernels the graph replays) with CUDA events around the Correlation call.
    n_launch, corr_events, ms_eager = 0, [], None
    if args.impl == "native":
        eager = getattr(net, "net", net)
        overlap = eager.flowNet.overlap_sd
        eager.flowNet.overlap_sd = False         # per-kernel timings: no second stream running beside the timed kernels
        run_step(eager, im1, im2, mb)
        launches.enabled = True
        ms_eager = timed(lambda: run_step(eager, im1, im2, mb), 1, 0, None, device)
        launches.enabled = False
        eager.flowNet.overlap_sd = overlap
        n_launch = launches.count * args.steps
        corr_events = launches.corr_events

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    line = {"metric": METRIC, "value": args.global_batch / (ms_dev * 1e-3), "unit": "pairs/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "flownet2_fwd+conf_512x1024 (BASELINE configs[3])", "global_batch": args.global_batch,
                       "per_gpu_batch": B_local, "micro_batch": mb, "frame": [H, W], "weights": "random-init",
                       "conv_math": "cudnn fp32 with TF32 allowed (torch default, same in the reference arm)",
                       "conv_layout": args.memory_format if args.impl == "native" else "contiguous",
                       "cuda_graph": bool(args.impl == "native" and not args.no_cuda_graph),
                       "sd_branch_on_second_stream": bool(args.impl == "native" and args.sd_overlap),
                       "l2": "inputs larger than L2 (per-rank frames %.0f MB, activations several GB per micro-batch)"
                             % (2 * B_local * 3 * H * W * 4 / 1e6),
                       "parallelism": "batch-sharded x%d, no collective on the data path" % world},
            "clocks": clocks,
            "e2e": {"value": args.global_batch / (ms_e2e * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": 2 * B_local * 3 * H * W * 4, "d2h_bytes_per_step": B_local * 3 * H * W * 4},
            "gpu_launches": n_launch}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["config"]["operators"] = ("the reference's own FlowNet2 class (models/flownet2_pytorch/models.py, byte-compiled "
                                       "unmodified into oracle/_ref/refpy) with its own wrappers and CUDA extensions rebuilt for sm_100"
                                       if ref_kind == "refclass" else
                                       "reference CUDA extensions rebuilt for sm_100 (oracle/_ref) in the restated FlowNet2, unfused glue")
        line["cpu_baseline"] = {"value": line["value"], "unit": "pairs/s", "cores": 0, "kind": "reference",
                                "sample": "reference operators are CUDA-only; this arm ran them on the GPU (see DESIGN.md)"}
    else:
        from ir2rgb_b200 import functional as F
        from ir2rgb_b200 import _lib
        ffma = F.ffma_peak_tflops()                  # the same probe right after the steps (board warm, usually power-capped)
        tc_on = bool(_lib.load().flowops_corr_get_impl() & 1)
        # roofline of the operator BASELINE.json's metric names: Correlation forward, timed live inside one eager step
        if corr_events:
            us = [ev[0].elapsed_time(ev[1]) * 1e3 for ev in corr_events]
            shp = corr_events[0][2]
            split = corr_events[0][3].startswith("correlation_planes_forward")
            flop = 2.0 * shp[0] * shp[2] * shp[3] * 441 * shp[1]          # algorithmic (useful) flops: 2*B*H*W*441*C
            mean_us = sum(us) / len(us)
            traffic, traffic_src = NCU_DRAM_BYTES.get((corr_events[0][3], tuple(shp), tc_on), (None, None))
            common = {"alg_flop_per_launch": flop, "us_per_launch": mean_us, "launches_timed": len(us),
                      "share_of_step": sum(us) / (ms_eager * 1e3), "traffic": traffic, "traffic_source": traffic_src,
                      "fp32_peak": {"idle": ffma_idle, "after_steps": ffma, "nominal": FP32_NOMINAL_TFLOPS,
                                    "source": "flowops_bench_ffma in this process, before the first step and after the timed steps"},
                      "timed_in": "one eagerly launched single-stream step of the same workload (the timed steps replay these kernels from a CUDA graph)"}
            if tc_on:
                # tcgen05 kernel (csrc/corr_tc.cu).  `achieved` counts ALGORITHMIC flops; the tensor pipe executes
                # 3 (3xTF32) x 1024/441 (dense 128 x 256 UMMA tiles around a banded contraction) = 6.97x as many.
                executed = flop * 3.0 * 1024.0 / 441.0
                line["roofline"] = dict(common, kernel="corr_fwd_tc (tcgen05 3xTF32; input planes are written by the conv3 epilogue kernel)",
                                        bound="tensor", achieved=flop / mean_us / 1e6, peak=pk["tf32_tflops"], unit="TFLOP/s",
                                        frac=flop / mean_us / 1e6 / pk["tf32_tflops"],
                                        frac_3xtf32=flop / mean_us / 1e6 / (pk["tf32_tflops"] / 3.0),
                                        peak_source=pk["source"] + ", sustained figure (kernel timed inside a long step)",
                                        executed_tflops=executed / mean_us / 1e6, frac_executed=executed / mean_us / 1e6 / pk["tf32_tflops"],
                                        executed_over_algorithmic=3.0 * 1024.0 / 441.0,
                                        x_fp32_pipe_peak=flop / mean_us / 1e6 / max(ffma_idle or ffma, ffma),
                                        note="frac = algorithmic / TF32 peak; frac_3xtf32 = against TF32 peak / 3, the ceiling of ANY "
                                             "fp32-accurate (3xTF32) tensor-core kernel; frac_executed = flops the tensor pipe executes / "
                                             "TF32 peak; x_fp32_pipe_peak > 1 means faster than any FP32-FMA kernel can be")
            else:
                line["roofline"] = dict(common, kernel="corr_fwd_fast (input planes are written by the conv3 epilogue kernel)" if split
                                        else "corr_fwd (corr_planarize + corr_fwd_fast)", bound="fp32",
                                        achieved=flop / mean_us / 1e6, peak=ffma, unit="TFLOP/s", frac=flop / mean_us / 1e6 / ffma,
                                        peak_idle=ffma_idle, peak_in_step=ffma, frac_vs_nominal=flop / mean_us / 1e6 / FP32_NOMINAL_TFLOPS,
                                        peak_source="flowops_bench_ffma measured on this GPU after the steps (nominal 74.5)")
        # the largest hand-written kernel family of the step by time (ncu launch list: bias_lrelu_nhwc ~12 % of GPU time):
        # the in-place bias + LeakyReLU epilogue that follows every convolution -- HBM-bound, 8 bytes per element
        if launches.epi_events:
            us = [ev[0].elapsed_time(ev[1]) * 1e3 for ev in launches.epi_events]
            nbytes = sum(ev[2] for ev in launches.epi_events)
            gbs = nbytes / sum(us) / 1e3
            line["roofline_epilogue"] = {"kernel": "bias_lrelu_nhwc (all in-place conv epilogues of the step that move at least 8 MB)", "bound": "hbm",
                                         "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                                         "traffic": None, "alg_bytes_per_launch": nbytes / len(us), "us_per_launch": sum(us) / len(us),
                                         "launches_timed": len(us), "share_of_step": sum(us) / (ms_eager * 1e3),
                                         "timed_in": "the same eagerly launched step; launches are back to back with cuDNN kernels, "
                                                     "so part of each tensor is still in L2 when its epilogue runs"}
        line["peaks"] = dict(pk, ffma_tflops=ffma, ffma_tflops_idle=ffma_idle, ffma_tflops_nominal=FP32_NOMINAL_TFLOPS)
        if not args.no_ops:
            try:
                line["ops"] = ops_table(pk, ffma)
            except Exception as e:      # the table is informative; never lose the headline line over it
                line["ops"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_port_pairs_per_s()
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
    if args.impl == "native" and world == 1 and not args.no_vid2vid:
        # BASELINE configs[4]: one vid2vid training iteration (FlowNet2 flow + warp losses through the new operators),
        # 1 GPU here; the 8-GPU data-parallel run is tools/bench_vid2vid_step.py under torchrun.  Own process: the
        # 365 M-parameter generator needs the memory this one still holds.
        try:
            del net, im1, im2
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_vid2vid_step.py"), "--iters", "5"],
                               capture_output=True, text=True, timeout=600)
            rec = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            line["extra"] = {"vid2vid_step": json.loads(rec[-1])} if rec else {"vid2vid_step": {"error": r.stderr[-400:]}}
        except Exception as e:
            line["extra"] = {"vid2vid_step": {"error": repr(e)}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
