"""BASELINE configs[4]: one ir2rgb / vid2vid training iteration at 512x1024 (composite generator ngf 128, batch norm, no
VGG loss, n_input_gen_frames 3, two temporal discriminator scales, one frame per GPU and iteration), random-init
G / D / FlowNet2, synthetic frames, one process per GPU with an NCCL gradient all-reduce per network.

    python tools/bench_vid2vid_step.py [--impl native|reference] [--iters 5]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_vid2vid_step.py

`--impl reference` runs the SAME generator / discriminator modules with the reference's operators on the flow path
(rebuilt CUDA extensions inside FlowNet2, ATen grid_sample chain for `resample`), so the two arms differ only in the
flow hot path.  Prints one JSON line: ms per iteration (max over ranks, CUDA events), per-phase times on rank 0,
the share of the flow path, and the all-reduce time.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=8, help="iterations before timing; both temporal scales are active from the 7th")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--ngf", type=int, default=128)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    torch.backends.cudnn.benchmark = True
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from ir2rgb_b200.train.vid2vid_step import Vid2VidStep
    torch.manual_seed(0)
    if args.impl == "native":
        from ir2rgb_b200.models.flownet import FlowNet
        from ir2rgb_b200.runtime import GraphedFlowNet
        net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[device.index], checkpoints_dir=".", name="c5").eval()
        net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
        graphed = GraphedFlowNet(net)
        flow_net = lambda a, b: graphed(a, b)
        resample = None                                   # libflowops (ir2rgb_b200.models.networks.resample)
    else:
        from oracle import torch_ref
        from oracle.harness import OracleFlowNet
        oracle_net = OracleFlowNet("ref", str(device))

        def flow_net(a, b):                               # flownet.py:20-36: fold frames into the batch
            n, t, c, h, w = a.shape
            f, cf = oracle_net(a.reshape(-1, c, h, w), b.reshape(-1, c, h, w))
            return f.view(n, t, 2, h, w), cf.view(n, t, 1, h, w)
        resample = torch_ref.networks_resample
    torch.manual_seed(0)                                  # identical G / D initialisation on every rank
    step = Vid2VidStep(flow_net, device, ngf=args.ngf, resample=resample, world_size=world)

    g = torch.Generator(device=device)
    g.manual_seed(1000 + rank)                            # every rank trains on its own frames
    H, W = args.height, args.width

    def window():
        return (2 * torch.rand(1, 3, 3, H, W, device=device, generator=g) - 1,
                2 * torch.rand(1, 3, 3, H, W, device=device, generator=g) - 1)

    for _ in range(args.warmup):
        out = step.step(*window())
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    phases = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        out = step.step(*window())
        for k, v in step.timer_ms().items():
            phases[k] = phases.get(k, 0.0) + v / args.iters
    e1.record()
    e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        flow_path = phases.get("flownet", 0.0) + phases.get("flow_losses", 0.0)
        allreduce = sum(v for k, v in phases.items() if k.startswith("allreduce_"))
        n_params = {"G": sum(p.numel() for p in step.netG.parameters()),
                    "D": sum(p.numel() for p in step.netD.parameters()) + sum(p.numel() for n in step.netD_T for p in n.parameters())}
        print(json.dumps({
            "metric": "vid2vid_train_iteration_ms", "unit": "ms", "higher_is_better": False, "impl": args.impl, "n_gpus": world,
            "value": ms.item(), "iterations_per_s_all_gpus": world * 1e3 / ms.item(), "iters": args.iters, "warmup": args.warmup,
            "config": {"workload": "vid2vid training iteration (BASELINE configs[4])", "frame": [H, W], "ngf": args.ngf, "norm": "batch",
                       "n_input_gen_frames": 3, "n_frames_D": 3, "n_scales_temporal": 2, "no_vgg": True, "batch_per_gpu": 1,
                       "params": n_params, "weights": "random-init", "data": "synthetic, seed = rank",
                       "flow_path": "libflowops" if args.impl == "native" else "reference CUDA extensions + ATen grid_sample chain"},
            "phases_ms_rank0": {k: round(v, 3) for k, v in sorted(phases.items())},
            "flow_path_ms": round(flow_path, 3), "flow_path_share": flow_path / ms.item(),
            "flow_path_note": "FlowNet2 + confidence (incl. the temporal scales' skipped-frame flows) and the two loss warps fwd; "
                              "the generator's own warp and all warp backwards are inside generator_fwd / generator_bwd",
            "allreduce_ms": round(allreduce, 3), "allreduce_bytes": 4 * (n_params["G"] + n_params["D"]),
            "temporal_scales_active": out["temporal_scales_active"],
            "losses": {k: (v.item() if hasattr(v, "item") else v) for k, v in out.items() if k in ("G", "D", "F_Flow", "F_Warp", "G_Warp")}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
