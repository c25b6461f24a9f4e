"""The flow path of ONE vid2vid training iteration (BASELINE configs[4]), isolated and timed.

Per iteration and per frame pair the reference's train loop executes on the flow path
  * FlowNet(real_B, real_B_prev) under no_grad            (train_vid2vid.py:65 -> flownet.py:20-57),
  * the generator's warp  resample(img_prev, flow)         (networks.py:207; gradients to flow and img_prev),
  * loss_F_Warp: resample(real_B_prev, flow)               (discriminator.py:120; gradient to flow),
  * loss_G_Warp: resample(fake_B_prev, flow_ref)           (discriminator.py:137; forward only, detached at :138),
with MaskedL1 losses on top (discriminator.py:118-138) -- i.e. 1 FlowNet2 forward + confidence, 3 grid_sample
warps forward and 2 backward, at batch 1 and 512x1024 (max_frames_per_gpu = 1).  The generator / discriminator
networks themselves are stock cuDNN modules outside the hot path and are replaced by
leaf tensors here: `flow` and `img_prev` stand for the generator outputs the gradients flow back into.

Runs one process per GPU (torchrun) or a single process; every rank does the same independent work (frame
pairs are independent: no collective on this path).  Prints one JSON line with the per-iteration time of the
new path and of the reference's (rebuilt CUDA extensions + ATen grid_sample chain) on the same GPU.

    python tools/flowpath_step.py [--iters 20] [--skip-ref]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def masked_l1(a, b, mask):
    """loss.MaskedL1Loss (models/loss.py): L1 between mask-weighted tensors."""
    mask = mask.expand_as(a)
    return torch.nn.functional.l1_loss(a * mask, b * mask)


def make_step(flow_net, resample, device, H, W, lambda_T=10.0, lambda_F=10.0):
    torch.manual_seed(1)
    real_B = (2 * torch.rand(1, 1, 3, H, W, device=device) - 1)
    real_B_prev = (real_B + 0.05 * torch.randn_like(real_B)).clamp(-1, 1)
    fake_B = (2 * torch.rand(1, 3, H, W, device=device) - 1)
    fake_B_prev = (2 * torch.rand(1, 3, H, W, device=device) - 1)
    # generator outputs the flow-path gradients reach
    flow = (3 * torch.randn(1, 2, H // 8, W // 8, device=device))
    flow = torch.nn.functional.interpolate(flow, size=(H, W), mode="bilinear", align_corners=False).contiguous().requires_grad_()
    img_prev = (2 * torch.rand(1, 3, H, W, device=device) - 1).requires_grad_()
    weight = torch.rand(1, 1, H, W, device=device)
    img_raw = 2 * torch.rand(1, 3, H, W, device=device) - 1

    def step():
        flow.grad = None
        img_prev.grad = None
        flow_ref, conf_ref = flow_net(real_B, real_B_prev)                   # no_grad inside
        flow_ref, conf_ref = flow_ref[:, 0], conf_ref[:, 0]
        img_warp = resample(img_prev, flow)                                  # networks.py:207
        img_final = img_raw * weight + img_warp * (1 - weight)               # networks.py:209-210
        loss_F_Flow = masked_l1(flow, flow_ref, conf_ref) * lambda_F         # discriminator.py:118
        real_B_warp = resample(real_B_prev[:, 0], flow)                      # :120
        loss_F_Warp = masked_l1(real_B_warp, real_B[:, 0], conf_ref) * lambda_T
        fake_B_warp_ref = resample(fake_B_prev, flow_ref)                    # :137
        loss_G_Warp = masked_l1(fake_B, fake_B_warp_ref.detach(), conf_ref) * lambda_T
        loss = loss_F_Flow + loss_F_Warp + loss_G_Warp + masked_l1(img_final, real_B[:, 0], conf_ref)
        loss.backward()
        return loss.detach(), flow.grad, img_prev.grad
    return step


def time_step(step, iters, device):
    for _ in range(3):
        step()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = step()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    torch.backends.cudnn.benchmark = True
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    from ir2rgb_b200.models import networks
    from ir2rgb_b200.models.flownet import FlowNet
    from ir2rgb_b200.runtime import GraphedFlowNet
    torch.manual_seed(0)
    net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[device.index], checkpoints_dir=".", name="step").eval()
    state = {k: v.clone() for k, v in net.flowNet.state_dict().items()}
    net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
    graphed = GraphedFlowNet(net)                    # batch 1: ~450 launches per FlowNet2 forward are launch-bound
    ms_new, out_new = time_step(make_step(graphed, networks.resample, device, args.height, args.width), args.iters, device)

    line = {"metric": "vid2vid_flow_path_ms_per_iteration", "unit": "ms", "higher_is_better": False, "n_gpus": world,
            "config": {"workload": "flow path of one vid2vid training iteration (BASELINE configs[4]): FlowNet2 fwd + conf, "
                                   "3 grid_sample warps fwd, 2 bwd, masked-L1 flow/warp losses; batch 1 per GPU",
                       "frame": [args.height, args.width], "weights": "random-init"},
            "value": ms_new, "iterations_per_s_all_gpus": world * 1e3 / ms_new}
    if not args.skip_ref:
        from oracle import ref_ext, torch_ref
        from oracle.harness import OracleFlowNet
        if ref_ext.available():
            oracle_net = OracleFlowNet("ref", str(device), state_dict=state)
            ref_net = lambda a, b: tuple(t.unsqueeze(1) for t in oracle_net(a[:, 0], b[:, 0]))      # 5-D in, 5-D out
            ms_ref, out_ref = time_step(make_step(ref_net, torch_ref.networks_resample, device, args.height, args.width),
                                        max(3, args.iters // 2), device)
            # parity of the warps and losses: feed BOTH paths the same reference flow / confidence (the two FlowNet2
            # bodies differ by TF32 conv noise, and the hard-thresholded confidence mask turns that into whole-pixel flips)
            fixed = tuple(t.clone() for t in graphed(*[2 * torch.rand(1, 1, 3, args.height, args.width, device=device) - 1] * 2))
            fixed_net = lambda a, b: fixed
            p_new = make_step(fixed_net, networks.resample, device, args.height, args.width)()
            p_ref = make_step(fixed_net, torch_ref.networks_resample, device, args.height, args.width)()
            rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
            line["reference"] = {"value": ms_ref, "operators": "reference CUDA extensions (oracle/_ref) + ATen grid_sample chain, eager",
                                 "parity_same_flow_ref": {"loss_rel_diff": abs(p_new[0].item() - p_ref[0].item()) / abs(p_ref[0].item()),
                                                          "grad_flow_max_rel_diff": rel(p_new[1], p_ref[1]),
                                                          "grad_img_max_rel_diff": rel(p_new[2], p_ref[2])},
                                 "end_to_end_loss_rel_diff": abs(out_new[0].item() - out_ref[0].item()) / abs(out_ref[0].item())}
            line["speedup_vs_reference"] = ms_ref / ms_new
    if dist is not None:
        t = torch.tensor([ms_new], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["value"] = t.item()
        line["iterations_per_s_all_gpus"] = world * 1e3 / t.item()
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
