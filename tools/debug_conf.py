import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200 import functional as F
from ir2rgb_b200.models.flownet import FlowNet
torch.manual_seed(9)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.deterministic = True
net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
net.flowNet = net.flowNet.to(memory_format=torch.channels_last)
a = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
b = 2 * torch.rand(2, 3, 128, 192, device="cuda") - 1
with torch.no_grad():
    data1 = torch.cat([a.unsqueeze(2), b.unsqueeze(2)], dim=2)
    flow = net.flowNet(data1)
    print("flow layout contiguous:", flow.is_contiguous(), "cl:", flow.is_contiguous(memory_format=torch.channels_last), "absmax", flow.abs().max().item(), "finite", torch.isfinite(flow).all().item())
    conf_f = F.warp_conf_forward(a, b, flow, 0.02)
    w = net.resample(b, flow)
    t = a - w
    s = torch.sum(t * t, dim=1, keepdim=True)
    conf_p = (s < 0.02).float()
    diff = conf_f != conf_p
    print("flips", diff.float().mean().item(), "n", diff.sum().item())
    w2 = F.warp_forward(b, flow.contiguous(), 0)
    print("warp equal", torch.equal(w, w2))
    s2 = ((a - w2)[:, 0] ** 2 + (a - w2)[:, 1] ** 2) + (a - w2)[:, 2] ** 2
    print("sum order a+b+c equal to torch.sum:", torch.equal(s2.unsqueeze(1), s), (s2.unsqueeze(1) - s).abs().max().item())
    idx = diff.nonzero()[:5]
    for i in idx:
        print(i.tolist(), s[tuple(i)].item(), s2.unsqueeze(1)[tuple(i)].item())
    from oracle import c_oracle as co
    ref = torch.from_numpy(co.resample2d_fwd(b.cpu().numpy(), flow.contiguous().cpu().numpy())).cuda()
    print("module vs oracle", torch.equal(w, ref), (w - ref).abs().max().item(), " functional vs oracle", torch.equal(w2, ref), (w2 - ref).abs().max().item())
    w3 = net.resample(b, flow.contiguous())
    print("module with contiguous flow vs oracle", torch.equal(w3, ref))
    fc = flow.contiguous()
    print("flow.contiguous strides", fc.stride(), fc.is_contiguous(), "flow strides", flow.stride())
    w4 = F.warp_forward(b, flow, 0)
    print("functional with CL flow vs oracle", torch.equal(w4, ref))
