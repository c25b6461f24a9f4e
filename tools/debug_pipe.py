import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ir2rgb_b200.models.flownet import FlowNet
from ir2rgb_b200.runtime import HostPipeline
torch.manual_seed(0)
torch.backends.cudnn.deterministic = True
torch.backends.cudnn.allow_tf32 = False
net = FlowNet(fp16=False, flownet_checkpoint_path=None, gpu_ids=[0], checkpoints_dir=".", name="t").eval()
torch.manual_seed(7)
h1 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory(); h2 = (2 * torch.rand(5, 3, 64, 128) - 1).pin_memory()
hf, hc = torch.empty(5, 2, 64, 128).pin_memory(), torch.empty(5, 1, 64, 128).pin_memory()
a, b = h1[0:2].cuda(), h2[0:2].cuda()
f1, _ = net(a, b); f2, _ = net(a, b)
print("determinism same call:", (f1 - f2).abs().max().item())
f3, _ = net(h1[0:2].cuda(), h2[0:2].cuda())
print("determinism new copies:", (f1 - f3).abs().max().item())
pipe = HostPipeline(net, torch.device("cuda", 0))
for rep in range(2):
    pipe(h1, h2, 2, hf, hc); torch.cuda.synchronize()
    for s in (0, 2, 4):
        f, c = net(h1[s:s + 2].cuda(), h2[s:s + 2].cuda())
        print("rep", rep, "s", s, "flow diff", (hf[s:s + 2] - f.cpu()).abs().max().item(), "conf diff", (hc[s:s+2]-c.cpu()).abs().max().item())
