import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
 "generic": "a=torch.randn(1,8,9,11,device='cuda'); o=F.correlation_forward(a,a,4,1,4,1,1); torch.cuda.synchronize(); print('generic ok', o.shape)",
 "fast_small": "a=torch.randn(2,16,8,12,device='cuda'); o=F.correlation_forward(a,a,20,1,20,1,2); torch.cuda.synchronize(); print('fast small ok', o.abs().sum().item())",
 "fast_w128": "a=torch.randn(1,8,64,128,device='cuda'); o=F.correlation_forward(a,a,20,1,20,1,2); torch.cuda.synchronize(); print('fast w128 ok', o.abs().sum().item())",
 "fast_c2": "a=torch.randn(8,256,48,64,device='cuda'); o=F.correlation_forward(a,a,20,1,20,1,2); torch.cuda.synchronize(); print('fast c2 ok', o.abs().sum().item())",
}
for name, code in CASES.items():
    env = dict(os.environ, FLOWOPS_DEBUG_SYNC="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import torch; from ir2rgb_b200 import functional as F; " + code], env=env, capture_output=True, text=True, timeout=300)
    print("==", name, "rc", r.returncode, (r.stdout + r.stderr).strip().splitlines()[-1:] )
