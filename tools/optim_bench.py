"""A few optimizer steps over the 362-tensor / 172 M-parameter table (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import synthetic as syn
from hmmc_b200.optimization import BertAdam

sizes = syn.ema_param_numels()
n = sum(sizes)
flat = torch.randn(n, device="cuda") * 0.02
gflat = torch.randn(n, device="cuda") * 1e-4
params = [torch.nn.Parameter(x) for x in torch.split(flat, sizes)]
grads = list(torch.split(gflat, sizes))
opt = BertAdam(params, lr=1e-7, warmup=0.1, t_total=1000, schedule="warmup_cosine", b2=0.98, weight_decay=0.2)
for _ in range(int(os.environ.get("REPS", "3"))):
    for p, g in zip(params, grads):
        p.grad = g
    opt.step(global_max_norm=1.0)
torch.cuda.synchronize()
print("grad norm", float(opt.last_grad_norm))
