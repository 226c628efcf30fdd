"""Launch list of one MLP forward + backward (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200.mlp import MLP

prec = os.environ.get("PREC", "bf16")
M, Din, Dh, Dout = 1536, 512, 4096, 512
m = MLP(Din, Dh, Dout, 2, precision=prec).cuda().train()
x = torch.randn(M, Din, device="cuda", requires_grad=True)
dy = torch.randn(M, Dout, device="cuda") * 0.05
for _ in range(int(os.environ.get("REPS", "3"))):
    x.grad = None
    m(x).backward(dy)
torch.cuda.synchronize()
