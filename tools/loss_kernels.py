"""Run the pre-train loss fwd+bwd a few times (eager) so ncu can list its kernels:
   ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none python tools/loss_kernels.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
b = int(os.environ.get("B", 256))
prec = os.environ.get("PREC", "bf16")
inp = syn.pretrain_inputs(b, F=12, D=512, seed=2)
qs = {n: cu(x) for n, x in syn.queues(1024, F=12, D=512, seed=3).items()}
t = {n: cu(x) for n, x in inp.items()}
NOGRAD = os.environ.get("NOGRAD") == "1"      # forward only: the S-GEMM epilogue without the E stores
for it in range(4):
    tt = {n: (x.requires_grad_(not NOGRAD) if n in ("v_fea", "title_fea", "frame_fea", "frame_pred") else x) for n, x in t.items()}
    for x in tt.values():
        x.grad = None
    total, parts = ops.pretrain_head(tt["v_fea"], tt["title_fea"], tt["frame_fea"], tt["frame_pred"], tt["v_fea_k"],
                                     tt["title_fea_k"], tt["frame_fea_k"], tt["frame_proj_k"], qs["queue_v_cross_ng"],
                                     qs["queue_title_cross_ng"], qs["queue_frame_proj_ng"], qs["queue_frame_cross_ng"],
                                     0.07, 0.05, 0.45, 0.45, True, prec)
    if not NOGRAD:
        total.backward()
torch.cuda.synchronize()
print("loss", float(total))
