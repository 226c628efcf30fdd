"""FAM block alone, eager, forward-only then forward+backward, for an ncu launch list."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
b, F, D, K = int(os.environ.get("B", 256)), 12, 512, 1024
inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
q = cu(inp["frame_pred"]).reshape(b * F, D)
k = cu(inp["frame_proj_k"]).reshape(b * F, D)
queue = cu(syn.queues(K, F=F, D=D, seed=3)["queue_frame_proj_ng"])
p = ops.resolve_precision(os.environ.get("PREC", "bf16"))
for need in (False, True, False, True):
    ops.infonce_raw(q, k, ops.POS_FRAME_NEIGHBOUR, b, F, F, queue, 0.07, 1.0, p, need)
torch.cuda.synchronize()
