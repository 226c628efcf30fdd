"""Where does the S-GEMM (InfoNCE epilogue) spend its time?  FAM block (R=1536 x Kq=12288 x K=512):
forward only (row sums, no E store) vs forward+backward (E store + U-GEMM), graph replay."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn
from hmmc_b200.graphs import GraphedStep

cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
b, F, D, K = int(os.environ.get("B", 128)), 12, 512, 1024
inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
q = cu(inp["frame_pred"]).reshape(b * F, D)
k = cu(inp["frame_proj_k"]).reshape(b * F, D)
queue = cu(syn.queues(K, F=F, D=D, seed=3)["queue_frame_proj_ng"])
for prec in ("bf16", "bf16x3"):
    p = ops.resolve_precision(prec)
    for need in (False, True):
        fn = lambda: ops.infonce_raw(q, k, ops.POS_FRAME_NEIGHBOUR, b, F, F, queue, 0.07, 1.0, p, need)
        g = GraphedStep(fn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("FAM block b=%d %-6s %-14s %7.1f us" % (b, prec, "fwd+bwd" if need else "fwd only", e0.elapsed_time(e1) / 50 * 1e3))
