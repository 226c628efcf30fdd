"""Device time of pack_rows + enqueue for a gathered batch (no collective): B = W*b keys."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn

F, D, K = 12, 512, 1024
dev = torch.device("cuda")
qs = [torch.from_numpy(x).to(dev) for x in syn.queues(K, F=F, D=D, seed=3).values()]
order = [0, 4, 3, 2, 1]   # v, tag, title, frame_cross, frame_proj  (QUEUE_NAMES order is v, fproj, fcross, title, tag)
bufs = [qs[0], qs[4], qs[3], qs[2], qs[1]]
ptr = torch.zeros(1, dtype=torch.long, device=dev)
for prec in ("bf16", "bf16x3", "fp32"):
    p = ops.resolve_precision(prec)
    for W, b in ((1, 128), (4, 128), (8, 128)):
        g = torch.randn(W * b, (3 + 2 * F) * D, device=dev)
        keys = [torch.randn(b, D, device=dev) for _ in range(3)] + [torch.randn(b, F, D, device=dev) for _ in range(2)]
        def run():
            ops.enqueue(g, W, b, F, D, bufs, ptr, 0, K, p)
        def runp():
            ops.pack_rows(keys)
        for name, fn in (("enqueue", run), ("pack_rows(b=128)", runp)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("%-6s B=%4d %-18s %7.1f us" % (prec, W * b, name, e0.elapsed_time(e1) / 20 * 1e3))
