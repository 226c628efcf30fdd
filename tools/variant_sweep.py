"""Time the pre-train loss (graph replay) and its kernels for one library variant built by tools/build_variants.py:
    VARIANT=A python tools/variant_sweep.py      (no VARIANT: the in-tree library)"""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import _lib
v = os.environ.get("VARIANT")
if v:
    _lib.LIB_PATH = os.path.join(ROOT, "hmmc_b200", "_variants", v + ".so")
from hmmc_b200 import ops, synthetic as syn
from hmmc_b200.graphs import GraphedStep

cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
qn = ("v_fea", "title_fea", "frame_fea", "frame_pred")
for b in (128, 256):
    inp = syn.pretrain_inputs(b, F=12, D=512, seed=2)
    qs = {n: cu(x) for n, x in syn.queues(1024, F=12, D=512, seed=3).items()}
    t = {n: cu(x) for n, x in inp.items()}
    for n in qn:
        t[n].requires_grad_(True)

    def run():
        for n in qn:
            t[n].grad = None
        total, _ = ops.pretrain_head(t["v_fea"], t["title_fea"], t["frame_fea"], t["frame_pred"], t["v_fea_k"],
                                     t["title_fea_k"], t["frame_fea_k"], t["frame_proj_k"], qs["queue_v_cross_ng"],
                                     qs["queue_title_cross_ng"], qs["queue_frame_proj_ng"], qs["queue_frame_cross_ng"],
                                     0.07, 0.05, 0.45, 0.45, True, "bf16")
        total.backward()
        return total
    g = GraphedStep(run)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
    print("variant %-8s b=%3d loss fwd+bwd %7.1f us  loss=%.6f" % (v or "in-tree", b, best, float(g.outputs)), flush=True)
