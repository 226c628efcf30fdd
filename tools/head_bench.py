"""Device time of the individual head ops (CUDA events, warm, resident inputs)."""
import os
import sys
import types
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import modeling, ops, retrieval
from hmmc_b200 import synthetic as syn


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
for B in (256, 1024):
    t, v, fr = [cu(x) for x in syn.finetune_inputs(B, seed=1)]
    for prec in ("fp32", "bf16x3", "bf16"):
        p = ops.resolve_precision(prec)
        us = timeit(lambda: ops.sym_ce_raw(t, v, fr, 100.0, 0.85, 0.15, p, True))
        print("fine-tune head fwd+bwd  B=%4d %-6s %8.1f us" % (B, prec, us))
T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=4)
T, V, Fr = cu(T), cu(V), cu(Fr)
for prec in ("fp32", "bf16x3", "bf16"):
    task = types.SimpleNamespace(local_rank=0, top_frames=2, use_frame_fea=True, head_precision=prec)
    m = modeling.BirdModel(modeling.default_cross_config(), task)
    us = timeit(lambda: ops.rank_count(retrieval.similarity_matrix(m, T, V, Fr)))
    print("eval 1000x1000x12 sim+rank %-6s %8.1f us" % (prec, us))

# pre-train head loss fwd+bwd only (no EMA, no enqueue)
for b in (128, 256):
    inp = syn.pretrain_inputs(b, F=12, D=512, seed=2)
    qs = {n: cu(x) for n, x in syn.queues(1024, F=12, D=512, seed=3).items()}
    t = {n: cu(x) for n, x in inp.items()}
    for prec in ("bf16", "bf16x3"):
        def run(prec=prec):
            tt = {n: (x.requires_grad_(True) if n in ("v_fea", "title_fea", "frame_fea", "frame_pred") else x) for n, x in t.items()}
            for x in tt.values():
                x.grad = None
            total, parts = ops.pretrain_head(tt["v_fea"], tt["title_fea"], tt["frame_fea"], tt["frame_pred"], tt["v_fea_k"],
                                             tt["title_fea_k"], tt["frame_fea_k"], tt["frame_proj_k"], qs["queue_v_cross_ng"],
                                             qs["queue_title_cross_ng"], qs["queue_frame_proj_ng"], qs["queue_frame_cross_ng"],
                                             0.07, 0.05, 0.45, 0.45, True, prec)
            total.backward()
        from hmmc_b200.graphs import GraphedStep
        g = GraphedStep(run)
        us = timeit(g.replay)
        fl = 2 * 2.0 * 512 * b * (12 * 1024 * 12 + 1024 * 12 + 12 * 1024 + 2 * 1024) * (3 if prec == "bf16x3" else 1)
        print("pre-train head loss fwd+bwd b=%3d %-6s %8.1f us  (%.0f TFLOP/s executed)" % (b, prec, us, fl / us / 1e6))
