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
