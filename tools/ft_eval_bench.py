"""Device time (CUDA-graph replay) of the fine-tune head at B = 256 and of the config-2 eval (1000 x 1000 x 12)."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import modeling, ops, retrieval, synthetic as syn
from hmmc_b200.graphs import GraphedStep

def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best

cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
for B in (256,):
    t, v, fr = [cu(x) for x in syn.finetune_inputs(B, seed=1)]
    for prec in ("bf16x3", "bf16"):
        p = ops.resolve_precision(prec)
        g = GraphedStep(lambda: ops.sym_ce_raw(t, v, fr, 100.0, 0.85, 0.15, p, True)[0])
        print("fine-tune head fwd+bwd B=%d %-6s graph replay %7.1f us  loss %.6f" % (B, prec, timeit(g.replay), float(g.outputs)), flush=True)
T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=4)
T, V, Fr = cu(T), cu(V), cu(Fr)
for prec in ("bf16x3", "bf16"):
    task = types.SimpleNamespace(local_rank=0, top_frames=2, use_frame_fea=True, head_precision=prec)
    m = modeling.BirdModel(modeling.default_cross_config(), task)
    g = GraphedStep(lambda: ops.rank_count(retrieval.similarity_matrix(m, T, V, Fr))[0])
    print("eval 1000x1000x12 sim + ranks %-6s graph replay %7.1f us" % (prec, timeit(g.replay)), flush=True)
