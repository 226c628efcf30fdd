"""Bisect which part of the head fails under CUDA-graph capture (one GPU)."""
import os, sys, types, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import modeling, ops, synthetic as syn
dev = torch.device("cuda", 0)
b, F, D, K = 8, 12, 128, int(os.environ.get("K", "128"))

def try_capture(name, fn, warm=1):
    try:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("OK   ", name, flush=True)
    except Exception as e:
        print("FAIL ", name, repr(e)[:160].replace("\n", " "), flush=True)
        torch.cuda.synchronize()

A = torch.randn(64, 128, device=dev); Bm = torch.randn(96, 128, device=dev)
try_capture("gemm_f32", lambda: ops.gemm_f32(A, Bm))
inp = syn.pretrain_inputs(b, F=F, D=D, seed=70)
qs = {n: torch.from_numpy(x).to(dev) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
t = {n: torch.from_numpy(x).to(dev) for n, x in inp.items()}
for prec in ("fp32", "bf16", "bf16x3"):
    p = ops.resolve_precision(prec)
    try_capture("infonce fwd %s" % prec, lambda: ops.infonce_raw(t["v_fea"], t["title_fea_k"], ops.POS_PAIR, b, 1, 1, qs["queue_title_cross_ng"], 0.07, 1.0, p, False))
    try_capture("infonce fwd+bwd %s" % prec, lambda: ops.infonce_raw(t["v_fea"], t["title_fea_k"], ops.POS_PAIR, b, 1, 1, qs["queue_title_cross_ng"], 0.07, 1.0, p, True))
    def head(prec=prec):
        tt = {n: (x.detach().requires_grad_(n in ("v_fea", "title_fea", "frame_fea", "frame_pred"))) for n, x in t.items()}
        total, _ = ops.pretrain_head(tt["v_fea"], tt["title_fea"], tt["frame_fea"], tt["frame_pred"], tt["v_fea_k"], tt["title_fea_k"],
                                     tt["frame_fea_k"], tt["frame_proj_k"], qs["queue_v_cross_ng"], qs["queue_title_cross_ng"],
                                     qs["queue_frame_proj_ng"], qs["queue_frame_cross_ng"], 0.07, 0.05, 0.45, 0.45, True, prec)
        total.backward()
        return total
    try_capture("pretrain_head fwd+bwd %s" % prec, head)
    for defer in (False, True):
        task = types.SimpleNamespace(local_rank=0, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                     contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision=prec, defer_enqueue=defer)
        m = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).to(dev)
        order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k", "frame_proj_k"]
        static = {n: torch.from_numpy(inp[n]).to(dev).requires_grad_(n in order[:4]) for n in order}
        def step():
            for n in order[:4]:
                static[n].grad = None
            loss = m.head_loss(*[static[n] for n in order])
            loss.backward()
            return loss
        try_capture("model.head_loss %s defer=%s (pending staged)" % (prec, defer), step)
        if defer:
            def warm_then_flush():
                step(); m.flush_pending_enqueue()
            warm_then_flush(); torch.cuda.synchronize()
            try_capture("model.head_loss %s defer=True after flush" % prec, step, warm=0)
