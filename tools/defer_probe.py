"""Step time of the pre-train head (EMA + loss + enqueue) under the immediate and the deferred key-exchange
schedule, eager and replayed from a CUDA graph.  One process = one GPU; under torchrun it uses NCCL.
    python tools/defer_probe.py            |  torchrun --nproc-per-node 2 tools/defer_probe.py"""
import os, sys, time, types
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import modeling, synthetic as syn
from hmmc_b200.graphs import GraphedStep

W = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
b, F, D = int(os.environ.get("B_PER", 128)), 12, 512
K = max(1024, b * W)

class Params(torch.nn.Module):
    def __init__(self, flat, sizes):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(x, requires_grad=False) for x in torch.split(flat, sizes)])
sizes = syn.ema_param_numels()
enc = Params(torch.randn(sum(sizes), device=dev), sizes); enc_k = Params(torch.randn(sum(sizes), device=dev), sizes)
order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k", "frame_proj_k"]
inp = syn.pretrain_inputs(b, F=F, D=D, seed=100 + rank)

PEER = None


def run(defer, graph, ema=True):
    task = types.SimpleNamespace(local_rank=local, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                 contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision="bf16",
                                 defer_enqueue=defer, peer_exchange=PEER)
    m = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).to(dev)
    m.model_pairs = [[enc, enc_k]]
    with torch.no_grad():
        for n, x in syn.queues(K, F=F, D=D, seed=3).items():
            getattr(m, n).copy_(torch.from_numpy(x))
    t = {n: torch.from_numpy(inp[n]).to(dev).requires_grad_(n in order[:4]) for n in order}
    def step():
        for n in order[:4]:
            t[n].grad = None
        m.start_pending_exchange()
        if ema:
            with torch.no_grad():
                m._momentum_update()
        loss = m.head_loss(*[t[n] for n in order])
        loss.backward()
        return loss
    for _ in range(3):
        loss = step()
    del loss
    fn = step
    if graph:
        g = GraphedStep(step, warmup=0)
        fn = g.replay
    for _ in range(5):
        fn()
    if W > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if os.environ.get("TRACE") and defer and ema:        # every rank replays (the exchange is collective)
        # kernel timeline of a few replays (CUPTI through torch.profiler): stream, start, duration
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                fn()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        emas = [e for e in evs if "ema_multi" in e.name]
        if len(emas) >= 3 and rank == 0:
            t0, t1 = emas[1].time_range.start, emas[2].time_range.start
            print("---- timeline of one replay (us from the start of the momentum update), peer=%s" % (PEER,))
            for e in evs:
                if t0 - 50 <= e.time_range.start < t1 - 50:
                    print("%8.1f  +%7.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]), flush=True)
    ms = torch.tensor([e0.elapsed_time(e1) / 100], device=dev)
    if W > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    m.flush_pending_enqueue()
    torch.cuda.synchronize()
    if graph:
        g.release()
    return float(ms)

def run_cfg(defer, peer, ema=True):
    global PEER
    PEER = peer
    return run(defer, True, ema)

for ema in (True, False):
    for name, defer, peer in (("immediate (gather + enqueue at the end of the step)", False, None),
                              ("deferred, exchange over peer memory", True, None),
                              ("deferred, NCCL all-gather", True, False)):
        if W == 1 and peer is False:
            continue
        ms = run_cfg(defer, peer, ema)
        if rank == 0:
            print("W=%d %s momentum update, %s: %.4f ms/step" % (W, "with" if ema else "WITHOUT", name, ms), flush=True)
if W > 1:
    dist.destroy_process_group()
