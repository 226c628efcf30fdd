"""Does a small kernel on a side stream run BESIDE the momentum update (a kernel that fills the GPU), or after it?
Eager and graph replay, default and high stream priority.   python tools/overlap_probe.py"""
import os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import modeling, synthetic as syn
dev = torch.device("cuda", 0)

class Params(torch.nn.Module):
    def __init__(self, flat, sizes):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(x, requires_grad=False) for x in torch.split(flat, sizes)])
sizes = syn.ema_param_numels()
enc = Params(torch.randn(sum(sizes), device=dev), sizes); enc_k = Params(torch.randn(sum(sizes), device=dev), sizes)
task = types.SimpleNamespace(local_rank=0, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                             contrast_num_negative=1024, max_frames=12, use_frame_fea=True, head_precision="bf16")
m = modeling.BirdPreTrainedModel(modeling.default_cross_config(), task).to(dev)
m.model_pairs = [[enc, enc_k]]
B = int(os.environ.get("B", 1024))
keys = [torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev), torch.randn(B, 512, device=dev),
        torch.randn(B, 12, 512, device=dev), torch.randn(B, 12, 512, device=dev)]

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
    return e0.elapsed_time(e1) / n

def ema():
    with torch.no_grad():
        m._momentum_update()
CHAIN = int(os.environ.get("CHAIN", 0))
tiny = torch.zeros(1 << 18, device=dev)
def small_op():
    if CHAIN:                                   # a chain of CHAIN dependent tiny kernels (~2-3 us each alone)
        for _ in range(CHAIN):
            tiny.add_(1.0)
        return
    with torch.no_grad():
        m._dequeue_and_enqueue(*keys)          # the real enqueue of B keys (W = 1: no gather)
print("ema alone %.4f ms   small alone %.4f ms" % (timeit(ema), timeit(small_op)), flush=True)

for prio in (0, -1):
    side = torch.cuda.Stream(device=dev, priority=prio)
    for first in ("side", "main"):
        def both():
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event(); fork.record(main); side.wait_event(fork)
            if first == "side":
                with torch.cuda.stream(side):
                    small_op()
                    done = torch.cuda.Event(); done.record(side)
                ema()
            else:
                ema()
                with torch.cuda.stream(side):
                    small_op()
                    done = torch.cuda.Event(); done.record(side)
            main.wait_event(done)
        t_eager = timeit(both)
        g = torch.cuda.CUDAGraph()
        cs = torch.cuda.Stream(device=dev)
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            both()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=cs, capture_error_mode="thread_local"):
                both()
        torch.cuda.synchronize()
        t_graph = timeit(g.replay)
        print("priority %2d, %s issued first: eager %.4f ms   graph %.4f ms" % (prio, first, t_eager, t_graph), flush=True)
