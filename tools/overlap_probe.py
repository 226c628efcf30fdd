"""Does the loss (tensor-bound GEMMs) hide under the EMA (HBM-bound) when both run at once?
Times EMA + loss sequentially and concurrently on two streams, for several SM budgets of the GEMM grids
and both issue orders."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn

dev = torch.device("cuda")
sizes = syn.ema_param_numels()
flat = torch.randn(sum(sizes), device=dev)
flat_k = torch.randn(sum(sizes), device=dev)
tab = ops.EmaTable(list(zip(torch.split(flat, sizes), torch.split(flat_k, sizes))))
b, F, D, K = 128, 12, 512, 1024
inp = {n: torch.from_numpy(x).to(dev) for n, x in syn.pretrain_inputs(b, F=F, D=D, seed=2).items()}
qs = {n: torch.from_numpy(x).to(dev) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
qn = ("v_fea", "title_fea", "frame_fea", "frame_pred")
for n in qn:
    inp[n].requires_grad_(True)


def loss():
    total, _ = ops.pretrain_head(inp["v_fea"], inp["title_fea"], inp["frame_fea"], inp["frame_pred"], inp["v_fea_k"],
                                 inp["title_fea_k"], inp["frame_fea_k"], inp["frame_proj_k"], qs["queue_v_cross_ng"],
                                 qs["queue_title_cross_ng"], qs["queue_frame_proj_ng"], qs["queue_frame_cross_ng"],
                                 0.07, 0.05, 0.45, 0.45, True, "bf16")
    return total


def ema():
    tab.run(0.99)


A = torch.cuda.Stream()
B = torch.cuda.Stream(priority=-1)
main = torch.cuda.current_stream()


def timeit(fn, reps=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def seq():
    with torch.no_grad():
        ema()
        loss()


def par(order):
    def run():
        e = torch.cuda.Event()
        e.record(main)
        A.wait_event(e)
        B.wait_event(e)
        with torch.no_grad():
            for who in order:
                if who == "loss":
                    with torch.cuda.stream(B):
                        loss()
                else:
                    with torch.cuda.stream(A):
                        ema()
        main.wait_stream(A)
        main.wait_stream(B)
    return run


with torch.no_grad():
    print("ema alone   %.1f us" % timeit(ema))
    print("loss alone  %.1f us (forward + fused backward kernels, no autograd)" % timeit(loss))
    print("sequential  %.1f us" % timeit(seq))
    for reserved in (0, 48, 74, 100, 120):
        ops.set_reserved_sms(reserved)
        for order in (("loss", "ema"), ("ema", "loss")):
            print("concurrent  reserved=%3d  issue order %s: %.1f us" % (reserved, "+".join(order), timeit(par(order))))
    ops.set_reserved_sms(0)
