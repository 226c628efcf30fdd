"""Feasibility of peer-mapped buffers between the ranks of one node (torchrun, NCCL):
symmetric memory (torch.distributed._symmetric_memory) and legacy CUDA IPC; timings of the NCCL all-gather of the
key block and of a peer-to-peer push of the same bytes."""
import os, sys, time
import torch, torch.distributed as dist
W = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
b, width = 128, 25 * 512 + 2 * 0
rows = W * b
def log(*a):
    if rank == 0: print(*a, flush=True)
def timeit(fn, n=30):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
send = torch.randn(b, width, device=dev)
gathered = torch.empty(rows, width, device=dev)
log("nccl all_gather_into_tensor %d x %d fp32 per rank, W=%d: %.1f us" % (b, width, W, 1e3 * timeit(lambda: dist.all_gather_into_tensor(gathered, send))))
ok_symm = False
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((rows, width), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    log("symm_mem ok: buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:8], "signal pads", len(hdl.signal_pad_ptrs))
    peers = [hdl.get_buffer(r, (rows, width), torch.float32) for r in range(W)]
    def push():
        for r in range(W):
            peers[(rank + r) % W][rank * b:(rank + 1) * b].copy_(send, non_blocking=True)
    us = 1e3 * timeit(push)
    dist.barrier(); torch.cuda.synchronize()
    want = [torch.empty_like(send) for _ in range(W)]
    dist.all_gather(want, send)
    err = max(float((t[r * b:(r + 1) * b] - want[r]).abs().max()) for r in range(W))
    log("symm_mem push of own block to %d peers (copy engine/kernels): %.1f us, max err %.1e" % (W, us, err))
    ok_symm = True
except Exception as e:  # noqa: BLE001
    log("symm_mem FAILED:", repr(e)[:300])
try:
    buf = torch.zeros(rows, width, device=dev)
    h = buf.untyped_storage()._share_cuda_()
    hs = [None] * W
    dist.all_gather_object(hs, h)
    views = []
    for r in range(W):
        if r == rank:
            views.append(buf)
        else:
            st = torch.UntypedStorage._new_shared_cuda(*hs[r])
            views.append(torch.empty(0, dtype=torch.float32, device=st.device).set_(st, 0, (rows, width)))
    log("legacy IPC ok, peer devices:", [str(v.device) for v in views])
    def push2():
        for r in range(W):
            views[(rank + r) % W][rank * b:(rank + 1) * b].copy_(send, non_blocking=True)
    us = 1e3 * timeit(push2)
    dist.barrier(); torch.cuda.synchronize()
    want = [torch.empty_like(send) for _ in range(W)]
    dist.all_gather(want, send)
    err = max(float((buf[r * b:(r + 1) * b] - want[r].to(dev)).abs().max()) for r in range(W))
    log("legacy IPC push: %.1f us, max err %.1e" % (us, err))
except Exception as e:  # noqa: BLE001
    log("legacy IPC FAILED:", repr(e)[:300])
dist.barrier()
dist.destroy_process_group()
