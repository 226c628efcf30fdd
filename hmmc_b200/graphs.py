"""CUDA-graph capture of a whole head step (EMA + loss forward/backward + key exchange + enqueue).

The step is a fixed sequence of ~8 launches with static shapes, and issuing it from Python costs
more host time than the GPU needs to run it once ranks and collectives are involved; a captured
graph replays it with one launch.  All device state the step touches (queues, queue_ptr, momentum
parameters, workspaces) is updated in place by the kernels, so replaying is equivalent to calling
the step again.

    g = GraphedStep(lambda: step(static_inputs))      # warm-up calls + capture
    for batch in loader:
        static_inputs.copy_(batch); g.replay(); use(g.outputs)

The warm-up calls are REAL steps: each one updates the momentum parameters, enqueues the step's keys and
advances queue_ptr, exactly like a call outside the graph (the capture pass itself executes nothing).  A
training loop that must not take extra steps passes ``warmup=0`` after having run its first steps eagerly
(those already sized the workspaces and built the operand copies the capture relies on).  Drop every reference
to the eager steps' loss tensors before capturing: a live loss keeps its autograd graph alive, whose AccumulateGrad
nodes belong to the stream of that eager step, and the capture's backward would have to synchronise with it -
which invalidates the capture.
"""
import gc

import torch


class GraphedStep:
    def __init__(self, fn, warmup=3):
        """``fn()`` runs one step on static input tensors and returns a tensor (or tuple of
        tensors) to keep, e.g. the loss.  Gradients land in the ``.grad`` of the static inputs."""
        self.fn = fn
        if warmup > 0:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    fn()
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process keep making CUDA calls while this one captures - the NCCL
        # watchdog polls the events of earlier collectives - and must not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.outputs = fn()

    def replay(self):
        self.graph.replay()
        return self.outputs

    def release(self):
        """Destroy the executable graph (and the NCCL work it holds) - call before
        torch.distributed.destroy_process_group(): a live graph that captured a collective keeps the
        communicator busy and the destroy never returns."""
        torch.cuda.synchronize()
        self.outputs = None
        self.fn = None
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
        gc.collect()
        torch.cuda.synchronize()
