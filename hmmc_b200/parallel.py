"""Process-group plumbing of the head: one process per GPU, torch.distributed (NCCL on
the B200 box, gloo in the CPU tests).  The only exchanges of the path are

  * the embedding all-gather of the fine-tune head and its SUM reduce-scatter backward
    (modules/modeling.py:25-36, 698-700; the backward lives in diffdist==0.1 which is
    not vendored -- contract: rank r receives the sum over ranks of gradient slice r);
  * the key all-gather of the pre-train enqueue (modules/modeling.py:249-258);
  * the [Nt] score / count exchange of the sharded-gallery eval.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


# The all-gather whose consumer is replicated on every rank needs no exchange in its backward
# (_AllGatherCatReplicated); False restores the generic SUM reduce-scatter (tests compare the two).
REPLICATED_GATHER_BWD = True


def _all_gather_into(out, x, async_op=False, group=None):
    """Flat all-gather where the backend has one (NCCL), the list form elsewhere (gloo).  Chosen by backend,
    never by catching an error: a rank whose collective failed must not issue a different one."""
    if dist.get_backend() == "nccl":
        return dist.all_gather_into_tensor(out, x, group=group, async_op=async_op)
    W = dist.get_world_size()
    return dist.all_gather(list(out.chunk(W, dim=0)), x, group=group, async_op=async_op)


def _reduce_scatter_sum(out, g):
    backend = dist.get_backend()
    if backend == "nccl":
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM)
    else:                                            # gloo has no reduce-scatter
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        W, r = world()
        b = g.shape[0] // W
        out.copy_(g[r * b:(r + 1) * b])


class _AllGatherCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        W, _ = world()
        x = x.contiguous()
        out = x.new_empty((W * x.shape[0],) + tuple(x.shape[1:]))
        _all_gather_into(out, x)
        ctx.b = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = g.new_empty((ctx.b,) + tuple(g.shape[1:]))
        _reduce_scatter_sum(out, g)
        return out


class _AllGatherCatReplicated(torch.autograd.Function):
    """All-gather whose consumer is evaluated IDENTICALLY on every rank (the fine-tune head: every rank
    computes the same global loss from the same gathered rows with deterministic kernels).
    The SUM reduce-scatter of the generic backward then adds W bit-identical copies of the same
    gradient, so rank r's result is W x its own rows of its own gradient: no communication."""

    @staticmethod
    def forward(ctx, x):
        W, rank = world()
        x = x.contiguous()
        out = x.new_empty((W * x.shape[0],) + tuple(x.shape[1:]))
        _all_gather_into(out, x)
        ctx.b, ctx.W, ctx.rank = x.shape[0], W, rank
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.b:(ctx.rank + 1) * ctx.b] * float(ctx.W)


def all_gather_cat_replicated(x):
    """all_gather_cat for a consumer that is replicated on all ranks (see _AllGatherCatReplicated).
    Assumption: every rank evaluates a bit-identical loss from the gathered rows (deterministic kernels, same
    inputs); parallel.REPLICATED_GATHER_BWD = False restores the reduce-scatter."""
    W, _ = world()
    if W == 1:
        return x.contiguous()
    if not REPLICATED_GATHER_BWD:
        return _AllGatherCat.apply(x)
    return _AllGatherCatReplicated.apply(x)


def all_gather_cat(x):
    """Differentiable concat-all-gather on dim 0."""
    W, _ = world()
    if W == 1:
        return x.contiguous()
    return _AllGatherCat.apply(x)


@torch.no_grad()
def all_gather_rows(x):
    """Plain all-gather of a [b, w] block -> [W*b, w] (keys for the enqueue; no gradient)."""
    W, _ = world()
    if W == 1:
        return x
    out = x.new_empty((W * x.shape[0],) + tuple(x.shape[1:]))
    _all_gather_into(out, x.contiguous())
    return out


_overlap_group = None


def overlap_group():
    """Process group for collectives that have to run BESIDE a kernel that fills the GPU (the deferred key
    exchange beside the momentum update).  Blocks of two kernels of equal priority are dispatched in launch
    order: the collective's blocks would only get SMs once the other kernel has no blocks left to place, i.e. it
    would run after it, not beside it (measured at 8 GPUs: 0.16 ms of the step exposed).  NCCL: a group whose
    internal stream has high priority; other backends: the default group.  Collective: every rank creates it at
    the same point (the first deferred exchange)."""
    global _overlap_group
    if _overlap_group is None:
        if dist.get_backend() == "nccl":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            _overlap_group = dist.new_group(backend="nccl", pg_options=opts)
        else:
            _overlap_group = dist.group.WORLD
    return _overlap_group


@torch.no_grad()
def all_gather_rows_into(out, x, group=None):
    """All-gather of a [b, w] block into a caller-owned [W*b, w] buffer, ordered on the CURRENT stream (the
    deferred key exchange issues it on a side stream; capturable in a CUDA graph)."""
    W, _ = world()
    if W == 1:
        if out.data_ptr() != x.data_ptr():
            out.copy_(x)
        return out
    _all_gather_into(out, x.contiguous(), group=group)
    return out


class PeerExchange:
    """The deferred key exchange over peer memory instead of an NCCL all-gather (include/hmmc_head.h,
    hmmc_peer_push_rows / hmmc_peer_wait): every rank of the node owns a two-slot receive buffer
    [2][W*b, width] and W flags, mapped into all its peers through symmetric memory; a rank pushes its rows into
    every buffer with plain stores over NVLink and raises its flag, the enqueue starts when all flags are up.
    Measured at 2 GPUs: NCCL all-gather of the 6.5 MB block 41 us, the push 24 us; and the push is an ordinary
    small-grid kernel that runs beside the momentum update.  Construction is collective."""

    MAX_PEERS = 16

    def __init__(self, b, width, device):
        import torch.distributed._symmetric_memory as symm
        W, r = world()
        self.W, self.rank, self.b, self.width = W, r, b, width
        self.slot_stride = W * b * width
        self.recv = symm.empty((2 * W * b, width), dtype=torch.float32, device=device)
        self.flags = symm.empty((self.MAX_PEERS,), dtype=torch.int32, device=device)
        self.recv.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(device)
        name = dist.group.WORLD.group_name
        self.buf_ptrs = [int(p) for p in symm.rendezvous(self.recv, name).buffer_ptrs]
        self.flag_ptrs = [int(p) for p in symm.rendezvous(self.flags, name).buffer_ptrs]
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)      # exchanges completed (device side)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()                       # every rank's flags are zero before anybody raises one

    @staticmethod
    def usable():
        W, _ = world()
        if W < 2 or W > PeerExchange.MAX_PEERS or dist.get_backend() != "nccl":
            return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception:  # noqa: BLE001
            return False
        return True

    def exchange(self, send):
        """Push `send` [b, width] to every rank and wait (on the current stream) for everybody's rows."""
        from . import ops
        ops.peer_push_rows(send, self.buf_ptrs, self.flag_ptrs, self.rank, self.slot_stride, self.epoch, self.done)
        ops.peer_wait(self.flags, self.W, self.epoch)

    def current(self):
        """The rows of the last completed exchange ([W*b, width] view of the slot; reads the counter: syncs)."""
        slot = (int(self.epoch.item()) - 1) & 1
        return self.recv[slot * self.W * self.b:(slot + 1) * self.W * self.b]


def same_node():
    """True when every rank of the default group runs on this host (collective): peer-mapped memory only exists
    between the GPUs of one node."""
    W, _ = world()
    if W == 1:
        return True
    import socket
    names = [None] * W
    dist.all_gather_object(names, socket.gethostname())
    return all(n == names[0] for n in names)


def make_peer_exchange(b, width, device):
    """PeerExchange if every rank can build one, else None on every rank (collective)."""
    ok, px = 0, None
    if PeerExchange.usable() and same_node():
        try:
            px = PeerExchange(b, width, device)
            ok = 1
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn("peer-memory key exchange unavailable, using the NCCL all-gather: %r" % (e,))
    W, _ = world()
    if W > 1:
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item())
    return px if ok else None


def shard_range(n, W=None, r=None):
    """Contiguous [begin, end) slice of n items owned by rank r (gallery / text sharding)."""
    if W is None:
        W, r = world()
    per = (n + W - 1) // W
    lo = min(r * per, n)
    return lo, min(lo + per, n)


@torch.no_grad()
def all_reduce_sum_(t):
    W, _ = world()
    if W > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


@torch.no_grad()
def all_gather_varlen(x, counts):
    """Concatenate per-rank 1-D blocks of different lengths (v2t ranks of the gallery shards)."""
    W, r = world()
    if W == 1:
        return x
    m = max(counts)
    pad = x.new_zeros(m)
    pad[:x.numel()] = x
    out = x.new_empty(W * m)
    _all_gather_into(out, pad)
    return torch.cat([out[i * m:i * m + counts[i]] for i in range(W)])
