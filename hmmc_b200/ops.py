"""Torch-facing wrappers of the C ABI: tensors in, tensors out, autograd where the
reference differentiates.  Device memory and streams come from torch; all arithmetic
happens in libhmmc_head.so on the caller's current CUDA stream.
"""
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib import (PREC_BF16, PREC_BF16X3, PREC_FP32, PRECISIONS, POS_FRAME_NEIGHBOUR, POS_FRAMES_TO_ONE,
                   POS_ONE_TO_FRAMES, POS_PAIR, HmmcError, hmmc_head_schedule, hmmc_pretrain_io, hmmc_queue)

DEFAULT_PRECISION = os.environ.get("HMMC_PRECISION", "bf16x3")


def resolve_precision(p=None):
    if p is None:
        p = DEFAULT_PRECISION
    if isinstance(p, str):
        if p not in PRECISIONS:
            raise ValueError("unknown precision %r (choose from %s)" % (p, sorted(PRECISIONS)))
        return PRECISIONS[p]
    return int(p)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32c(t, name):
    """fp32 contiguous CUDA view of ``t`` (the reference's encoders end in .float())."""
    if not t.is_cuda:
        raise HmmcError("%s must be a CUDA tensor: libhmmc_head has no CPU path" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_workspaces = {}          # (device index, stream handle) -> [buffer, baked into a CUDA graph?]
_graph_pinned = []        # buffers a capture has baked into a graph: kept alive for the life of the process


def workspace(device, nbytes):
    """Scratch buffer of the CURRENT stream (calls on one stream are ordered, so they may share it; calls on
    different streams - side streams, eval threads, a capture stream - never do).  A buffer that a CUDA-graph
    capture has seen is never freed when a later call needs a larger one: replays keep writing to it."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ent = _workspaces.get(key)
    if ent is None or ent[0].numel() < nbytes:
        if ent is not None and ent[1]:
            _graph_pinned.append(ent[0])
        ent = [torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device), False]
        _workspaces[key] = ent
    if torch.cuda.is_current_stream_capturing():
        ent[1] = True
    return ent[0]


def device_check():
    _lib.check(_lib.load().hmmc_device_check(), "hmmc_device_check")


# ----------------------------------------------------------------------------- primitives

def rownorm_pack(x, eps, planes, want_xhat=False, want_packed=True):
    lib = _lib.load()
    x = _f32c(x, "x")
    R, D = x.shape
    xhat = torch.empty_like(x) if want_xhat else None
    inv = torch.empty(R, dtype=torch.float32, device=x.device)
    packed = torch.empty(R, planes * D, dtype=torch.bfloat16, device=x.device) if want_packed else None
    _lib.check(lib.hmmc_rownorm_pack(_p(x), R, D, D, float(eps), planes, _p(xhat), _p(inv), _p(packed),
                                     planes * D, _stream()), "hmmc_rownorm_pack")
    return xhat, inv, packed


def gemm_f32(A, B, alpha=1.0):
    """C = alpha * A @ B.T for 2-D fp32 tensors of any strides (CUDA-core path)."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    _lib.check(lib.hmmc_gemm_f32(_p(A), A.stride(0), A.stride(1), _p(B), B.stride(0), B.stride(1), _p(C), N,
                                 M, N, K, float(alpha), _stream()), "hmmc_gemm_f32")
    return C


def umma_gemm_nt(Ap, Bp, K, planes, alpha=1.0, tiling=0):
    """C = alpha * A . B^T on tcgen05 from plane-packed bf16 operands [rows, planes*K].
    tiling: 0 = chosen from the shape, 128 / 256 = single-CTA tile width, 512 = CTA-pair kernel."""
    lib = _lib.load()
    M, N = Ap.shape[0], Bp.shape[0]
    C = torch.empty(M, N, dtype=torch.float32, device=Ap.device)
    _lib.check(lib.hmmc_umma_gemm_nt_tiled(_p(Ap), Ap.stride(0), _p(Bp), Bp.stride(0), _p(C), N, M, N, K, planes,
                                           float(alpha), int(tiling), _stream()), "hmmc_umma_gemm_nt_tiled")
    return C


# ----------------------------------------------------------------------------- queues

class QueueState:
    """Derived bf16 operand copies of one negative queue buffer ([D, Kq] fp32, the reference's state-dict
    layout, modules/modeling.py:138-149).  One state per buffer (held weakly), with the copies of each plane
    count (1 = bf16, 2 = bf16x3) built on first use.  Staleness is tracked explicitly: `gen` counts writes to the
    buffer made by the enqueue kernel (which keeps only the plane count in use in step), `buf._version` catches
    writes made through torch (load_state_dict, copy_)."""

    def __init__(self, buf):
        assert buf.dim() == 2 and buf.dtype == torch.float32 and buf.is_cuda and buf.is_contiguous()
        self.ref = weakref.ref(buf)
        self.D, self.Kq = buf.shape
        self.gen = 0
        self.graph_planes = None      # plane count a captured enqueue keeps in step during replays
        self.packs = {}               # planes -> [pack_kd, pack_dk, gen, version]

    def pack(self, buf, planes):
        ent = self.packs.get(planes)
        if ent is None:
            ent = [torch.empty(self.Kq, planes * self.D, dtype=torch.bfloat16, device=buf.device),
                   torch.empty(self.D, planes * self.Kq, dtype=torch.bfloat16, device=buf.device), -1, -1]
            self.packs[planes] = ent
        # replays of a captured enqueue change the buffer without passing through Python: copies of the
        # other plane count are rebuilt on every use from then on
        unsure = self.graph_planes is not None and self.graph_planes != planes
        if ent[2] != self.gen or ent[3] != buf._version or unsure:
            q = hmmc_queue(buf.data_ptr(), ent[0].data_ptr(), ent[1].data_ptr(), self.D, self.Kq, planes, 0)
            _lib.check(_lib.load().hmmc_queue_pack(ctypes.byref(q), _stream()), "hmmc_queue_pack")
            ent[2], ent[3] = self.gen, buf._version
        return ent

    def repack(self, buf, planes):
        """Rebuild the copies of `planes` from the buffer unconditionally; returns (pack_kd, pack_dk)."""
        if planes in self.packs:
            self.packs[planes][2] = -1
        ent = self.pack(buf, planes)
        return ent[0], ent[1]

    def struct(self, buf, planes):
        ent = self.pack(buf, planes)
        return hmmc_queue(buf.data_ptr(), ent[0].data_ptr(), ent[1].data_ptr(), self.D, self.Kq, planes, 0)

    def wrote(self, buf, planes):
        """The enqueue kernel updated the buffer and the copies of `planes` (None: the buffer only)."""
        self.gen += 1
        if planes in self.packs:
            self.packs[planes][2], self.packs[planes][3] = self.gen, buf._version
        if torch.cuda.is_current_stream_capturing():
            self.graph_planes = planes


_queue_states = {}        # id(buffer) -> QueueState, dropped when the buffer dies


def queue_state(buf):
    st = _queue_states.get(id(buf))
    if st is None or st.ref() is not buf:
        if not buf.is_contiguous():
            raise HmmcError("queue buffers must be contiguous [D, Kq] tensors")
        st = QueueState(buf)
        key = id(buf)
        _queue_states[key] = st
        weakref.finalize(buf, _queue_states.pop, key, None)
    return st


def _queue_struct(buf, prec):
    """(hmmc_queue for this precision, state or None); FP32 precision needs no copies."""
    if prec == PREC_FP32:
        D, Kq = buf.shape
        return hmmc_queue(buf.data_ptr(), 0, 0, D, Kq, 1, 0), queue_state(buf)
    st = queue_state(buf)
    return st.struct(buf, 2 if prec == PREC_BF16X3 else 1), st


# ----------------------------------------------------------------------------- InfoNCE vs queue

def infonce_raw(q2d, keys2d, pos_mode, b, Fq, Fk, queue_buf, temperature, weight, prec, need_grad):
    """Returns (loss 0-d tensor, dq or None)."""
    lib = _lib.load()
    D = q2d.shape[1]
    qs, keep = _queue_struct(queue_buf, prec)
    R = b * Fq
    nbytes = lib.hmmc_infonce_workspace_bytes(R, D, qs.Kq, prec)
    ws = workspace(q2d.device, nbytes)
    loss = torch.zeros((), dtype=torch.float32, device=q2d.device)
    dq = torch.empty_like(q2d) if need_grad else None
    _lib.check(lib.hmmc_infonce_queue_fwd_bwd(_p(q2d), _p(keys2d), pos_mode, b, Fq, Fk, D, ctypes.byref(qs),
                                              float(temperature), float(weight), prec, _p(loss), _p(dq),
                                              _p(ws), ws.numel(), _stream()), "hmmc_infonce_queue_fwd_bwd")
    return loss, dq


class _InfoNCEFn(torch.autograd.Function):
    """loss = weight * sum over the positive terms of mean-CE([l+, l_neg]/T, 0); forward
    and backward run in one pass (the gradient w.r.t. q is produced with the loss)."""

    @staticmethod
    def forward(ctx, q, keys, queue_buf, pos_mode, b, Fq, Fk, temperature, weight, prec):
        q2d = _f32c(q, "q").reshape(b * Fq, -1)
        k2d = _f32c(keys, "k").reshape(b * Fk, -1)
        need = q.requires_grad
        loss, dq = infonce_raw(q2d, k2d, pos_mode, b, Fq, Fk, queue_buf, temperature, weight, prec, need)
        ctx.q_shape = q.shape
        ctx.q_dtype = q.dtype
        if need:
            ctx.save_for_backward(dq)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dq,) = ctx.saved_tensors
        gq = (dq * g).reshape(ctx.q_shape).to(ctx.q_dtype)
        return gq, None, None, None, None, None, None, None, None, None


def infonce(q, keys, queue_buf, pos_mode, b, Fq, Fk, temperature, weight=1.0, precision=None):
    return _InfoNCEFn.apply(q, keys, queue_buf, pos_mode, b, Fq, Fk, float(temperature), float(weight),
                            resolve_precision(precision))


def scale_inplace(tensors, g):
    """t *= g for every tensor (g: 0-d device tensor), one launch."""
    lib = _lib.load()
    ts = [t for t in tensors if t is not None and t.numel()]
    if not ts:
        return tensors
    g = g.reshape(()).to(device=ts[0].device, dtype=torch.float32)
    n = len(ts)
    ptrs = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in ts])
    nums = (ctypes.c_int64 * n)(*[t.numel() for t in ts])
    _lib.check(lib.hmmc_scale_tensors(ptrs, nums, n, _p(g), _stream()), "hmmc_scale_tensors")
    return tensors


class _PretrainHeadFn(torch.autograd.Function):
    """w_fam*FAM + w_vtm*VTM + w_ftm*FTM of the pre-train head, forward and backward in one C
    call (four launches).  Returns (total, parts[3]); only ``total`` is differentiable."""

    @staticmethod
    def forward(ctx, v_fea, title_fea, frame_fea, frame_pred, v_fea_k, title_fea_k, frame_fea_k, frame_proj_k,
                q_v, q_title, q_frame_proj, q_frame_cross, temperature, w_fam, w_vtm, w_ftm, use_frame_fea, prec,
                release_event=None):
        lib = _lib.load()
        b, F, D = frame_fea.shape
        qin = [v_fea, title_fea, frame_fea, frame_pred]
        t = [_f32c(x, "embedding") for x in qin + [v_fea_k, title_fea_k, frame_fea_k, frame_proj_k]]
        need = [x.requires_grad for x in qin]
        any_grad = any(need)
        grads = [torch.empty_like(x) if any_grad else None for x in t[:4]]
        io = hmmc_pretrain_io(*[x.data_ptr() for x in t], *[(g.data_ptr() if g is not None else 0) for g in grads])
        structs = [_queue_struct(qb, prec) for qb in (q_v, q_title, q_frame_proj, q_frame_cross)]
        K = q_v.shape[1]
        nbytes = lib.hmmc_pretrain_head_workspace_bytes(b, F, D, K, prec)
        ws = workspace(t[0].device, nbytes)
        losses = torch.empty(4, dtype=torch.float32, device=t[0].device)
        sched = hmmc_head_schedule(0, 0, None)        # both phases, all SMs
        if release_event is not None:
            release_event.record()               # materialises the handle; the library records it again later
            sched.queues_released = release_event.cuda_event
        _lib.check(lib.hmmc_pretrain_head_fwd_bwd_sched(ctypes.byref(io), b, F, D,
                                                        *[ctypes.byref(s[0]) for s in structs], float(temperature),
                                                        float(w_fam), float(w_vtm), float(w_ftm),
                                                        int(bool(use_frame_fea)), prec, _p(losses), ctypes.byref(sched),
                                                        _p(ws), ws.numel(), _stream()),
                   "hmmc_pretrain_head_fwd_bwd_sched")
        ctx.need = need
        ctx.meta = [(x.shape, x.dtype) for x in qin]
        if any_grad:
            ctx.save_for_backward(*grads)
        total, parts = losses[0], losses[1:]
        ctx.mark_non_differentiable(parts)
        return total, parts

    @staticmethod
    def backward(ctx, g, _gparts):
        grads = list(ctx.saved_tensors)
        if getattr(ctx, "consumed", False):
            raise HmmcError("the fused head's gradients were already consumed (retain_graph is not supported)")
        ctx.consumed = True
        scale_inplace([grads[i] for i in range(4) if ctx.need[i]], g)      # in place: the buffers are ours
        out = [grads[i].reshape(ctx.meta[i][0]).to(ctx.meta[i][1]) if ctx.need[i] else None for i in range(4)]
        return tuple(out) + (None,) * 15


def pretrain_head(v_fea, title_fea, frame_fea, frame_pred, v_fea_k, title_fea_k, frame_fea_k, frame_proj_k,
                  q_v, q_title, q_frame_proj, q_frame_cross, temperature, w_fam, w_vtm, w_ftm, use_frame_fea=True,
                  precision=None, release_event=None):
    """release_event: a torch.cuda.Event the library records on the current stream right after the last kernel
    that reads the queues (the enqueue may then overlap the rest of the loss on another stream)."""
    return _PretrainHeadFn.apply(v_fea, title_fea, frame_fea, frame_pred, v_fea_k, title_fea_k, frame_fea_k,
                                 frame_proj_k, q_v, q_title, q_frame_proj, q_frame_cross, float(temperature),
                                 float(w_fam), float(w_vtm), float(w_ftm), bool(use_frame_fea),
                                 resolve_precision(precision), release_event)


class PretrainHeadState:
    """What pretrain_head_begin leaves for pretrain_head_end (buffers, argument structs, the workspace)."""
    __slots__ = ("t", "grads", "structs", "keep", "ws", "losses", "dims", "args", "need", "meta", "queues")


def _head_call(state, keys, phase, stream, reserved_sms=0):
    lib = _lib.load()
    b, F, D, K, prec = state.dims
    kp = [k.data_ptr() for k in keys] if keys is not None else [0, 0, 0, 0]
    io = hmmc_pretrain_io(*[x.data_ptr() for x in state.t], *kp,
                          *[(g.data_ptr() if g is not None else 0) for g in state.grads])
    temperature, w_fam, w_vtm, w_ftm, use_frame_fea = state.args
    sched = hmmc_head_schedule(phase, int(reserved_sms), None)
    _lib.check(lib.hmmc_pretrain_head_fwd_bwd_sched(ctypes.byref(io), b, F, D, *[ctypes.byref(s) for s in state.structs],
                                                    float(temperature), float(w_fam), float(w_vtm), float(w_ftm),
                                                    int(bool(use_frame_fea)), prec, _p(state.losses),
                                                    ctypes.byref(sched), _p(state.ws), state.ws.numel(), stream),
               "hmmc_pretrain_head_fwd_bwd_sched")


def pretrain_head_begin(v_fea, title_fea, frame_fea, frame_pred, q_v, q_title, q_frame_proj, q_frame_cross,
                        temperature, w_fam, w_vtm, w_ftm, use_frame_fea=True, precision=None, stream=None,
                        reserved_sms=0):
    """First half of the fused pre-train head: normalise the queries and run both GEMM passes against the
    queues (everything that does not need the keys), issued on `stream` (a torch.cuda.Stream; default: the
    current one), with the persistent GEMM grids leaving `reserved_sms` SMs free for whatever runs beside
    them.  All buffers are allocated on the CURRENT stream; the caller orders `stream` after the queries and
    joins it before pretrain_head_end.  In the reference the queries exist before
    _momentum_update() and the key encoders run (modules/modeling.py:340-377)."""
    prec = resolve_precision(precision)
    st = PretrainHeadState()
    b, F, D = frame_fea.shape
    qin = [v_fea, title_fea, frame_fea, frame_pred]
    st.t = [_f32c(x, "embedding") for x in qin]
    st.need = [x.requires_grad for x in qin] if torch.is_grad_enabled() else [False] * 4
    any_grad = any(st.need)
    st.grads = [torch.empty_like(x) if any_grad else None for x in st.t]
    pairs = [_queue_struct(qb, prec) for qb in (q_v, q_title, q_frame_proj, q_frame_cross)]
    st.structs = [p_[0] for p_ in pairs]
    st.keep = pairs
    K = q_v.shape[1]
    lib = _lib.load()
    nbytes = lib.hmmc_pretrain_head_workspace_bytes(b, F, D, K, prec)
    st.ws = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=st.t[0].device)    # private: must survive until end
    st.losses = torch.empty(4, dtype=torch.float32, device=st.t[0].device)
    st.dims = (b, F, D, K, prec)
    st.args = (temperature, w_fam, w_vtm, w_ftm, use_frame_fea)
    st.meta = [(x.shape, x.dtype) for x in qin]
    st.queues = (q_v, q_title, q_frame_proj, q_frame_cross)
    if stream is not None:
        ready = torch.cuda.Event()          # queries, fresh buffers and (re)packed queues are ready here
        ready.record()
        stream.wait_event(ready)
    _head_call(st, None, 1, stream.cuda_stream if stream is not None else _stream(), reserved_sms)
    return st


class _PretrainHeadEndFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_fea, title_fea, frame_fea, frame_pred, state, v_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        keys = [_f32c(x, "key embedding") for x in (v_fea_k, title_fea_k, frame_fea_k, frame_proj_k)]
        _head_call(state, keys, 2, _stream())
        ctx.need = state.need
        ctx.meta = state.meta
        if any(state.need):
            ctx.save_for_backward(*state.grads)
        total, parts = state.losses[0], state.losses[1:]
        ctx.mark_non_differentiable(parts)
        return total, parts

    @staticmethod
    def backward(ctx, g, _gparts):
        grads = list(ctx.saved_tensors)
        if getattr(ctx, "consumed", False):
            raise HmmcError("the fused head's gradients were already consumed (retain_graph is not supported)")
        ctx.consumed = True
        scale_inplace([grads[i] for i in range(4) if ctx.need[i]], g)
        out = [grads[i].reshape(ctx.meta[i][0]).to(ctx.meta[i][1]) if ctx.need[i] else None for i in range(4)]
        return tuple(out) + (None,) * 5


def pretrain_head_end(state, v_fea, title_fea, frame_fea, frame_pred, v_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
    """Second half: positives, losses and gradients, on the current stream.  The four query tensors are passed
    again only so that autograd routes the gradients to them.  Returns (total, parts[3])."""
    return _PretrainHeadEndFn.apply(v_fea, title_fea, frame_fea, frame_pred, state, v_fea_k, title_fea_k, frame_fea_k,
                                    frame_proj_k)


# ----------------------------------------------------------------------------- EMA / enqueue

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class EmaTable:
    """Device-side pointer table of the (param_k, param) pairs of _momentum_update."""

    def __init__(self, pairs):
        lib = _lib.load()
        be = lib.hmmc_ema_block_elems()
        pk_ptrs, p_ptrs, numels, dts, offs = [], [], [], [], [0]
        dev = None
        for p, pk in pairs:
            if p.dtype not in _DT or pk.dtype != p.dtype:
                raise HmmcError("EMA: unsupported / mismatching dtypes %s vs %s" % (p.dtype, pk.dtype))
            if not (p.is_cuda and pk.is_cuda and p.is_contiguous() and pk.is_contiguous()):
                raise HmmcError("EMA: parameters must be contiguous CUDA tensors")
            if p.numel() == 0:
                continue
            dev = pk.device
            pk_ptrs.append(pk.data_ptr())
            p_ptrs.append(p.data_ptr())
            numels.append(p.numel())
            dts.append(_DT[p.dtype])
            offs.append(offs[-1] + (p.numel() + be - 1) // be)
        self.n = len(numels)
        self.total_blocks = offs[-1]
        self.total_elems = sum(numels)
        self.signature = (tuple(pk_ptrs), tuple(p_ptrs))
        if self.n:
            u64 = lambda v: torch.tensor(v, dtype=torch.int64).to(dev)   # pointers < 2^63
            self.pk = u64(pk_ptrs)
            self.p = u64(p_ptrs)
            self.numels = u64(numels)
            self.dtypes = torch.tensor(dts, dtype=torch.int32).to(dev)
            self.offs = u64(offs)

    def still_valid(self, pairs):
        """True while every (param, param_k) pair still lives where the device table points (.half(), .to(),
        load_state_dict(assign=True) or `param_k.data = ...` re-allocate): a host-side compare of all pointers."""
        live = [(p, pk) for p, pk in pairs if p.numel()]
        return (tuple(pk.data_ptr() for _, pk in live), tuple(p.data_ptr() for p, _ in live)) == self.signature

    def run(self, m):
        if not self.n:
            return
        import numpy as np
        m32 = float(np.float32(m))
        omm32 = float(np.float32(1.0 - m))      # (1. - m) in python double, then fp32 (modeling.py:242)
        _lib.check(_lib.load().hmmc_ema_multi(_p(self.pk), _p(self.p), _p(self.numels), _p(self.dtypes),
                                              _p(self.offs), self.n, self.total_blocks, m32, omm32, _stream()),
                   "hmmc_ema_multi")


def pack_rows(tensors, out=None, staged=None, norm_dim=0):
    """[rows, w_i] blocks -> one [rows, sum w_i] fp32 buffer (send buffer of the all-gather).
    staged: optional int32[1] device tensor the kernel sets to 1 ("keys staged", consumed by enqueue).
    norm_dim = D: every D-vector is written L2-normalised (the enqueue's normalisation, done at the source)."""
    lib = _lib.load()
    ts = [_f32c(t, "block").reshape(t.shape[0], -1) for t in tensors]
    rows = ts[0].shape[0]
    widths = [t.shape[1] for t in ts]
    if out is None:
        out = torch.empty(rows, sum(widths), dtype=torch.float32, device=ts[0].device)
    n = len(ts)
    ptrs = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in ts])
    ws = (ctypes.c_int32 * n)(*widths)
    _lib.check(lib.hmmc_pack_rows(ptrs, ws, n, rows, _p(out), _p(staged), int(norm_dim), _stream()), "hmmc_pack_rows")
    return out


class _PackRowsFn(torch.autograd.Function):
    """pack_rows with a gradient: one launch forward, one launch (unpack_rows) backward."""

    @staticmethod
    def forward(ctx, *tensors):
        ctx.shapes = [tuple(t.shape) for t in tensors]
        ctx.dtypes = [t.dtype for t in tensors]
        return pack_rows(tensors)

    @staticmethod
    def backward(ctx, g):
        rows = ctx.shapes[0][0]
        widths = [int(torch.Size(s).numel() // rows) for s in ctx.shapes]
        outs = unpack_rows(_f32c(g, "packed gradient"), widths)
        return tuple(o.reshape(s).to(d) for o, s, d in zip(outs, ctx.shapes, ctx.dtypes))


def pack_rows_autograd(tensors):
    """[rows, ...] blocks -> [rows, sum of widths] fp32, differentiable (send buffer of a differentiable gather)."""
    return _PackRowsFn.apply(*tensors)


def unpack_rows(packed, widths):
    lib = _lib.load()
    rows = packed.shape[0]
    outs = [torch.empty(rows, w, dtype=torch.float32, device=packed.device) for w in widths]
    n = len(outs)
    ptrs = (ctypes.c_uint64 * n)(*[t.data_ptr() for t in outs])
    ws = (ctypes.c_int32 * n)(*widths)
    _lib.check(lib.hmmc_unpack_rows(_p(packed), ptrs, ws, n, rows, _stream()), "hmmc_unpack_rows")
    return outs


def enqueue(gathered, W, b, F, D, queue_bufs5, queue_ptr, ptr_host, K, prec, direct=None, staged=None, slot=None,
            prenormalised=False):
    """queue_bufs5 order: v, tag, title, frame_cross, frame_proj.  ``direct`` = the five key
    tensors themselves (single process: no gather, no packed copy).  staged: the mark pack_rows set
    (deferred schedule): the enqueue happens only while it is set, and clears it.  slot = (epoch, stride):
    ``gathered`` is the two-slot receive buffer of the peer exchange, the slot is chosen on the device.
    prenormalised: the gathered rows are unit vectors already (pack_rows(norm_dim=D))."""
    lib = _lib.load()
    arr = (hmmc_queue * 5)()
    keep = []
    for i, buf in enumerate(queue_bufs5):
        qs, st = _queue_struct(buf, prec)
        arr[i] = qs
        keep.append(st)
    dev = queue_bufs5[0].device
    scratch = torch.empty((3 + 2 * F) * W * b, dtype=torch.float32, device=dev)     # key norms
    if direct is not None:
        _lib.check(lib.hmmc_enqueue_norm_direct(*[_p(t) for t in direct], W * b, F, D, arr, _p(queue_ptr),
                                                int(ptr_host), K, _p(scratch), _stream()), "hmmc_enqueue_norm_direct")
    else:
        epoch, stride = slot if slot is not None else (None, 0)
        _lib.check(lib.hmmc_enqueue_norm(_p(gathered), W, b, F, D, arr, _p(queue_ptr), int(ptr_host), K, _p(scratch),
                                         _p(staged), _p(epoch), int(stride), int(bool(prenormalised)), _stream()),
                   "hmmc_enqueue_norm")
    planes = None if prec == PREC_FP32 else (2 if prec == PREC_BF16X3 else 1)
    for st, buf in zip(keep, queue_bufs5):
        st.wrote(buf, planes)               # the kernel kept the copies of this plane count in step


def peer_push_rows(send, peer_bufs, peer_flags, rank, slot_stride, epoch, done_counter):
    """Copy this rank's packed rows into block `rank` of the current slot of every rank's receive buffer and raise
    this rank's flag everywhere (hmmc_peer_push_rows).  peer_bufs / peer_flags: the W base addresses (ints)."""
    lib = _lib.load()
    W = len(peer_bufs)
    bufs = (ctypes.c_uint64 * W)(*[int(x) for x in peer_bufs])
    flags = (ctypes.c_uint64 * W)(*[int(x) for x in peer_flags])
    _lib.check(lib.hmmc_peer_push_rows(_p(send), send.numel(), bufs, flags, W, int(rank), int(slot_stride), _p(epoch),
                                       _p(done_counter), _stream()), "hmmc_peer_push_rows")


def peer_wait(my_flags, W, epoch):
    """Block the current stream until every rank's rows of the current exchange have landed, then count it."""
    lib = _lib.load()
    _lib.check(lib.hmmc_peer_wait(_p(my_flags), int(W), _p(epoch), _stream()), "hmmc_peer_wait")


# ----------------------------------------------------------------------------- fine-tune head

def loose_similarity_raw(seq, vis, scale, prec):
    lib = _lib.load()
    Bt, D = seq.shape
    if vis.dim() == 3:
        Bv, Fv = vis.shape[0], vis.shape[1]
    else:
        Bv, Fv = vis.shape[0], 1
    out = torch.empty(Bt, Bv * Fv, dtype=torch.float32, device=seq.device)
    nbytes = lib.hmmc_similarity_workspace_bytes(Bt, Bv, Fv, D, prec)
    ws = workspace(seq.device, nbytes)
    _lib.check(lib.hmmc_loose_similarity_fwd(_p(seq), Bt, _p(vis), Bv, Fv, D, float(scale), prec, _p(out),
                                             _p(ws), ws.numel(), _stream()), "hmmc_loose_similarity_fwd")
    return out.view(Bt, Bv, Fv) if vis.dim() == 3 else out


class _LooseSimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, vis, scale, prec):
        s = _f32c(seq, "sequence_output")
        v = _f32c(vis, "visual_output")
        ctx.save_for_backward(s, v)
        ctx.scale = scale
        ctx.prec = prec
        ctx.dtypes = (seq.dtype, vis.dtype)
        return loose_similarity_raw(s, v, scale, prec)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        s, v = ctx.saved_tensors
        Bt, D = s.shape
        Bv, Fv = (v.shape[0], v.shape[1]) if v.dim() == 3 else (v.shape[0], 1)
        g = _f32c(g, "grad").reshape(Bt, Bv * Fv)
        ds = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        dv = torch.empty_like(v) if ctx.needs_input_grad[1] else None
        nbytes = lib.hmmc_similarity_workspace_bytes(Bt, Bv, Fv, D, ctx.prec)
        ws = workspace(s.device, nbytes)
        _lib.check(lib.hmmc_loose_similarity_bwd(_p(s), Bt, _p(v), Bv, Fv, D, float(ctx.scale), _p(g), _p(ds),
                                                 _p(dv), _p(ws), ws.numel(), _stream()),
                   "hmmc_loose_similarity_bwd")
        if ds is not None:
            ds = ds.to(ctx.dtypes[0])
        if dv is not None:
            dv = dv.to(ctx.dtypes[1])
        return ds, dv, None, None


def loose_similarity(seq, vis, scale, precision=None):
    return _LooseSimFn.apply(seq, vis, float(scale), resolve_precision(precision))


class _CrossEnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim):
        lib = _lib.load()
        if sim.dim() != 2 or sim.shape[0] != sim.shape[1]:
            raise HmmcError("CrossEn expects a square similarity matrix, got %s" % (tuple(sim.shape),))
        S = sim if (sim.is_cuda and sim.dtype == torch.float32 and sim.stride(1) == 1) else _f32c(sim, "sim_matrix")
        B = S.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=S.device)
        need = sim.requires_grad
        dS = torch.empty(B, B, dtype=torch.float32, device=S.device) if need else None
        scratch = torch.empty(B, dtype=torch.float32, device=S.device)
        _lib.check(lib.hmmc_cross_en_fwd_bwd(_p(S), S.stride(0), B, _p(loss), _p(dS), B, _p(scratch), _stream()),
                   "hmmc_cross_en_fwd_bwd")
        if need:
            ctx.save_for_backward(dS)
        ctx.dtype = sim.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dS,) = ctx.saved_tensors
        return (dS * g).to(ctx.dtype)


def cross_en(sim):
    return _CrossEnFn.apply(sim)


_warned_slow_symce = set()


def _warn_symce_path(B, D, prec):
    """The tensor-core tiling of the fine-tune head takes global batches that are a multiple of 32 (bf16x3) or 64
    (bf16) with D % 64 == 0 and D <= 1024 (symce_tensor_ok in csrc/finetune.cu); anything else runs the exact
    CUDA-core path, correct but several times slower.  Say so once per shape instead of being silently slow."""
    if prec == PREC_FP32:
        return
    ok = D % 64 == 0 and D <= 1024 and (B % 32 == 0 if prec == PREC_BF16X3 else B % 64 == 0)
    if not ok and (B, D, prec) not in _warned_slow_symce:
        _warned_slow_symce.add((B, D, prec))
        import warnings
        warnings.warn("hmmc_b200: fine-tune head with global batch %d, D %d takes the CUDA-core path (the tensor-core "
                      "path needs batch %% %d == 0, D %% 64 == 0, D <= 1024): expect several times the step time"
                      % (B, D, 32 if prec == PREC_BF16X3 else 64), RuntimeWarning, stacklevel=3)


def sym_ce_raw(text, video, frames, scale, w_vtm, w_ftm, prec, need_grad):
    """Fused fine-tune head on gathered embeddings; returns (loss, dtext, dvideo, dframes)."""
    lib = _lib.load()
    B, D = text.shape
    _warn_symce_path(B, D, prec)
    F = frames.shape[1] if frames is not None else 0
    nbytes = lib.hmmc_sym_ce_workspace_bytes(B, F, D, prec)
    ws = workspace(text.device, nbytes)
    loss = torch.empty((), dtype=torch.float32, device=text.device)
    dt = torch.empty_like(text) if need_grad else None
    dv = torch.empty_like(video) if (need_grad and video is not None) else None
    df = torch.empty_like(frames) if (need_grad and frames is not None) else None
    _lib.check(lib.hmmc_sym_ce_fwd_bwd(_p(text), _p(video), _p(frames), B, F, D, float(scale), float(w_vtm),
                                       float(w_ftm), prec, _p(loss), _p(dt), _p(dv), _p(df), _p(ws), ws.numel(),
                                       _stream()), "hmmc_sym_ce_fwd_bwd")
    return loss, dt, dv, df


class _SymCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text, video, frames, scale, w_vtm, w_ftm, prec):
        t = _f32c(text, "query_output")
        v = _f32c(video, "visual_output") if video is not None else None
        f = _f32c(frames, "frame_output") if frames is not None else None
        need = any(x is not None and x.requires_grad for x in (text, video, frames))
        loss, dt, dv, df = sym_ce_raw(t, v, f, scale, w_vtm, w_ftm, prec, need)
        ctx.has = (video is not None, frames is not None)
        if need:
            ctx.save_for_backward(*[x for x in (dt, dv, df) if x is not None])
        return loss

    @staticmethod
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        if getattr(ctx, "consumed", False):
            raise HmmcError("the fused head's gradients were already consumed (retain_graph is not supported)")
        ctx.consumed = True
        scale_inplace(saved, g)                   # one launch (a no-op for loss.backward()); the buffers are ours
        dt = saved.pop(0)
        dv = saved.pop(0) if ctx.has[0] else None
        df = saved.pop(0) if ctx.has[1] else None
        return dt, dv, df, None, None, None, None


def sym_ce(text, video, frames, scale, w_vtm, w_ftm, precision=None):
    return _SymCeFn.apply(text, video, frames, float(scale), float(w_vtm), float(w_ftm),
                          resolve_precision(precision))


def sym_ce_packed_raw(packed, F, D, scale, w_vtm, w_ftm, prec, need_grad):
    """Fused fine-tune head on the all-gather's own layout: rows [text | video | frames(F*D)].
    Returns (loss, dpacked) with dpacked in the same layout (what the gather's backward consumes)."""
    lib = _lib.load()
    B = packed.shape[0]
    if packed.shape[1] != (2 + F) * D:
        raise HmmcError("sym_ce_packed: row width %d is not (2+F)*D = %d" % (packed.shape[1], (2 + F) * D))
    _warn_symce_path(B, D, prec)
    nbytes = lib.hmmc_sym_ce_packed_workspace_bytes(B, F, D, prec)
    ws = workspace(packed.device, nbytes)
    loss = torch.empty((), dtype=torch.float32, device=packed.device)
    dp = torch.empty_like(packed) if need_grad else None
    _lib.check(lib.hmmc_sym_ce_packed_fwd_bwd(_p(packed), B, F, D, float(scale), float(w_vtm), float(w_ftm), prec,
                                              _p(loss), _p(dp), _p(ws), ws.numel(), _stream()),
               "hmmc_sym_ce_packed_fwd_bwd")
    return loss, dp


class _SymCePackedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, F, D, scale, w_vtm, w_ftm, prec):
        loss, dp = sym_ce_packed_raw(_f32c(packed, "packed"), F, D, scale, w_vtm, w_ftm, prec, packed.requires_grad)
        if dp is not None:
            ctx.save_for_backward(dp)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dp,) = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            raise HmmcError("the fused head's gradients were already consumed (retain_graph is not supported)")
        ctx.consumed = True
        scale_inplace([dp], g)                    # in place: the buffer is ours
        return dp, None, None, None, None, None, None


def sym_ce_packed(packed, F, D, scale, w_vtm, w_ftm, precision=None):
    return _SymCePackedFn.apply(packed, int(F), int(D), float(scale), float(w_vtm), float(w_ftm),
                                resolve_precision(precision))


# ----------------------------------------------------------------------------- eval

def sim_topk(text, video, frames, scale, top_k, precision=None, want_sim=True, want_fsim=True, combine=False):
    """(sim [Nt,Nv], fsim [Nt,Nv]) of one text block against one gallery block.
    combine=True: returns (sim + fsim, None), the sum formed inside the kernel where the fused tiles apply
    (main_task_retrieval.py:512-513)."""
    lib = _lib.load()
    prec = resolve_precision(precision)
    text = _f32c(text, "text")
    Nt, D = text.shape
    video = _f32c(video, "video") if video is not None else None
    frames = _f32c(frames, "frames") if frames is not None else None
    Nv = video.shape[0] if video is not None else frames.shape[0]
    F = frames.shape[1] if frames is not None else 0
    sim = torch.empty(Nt, Nv, dtype=torch.float32, device=text.device) if (want_sim and video is not None) else None
    fsim = torch.empty(Nt, Nv, dtype=torch.float32, device=text.device) if (want_fsim and frames is not None) else None
    nbytes = lib.hmmc_sim_topk_workspace_bytes(Nt, Nv, F, D, prec)
    ws = workspace(text.device, nbytes)
    fused_sum = (combine and sim is not None and fsim is not None and prec != PREC_FP32
                 and lib.hmmc_eval_fused_supported(F, D, int(top_k)))
    _lib.check(lib.hmmc_sim_topk_fwd(_p(text), Nt, _p(video), _p(frames), Nv, F, D, float(scale), int(top_k), prec,
                                     _p(sim), _p(sim if fused_sum else fsim), Nv, _p(ws), ws.numel(), _stream()),
               "hmmc_sim_topk_fwd")
    if combine:
        if fused_sum or fsim is None:
            return sim, None
        return sim + fsim, None
    return sim, fsim


_diag_index = {}


def _diagonal_index(n, dev):
    """gt = 0..n-1 and group_start = 0..n of the square (one caption per video) case, built once per size and
    device instead of two fill kernels per call."""
    key = (n, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _diag_index:
        if len(_diag_index) > 16:
            _diag_index.clear()
        _diag_index[key] = (torch.arange(n, dtype=torch.int32, device=dev), torch.arange(n + 1, dtype=torch.int32, device=dev))
    return _diag_index[key]


def rank_count(sim, gt=None, group_start=None, want_t2v=True, want_v2t=True):
    """Integer rank vectors (int32, on the device) of a materialised similarity matrix."""
    lib = _lib.load()
    sim = _f32c(sim, "sim")
    Nt, Nv = sim.shape
    dev = sim.device
    if gt is None:
        if Nt != Nv:
            raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,1)" % (Nt, Nv, Nt))
        gt, group_start = _diagonal_index(Nt, dev)
    gt = gt.to(device=dev, dtype=torch.int32).contiguous()
    group_start = group_start.to(device=dev, dtype=torch.int32).contiguous()
    t2v = torch.empty(Nt, dtype=torch.int32, device=dev) if want_t2v else None
    v2t = torch.empty(Nv, dtype=torch.int32, device=dev) if want_v2t else None
    theta = torch.empty(Nv, dtype=torch.float32, device=dev)
    _lib.check(lib.hmmc_rank_count(_p(sim), sim.stride(0), Nt, Nv, _p(gt), _p(group_start), _p(t2v), _p(v2t),
                                   _p(theta), _stream()), "hmmc_rank_count")
    return t2v, v2t


def group_max(sim, group_start):
    """tensor_video_to_text_sim on the device: out[j, g] = max_{s in g} sim[s, j]."""
    lib = _lib.load()
    sim = _f32c(sim, "sim")
    Nt, Nv = sim.shape
    G = group_start.numel() - 1
    out = torch.empty(Nv, G, dtype=torch.float32, device=sim.device)
    gs = group_start.to(device=sim.device, dtype=torch.int32).contiguous()
    _lib.check(lib.hmmc_group_max(_p(sim), sim.stride(0), Nv, G, _p(gs), _p(out), _stream()), "hmmc_group_max")
    return out
