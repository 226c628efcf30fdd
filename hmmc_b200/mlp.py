"""Host mirror of the reference's projector / predictor MLP (modules/modeling.py:788-807; SURVEY.md
§8(f) row N1), backed by the tcgen05 GEMM engine of libhmmc_head.so.

Same constructor, sub-module layout and state-dict keys as the reference class
(``linear_hidden.1`` = Linear, ``linear_hidden.2`` = BatchNorm1d / SyncBatchNorm, ``linear_out``), so
checkpoints load unchanged and ``nn.SyncBatchNorm.convert_sync_batchnorm(mlp)`` (modeling.py:127-129)
keeps working: with a SyncBatchNorm inside and an initialised process group the batch statistics and
their backward sums are all-reduced over the ranks (one 2*inner_dim double exchange per direction).

No CPU or PyTorch fallback: inputs must be fp32 CUDA tensors; num_layers must be 2 (the reference's
cross_config.json sets proj_num_layers = pred_num_layers = 2).
"""
import ctypes

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from .ops import HmmcError, PREC_FP32, PREC_BF16X3, _f32c, _p, _stream, resolve_precision


def _params_struct(W1, b1, gamma, beta, W2, b2, rm, rv):
    ptr = lambda t: t.data_ptr() if t is not None else 0
    return _lib.hmmc_mlp_params(ptr(W1), ptr(b1), ptr(gamma), ptr(beta), ptr(W2), ptr(b2), ptr(rm), ptr(rv))


def _exchange_view(buf, ptr, n):
    """float64 view of the n-element exchange buffer the library placed inside `buf`."""
    off = int(ptr.value) - buf.data_ptr()
    return buf[off:off + 8 * n].view(torch.float64)


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, gamma, beta, W2, b2, running_mean, running_var, eps, momentum, training, prec, group,
                need):
        lib = _lib.load()
        x = _f32c(x, "x")
        ts = [_f32c(t, "parameter") for t in (W1, b1, gamma, beta, W2, b2)]
        M, Din = x.shape
        Dh, Dout = W1.shape[0], W2.shape[0]
        world = dist.get_world_size(group) if (training and group is not None) else 1
        nbytes = lib.hmmc_mlp_ctx_bytes(M, Din, Dh, Dout, prec, int(need))
        buf = torch.empty(nbytes, dtype=torch.uint8, device=x.device)       # lives until the backward
        ps = _params_struct(*ts, running_mean, running_var)
        sptr = ctypes.c_void_p()
        _lib.check(lib.hmmc_mlp_fwd_a(_p(x), M, Din, Dh, Dout, ctypes.byref(ps), prec, int(need), _p(buf), buf.numel(),
                                      ctypes.byref(sptr), _stream()), "hmmc_mlp_fwd_a")
        if world > 1:
            dist.all_reduce(_exchange_view(buf, sptr, 2 * Dh), group=group)
        y = torch.empty(M, Dout, dtype=torch.float32, device=x.device)
        _lib.check(lib.hmmc_mlp_fwd_b(M, Din, Dh, Dout, ctypes.byref(ps), float(eps), float(momentum), float(world * M),
                                      int(training), prec, int(need), _p(buf), buf.numel(), _p(y), _stream()),
                   "hmmc_mlp_fwd_b")
        if need:
            ctx.save_for_backward(buf, *ts)
            ctx.dims = (M, Din, Dh, Dout, prec, world)
            ctx.group = group
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        buf, W1, b1, gamma, beta, W2, b2 = ctx.saved_tensors
        M, Din, Dh, Dout, prec, world = ctx.dims
        if getattr(ctx, "consumed", False):
            raise HmmcError("the MLP's saved context was already consumed (retain_graph is not supported)")
        ctx.consumed = True
        dy = _f32c(dy, "grad_output")
        need = ctx.needs_input_grad
        new = lambda flag, *shape: torch.empty(*shape, dtype=torch.float32, device=dy.device) if flag else None
        dx, dW1, db1 = new(need[0], M, Din), new(need[1], Dh, Din), new(need[2], Dh)
        dgamma, dbeta = new(need[3], Dh), new(need[4], Dh)
        dW2, db2 = new(need[5], Dout, Dh), new(need[6], Dout)
        ps = _params_struct(W1, b1, gamma, beta, W2, b2, None, None)
        sptr = ctypes.c_void_p()
        _lib.check(lib.hmmc_mlp_bwd_a(_p(dy), M, Din, Dh, Dout, ctypes.byref(ps), prec, _p(buf), buf.numel(), _p(dW2),
                                      _p(db2), ctypes.byref(sptr), _stream()), "hmmc_mlp_bwd_a")
        if world > 1:
            dist.all_reduce(_exchange_view(buf, sptr, 2 * Dh), group=ctx.group)
        _lib.check(lib.hmmc_mlp_bwd_b(M, Din, Dh, Dout, ctypes.byref(ps), float(world * M), prec, _p(buf), buf.numel(),
                                      _p(dx), _p(dW1), _p(db1), _p(dgamma), _p(dbeta), _stream()), "hmmc_mlp_bwd_b")
        return dx, dW1, db1, dgamma, dbeta, dW2, db2, None, None, None, None, None, None, None, None


def mlp_forward(x, W1, b1, gamma, beta, W2, b2, running_mean=None, running_var=None, eps=1e-5, momentum=0.1,
                training=True, precision=None, group=None):
    """Functional form; `group` = process group whose ranks share the batch statistics (None: local)."""
    prec = resolve_precision(precision)
    if prec == PREC_FP32:
        prec = PREC_BF16X3          # the fp32-parity mode of the tensor cores; there is no CUDA-core MLP path
    # autograd.Function.forward runs with gradients disabled: decide here whether a backward can follow
    need = torch.is_grad_enabled() and any(t.requires_grad for t in (x, W1, b1, gamma, beta, W2, b2))
    return _MlpFn.apply(x, W1, b1, gamma, beta, W2, b2, running_mean, running_var, eps, momentum, training, prec, group,
                        need)


class MLP(nn.Module):
    """modules/modeling.py:788-807 (same arguments, same parameter / buffer names)."""

    def __init__(self, in_dim=512, inner_dim=4096, out_dim=512, num_layers=2, precision=None):
        super(MLP, self).__init__()
        linear_hidden = [nn.Identity()]
        for i in range(num_layers - 1):
            linear_hidden.append(nn.Linear(in_dim if i == 0 else inner_dim, inner_dim))
            linear_hidden.append(nn.BatchNorm1d(inner_dim))
            linear_hidden.append(nn.ReLU(inplace=True))
        self.linear_hidden = nn.Sequential(*linear_hidden)
        self.linear_out = nn.Linear(in_dim if num_layers == 1 else inner_dim,
                                    out_dim) if num_layers >= 1 else nn.Identity()
        self.num_layers = num_layers
        self.precision = precision

    def forward(self, x):
        if self.num_layers != 2:
            raise HmmcError("MLP: only the reference's num_layers = 2 configuration is implemented")
        lin1, bn = self.linear_hidden[1], self.linear_hidden[2]
        if x.dim() != 2:
            raise HmmcError("MLP: expected a [rows, %d] input, got %s" % (lin1.in_features, tuple(x.shape)))
        if bn.momentum is None or not bn.affine:
            raise HmmcError("MLP: the BatchNorm must be affine with a fixed momentum")
        training = self.training or not bn.track_running_stats
        group = None
        if training and isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
            group = bn.process_group if bn.process_group is not None else dist.group.WORLD
            if dist.get_world_size(group) == 1:
                group = None
        y = mlp_forward(x, lin1.weight, lin1.bias, bn.weight, bn.bias, self.linear_out.weight, self.linear_out.bias,
                        bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, bn.eps, bn.momentum, training,
                        self.precision, group)
        if training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return y


class _VisualTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, temporal, original):
        o = _f32c(original, "visual_hidden_original")
        t = _f32c(temporal, "visual_hidden") if temporal is not None else None
        B, F, D = o.shape
        out = torch.empty(B, D, dtype=torch.float32, device=o.device)
        _lib.check(_lib.load().hmmc_visual_tail_fwd(_p(t), _p(o), B, F, D, _p(out), _stream()), "hmmc_visual_tail_fwd")
        ctx.save_for_backward(o, *([t] if t is not None else []))
        ctx.has_t = t is not None
        return out

    @staticmethod
    def backward(ctx, g):
        o = ctx.saved_tensors[0]
        t = ctx.saved_tensors[1] if ctx.has_t else None
        B, F, D = o.shape
        dh = torch.empty_like(o)
        _lib.check(_lib.load().hmmc_visual_tail_bwd(_p(t), _p(o), _p(_f32c(g, "grad_output")), B, F, D, _p(dh), _stream()),
                   "hmmc_visual_tail_bwd")
        return (dh if ctx.has_t else None), dh


def visual_tail(visual_hidden, visual_hidden_original):
    """modules/module_cross.py:207-213: ``visual_hidden`` = output of the temporal transformer (None when
    use_temp is off), ``visual_hidden_original`` = the per-frame CLIP features [bs, frames, D].
    Returns (visual_output [bs, D], frame_output) exactly as VisualEncoder.forward does."""
    return _VisualTailFn.apply(visual_hidden, visual_hidden_original), visual_hidden_original
