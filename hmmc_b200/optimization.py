"""Host mirror of the reference's modules/optimization.py.

Same names, constructor arguments, validation, state keys (``step``, ``next_m``, ``next_v``) and
``get_lr`` as the reference's ``BertAdam`` (modules/optimization.py:52-168), so optimizer
checkpoints load unchanged.  ``step()`` runs the whole parameter list in three launches of
``hmmc_bert_adam_multi`` instead of a Python loop over parameters; passing
``global_max_norm`` folds the ``torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)`` call
that precedes ``optimizer.step()`` in the training loops (main_pretrain.py:277,
main_task_retrieval.py:291) into the same pass.

No CPU fallback: parameters must be contiguous fp32 CUDA tensors.
"""
import math

import numpy as np
import torch
from torch.optim import Optimizer
from torch.optim.optimizer import required

from . import _lib
from .ops import HmmcError, _p, _stream, workspace


def warmup_cosine(x, warmup=0.002):
    # modules/optimization.py:26-29
    if x < warmup:
        return x / warmup
    return 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    # modules/optimization.py:31-36
    if x < warmup:
        return x / warmup
    return 1.0


def warmup_linear(x, warmup=0.002):
    # modules/optimization.py:38-43
    if x < warmup:
        return x / warmup
    return max((x - 1.) / (warmup - 1.), 0)


SCHEDULES = {
    'warmup_cosine': warmup_cosine,
    'warmup_constant': warmup_constant,
    'warmup_linear': warmup_linear,
}


class _TensorTable:
    """Device pointer tables (p, grad, next_m, next_v) + block table of one parameter list."""

    def __init__(self, columns, numels, device):
        lib = _lib.load()
        be = lib.hmmc_ema_block_elems()
        offs = [0]
        for ne in numels:
            offs.append(offs[-1] + (ne + be - 1) // be)
        self.n = len(numels)
        self.total_blocks = offs[-1]
        self.total_elems = sum(numels)
        self.device = device
        i64 = lambda v: torch.tensor(v, dtype=torch.int64).to(device)
        self.numels = i64(numels)
        self.offs = i64(offs)
        self.dtypes = torch.zeros(self.n, dtype=torch.int32, device=device)
        self.cols = [i64(c) for c in columns]
        self.sigs = [tuple(c) for c in columns]
        self.ws_bytes = lib.hmmc_bert_adam_workspace_bytes(self.n, self.total_blocks)

    def refresh(self, k, ptrs):
        """Re-upload column k when its pointers changed (gradients are re-allocated by zero_grad)."""
        sig = tuple(ptrs)
        if sig != self.sigs[k]:
            self.cols[k] = torch.tensor(ptrs, dtype=torch.int64).to(self.device)
            self.sigs[k] = sig


def _check_tensor(t, what):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise HmmcError("%s must be a contiguous fp32 CUDA tensor (got %s, %s)" % (what, t.dtype, t.device))


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ for the fp32 CUDA gradients of `parameters`, one norm pass and
    one scale pass over all of them (main_pretrain.py:277).  Returns the total norm (0-d tensor)."""
    if float(norm_type) != 2.0:
        raise HmmcError("clip_grad_norm_: only the 2-norm is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None and p.grad.numel()]
    if not grads:
        return torch.tensor(0.)
    for g in grads:
        _check_tensor(g, "gradient")
    dev = grads[0].device
    tab = _TensorTable([[g.data_ptr() for g in grads]], [g.numel() for g in grads], dev)
    norms = torch.empty(tab.n + 1, dtype=torch.float32, device=dev)
    ws = workspace(dev, tab.ws_bytes)
    _lib.check(_lib.load().hmmc_clip_grad_norm_multi(_p(tab.cols[0]), _p(tab.numels), _p(tab.offs), tab.n,
                                                     tab.total_blocks, float(max_norm), _p(norms), _p(ws),
                                                     ws.numel(), _stream()), "hmmc_clip_grad_norm_multi")
    return norms[tab.n]


class BertAdam(Optimizer):
    """BERT version of Adam with the weight-decay fix; arguments as modules/optimization.py:52-87.

    step(closure=None, global_max_norm=None, write_back_grads=False)
        global_max_norm: also apply clip_grad_norm_(all parameters of this optimizer, global_max_norm)
        first, in the same pass.  write_back_grads: store the clipped gradients into p.grad like the
        reference's in-place clipping does (the training loops zero them right after, so the default
        skips that write).
    After a step, ``last_grad_norm`` is a 0-d device tensor with the total gradient norm.

    --enable_amp (main_pretrain.py:267-284): ``scaler.step(optimizer)`` of torch.cuda.amp.GradScaler finds
    ``_step_supports_amp_scaling`` and hands the step its ``grad_scale`` / ``found_inf`` tensors instead of
    unscaling the gradients itself.  The step is skipped when an inf / NaN was found (one host read, as the
    reference's unfused GradScaler path does) and otherwise multiplies every gradient by 1/scale inside the norm
    and update kernels: the arithmetic equals unscale_ followed by step(), without the extra pass.
    """
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=required, warmup=-1, t_total=-1, schedule='warmup_linear',
                 b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01, max_grad_norm=1.0):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        if not 0.0 <= b1 < 1.0:
            raise ValueError("Invalid b1 parameter: {} - should be in [0.0, 1.0[".format(b1))
        if not 0.0 <= b2 < 1.0:
            raise ValueError("Invalid b2 parameter: {} - should be in [0.0, 1.0[".format(b2))
        if not e >= 0.0:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(e))
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total,
                        b1=b1, b2=b2, e=e, weight_decay=weight_decay,
                        max_grad_norm=max_grad_norm)
        super(BertAdam, self).__init__(params, defaults)
        self._table = None
        self._table_key = None
        self._hyper_dev = None
        self._ring, self._ring_pos = None, 0
        self.last_grad_norm = None
        self._norms = None

    def get_lr(self):
        lr = []
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    return [0]
                lr.append(self._lr_scheduled(group, state['step']))
        return lr

    @staticmethod
    def _lr_scheduled(group, step):
        if group['t_total'] != -1:
            schedule_fct = SCHEDULES[group['schedule']]
            return group['lr'] * schedule_fct(step / group['t_total'], group['warmup'])
        return group['lr']

    def _live(self):
        live = []
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None or p.numel() == 0:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                state = self.state[p]
                if len(state) == 0:
                    _check_tensor(p.data, "parameter")
                    state['step'] = 0
                    state['next_m'] = torch.zeros_like(p.data)
                    state['next_v'] = torch.zeros_like(p.data)
                live.append((group, p, state))
        return live

    def step(self, closure=None, global_max_norm=None, write_back_grads=False):
        loss = None
        if closure is not None:
            loss = closure()
        live = self._live()
        if not live:
            return loss
        dev = live[0][1].device
        # torch.cuda.amp.GradScaler protocol (set on the optimizer around scaler.step())
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        if found_inf is not None and float(found_inf) > 0:
            return loss                       # GradScaler skips the step; scaler.update() lowers the scale
        inv_scale = None
        if grad_scale is not None:
            inv_scale = grad_scale.detach().to(device=dev, dtype=torch.float64).reciprocal().to(torch.float32).reshape(1)
        p_ptrs = [p.data_ptr() for _, p, _ in live]
        g_ptrs = [p.grad.data_ptr() for _, p, _ in live]
        m_ptrs = [s['next_m'].data_ptr() for _, _, s in live]
        v_ptrs = [s['next_v'].data_ptr() for _, _, s in live]
        key = tuple(p_ptrs)
        if self._table is None or self._table_key != key:
            for _, p, state in live:
                _check_tensor(p.data, "parameter")
                _check_tensor(state['next_m'], "next_m")
                _check_tensor(state['next_v'], "next_v")
            self._table = _TensorTable([p_ptrs, [0] * len(live), m_ptrs, v_ptrs], [p.numel() for _, p, _ in live], dev)
            self._table_key = key
            self._hyper_dev = torch.empty(len(live), 8, dtype=torch.float32, device=dev)
            self._norms = torch.empty(len(live) + 1, dtype=torch.float32, device=dev)
            self._ring = [[torch.empty(len(live), 8, dtype=torch.float32).pin_memory(), None] for _ in range(4)]
            self._ring_pos = 0
        tab = self._table
        if tuple(g_ptrs) != tab.sigs[1]:
            for _, p, _ in live:
                _check_tensor(p.grad, "gradient")
            tab.refresh(1, g_ptrs)
        tab.refresh(2, m_ptrs)
        tab.refresh(3, v_ptrs)
        # one hyper-parameter row per distinct (group, step); python doubles become fp32 exactly where
        # the reference's tensor-times-scalar ops cast them
        uniq, rows = {}, []
        idx = [uniq.setdefault((id(group), state['step']), len(uniq)) for group, _, state in live]
        for group, _, state in live:
            k = (id(group), state['step'])
            if uniq[k] == len(rows):
                b1, b2 = group['b1'], group['b2']
                rows.append((self._lr_scheduled(group, state['step']), group['weight_decay'], b1, 1 - b1, b2, 1 - b2,
                             group['e'], group['max_grad_norm']))
        h = np.asarray(rows, dtype=np.float64).astype(np.float32)[np.asarray(idx)]
        # upload through a small ring of pinned buffers (a pageable copy would synchronise the stream);
        # a slot is reused only after the copy that last read it has completed
        slot = self._ring[self._ring_pos % len(self._ring)]
        self._ring_pos += 1
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0].copy_(torch.from_numpy(h))
        self._hyper_dev.copy_(slot[0], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        ws = workspace(dev, tab.ws_bytes)
        gmax = float(global_max_norm) if global_max_norm is not None and global_max_norm > 0 else 0.0
        _lib.check(_lib.load().hmmc_bert_adam_multi(
            _p(tab.cols[0]), _p(tab.cols[1]), _p(tab.cols[2]), _p(tab.cols[3]), _p(tab.numels), _p(tab.dtypes),
            _p(tab.offs), tab.n, tab.total_blocks, _p(self._hyper_dev), gmax, 1 if write_back_grads else 0,
            _p(inv_scale), _p(self._norms), _p(ws), ws.numel(), _stream()), "hmmc_bert_adam_multi")
        self.last_grad_norm = self._norms[tab.n]
        for _, _, state in live:
            state['step'] += 1
        return loss
