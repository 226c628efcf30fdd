"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's optimizer step
(SURVEY.md §8(f) row N3).  Imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs,
never by the product path.

Restates, one rounded fp32 op per line:
  * torch.nn.utils.clip_grad_norm_(model.parameters(), G)       main_pretrain.py:277
  * BertAdam.step                                               modules/optimization.py:103-168
  * warmup_cosine / warmup_constant / warmup_linear             modules/optimization.py:26-49

Pins (tests/golden/optim.npz, written by oracle/gen_golden.py running the reference's own
BertAdam + torch's clip_grad_norm_ on CPU):
  * which ops round once: `add_(grad, alpha=1-b1)` and `addcmul_(grad, grad, value=1-b2)` are
    single fused multiply-adds in the reference run (`fma(alpha, g, m)`, `fma(value*g, g, v)`);
    every other op rounds separately.  With that, next_m / next_v are BIT-EXACT against the golden
    on steps where no clipping is active.
  * torch's vectorised CPU sqrt is not correctly rounded (0.6 % of elements are 1 ulp off an IEEE
    sqrt); this file and the CUDA kernel use the IEEE sqrt, so p agrees with the golden to 1 ulp.
  * on clipped steps the gradient norm enters; the reference sums squares in fp32 in an order that
    cannot be restated, so those steps agree to ~1e-7 relative (norms are computed in float64 here).
"""
import math

import numpy as np

f32 = np.float32


def warmup_cosine(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 1.0


def warmup_linear(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return max((x - 1.) / (warmup - 1.), 0)


SCHEDULES = {'warmup_cosine': warmup_cosine, 'warmup_constant': warmup_constant, 'warmup_linear': warmup_linear}


def lr_scheduled(group, step):
    # modules/optimization.py:156-161
    if group['t_total'] != -1:
        return group['lr'] * SCHEDULES[group['schedule']](step / group['t_total'], group['warmup'])
    return group['lr']


def _r(x):
    return np.asarray(x, dtype=np.float32)


def _fma(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 is exact in fp64; one fp64 add, then one
    rounding to fp32 (double rounding can differ from a true FMA only with probability ~2^-29)."""
    return _r(np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64))


def grad_norm(g):
    return math.sqrt(float(np.sum(np.asarray(g, np.float64) ** 2)))


def clip_coef(max_norm, total_norm):
    """clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max=1.0), all fp32."""
    c = f32(max_norm) / (f32(total_norm) + f32(1e-6))
    return f32(min(c, f32(1.0)))


def clip_and_step(params, grads, ms, vs, steps, groups, group_of, global_max_norm=None, norms=None):
    """One  clip_grad_norm_(all, global_max_norm); BertAdam.step()  over lists of fp32 arrays.

    groups: list of dicts (lr, schedule, warmup, t_total, b1, b2, e, weight_decay, max_grad_norm);
    group_of[i]: group index of tensor i; steps[i]: state['step'] of tensor i (advanced in place).
    norms: optional (per_tensor_norms, total_norm) to use instead of computing them (lets a test feed
    the device-computed norms and compare everything after them bit for bit).
    Returns (new_params, new_ms, new_vs, clipped_grads, total_norm).
    """
    n = len(params)
    if norms is None:
        per = [grad_norm(g) for g in grads]
        total = math.sqrt(sum(x * x for x in per))
    else:
        per, total = list(norms[0]), float(norms[1])
    cg = f32(1.0)
    if global_max_norm is not None and global_max_norm > 0:
        cg = clip_coef(global_max_norm, total)
    out_p, out_m, out_v, out_g = [], [], [], []
    for i in range(n):
        grp = groups[group_of[i]]
        p, g, m, v = _r(params[i]), _r(grads[i]), _r(ms[i]), _r(vs[i])
        g = _r(g * cg)                                               # global clip_grad_norm_
        ct = f32(1.0)
        if grp['max_grad_norm'] > 0:                                 # optimization.py:135-136
            ct = clip_coef(grp['max_grad_norm'], f32(per[i]) * cg)
        g = _r(g * ct)
        b1, b2 = grp['b1'], grp['b2']
        m = _fma(f32(1 - b1), g, _r(m * f32(b1)))                    # :141
        v = _fma(_r(f32(1 - b2) * g), g, _r(v * f32(b2)))            # :143
        u = _r(m / _r(np.sqrt(v) + f32(grp['e'])))                   # :144
        if grp['weight_decay'] > 0.0:
            u = _r(u + _r(f32(grp['weight_decay']) * p))             # :153-154
        lr = lr_scheduled(grp, steps[i])                             # :156-161
        p = _r(p + (-_r(f32(lr) * u)))                               # :163-164
        steps[i] += 1
        out_p.append(p); out_m.append(m); out_v.append(v); out_g.append(g)
    return out_p, out_m, out_v, out_g, f32(total)
