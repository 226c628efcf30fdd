"""Parity of the multi-tensor clip_grad_norm_ + BertAdam step against the oracle
(oracle/optim_oracle.py) and the golden vectors of the reference's own BertAdam (tests/golden/optim.npz)."""
import numpy as np
import pytest
import torch

from hmmc_b200 import synthetic as syn
from hmmc_b200.optimization import BertAdam, clip_grad_norm_
from oracle import optim_oracle as OO
from gpu_util import cu

pytestmark = pytest.mark.gpu

CASES = [("pretrain", 6, 1.0), ("plain", 3, None), ("linear", 4, None)]


def _make(case):
    groups = syn.optim_groups(case)
    params = [torch.nn.Parameter(cu(x)) for x in syn.optim_tensors()]
    pg = [dict(g, params=[p for p, gi in zip(params, syn.OPTIM_GROUP_OF) if gi == k]) for k, g in enumerate(groups)]
    return groups, params, BertAdam(pg, lr=groups[0]['lr'])


def _order(params, opt):
    """Index of each parameter in the optimizer's walk order (group by group)."""
    flat = [p for g in opt.param_groups for p in g['params']]
    ids = [id(p) for p in flat]
    return [ids.index(id(p)) for p in params]


@pytest.mark.parametrize("case,nsteps,gmax", CASES)
def test_bert_adam_bit_exact_vs_oracle(case, nsteps, gmax):
    """Given the device-computed gradient norms, p / next_m / next_v equal the oracle bit for bit on
    every step; the norms themselves agree with a float64 sum to 1e-6."""
    groups, params, opt = _make(case)
    op = syn.optim_tensors()
    om = [np.zeros_like(x) for x in op]
    ov = [np.zeros_like(x) for x in op]
    steps = [0] * len(op)
    pos = _order(params, opt)
    for st in range(nsteps):
        grads = syn.optim_grads(st)
        for p, g in zip(params, grads):
            p.grad = cu(g)
        opt.step(global_max_norm=gmax, write_back_grads=True)
        norms = opt._norms.cpu().numpy()
        per = [norms[pos[i]] for i in range(len(op))]
        ref_per = [OO.grad_norm(g) for g in grads]
        np.testing.assert_allclose(per, ref_per, rtol=1e-6)
        np.testing.assert_allclose(float(opt.last_grad_norm), np.sqrt(sum(x * x for x in ref_per)), rtol=1e-6)
        op, om, ov, og, _ = OO.clip_and_step(op, grads, om, ov, steps, groups, syn.OPTIM_GROUP_OF, gmax,
                                             norms=(per, norms[len(op)]))
        for i, p in enumerate(params):
            assert np.array_equal(p.detach().cpu().numpy(), op[i]), (st, i)
            assert np.array_equal(opt.state[p]['next_m'].cpu().numpy(), om[i]), (st, i)
            assert np.array_equal(opt.state[p]['next_v'].cpu().numpy(), ov[i]), (st, i)
            assert np.array_equal(p.grad.cpu().numpy(), og[i]), (st, i)
            assert opt.state[p]['step'] == st + 1
        opt.zero_grad()


@pytest.mark.parametrize("case,nsteps,gmax", CASES)
def test_bert_adam_vs_reference_golden(golden, case, nsteps, gmax):
    """Against the reference's own BertAdam run on CPU: bit-exact moments when nothing clips,
    ~1e-7 relative (norm summation order, torch's non-IEEE CPU sqrt) otherwise."""
    g = golden("optim")
    _, params, opt = _make(case)
    for st in range(nsteps):
        for p, gr in zip(params, syn.optim_grads(st)):
            p.grad = cu(gr)
        opt.step(global_max_norm=gmax)
        if gmax is not None:
            np.testing.assert_allclose(float(opt.last_grad_norm), g[case + "_totals"][st], rtol=1e-6)
        opt.zero_grad()
    last = nsteps - 1
    for i, p in enumerate(params):
        m = opt.state[p]['next_m'].cpu().numpy()
        v = opt.state[p]['next_v'].cpu().numpy()
        pn = p.detach().cpu().numpy()
        gm, gv, gp = g["%s_m%d" % (case, i)], g["%s_v%d" % (case, i)], g["%s_p%d_s%d" % (case, i, last)]
        if case == "plain":
            assert np.array_equal(m, gm) and np.array_equal(v, gv), i
            d = np.abs(pn.astype(np.float64) - gp)
            assert (d <= 2.0 ** -23 * np.maximum(np.abs(gp), 2.0 ** -6)).all() and (d > 0).mean() < 0.03, i
        else:
            for got, ref in ((m, gm), (v, gv), (pn, gp)):
                np.testing.assert_allclose(got, ref, rtol=3e-6, atol=1e-6 * float(np.abs(ref).max()))


def test_get_lr_and_state_dict_keys(golden):
    g = golden("optim")
    _, params, opt = _make("pretrain")
    for p in params:
        p.grad = torch.zeros_like(p)
    assert opt.get_lr() == [0]                       # before the first step (optimization.py:95-96)
    for st in range(6):
        for p, gr in zip(params, syn.optim_grads(st)):
            p.grad = cu(gr)
        opt.step(global_max_norm=1.0)
    np.testing.assert_allclose(sorted(set(opt.get_lr())), g["pretrain_lrs"][-1], rtol=1e-15)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "next_m", "next_v"}
    assert sd["param_groups"][0]["schedule"] == "warmup_cosine"


def test_params_without_grad_are_skipped():
    _, params, opt = _make("plain")
    before = [p.detach().clone() for p in params]
    for i, (p, gr) in enumerate(zip(params, syn.optim_grads(0))):
        p.grad = cu(gr) if i % 2 == 0 else None
    opt.step()
    for i, p in enumerate(params):
        if i % 2 == 0:
            assert not torch.equal(p.detach(), before[i]) and opt.state[p]['step'] == 1
        else:
            assert torch.equal(p.detach(), before[i]) and len(opt.state[p]) == 0
    # the set of live parameters changes: the table is rebuilt, counters advance independently
    for p, gr in zip(params, syn.optim_grads(1)):
        p.grad = cu(gr)
    opt.step()
    assert [opt.state[p]['step'] for p in params] == [2, 1, 2, 1, 2, 1, 2, 1]


def test_clip_grad_norm_standalone():
    grads = syn.optim_grads(0)
    params = [torch.nn.Parameter(cu(x)) for x in syn.optim_tensors()]
    for p, g in zip(params, grads):
        p.grad = cu(g)
    total = clip_grad_norm_(params, 1.0)
    ref_total = np.sqrt(sum(OO.grad_norm(g) ** 2 for g in grads))
    np.testing.assert_allclose(float(total), ref_total, rtol=1e-6)
    coef = OO.clip_coef(1.0, float(total))
    assert coef < 1
    for p, g in zip(params, grads):
        assert np.array_equal(p.grad.cpu().numpy(), (g * coef).astype(np.float32))
    # below the threshold the coefficient clamps to exactly 1: gradients untouched
    small = [g * np.float32(1e-3) for g in grads]
    for p, g in zip(params, small):
        p.grad = cu(g)
    clip_grad_norm_(params, 1.0)
    for p, g in zip(params, small):
        assert np.array_equal(p.grad.cpu().numpy(), g)


def test_rejects_what_it_cannot_run():
    from hmmc_b200.ops import HmmcError
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(HmmcError):
        BertAdam([p], lr=1e-3).step()
    q = torch.nn.Parameter(torch.zeros(8, device="cuda", dtype=torch.float16))
    q.grad = torch.ones_like(q)
    with pytest.raises(HmmcError):
        BertAdam([q], lr=1e-3).step()


def test_large_table_many_blocks():
    """A few hundred tensors of ragged sizes (block table, unaligned tails) against torch ops on the GPU."""
    rs = np.random.RandomState(5)
    sizes = [int(s) for s in rs.randint(1, 40000, size=200)] + [8192 * 3, 8192 * 5 + 1]
    params = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.05) for s in sizes]
    params.append(torch.nn.Parameter((torch.randn(10001, device="cuda") * 0.05)[1:]))   # 4-byte aligned only
    ref_p = [p.detach().clone() for p in params]
    opt = BertAdam(params, lr=1e-3, warmup=-1, t_total=-1, weight_decay=0.01, max_grad_norm=-1)
    ref_m = [torch.zeros_like(p) for p in params]
    ref_v = [torch.zeros_like(p) for p in params]
    for st in range(2):
        for p in params:
            p.grad = torch.randn_like(p) * 1e-3
        for i, p in enumerate(params):
            g = p.grad
            ref_m[i].mul_(0.9).add_(g, alpha=1 - 0.9)
            ref_v[i].mul_(0.999).addcmul_(g, g, value=1 - 0.999)
            upd = ref_m[i] / (ref_v[i].sqrt() + 1e-6)
            upd += 0.01 * ref_p[i]
            ref_p[i].add_(-(1e-3 * upd))
        opt.step()
    for i, p in enumerate(params):
        torch.testing.assert_close(p.detach(), ref_p[i], rtol=2e-6, atol=1e-8)
        torch.testing.assert_close(opt.state[p]['next_m'], ref_m[i], rtol=2e-6, atol=1e-10)


def test_grad_scaler_step_equals_unscaled_step():
    """--enable_amp (main_pretrain.py:258-284): scaler.scale(loss).backward(); clip_grad_norm_; scaler.step(opt);
    scaler.update().  GradScaler hands BertAdam its scale (`_step_supports_amp_scaling`); the unscale happens inside
    the norm and update kernels.  With a power-of-two scale the result equals the fp32 run bit for bit; a step that
    saw an inf is skipped (parameters, moments and step counters untouched) and the scale is halved."""
    _, pa, opt_a = _make("pretrain")
    _, pb, opt_b = _make("pretrain")
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0, growth_interval=1000)
    for st in range(3):
        grads = [cu(g) for g in syn.optim_grads(st)]
        for p, g in zip(pa, grads):
            p.grad = g.clone()
        opt_a.step()
        loss = sum((p * g).sum() for p, g in zip(pb, grads))      # d loss / d p = g
        scaler.scale(loss).backward()
        assert float(pb[0].grad.abs().max()) > 100 * float(grads[0].abs().max())      # gradients really are scaled
        scaler.step(opt_b)
        scaler.update()
        opt_b.zero_grad()
        for a, b in zip(pa, pb):
            assert torch.equal(a.detach(), b.detach()), st
            assert torch.equal(opt_a.state[a]['next_v'], opt_b.state[b]['next_v']), st
    assert float(scaler.get_scale()) == 65536.0
    # an overflowing gradient: the step is skipped and the scale backs off
    before = [p.detach().clone() for p in pb]
    steps_before = [opt_b.state[p]['step'] for p in pb]
    grads = [cu(g) for g in syn.optim_grads(5)]
    loss = sum((p * g).sum() for p, g in zip(pb, grads))
    scaler.scale(loss).backward()
    pb[1].grad.view(-1)[0] = float("inf")
    scaler.step(opt_b)
    scaler.update()
    assert float(scaler.get_scale()) == 32768.0
    for p, b0, s0 in zip(pb, before, steps_before):
        assert torch.equal(p.detach(), b0) and opt_b.state[p]['step'] == s0
