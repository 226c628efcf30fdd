"""Pin the numpy oracle (oracle/head_oracle.py) against outputs of the
reference's own code (tests/golden/*.npz, written by oracle/gen_golden.py)."""
import numpy as np
import pytest

from hmmc_b200 import synthetic as syn
from oracle import head_oracle as O

KEYS = ["R1", "R5", "R10", "MR", "MeanR"]
TVK = ["R1", "R5", "R10", "MedianR", "MeanR", "Std_Rank", "MR"]


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_metrics_square(golden):
    g = golden("metrics")
    rs = np.random.RandomState(0)
    x = rs.randn(1000, 1000).astype(np.float32)
    x[np.arange(1000), np.arange(1000)] += 2.0
    a = O.compute_metrics(x)
    b = O.compute_metrics(x.T)
    assert [a[k] for k in KEYS] == list(g["sq_t2v"])
    assert [b[k] for k in KEYS] == list(g["sq_v2t"])
    # rank-count form reproduces the sort/where form exactly
    assert O.metrics_from_ranks(O.ranks_square(x)) == a
    ties = np.array([[1, 1, 0], [0, 2, 3], [0, 0, 5]], dtype=np.float32)
    c = O.compute_metrics(ties)
    assert [c[k] for k in KEYS] == list(g["ties"])


def test_metrics_nonsquare_raises():
    with pytest.raises(ValueError):
        O.compute_metrics(np.zeros((20, 10), dtype=np.float32))


def test_metrics_multi_sentence(golden):
    g = golden("metrics")
    rs = np.random.RandomState(int(g["ms_seed"]))
    V = 50
    per = rs.randint(1, 6, size=V)
    assert (per == g["ms_per"]).all()
    S = int(per.sum())
    gt = np.repeat(np.arange(V), per)
    sim = rs.randn(S, V).astype(np.float32)
    sim[np.arange(S), gt] += 1.5
    cut = (np.cumsum(per) - 1).tolist()
    tv, vt = O.logging_rank(sim, True, cut)
    assert [tv[k] for k in TVK] == list(g["ms_tv"])
    assert [vt[k] for k in KEYS] == list(g["ms_vt"])
    # integer-rank form
    t2v, v2t = O.ranks_multi_sentence(sim, gt)
    assert O.metrics_from_ranks(v2t) == vt
    assert float(np.mean(t2v + 1)) == tv["MeanR"]


def test_similarity(golden):
    g = golden("similarity")
    s2 = O.loose_similarity(g["q"], g["v"])
    s3 = O.loose_similarity(g["q"], g["fr"])
    assert s2.shape == g["s2"].shape and s3.shape == g["s3"].shape
    np.testing.assert_allclose(s2, g["s2"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(s3, g["s3"], rtol=0, atol=2e-5)
    assert abs(O.cross_en(g["s2"][:7, :7]) - float(g["ce"])) < 1e-5


@pytest.mark.parametrize("B", [32, 256])
def test_finetune(golden, B):
    g = golden("finetune_B%d" % B)
    t, v, fr = syn.finetune_inputs(B, seed=int(g["seed"]))
    l32 = O.finetune_loss(t, v, fr)
    assert abs(l32 - float(g["loss"])) / float(g["loss"]) < 2e-6
    l64, dt, dv, dfr = O.finetune_loss_and_grads(t, v, fr)
    assert abs(l64 - float(g["loss"])) / float(g["loss"]) < 2e-6
    gn = [np.linalg.norm(dt), np.linalg.norm(dv), np.linalg.norm(dfr)]
    np.testing.assert_allclose(gn, g["gnorm"], rtol=1e-4)   # the fp32 reference itself carries ~2e-5 noise at scale 100
    assert _rel(dt[:g["dt"].shape[0]], g["dt"]) < 1e-4
    assert _rel(dv[:g["dv"].shape[0]], g["dv"]) < 1e-4
    assert _rel(dfr[:g["dfr"].shape[0]], g["dfr"]) < 1e-4


def test_contrastive_small(golden):
    g = golden("contrastive_small")
    l = O.contrastive_loss(g["q"], g["k"], g["queue"], float(g["T"]))
    assert abs(l - float(g["loss"])) < 2e-6
    l64, dq = O.contrastive_loss_and_grad(g["q"], g["k"], g["queue"], float(g["T"]))
    assert abs(l64 - float(g["loss"])) < 2e-6
    assert _rel(dq, g["dq"]) < 1e-5


@pytest.mark.parametrize("tag", ["small", "b32"])
def test_pretrain(golden, tag):
    g = golden("pretrain_" + tag)
    b, F, D, K, T = int(g["b"]), int(g["F"]), int(g["D"]), int(g["K"]), float(g["T"])
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = syn.queues(K, F=F, D=D, seed=3)
    fam, vtm, ftm = O.pretrain_loss_parts(inp, qs, T)
    assert abs(fam - float(g["fam"])) / float(g["fam"]) < 3e-6
    assert abs(vtm - float(g["vtm"])) / float(g["vtm"]) < 3e-6
    assert abs(ftm - float(g["ftm"])) / float(g["ftm"]) < 3e-6
    loss, grads = O.pretrain_loss_and_grads(inp, qs, T)
    assert abs(loss - float(g["loss"])) / float(g["loss"]) < 3e-6
    for n in ("v_fea", "title_fea", "frame_fea", "frame_pred"):
        ref = g["d_" + n]
        assert _rel(grads[n][:ref.shape[0]], ref) < 3e-5, n
        assert abs(np.linalg.norm(grads[n]) - float(g["gn_" + n])) / float(g["gn_" + n]) < 3e-5
    # enqueue
    ptr = O.dequeue_and_enqueue(qs, 0, inp["v_fea_k"], inp["tag_fea_k"], inp["title_fea_k"],
                                inp["frame_fea_k"], inp["frame_proj_k"], K)
    assert ptr == int(g["ptr"])
    for n in syn.QUEUE_NAMES:
        ref = g["after_" + n]
        np.testing.assert_allclose(qs[n][:, :ref.shape[1]], ref, rtol=0, atol=1e-7)
        if "sum_" + n in g.files:
            np.testing.assert_allclose(qs[n].astype(np.float64).sum(axis=1), g["sum_" + n], rtol=0, atol=1e-5)


def test_enqueue_wrap(golden):
    g = golden("enqueue_wrap")
    qs = syn.queues(16, F=4, D=64, seed=3)
    ptr, ptrs = 0, []
    for step in range(3):
        k = syn.pretrain_inputs(8, F=4, D=64, seed=20 + step)
        ptr = O.dequeue_and_enqueue(qs, ptr, k["v_fea_k"], k["tag_fea_k"], k["title_fea_k"],
                                    k["frame_fea_k"], k["frame_proj_k"], 16)
        ptrs.append(ptr)
    assert ptrs == list(g["ptrs"])
    for n in syn.QUEUE_NAMES:
        np.testing.assert_allclose(qs[n], g["after_" + n], rtol=0, atol=1e-7)
    with pytest.raises(ValueError):
        k = syn.pretrain_inputs(12, F=4, D=64, seed=1)
        O.dequeue_and_enqueue(qs, 8, k["v_fea_k"], k["tag_fea_k"], k["title_fea_k"],
                              k["frame_fea_k"], k["frame_proj_k"], 16)


def test_ema_bit_exact(golden):
    g = golden("ema")
    ps, pks = syn.ema_tensors()
    for _ in range(int(g["steps"])):
        pks = O.momentum_update(ps, pks, float(g["m"]))
    for i, pk in enumerate(pks):
        ref = g["out%d" % i]
        assert pk.dtype == ref.dtype
        assert np.array_equal(pk.view(np.uint8), ref.view(np.uint8)), i


def test_eval_1k(golden):
    g = golden("eval_1k")
    T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=int(g["seed"]))
    k = int(g["top_frames"])
    tl = lambda x: [x[i:i + 256] for i in range(0, x.shape[0], 256)]
    a, b, c = O.run_on_single_gpu(tl(T), tl(V), [np.zeros_like(x) for x in tl(V)], tl(Fr), k)
    sim = np.concatenate(a, 0)
    simf = np.concatenate(c, 0)
    assert bool(g["nan_title"]) and np.isnan(np.concatenate(b, 0)).all()
    np.testing.assert_allclose(sim[:4], g["sim_rows"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(simf[:4], g["simf_rows"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(np.diag(sim), g["sim_diag"], rtol=0, atol=3e-5)
    tot = sim + simf
    tv, vt = O.logging_rank(tot, False, None)
    assert [tv[k_] for k_ in KEYS] == list(g["tv"])
    assert [vt[k_] for k_ in KEYS] == list(g["vt"])
    assert (O.ranks_square(tot) == g["ranks_t2v"]).all()
    assert (O.ranks_square(tot.T) == g["ranks_v2t"]).all()
    np.testing.assert_allclose(O.eval_scores(T, V, Fr, k), tot, rtol=0, atol=1e-6)


def test_eval_multi(golden):
    g = golden("eval_multi")
    per = g["per"]
    T, V, Fr, gt, cut = syn.eval_inputs(int(per.sum()), 60, seed=int(g["seed"]), per_video=per)
    tot = O.eval_scores(T, V, Fr, int(g["top_frames"]), tile=128)
    np.testing.assert_allclose(tot, g["tot"], rtol=0, atol=3e-5)
    tv, vt = O.logging_rank(g["tot"], True, cut)
    assert [tv[k] for k in TVK] == list(g["tv"])
    assert [vt[k] for k in KEYS] == list(g["vt"])
    t2v, v2t = O.ranks_multi_sentence(g["tot"], gt)
    assert O.metrics_from_ranks(v2t) == vt
    assert float(np.mean(t2v + 1)) == tv["MeanR"]


def test_dist_collect_contract():
    rs = np.random.RandomState(3)
    W, b = 4, 3
    xs = [rs.randn(b, 5) for _ in range(W)]
    full = O.dist_collect_emulated(xs)
    assert full.shape == (W * b, 5) and np.array_equal(full[b:2 * b], xs[1])
    grads = [rs.randn(W * b, 5) for _ in range(W)]
    back = O.dist_collect_emulated_backward(grads, b)
    np.testing.assert_allclose(back[2], sum(g[2 * b:3 * b] for g in grads))


def test_torch_port(golden):
    """The CPU-baseline port (oracle/torch_port.py) reproduces the reference's outputs."""
    import torch
    from oracle import torch_port as P
    g = golden("pretrain_small")
    b, F, D, K, T = int(g["b"]), int(g["F"]), int(g["D"]), int(g["K"]), float(g["T"])
    inp = {n: torch.from_numpy(x).requires_grad_(n in ("v_fea", "title_fea", "frame_fea", "frame_pred"))
           for n, x in syn.pretrain_inputs(b, F=F, D=D, seed=2).items()}
    qs = {n: torch.from_numpy(x) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
    loss, ptr = P.pretrain_step(inp, qs, 0, K, T)
    assert abs(loss - float(g["loss"])) < 1e-5 and ptr == int(g["ptr"])
    for n in ("v_fea", "title_fea", "frame_fea", "frame_pred"):
        assert _rel(inp[n].grad.numpy(), g["d_" + n]) < 1e-5
    for n in syn.QUEUE_NAMES:
        np.testing.assert_allclose(qs[n].numpy(), g["after_" + n], rtol=0, atol=1e-7)
    g = golden("finetune_B32")
    t, v, fr = [torch.from_numpy(x).requires_grad_(True) for x in syn.finetune_inputs(32, seed=1)]
    assert abs(P.finetune_step(t, v, fr) - float(g["loss"])) < 1e-4
    assert _rel(t.grad.numpy(), g["dt"]) < 1e-5
    g = golden("eval_1k")
    Tn, Vn, Fn, _, _ = syn.eval_inputs(1000, 1000, seed=4)
    sim, (tv, vt) = P.eval_sim_and_rank(torch.from_numpy(Tn), torch.from_numpy(Vn), torch.from_numpy(Fn), 2)
    assert tv["R1"] == g["tv"][0] and vt["R1"] == g["vt"][0] and tv["MeanR"] == g["tv"][4]


def run_optim_oracle(case, nsteps, gmax, keep):
    from oracle import optim_oracle as OO
    groups = syn.optim_groups(case)
    p = syn.optim_tensors()
    m = [np.zeros_like(x) for x in p]
    v = [np.zeros_like(x) for x in p]
    steps = [0] * len(p)
    kept, totals = {}, []
    for st in range(nsteps):
        p, m, v, _, tot = OO.clip_and_step(p, syn.optim_grads(st), m, v, steps, groups, syn.OPTIM_GROUP_OF, gmax)
        totals.append(tot)
        if st in keep:
            kept[st] = p
    lrs = sorted({OO.lr_scheduled(groups[gi], steps[i]) for i, gi in enumerate(syn.OPTIM_GROUP_OF)})
    return kept, m, v, totals, lrs


def test_optim_oracle_unclipped_is_bit_exact(golden):
    """No clipping anywhere ('plain'): next_m / next_v bit for bit, p within 1 ulp (torch's CPU sqrt
    is not correctly rounded, see oracle/optim_oracle.py)."""
    g = golden("optim")
    kept, m, v, _, lrs = run_optim_oracle("plain", 3, None, (0, 1, 2))
    n = len(m)
    for i in range(n):
        assert np.array_equal(m[i], g["plain_m%d" % i]), i
        assert np.array_equal(v[i], g["plain_v%d" % i]), i
    for st in (0, 1, 2):
        for i in range(n):
            ref = g["plain_p%d_s%d" % (i, st)]
            d = np.abs(kept[st][i].astype(np.float64) - ref)
            one_ulp = 2.0 ** -23 * np.maximum(np.abs(ref), 2.0 ** -6)     # of p before the (tiny) update
            assert (d <= one_ulp).all() and (d > 0).mean() < 0.02, (st, i, d.max(), (d > 0).mean())
    assert np.allclose(lrs, g["plain_lrs"][-1], rtol=1e-15)


@pytest.mark.parametrize("case,nsteps,gmax,keep", [("pretrain", 6, 1.0, (0, 5)), ("linear", 4, None, (3,))])
def test_optim_oracle_clipped(golden, case, nsteps, gmax, keep):
    """Global (clip_grad_norm_) and per-parameter (inside BertAdam.step) clipping active."""
    g = golden("optim")
    kept, m, v, totals, lrs = run_optim_oracle(case, nsteps, gmax, keep)
    if gmax is not None:
        np.testing.assert_allclose(np.array(totals, np.float64), g[case + "_totals"], rtol=1e-6)
    for i in range(len(m)):
        # ~1e-7 relative from the norm (see oracle/optim_oracle.py); atol covers cancelled elements
        for got, ref in [(m[i], g["%s_m%d" % (case, i)]), (v[i], g["%s_v%d" % (case, i)])] + \
                [(kept[st][i], g["%s_p%d_s%d" % (case, i, st)]) for st in keep]:
            np.testing.assert_allclose(got, ref, rtol=3e-6, atol=1e-6 * float(np.abs(ref).max()))
    np.testing.assert_allclose(lrs, g[case + "_lrs"][-1], rtol=1e-15)


def test_optim_torch_port(golden):
    """The CPU-baseline port of the optimizer step is the reference's op sequence."""
    import torch
    from oracle import torch_port as P
    g = golden("optim")
    groups = syn.optim_groups("pretrain")
    params = [torch.from_numpy(x.copy()) for x in syn.optim_tensors()]
    state = P.bert_adam_state(params)
    for st in range(6):
        grads = [torch.from_numpy(x) for x in syn.optim_grads(st)]
        P.clip_and_bert_adam_step(params, grads, state, groups, syn.OPTIM_GROUP_OF, 1.0)
    for i, p in enumerate(params):
        assert np.array_equal(p.numpy(), g["pretrain_p%d_s5" % i]), i
        assert np.array_equal(state[i]['next_m'].numpy(), g["pretrain_m%d" % i]), i


def test_mlp_oracle_vs_reference_class(golden):
    """oracle/mlp_oracle.py against the reference's own MLP (Linear -> BatchNorm1d -> ReLU -> Linear) run on
    CPU in fp32: outputs, every gradient, the running statistics and the eval-mode forward."""
    from oracle import mlp_oracle as MO
    g = golden("mlp")
    c = syn.mlp_case()
    y, cache = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    assert _rel(y, g["y"]) < 2e-6
    grads = MO.backward(c["dy"], cache, c["W1"], c["gamma"], c["W2"])
    for k in ("dx", "dW1", "dgamma", "dbeta", "dW2", "db2"):
        assert _rel(grads[k], g[k]) < 5e-6, k
    # db1 is zero in exact arithmetic (the bias cancels inside the BatchNorm): only rounding noise on both sides
    assert np.abs(grads["db1"]).max() < 1e-12 and np.abs(g["db1"]).max() < 1e-6
    rm, rv = MO.running_stats(cache, c["rm"], c["rv"])
    assert _rel(rm, g["rm"]) < 1e-6 and _rel(rv, g["rv"]) < 1e-6
    y_eval, _ = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"], running=(rm, rv))
    assert _rel(y_eval, g["y_eval"]) < 2e-6


def test_mlp_mirror_state_dict_keys(golden):
    from hmmc_b200.mlp import MLP
    g = golden("mlp")
    m = MLP(in_dim=64, inner_dim=128, out_dim=64, num_layers=2)
    assert sorted(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    import torch
    with pytest.raises(Exception):
        m(torch.zeros(4, 64))            # CPU tensors: there is no CPU path


def test_visual_tail_oracle_vs_reference(golden):
    from oracle import mlp_oracle as MO
    g = golden("visual_tail")
    orig = g["orig"].reshape(6, 12, 64)
    temp = g["temp"].transpose(1, 0, 2)                 # LND -> NLD as VisualEncoder.forward permutes it
    for tag, t in (("temp", temp), ("notemp", None)):
        out, h, n = MO.visual_tail(t, orig)
        assert _rel(out, g["out_" + tag]) < 2e-6
        dh = MO.visual_tail_backward(g["g"], h, n)
        assert _rel(dh.reshape(72, 64), g["dorig_" + tag]) < 5e-6
        if t is not None:
            assert _rel(dh.transpose(1, 0, 2), g["dtemp"]) < 5e-6
