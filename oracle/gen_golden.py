"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

Each fixture stores the recipe parameters (inputs are re-drawn from
``hmmc_b200.synthetic`` with the same seeds on the checking side, or stored in
full when small) and the outputs of the reference's own functions.  TEST
INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hmmc_b200 import synthetic as syn      # noqa: E402
from oracle import ref_shim                 # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _t(x, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).clone()
    if grad:
        t.requires_grad_(True)
    return t


def gen_metrics(M, metrics, R):
    rs = np.random.RandomState(0)
    x = rs.randn(1000, 1000).astype(np.float32)
    x[np.arange(1000), np.arange(1000)] += 2.0
    a = metrics.compute_metrics(x)
    b = metrics.compute_metrics(x.T)
    ties = np.array([[1, 1, 0], [0, 2, 3], [0, 0, 5]], dtype=np.float32)
    c = metrics.compute_metrics(ties)
    keys = ["R1", "R5", "R10", "MR", "MeanR"]
    # multi-sentence
    rs = np.random.RandomState(7)
    V = 50
    per = rs.randint(1, 6, size=V)
    S = int(per.sum())
    gt = np.repeat(np.arange(V), per)
    sim = rs.randn(S, V).astype(np.float32)
    sim[np.arange(S), gt] += 1.5
    cut = (np.cumsum(per) - 1).tolist()

    class _L:
        def info(self, *a, **k):
            pass
    # logging_rank returns only tv; run the two halves exactly as it does
    cut2 = [c_ + 1 for c_ in cut]
    maxlen = max(e - s for s, e in zip([0] + cut2[:-1], cut2))
    blocks = [np.concatenate((sim[s:e], np.full((maxlen - e + s, V), -np.inf)), axis=0)
              for s, e in zip([0] + cut2[:-1], cut2)]
    sim3 = np.stack(blocks, axis=0)
    tv = metrics.tensor_text_to_video_metrics(sim3)
    vt = metrics.compute_metrics(metrics.tensor_video_to_text_sim(sim3))
    tv_lr = metrics.logging_rank(sim, True, cut, _L())
    assert tv_lr == tv
    tvk = ["R1", "R5", "R10", "MedianR", "MeanR", "Std_Rank", "MR"]
    np.savez(os.path.join(OUT, "metrics.npz"),
             sq_t2v=np.array([a[k] for k in keys], dtype=np.float64),
             sq_v2t=np.array([b[k] for k in keys], dtype=np.float64),
             ties=np.array([c[k] for k in keys], dtype=np.float64),
             ms_per=per, ms_seed=7,
             ms_tv=np.array([tv[k] for k in tvk], dtype=np.float64),
             ms_vt=np.array([vt[k] for k in keys], dtype=np.float64))
    print("metrics: sq", a, "| ms tv", tv, "| vt", vt)


def gen_similarity(M, metrics, R):
    s = ref_shim.finetune_self()
    rs = np.random.RandomState(11)
    q = rs.randn(9, 64).astype(np.float32)
    v = rs.randn(7, 64).astype(np.float32)
    fr = rs.randn(7, 5, 64).astype(np.float32)
    s2 = s.loose_similarity(_t(q), _t(v)).numpy()
    s3 = s.loose_similarity(_t(q), _t(fr)).numpy()
    ce = float(M.CrossEn()(_t(s2[:7, :7])))
    np.savez(os.path.join(OUT, "similarity.npz"), q=q, v=v, fr=fr, s2=s2, s3=s3, ce=ce)
    print("similarity:", s2.shape, s3.shape, ce)


def gen_finetune(M, metrics, R):
    for B, full in ((32, True), (256, False)):
        t, v, fr = syn.finetune_inputs(B, seed=1)
        tt, tv, tf = _t(t, True), _t(v, True), _t(fr, True)
        loss = ref_shim.finetune_forward(tt, tv, tf)
        loss.backward()
        out = dict(B=B, seed=1, loss=float(loss),
                   gnorm=np.array([tt.grad.norm().item(), tv.grad.norm().item(), tf.grad.norm().item()]))
        if full:
            out.update(dt=tt.grad.numpy(), dv=tv.grad.numpy(), dfr=tf.grad.numpy())
        else:
            out.update(dt=tt.grad.numpy()[:8], dv=tv.grad.numpy()[:8], dfr=tf.grad.numpy()[:2])
        np.savez_compressed(os.path.join(OUT, "finetune_B%d.npz" % B), **out)
        print("finetune B=%d loss %.6f gnorm %s" % (B, float(loss), out["gnorm"]))


def gen_contrastive(M, metrics, R):
    """contrastive_loss on a small problem, all inputs stored."""
    rs = np.random.RandomState(13)
    b, D, Kq = 6, 32, 40
    q = rs.randn(b, D).astype(np.float32)
    k = (rs.randn(b, D) + 0.5 * q).astype(np.float32)
    queue = syn.normalize_cols(rs.randn(D, Kq).astype(np.float32))
    s = ref_shim.pretrain_self({}, K=Kq, F=1)
    tq = _t(q, True)
    loss = s.contrastive_loss(tq, _t(k), _t(queue))
    loss.backward()
    np.savez(os.path.join(OUT, "contrastive_small.npz"), q=q, k=k, queue=queue, T=0.07,
             loss=float(loss), dq=tq.grad.numpy())
    print("contrastive small loss", float(loss))


def gen_pretrain(M, metrics, R):
    for tag, b, F, D, K, full in (("small", 8, 4, 64, 64, True), ("b32", 32, 12, 512, 1024, False)):
        inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
        qs = syn.queues(K, F=F, D=D, seed=3)
        tin = {n: _t(x, n in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for n, x in inp.items()}
        tq = {n: _t(x) for n, x in qs.items()}
        # head losses one by one (same bound methods forward() uses)
        s0 = ref_shim.pretrain_self({n: x.clone() for n, x in tq.items()}, K, F)
        with torch.no_grad():
            fam = float(s0.frame_self_loss(tin["frame_pred"], tin["frame_proj_k"], s0.queue_frame_proj_ng))
            vtm = float(s0.contrastive_loss(tin["v_fea"], tin["title_fea_k"], s0.queue_title_cross_ng)
                        + s0.contrastive_loss(tin["title_fea"], tin["v_fea_k"], s0.queue_v_cross_ng))
            ftm = float(s0.frame_cross_loss(tin["frame_fea"], tin["frame_fea_k"], s0.queue_frame_cross_ng,
                                            tin["title_fea"], tin["title_fea_k"], s0.queue_title_cross_ng))
        loss, s = ref_shim.pretrain_forward(tin, tq, K, F)
        loss.backward()
        out = dict(b=b, F=F, D=D, K=K, T=0.07, loss=float(loss), fam=fam, vtm=vtm, ftm=ftm,
                   ptr=int(s.queue_ptr))
        gn = {}
        for n in ("v_fea", "title_fea", "frame_fea", "frame_pred"):
            g = tin[n].grad.numpy()
            gn[n] = float(np.linalg.norm(g))
            out["d_" + n] = g if (full or g.ndim == 2) else g[:4]
            out["gn_" + n] = gn[n]
        for n in syn.QUEUE_NAMES:
            qa = getattr(s, n).numpy()
            if full:
                out["after_" + n] = qa
            else:
                Fq = F if "frame" in n else 1
                out["after_" + n] = qa[:, : 3 * Fq]            # first three enqueued entries
                out["sum_" + n] = qa.astype(np.float64).sum(axis=1)
        np.savez_compressed(os.path.join(OUT, "pretrain_%s.npz" % tag), **out)
        print("pretrain %s loss %.6f fam %.6f vtm %.6f ftm %.6f ptr %d gn %s" % (tag, float(loss), fam, vtm, ftm, out["ptr"], gn))
    # second enqueue step on the small case: pointer advance and wrap-around
    inp = syn.pretrain_inputs(8, F=4, D=64, seed=2)
    qs = {n: _t(x) for n, x in syn.queues(16, F=4, D=64, seed=3).items()}
    s = ref_shim.pretrain_self(qs, 16, 4)
    ptrs = []
    for step in range(3):
        keys = syn.pretrain_inputs(8, F=4, D=64, seed=20 + step)
        s._dequeue_and_enqueue(_t(keys["v_fea_k"]), _t(keys["tag_fea_k"]), _t(keys["title_fea_k"]),
                               _t(keys["frame_fea_k"]), _t(keys["frame_proj_k"]))
        ptrs.append(int(s.queue_ptr))
    out = {"after_" + n: getattr(s, n).numpy() for n in syn.QUEUE_NAMES}
    np.savez_compressed(os.path.join(OUT, "enqueue_wrap.npz"), ptrs=np.array(ptrs), **out)
    print("enqueue ptrs", ptrs)


def gen_ema(M, metrics, R):
    ps, pks = syn.ema_tensors()

    class _Mod:
        def __init__(self, xs):
            self.xs = [torch.nn.Parameter(_t(x), requires_grad=False) for x in xs]

        def parameters(self):
            return self.xs
    s = types.SimpleNamespace(contrast_momentum=0.99)
    a, b = _Mod(ps), _Mod(pks)
    s.model_pairs = [[a, b]]
    for _ in range(3):
        M.BirdPreTrainedModel._momentum_update(s)
    out = {"out%d" % i: p.data.numpy() for i, p in enumerate(b.xs)}
    # copy_params
    c = _Mod(pks)
    s.model_pairs = [[a, c]]
    M.BirdPreTrainedModel.copy_params(s)
    assert all(torch.equal(x.data, y.data) for x, y in zip(a.xs, c.xs))
    np.savez(os.path.join(OUT, "ema.npz"), steps=3, m=0.99, **out)
    print("ema ok", [o.dtype for o in out.values()])


def gen_eval(M, metrics, R):
    # square 1000 x 1000 x 12 (config 2), tiles of 256, top_frames 2
    T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=4)
    s = ref_shim.finetune_self(top_frames=2)
    tl = lambda x, n=256: [_t(x[i:i + n]) for i in range(0, x.shape[0], n)]
    with torch.no_grad():
        a, b, c = R._run_on_single_gpu(s, tl(T), tl(V), [torch.zeros_like(x) for x in tl(V)], tl(Fr))
    sim = np.concatenate(a, axis=0)
    simf = np.concatenate(c, axis=0)
    tot = sim + simf
    tv = metrics.compute_metrics(tot)
    vt = metrics.compute_metrics(tot.T)
    keys = ["R1", "R5", "R10", "MR", "MeanR"]
    ranks = (tot > np.diag(tot)[:, None]).sum(1)
    np.savez_compressed(os.path.join(OUT, "eval_1k.npz"), seed=4, top_frames=2,
                        sim_rows=sim[:4], simf_rows=simf[:4], sim_diag=np.diag(sim), simf_diag=np.diag(simf),
                        tv=np.array([tv[k] for k in keys]), vt=np.array([vt[k] for k in keys]),
                        ranks_t2v=ranks, ranks_v2t=(tot.T > np.diag(tot)[:, None]).sum(1),
                        nan_title=bool(np.isnan(np.concatenate(b, axis=0)).all()))
    print("eval 1k tv", tv, "vt", vt)
    # multi-sentence 300 texts x 60 videos x 12, top_frames 3
    rs = np.random.RandomState(9)
    per = rs.randint(1, 10, size=60)
    per[-1] += 300 - per.sum() if per.sum() < 300 else 0
    while per.sum() > 300:
        per[np.argmax(per)] -= 1
    T, V, Fr, gt, cut = syn.eval_inputs(int(per.sum()), 60, seed=5, per_video=per)
    s = ref_shim.finetune_self(top_frames=3)
    with torch.no_grad():
        a, b, c = R._run_on_single_gpu(s, tl(T, 128), tl(V, 32), [torch.zeros_like(x) for x in tl(V, 32)], tl(Fr, 32))
    tot = np.concatenate(a, axis=0) + np.concatenate(c, axis=0)

    class _L:
        def info(self, *a, **k):
            pass
    cut2 = [c_ + 1 for c_ in cut]
    maxlen = max(e - s_ for s_, e in zip([0] + cut2[:-1], cut2))
    sim3 = np.stack([np.concatenate((tot[s_:e], np.full((maxlen - e + s_, 60), -np.inf)), axis=0)
                     for s_, e in zip([0] + cut2[:-1], cut2)], axis=0)
    tvm = metrics.tensor_text_to_video_metrics(sim3)
    vtm = metrics.compute_metrics(metrics.tensor_video_to_text_sim(sim3))
    tvk = ["R1", "R5", "R10", "MedianR", "MeanR", "Std_Rank", "MR"]
    np.savez_compressed(os.path.join(OUT, "eval_multi.npz"), seed=5, top_frames=3, per=per,
                        tot=tot, tv=np.array([tvm[k] for k in tvk]), vt=np.array([vtm[k] for k in keys]))
    print("eval multi tv", tvm, "vt", vtm)


def gen_eval_epoch(M, metrics, R):
    """The reference's eval_epoch (main_task_retrieval.py:358-524) on CPU with index-lookup stub encoders."""
    import logging
    R.logger = logging.getLogger("gen_golden")
    out = {}
    for tag, multi, batch in (("square", False, 64), ("multi", True, 32)):
        T, V, Fr, batches, cut = syn.eval_epoch_case(multi, batch)
        s = ref_shim.finetune_self(top_frames=2)

        class _Txt:
            logit_scale = s.text_encoder.logit_scale

            def __call__(self, ids, mask):
                return _t(T)[ids[:, 0]]
        s.text_encoder = _Txt()
        s.visual_encoder = lambda video, video_frame: (_t(V)[video[:, 0]], _t(Fr)[video[:, 0]])
        s.to = lambda device: s
        s.eval = lambda: None

        class _DS:
            multi_sentence_per_video = multi
            cut_off_points = cut
            sentence_num = T.shape[0]
            video_num = V.shape[0]

        class _DL(list):
            dataset = _DS()
        dl = _DL([(torch.from_numpy(ci)[:, None], torch.ones(len(ci), 1, dtype=torch.long),
                   torch.from_numpy(vi)[:, None], torch.full((len(ci),), 12)) for ci, vi in batches])
        args = types.SimpleNamespace(task="retrieval", use_frame_fea=True)
        tv = R.eval_epoch(args, s, dl, torch.device("cpu"), 1)
        out[tag + "_keys"] = np.array(sorted(tv.keys()))
        out[tag + "_vals"] = np.array([tv[k] for k in sorted(tv.keys())], dtype=np.float64)
        print("eval_epoch", tag, tv)
    np.savez(os.path.join(OUT, "eval_epoch.npz"), **out)


def gen_mlp(M, metrics, R):
    """The reference's MLP class (modules/modeling.py:788-807) on CPU: training-mode forward + backward,
    running statistics afterwards, and an eval-mode forward."""
    c = syn.mlp_case()
    m = M.MLP(in_dim=64, inner_dim=128, out_dim=64, num_layers=2)
    lin1, bn = m.linear_hidden[1], m.linear_hidden[2]
    with torch.no_grad():
        lin1.weight.copy_(_t(c["W1"])); lin1.bias.copy_(_t(c["b1"]))
        bn.weight.copy_(_t(c["gamma"])); bn.bias.copy_(_t(c["beta"]))
        bn.running_mean.copy_(_t(c["rm"])); bn.running_var.copy_(_t(c["rv"]))
        m.linear_out.weight.copy_(_t(c["W2"])); m.linear_out.bias.copy_(_t(c["b2"]))
    m.train()
    x = _t(c["x"], grad=True)
    y = m(x)
    y.backward(_t(c["dy"]))
    out = dict(y=y.detach().numpy(), dx=x.grad.numpy(), dW1=lin1.weight.grad.numpy(), db1=lin1.bias.grad.numpy(),
               dgamma=bn.weight.grad.numpy(), dbeta=bn.bias.grad.numpy(), dW2=m.linear_out.weight.grad.numpy(),
               db2=m.linear_out.bias.grad.numpy(), rm=bn.running_mean.numpy().copy(), rv=bn.running_var.numpy().copy(),
               nbt=int(bn.num_batches_tracked), keys=np.array(sorted(m.state_dict().keys())))
    m.eval()
    with torch.no_grad():
        out["y_eval"] = m(_t(c["x"])).numpy()
    np.savez_compressed(os.path.join(OUT, "mlp.npz"), **out)
    print("mlp golden: |y|", float(np.abs(out["y"]).mean()), "keys", list(out["keys"]))


def gen_visual_tail(M, metrics, R):
    """VisualEncoder.forward (modules/module_cross.py:178-216) with stub CLIP / temporal transformer that
    return fixed tensors: the residual + normalise + mean tail is the reference's own code."""
    import modules.module_cross as MC
    rs = np.random.RandomState(41)
    bs, F, D = 6, 12, 64
    orig = _t(rs.randn(bs * F, D).astype(np.float32), grad=True)
    temp = _t(rs.randn(F, bs, D).astype(np.float32), grad=True)        # LND, as the temporal transformer returns
    g = rs.randn(bs, D).astype(np.float32)
    out = {}
    for use_temp in (True, False):
        for t in (orig, temp):
            t.grad = None
        s = types.SimpleNamespace(use_temp=use_temp, encode_image=lambda video, video_frame: orig,
                                  frame_position_embeddings=torch.nn.Embedding(48, D),
                                  temporal_transformer=lambda x, mask: temp)
        vo, fo = MC.VisualEncoder.forward(s, torch.zeros(bs, F, 3, 2, 2), F)
        vo.backward(_t(g))
        tag = "temp" if use_temp else "notemp"
        out["out_" + tag] = vo.detach().numpy()
        out["dorig_" + tag] = orig.grad.numpy().copy()
        if use_temp:
            out["dtemp"] = temp.grad.numpy().copy()
        assert fo.shape == (bs, F, D)
    np.savez(os.path.join(OUT, "visual_tail.npz"), orig=orig.detach().numpy(), temp=temp.detach().numpy(), g=g, **out)
    print("visual tail golden", out["out_temp"].shape)


OPTIM_CASES = (("pretrain", 6, 1.0, (0, 5)), ("plain", 3, None, (0, 1, 2)), ("linear", 4, None, (3,)))


def gen_optim(M, metrics, R):
    """clip_grad_norm_ + the reference's BertAdam (modules/optimization.py), a few steps on CPU."""
    from modules.optimization import BertAdam
    out = {}
    for case, nsteps, gmax, keep in OPTIM_CASES:
        groups = syn.optim_groups(case)
        params = [torch.nn.Parameter(_t(x)) for x in syn.optim_tensors()]
        pg = [dict(g, params=[p for p, gi in zip(params, syn.OPTIM_GROUP_OF) if gi == k]) for k, g in enumerate(groups)]
        opt = BertAdam(pg, lr=groups[0]['lr'])
        totals, lrs = [], []
        for st in range(nsteps):
            for p, g in zip(params, syn.optim_grads(st)):
                p.grad = _t(g)
            if gmax is not None:
                totals.append(float(torch.nn.utils.clip_grad_norm_(params, gmax)))
            opt.step()
            lrs.append(sorted(set(opt.get_lr())))
            if st in keep:
                for i, p in enumerate(params):
                    out["%s_p%d_s%d" % (case, i, st)] = p.data.numpy().copy()
            opt.zero_grad()
        for i, p in enumerate(params):
            out["%s_m%d" % (case, i)] = opt.state[p]['next_m'].numpy().copy()
            out["%s_v%d" % (case, i)] = opt.state[p]['next_v'].numpy().copy()
        out["%s_totals" % case] = np.array(totals, dtype=np.float64)
        out["%s_lrs" % case] = np.array(lrs, dtype=np.float64)
        print("optim", case, "steps", nsteps, "total norms", totals, "lr", lrs[-1])
    np.savez_compressed(os.path.join(OUT, "optim.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    M, metrics, R = ref_shim.load()
    only = sys.argv[1:]
    for fn in (gen_metrics, gen_similarity, gen_finetune, gen_contrastive, gen_pretrain, gen_ema, gen_eval, gen_optim, gen_eval_epoch, gen_mlp, gen_visual_tail):
        if only and fn.__name__[4:] not in only:
            continue
        fn(M, metrics, R)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
