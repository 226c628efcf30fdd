"""CPU oracle: a numpy restatement of the HMMC contrastive head.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module; ``hmmc_b200/`` never does (the product path fails loudly without its
CUDA library instead of falling back to this).

Every function restates one reference function and cites it (paths relative to
the reference root).  The reference has no tests and no golden vectors of its
own, so the oracle is pinned against outputs of the reference's
own code executed in the build container: ``oracle/gen_golden.py`` imports the
unmodified reference through ``oracle/ref_shim.py``, writes ``tests/golden/*.npz``
and ``tests/test_oracle_golden.py`` checks this file against them.
The one piece that stays **parity unpinned** is the backward of the
differentiable all-gather (third-party ``diffdist==0.1``, requirements.txt:3,
not vendored): the contract restated in :func:`dist_collect_emulated_backward`
is "forward = concat in rank order, backward to rank r = sum over ranks of their
gradient slice r".

All functions take a ``dtype`` (float32 = as the reference runs, float64 = for
tight gradient checks).
"""
import numpy as np

# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------


def _f_normalize(x, axis, eps=1e-12):
    """torch.nn.functional.normalize: x / max(||x||_2, eps)."""
    n = np.sqrt((x * x).sum(axis=axis, keepdims=True))
    return x / np.maximum(n, x.dtype.type(eps))


def _log_softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=axis, keepdims=True))


def logit_scale_value(logit_scale=4.6052, dtype=np.float32):
    """min(exp(logit_scale), 100)  -- modules/modeling.py:216-217."""
    return np.minimum(np.exp(dtype(logit_scale)), dtype(100.0))


# ----------------------------------------------------------------------------
# fine-tune head: loose_similarity / CrossEn / frame_loss / combine
# ----------------------------------------------------------------------------


def loose_similarity(seq, vis, logit_scale=4.6052, dtype=np.float32):
    """modules/modeling.py:207-229.  2-D vis -> [Bt,Bv]; 3-D vis [Bv,F,D] ->
    [Bt,Bv,F].  No eps in the normalisation (0/0 = NaN is the spec)."""
    seq = np.asarray(seq, dtype=dtype)
    vis = np.asarray(vis, dtype=dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        vis = vis / np.sqrt((vis * vis).sum(axis=-1, keepdims=True))
        seq = seq / np.sqrt((seq * seq).sum(axis=-1, keepdims=True))
    s = logit_scale_value(logit_scale, dtype)
    if vis.ndim == 2:
        return s * (seq @ vis.T)
    # [Bv,F,D] -> matmul(seq [Bt,D], vis^T [Bv,D,F]) = [Bv,Bt,F] -> permute(1,0,2)
    out = s * np.einsum("td,vfd->vtf", seq, vis)
    return np.transpose(out, (1, 0, 2))


def cross_en(sim):
    """CrossEn.forward, modules/until_module.py:196-205."""
    lp = _log_softmax(sim, axis=-1)
    return -np.diag(lp).mean()


def frame_loss(query, frames, logit_scale=4.6052, dtype=np.float32):
    """BirdModel.frame_loss, modules/modeling.py:665-673 (the live loop)."""
    F = frames.shape[1]
    loss = dtype(0.0)
    for i in range(F):
        sim = loose_similarity(query, frames[:, i, :], logit_scale, dtype)
        loss = loss + (cross_en(sim) + cross_en(sim.T)) / F
    return loss


def finetune_loss(t, v, fr, logit_scale=4.6052, w_vtm=0.85, w_ftm=0.15,
                  use_frame_fea=True, dtype=np.float32):
    """BirdModel.forward loss after the all-gather, modules/modeling.py:702-709."""
    loss = dtype(0.0)
    if use_frame_fea:
        loss = loss + dtype(w_ftm) * frame_loss(t, fr, logit_scale, dtype)
    sim = loose_similarity(t, v, logit_scale, dtype)
    loss = loss + dtype(w_vtm) * (cross_en(sim) + cross_en(sim.T))
    return loss


def _sym_ce_grad(that, vhat, s):
    """loss and dL/dthat, dL/dvhat of CE(S)+CE(S^T), S = s*that@vhat^T
    (SURVEY.md §8a': G = (P_row - I + P_col - I)/B)."""
    B = that.shape[0]
    S = s * (that @ vhat.T)
    lr = _log_softmax(S, axis=1)
    lc = _log_softmax(S, axis=0)
    loss = -(np.diag(lr).mean() + np.diag(lc).mean())
    G = (np.exp(lr) + np.exp(lc) - 2.0 * np.eye(B, dtype=S.dtype)) / B
    return loss, s * (G @ vhat), s * (G.T @ that)


def _unnormalize_grad(x, g_hat):
    """Chain rule through x_hat = x/||x|| (no eps): (I - x_hat x_hat^T) g / ||x||."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    xh = x / n
    return (g_hat - xh * (xh * g_hat).sum(axis=-1, keepdims=True)) / n


def finetune_loss_and_grads(t, v, fr, logit_scale=4.6052, w_vtm=0.85, w_ftm=0.15,
                            dtype=np.float64):
    """Analytic forward + backward of :func:`finetune_loss`; returns
    (loss, dt, dv, dfr)."""
    t = np.asarray(t, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    fr = np.asarray(fr, dtype=dtype)
    s = logit_scale_value(logit_scale, dtype)
    F = fr.shape[1]
    that = t / np.sqrt((t * t).sum(-1, keepdims=True))
    vhat = v / np.sqrt((v * v).sum(-1, keepdims=True))
    fhat = fr / np.sqrt((fr * fr).sum(-1, keepdims=True))
    l0, gt0, gv0 = _sym_ce_grad(that, vhat, s)
    loss = w_vtm * l0
    g_that = w_vtm * gt0
    g_vhat = w_vtm * gv0
    g_fhat = np.zeros_like(fhat)
    for i in range(F):
        li, gti, gfi = _sym_ce_grad(that, fhat[:, i, :], s)
        loss = loss + (w_ftm / F) * li
        g_that = g_that + (w_ftm / F) * gti
        g_fhat[:, i, :] = (w_ftm / F) * gfi
    return (dtype(loss), _unnormalize_grad(t, g_that), _unnormalize_grad(v, g_vhat),
            _unnormalize_grad(fr, g_fhat))


# ----------------------------------------------------------------------------
# pre-train head: InfoNCE vs queue, FAM / VTM / FTM
# ----------------------------------------------------------------------------


def contrastive_loss(q, k, queue, T=0.07, dtype=np.float32):
    """BirdPreTrainedModel.contrastive_loss, modules/modeling.py:286-313.
    q,k [b,D]; queue [D,Kq] (state-dict layout)."""
    q = _f_normalize(np.asarray(q, dtype=dtype), 1)
    k = _f_normalize(np.asarray(k, dtype=dtype), 1)
    queue = np.asarray(queue, dtype=dtype)
    l_pos = np.diag(q @ k.T).reshape(-1, 1)
    l_neg = q @ queue
    logits = np.concatenate([l_pos, l_neg], axis=1) / dtype(T)
    return -_log_softmax(logits, axis=1)[:, 0].mean()


def contrastive_loss_and_grad(q, k, queue, T=0.07, dtype=np.float64):
    """loss and dL/dq (only q receives a gradient: k is under no_grad and the
    queue is detached, modules/modeling.py:302,368-378)."""
    q = np.asarray(q, dtype=dtype)
    k = np.asarray(k, dtype=dtype)
    queue = np.asarray(queue, dtype=dtype)
    b = q.shape[0]
    nq = np.maximum(np.sqrt((q * q).sum(1, keepdims=True)), 1e-12)
    qh = q / nq
    kh = _f_normalize(k, 1)
    l_pos = (qh * kh).sum(1, keepdims=True) / T
    l_neg = (qh @ queue) / T
    logits = np.concatenate([l_pos, l_neg], axis=1)
    lp = _log_softmax(logits, axis=1)
    loss = -lp[:, 0].mean()
    p = np.exp(lp)
    g_hat = ((p[:, :1] - 1.0) * kh + p[:, 1:] @ queue.T) / (b * T)
    g = (g_hat - qh * (qh * g_hat).sum(1, keepdims=True)) / nq
    return dtype(loss), g


def frame_self_loss(frame_fea, frame_fea_k, queue, T=0.07, dtype=np.float32):
    """FAM, modules/modeling.py:315-323."""
    F = frame_fea.shape[1]
    loss = dtype(0.0)
    for i in range(F - 1):
        loss = loss + contrastive_loss(frame_fea[:, i], frame_fea_k[:, i + 1], queue, T, dtype) \
            + contrastive_loss(frame_fea[:, i + 1], frame_fea_k[:, i], queue, T, dtype)
    return loss / (F - 1)


def frame_cross_loss(frame_fea, frame_fea_k, queue_frame, text_fea, text_fea_k, queue_text,
                     T=0.07, dtype=np.float32):
    """FTM, modules/modeling.py:325-332."""
    F = frame_fea.shape[1]
    loss = dtype(0.0)
    for i in range(F):
        loss = loss + contrastive_loss(text_fea, frame_fea_k[:, i], queue_frame, T, dtype) \
            + contrastive_loss(frame_fea[:, i], text_fea_k, queue_text, T, dtype)
    return loss / F


PRETRAIN_WEIGHTS = dict(FAM=0.05, VTM=0.45, FTM=0.45, MLM=0.05)  # modules/cross-base/cross_config.json


def pretrain_loss_parts(inp, queues, T=0.07, dtype=np.float32):
    """The three head losses of BirdPreTrainedModel.forward,
    modules/modeling.py:385-400 (dataset != 'bird' branch)."""
    fam = frame_self_loss(inp["frame_pred"], inp["frame_proj_k"], queues["queue_frame_proj_ng"], T, dtype)
    vtm = contrastive_loss(inp["v_fea"], inp["title_fea_k"], queues["queue_title_cross_ng"], T, dtype) \
        + contrastive_loss(inp["title_fea"], inp["v_fea_k"], queues["queue_v_cross_ng"], T, dtype)
    ftm = frame_cross_loss(inp["frame_fea"], inp["frame_fea_k"], queues["queue_frame_cross_ng"],
                           inp["title_fea"], inp["title_fea_k"], queues["queue_title_cross_ng"], T, dtype)
    return fam, vtm, ftm


def pretrain_loss(inp, queues, T=0.07, weights=PRETRAIN_WEIGHTS, mlm=0.0, dtype=np.float32):
    """modules/modeling.py:424 with the MLM term supplied by the caller."""
    fam, vtm, ftm = pretrain_loss_parts(inp, queues, T, dtype)
    return dtype(weights["FAM"]) * fam + dtype(weights["VTM"]) * vtm + dtype(weights["FTM"]) * ftm \
        + dtype(weights["MLM"]) * dtype(mlm)


def pretrain_loss_and_grads(inp, queues, T=0.07, weights=PRETRAIN_WEIGHTS, dtype=np.float64):
    """Analytic loss and gradients w.r.t. v_fea, title_fea, frame_fea, frame_pred."""
    F = inp["frame_fea"].shape[1]
    g = {n: np.zeros(inp[n].shape, dtype=dtype) for n in ["v_fea", "title_fea", "frame_fea", "frame_pred"]}
    loss = dtype(0.0)
    # FAM
    w = weights["FAM"] / (F - 1)
    Q = queues["queue_frame_proj_ng"]
    for i in range(F - 1):
        l, gi = contrastive_loss_and_grad(inp["frame_pred"][:, i], inp["frame_proj_k"][:, i + 1], Q, T, dtype)
        loss += w * l
        g["frame_pred"][:, i] += w * gi
        l, gi = contrastive_loss_and_grad(inp["frame_pred"][:, i + 1], inp["frame_proj_k"][:, i], Q, T, dtype)
        loss += w * l
        g["frame_pred"][:, i + 1] += w * gi
    # VTM
    w = weights["VTM"]
    l, gi = contrastive_loss_and_grad(inp["v_fea"], inp["title_fea_k"], queues["queue_title_cross_ng"], T, dtype)
    loss += w * l
    g["v_fea"] += w * gi
    l, gi = contrastive_loss_and_grad(inp["title_fea"], inp["v_fea_k"], queues["queue_v_cross_ng"], T, dtype)
    loss += w * l
    g["title_fea"] += w * gi
    # FTM
    w = weights["FTM"] / F
    for i in range(F):
        l, gi = contrastive_loss_and_grad(inp["title_fea"], inp["frame_fea_k"][:, i],
                                          queues["queue_frame_cross_ng"], T, dtype)
        loss += w * l
        g["title_fea"] += w * gi
        l, gi = contrastive_loss_and_grad(inp["frame_fea"][:, i], inp["title_fea_k"],
                                          queues["queue_title_cross_ng"], T, dtype)
        loss += w * l
        g["frame_fea"][:, i] += w * gi
    return loss, g


# ----------------------------------------------------------------------------
# momentum encoder EMA and queue enqueue
# ----------------------------------------------------------------------------


def momentum_update(params, params_k, m=0.99):
    """_momentum_update, modules/modeling.py:238-242: p_k <- p_k*m + p*(1-m),
    evaluated in the tensor's own dtype with python-float scalars: three
    separately rounded ops (torch computes each in fp32 opmath and rounds to the
    storage dtype)."""
    out = []
    for p, pk in zip(params, params_k):
        dt = pk.dtype
        a = (pk.astype(np.float32) * np.float32(m)).astype(dt)
        b = (p.astype(np.float32) * np.float32(1.0 - m)).astype(dt)
        out.append((a.astype(np.float32) + b.astype(np.float32)).astype(dt))
    return out


def dequeue_and_enqueue(queues, ptr, v_k, tag_k, title_k, frame_fea_k, frame_proj_k, K):
    """_dequeue_and_enqueue, modules/modeling.py:244-284 (after the gather).
    ``queues`` maps buffer name -> [D,Kq] array and is updated in place; returns
    the new pointer."""
    v_k = _f_normalize(np.asarray(v_k, np.float32), 1)
    tag_k = _f_normalize(np.asarray(tag_k, np.float32), 1)
    title_k = _f_normalize(np.asarray(title_k, np.float32), 1)
    frame_fea_k = _f_normalize(np.asarray(frame_fea_k, np.float32), 2)
    frame_proj_k = _f_normalize(np.asarray(frame_proj_k, np.float32), 2)
    B = v_k.shape[0]
    F = frame_fea_k.shape[1]
    D = v_k.shape[1]
    if ptr + B > K:
        raise ValueError("enqueue past the end of the queue: ptr %d + batch %d > K %d" % (ptr, B, K))
    queues["queue_v_cross_ng"][:, ptr:ptr + B] = v_k.T
    queues["queue_tag_cross_ng"][:, ptr:ptr + B] = tag_k.T
    queues["queue_title_cross_ng"][:, ptr:ptr + B] = title_k.T
    queues["queue_frame_proj_ng"][:, ptr * F:(ptr + B) * F] = frame_proj_k.reshape(-1, D).T
    queues["queue_frame_cross_ng"][:, ptr * F:(ptr + B) * F] = frame_fea_k.reshape(-1, D).T
    return (ptr + B) % K


def dist_collect_emulated(xs):
    """dist_collect forward, modules/modeling.py:25-36: concat in rank order."""
    return np.concatenate(list(xs), axis=0)


def dist_collect_emulated_backward(grads_per_rank, b):
    """diffdist all_gather backward (PARITY UNPINNED, see header): rank r gets
    the sum over ranks of slice r of their gradients."""
    tot = np.sum(np.stack(list(grads_per_rank), axis=0), axis=0)
    return [tot[r * b:(r + 1) * b] for r in range(len(grads_per_rank))]


# ----------------------------------------------------------------------------
# eval similarity
# ----------------------------------------------------------------------------


def run_on_single_gpu(q_list, v_list, title_list, frame_list, top_frames, logit_scale=4.6052,
                      dtype=np.float32):
    """_run_on_single_gpu, main_task_retrieval.py:321-357: returns three lists
    of per-text-tile row blocks."""
    sim_matrix, sim_title, sim_frame = [], [], []
    for q in q_list:
        row, trow, frow = [], [], []
        for v, ti, fr in zip(v_list, title_list, frame_list):
            row.append(loose_similarity(q, v, logit_scale, dtype))
            trow.append(loose_similarity(q, ti, logit_scale, dtype))
            fl = loose_similarity(q, fr, logit_scale, dtype)
            fl = -np.sort(-fl, axis=2)[:, :, :top_frames]        # torch.topk(...)[0]
            frow.append(fl.mean(axis=2, dtype=dtype))
        sim_matrix.append(np.concatenate(row, axis=-1))
        sim_title.append(np.concatenate(trow, axis=-1))
        sim_frame.append(np.concatenate(frow, axis=1))
    return sim_matrix, sim_title, sim_frame


def eval_scores(T, V, Fr, top_frames, logit_scale=4.6052, tile=256, dtype=np.float32):
    """sim + sim_frame over the whole set (main_task_retrieval.py:490-513,
    task 'retrieval' with --use_frame_fea)."""
    ql = [T[i:i + tile] for i in range(0, T.shape[0], tile)]
    vl = [V[i:i + tile] for i in range(0, V.shape[0], tile)]
    fl = [Fr[i:i + tile] for i in range(0, Fr.shape[0], tile)]
    tl = [np.ones_like(x) for x in vl]      # title block is discarded for task 'retrieval'
    a, _, c = run_on_single_gpu(ql, vl, tl, fl, top_frames, logit_scale, dtype)
    return np.concatenate(a, axis=0) + np.concatenate(c, axis=0)


# ----------------------------------------------------------------------------
# rank metrics
# ----------------------------------------------------------------------------


def compute_metrics(x):
    """metrics.py:12-39 (sort / where form, ties emit duplicates exactly as the
    reference does)."""
    sx = np.sort(-x, axis=1)
    d = np.diag(-x)[:, np.newaxis]
    ind = np.where((sx - d) == 0)[1]
    return {"R1": float(np.sum(ind == 0)) * 100 / len(ind),
            "R5": float(np.sum(ind < 5)) * 100 / len(ind),
            "R10": float(np.sum(ind < 10)) * 100 / len(ind),
            "MR": np.median(ind) + 1,
            "MeanR": np.mean(ind) + 1}


def ranks_square(x):
    """rank_i = #{j : x_ij > x_ii}  (equals metrics.py:20-28 on tie-free input)."""
    return (x > np.diag(x)[:, None]).sum(axis=1).astype(np.int64)


def metrics_from_ranks(ranks):
    """The five scalars of metrics.py:32-37 from an integer rank vector."""
    ind = np.asarray(ranks)
    return {"R1": float(np.sum(ind == 0)) * 100 / len(ind),
            "R5": float(np.sum(ind < 5)) * 100 / len(ind),
            "R10": float(np.sum(ind < 10)) * 100 / len(ind),
            "MR": np.median(ind) + 1,
            "MeanR": np.mean(ind) + 1}


def pad_multi_sentence(sim, cut_off_points):
    """logging_rank's reshape, metrics.py:102-112: [Nt,Nv] -> [V,maxlen,Nv] with
    -inf padding.  ``cut_off_points`` is the already-shifted (-1) list."""
    ends = [c + 1 for c in cut_off_points]
    starts = [0] + ends[:-1]
    maxlen = max(e - s for s, e in zip(starts, ends))
    blocks = []
    for s, e in zip(starts, ends):
        pad = np.full((maxlen - (e - s), sim.shape[1]), -np.inf, dtype=sim.dtype)
        blocks.append(np.concatenate((sim[s:e], pad), axis=0))
    return np.stack(blocks, axis=0)


def tensor_text_to_video_metrics(sim3):
    """metrics.py:49-76 restated with a rank count instead of the double
    argsort (identical on tie-free input).  Percentages are float32, MedianR is
    the lower median (torch.median), MeanR / Std_Rank are float64 numpy."""
    V, L, Nv = sim3.shape
    ranks = []
    for l in range(L):
        blk = sim3[:, l, :]                       # [V, Nv]; row v is a caption of video v
        d = blk[np.arange(V), np.arange(V)]
        r = (blk > d[:, None]).sum(axis=1)
        ranks.append(r)
    ranks = np.stack(ranks, axis=0).reshape(-1)    # flatten(diagonal(second_argsort, 1, 2)) order: [L, V]
    diag = np.stack([sim3[np.arange(V), l, np.arange(V)] for l in range(L)], axis=0).reshape(-1)
    # torch.diagonal(sim, dim1=0, dim2=2) -> [L, V] as well
    valid = ~(np.isinf(diag) | np.isnan(diag))
    vr = ranks[valid].astype(np.int64)
    n = np.float32(len(vr))
    res = {}
    for k in (1, 5, 10):
        res["R%d" % k] = float(np.float32(int(np.sum(vr < k)) * 100) / n)   # int64*100 / len -> float32
    srt = np.sort(vr + 1)
    res["MedianR"] = float(srt[(len(srt) - 1) // 2])
    res["MeanR"] = float(np.mean(vr + 1))
    res["Std_Rank"] = float(np.std(vr + 1))
    res["MR"] = res["MedianR"]
    return res


def tensor_video_to_text_sim(sim3):
    """metrics.py:79-86: NaN -> -inf, max over the sentence axis, transpose."""
    x = np.array(sim3, copy=True)
    x[x != x] = -np.inf
    return x.max(axis=1).T


def ranks_multi_sentence(sim, gt):
    """Integer ranks of the multi-sentence layout (SURVEY.md §8a' 'Ranks'):
    t2v rank_s = #{j : sim_sj > sim_{s,gt(s)}};
    v2t M_jg = max_{s in g} sim_sj, rank_j = #{g : M_jg > M_jj}."""
    Nt, Nv = sim.shape
    gt = np.asarray(gt)
    d = sim[np.arange(Nt), gt]
    t2v = (sim > d[:, None]).sum(axis=1).astype(np.int64)
    M = np.full((Nv, Nv), -np.inf, dtype=sim.dtype)     # M[g, j]
    np.maximum.at(M, gt, sim)
    Mt = M.T                                            # [j, g]
    v2t = (Mt > np.diag(Mt)[:, None]).sum(axis=1).astype(np.int64)
    return t2v, v2t


def logging_rank(sim, multi_sentence, cut_off_points):
    """metrics.py:89-143 without the logging; returns (tv_metrics, vt_metrics)."""
    if multi_sentence:
        sim3 = pad_multi_sentence(sim, cut_off_points)
        return tensor_text_to_video_metrics(sim3), compute_metrics(tensor_video_to_text_sim(sim3))
    return compute_metrics(sim), compute_metrics(sim.T)
