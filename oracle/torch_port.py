"""CPU baseline: a torch restatement of the reference head, op for op as written.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline, --impl reference and --impl reference-gpu legs;
device-agnostic, so the same op sequence on `cuda` tensors is the reference's GPU eager path).
The reference itself is Python and does not exist on the GPU box, so the CPU arm times
this port: the same ATen op sequence as modules/modeling.py (normalize, b x b matmul +
diag, queue.clone(), matmul, cat, /T, cross_entropy, autograd backward), including the
per-call redundancy the new kernels remove.  Pinned against tests/golden in
tests/test_oracle_golden.py::test_torch_port.
"""
import torch
import torch.nn.functional as F


def contrastive_loss(q, k, queue, T):
    """modules/modeling.py:286-313."""
    q = F.normalize(q.squeeze(), dim=1)
    k = F.normalize(k.squeeze(), dim=1)
    bs = q.size(0)
    l_pos = torch.diag(torch.matmul(q, k.T)).reshape([bs, -1])
    l_neg = torch.matmul(q, queue.clone().detach())
    logits = torch.cat([l_pos, l_neg], dim=1)
    logits /= T
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)   # the reference: .cuda()
    return F.cross_entropy(logits, labels)


def frame_self_loss(frame_fea, frame_fea_k, queue, T):
    """modules/modeling.py:315-323."""
    loss = 0.
    n = frame_fea.size(1)
    for i in range(n - 1):
        loss = loss + contrastive_loss(frame_fea[:, i, :], frame_fea_k[:, i + 1, :], queue, T) \
            + contrastive_loss(frame_fea[:, i + 1, :], frame_fea_k[:, i, :], queue, T)
    return loss / (n - 1)


def frame_cross_loss(frame_fea, frame_fea_k, queue_frame, text_fea, text_fea_k, queue_text, T):
    """modules/modeling.py:325-332."""
    loss = 0.
    n = frame_fea.size(1)
    for i in range(n):
        loss = loss + contrastive_loss(text_fea, frame_fea_k[:, i, :], queue_frame, T) \
            + contrastive_loss(frame_fea[:, i, :], text_fea_k, queue_text, T)
    return loss / n


def momentum_update(params, params_k, m):
    """modules/modeling.py:238-242 (rebinding .data like the reference)."""
    for p, pk in zip(params, params_k):
        pk.data = pk.data * m + p.data * (1. - m)


def dequeue_and_enqueue(queues, ptr, keys, K):
    """modules/modeling.py:244-284, W = 1.  queues: dict name -> [D,Kq]; returns new ptr."""
    v = F.normalize(keys["v_fea_k"], dim=1)
    tag = F.normalize(keys["tag_fea_k"], dim=1)
    title = F.normalize(keys["title_fea_k"], dim=1)
    ff = F.normalize(keys["frame_fea_k"], dim=2)
    fp = F.normalize(keys["frame_proj_k"], dim=2)
    B, n = v.size(0), ff.size(1)
    ff = ff.view(-1, ff.size(-1))
    fp = fp.view(-1, fp.size(-1))
    queues["queue_v_cross_ng"][:, ptr:ptr + B] = v.T
    queues["queue_tag_cross_ng"][:, ptr:ptr + B] = tag.T
    queues["queue_title_cross_ng"][:, ptr:ptr + B] = title.T
    queues["queue_frame_proj_ng"][:, ptr * n:(ptr + B) * n] = fp.T
    queues["queue_frame_cross_ng"][:, ptr * n:(ptr + B) * n] = ff.T
    return (ptr + B) % K


def pretrain_step(inp, queues, ptr, K, T=0.07, weights=(0.05, 0.45, 0.45), ema=None, momentum=0.99):
    """One head step as BirdPreTrainedModel.forward runs it (modules/modeling.py:369-424) plus
    backward: EMA, FAM + VTM + FTM, enqueue.  ``inp`` tensors that need grad must have
    requires_grad set.  Returns (loss value, new ptr)."""
    if ema is not None:
        with torch.no_grad():
            momentum_update(ema[0], ema[1], momentum)
    fam = frame_self_loss(inp["frame_pred"], inp["frame_proj_k"], queues["queue_frame_proj_ng"], T)
    vtm = contrastive_loss(inp["v_fea"], inp["title_fea_k"], queues["queue_title_cross_ng"], T) \
        + contrastive_loss(inp["title_fea"], inp["v_fea_k"], queues["queue_v_cross_ng"], T)
    ftm = frame_cross_loss(inp["frame_fea"], inp["frame_fea_k"], queues["queue_frame_cross_ng"], inp["title_fea"],
                           inp["title_fea_k"], queues["queue_title_cross_ng"], T)
    with torch.no_grad():
        ptr = dequeue_and_enqueue(queues, ptr, inp, K)
    loss = weights[0] * fam + weights[1] * vtm + weights[2] * ftm
    loss.backward()
    return float(loss), ptr


def loose_similarity(seq, vis, scale=100.0):
    """modules/modeling.py:207-229."""
    vis = vis.squeeze()
    vis = vis / vis.norm(dim=-1, keepdim=True)
    seq = seq.squeeze()
    seq = seq / seq.norm(dim=-1, keepdim=True)
    if vis.dim() == 2:
        return scale * torch.matmul(seq, vis.t())
    return (scale * torch.matmul(seq, vis.permute(0, 2, 1))).permute(1, 0, 2)


def cross_en(sim):
    """modules/until_module.py:196-205."""
    return (-torch.diag(F.log_softmax(sim, dim=-1))).mean()


def finetune_step(t, v, fr, w_vtm=0.85, w_ftm=0.15):
    """modules/modeling.py:665-673, 702-709 + backward."""
    n = fr.size(1)
    loss = 0.
    for i in range(n):
        s = loose_similarity(t, fr[:, i, :])
        loss = loss + w_ftm * (cross_en(s) + cross_en(s.T)) / n
    s = loose_similarity(t, v)
    loss = loss + w_vtm * (cross_en(s) + cross_en(s.T))
    loss.backward()
    return float(loss)


def eval_sim_and_rank(T, V, Fr, top_k, tile=256):
    """_run_on_single_gpu + `sim += sim_frame` + compute_metrics both ways
    (main_task_retrieval.py:321-357, 512-513; metrics.py:12-39), square layout."""
    import numpy as np
    rows = []
    with torch.no_grad():
        for i in range(0, T.shape[0], tile):
            row, frow = [], []
            for j in range(0, V.shape[0], tile):
                a = loose_similarity(T[i:i + tile], V[j:j + tile]).cpu().numpy()
                fl = loose_similarity(T[i:i + tile], Fr[j:j + tile])
                fl = torch.mean(torch.topk(fl, k=top_k, dim=2)[0], dim=2).cpu().numpy()
                row.append(a)
                frow.append(fl)
            rows.append(np.concatenate(row, axis=-1) + np.concatenate(frow, axis=1))
    sim = np.concatenate(rows, axis=0)
    out = []
    for x in (sim, sim.T):
        sx = np.sort(-x, axis=1)
        d = np.diag(-x)[:, np.newaxis]
        ind = np.where((sx - d) == 0)[1]
        out.append({"R1": float(np.sum(ind == 0)) * 100 / len(ind), "MR": np.median(ind) + 1,
                    "MeanR": np.mean(ind) + 1})
    return sim, out


# ----------------------------------------------------------------------------- optimizer (N3)
def bert_adam_state(params):
    # modules/optimization.py:124-130
    return [{'step': 0, 'next_m': torch.zeros_like(p), 'next_v': torch.zeros_like(p)} for p in params]


def _clip_(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ on a list of gradient tensors (2-norm), in place."""
    norms = [torch.linalg.vector_norm(g, 2.0) for g in grads]
    total = torch.linalg.vector_norm(torch.stack(norms), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def clip_and_bert_adam_step(params, grads, state, groups, group_of, global_max_norm=None):
    """main_pretrain.py:277 + modules/optimization.py:103-168 as the reference runs them: one
    parameter at a time, every elementwise op a separate pass over the tensor."""
    from oracle.optim_oracle import lr_scheduled
    grads = [g.clone() for g in grads]
    total = None
    if global_max_norm is not None:
        total = _clip_(grads, global_max_norm)
    for i, (p, grad) in enumerate(zip(params, grads)):
        group = groups[group_of[i]]
        st = state[i]
        next_m, next_v = st['next_m'], st['next_v']
        beta1, beta2 = group['b1'], group['b2']
        if group['max_grad_norm'] > 0:
            _clip_([grad], group['max_grad_norm'])
        next_m.mul_(beta1).add_(grad, alpha=1 - beta1)
        next_v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        update = next_m / (next_v.sqrt() + group['e'])
        if group['weight_decay'] > 0.0:
            update += group['weight_decay'] * p
        update_with_lr = lr_scheduled(group, st['step']) * update
        p.add_(-update_with_lr)
        st['step'] += 1
    return total
