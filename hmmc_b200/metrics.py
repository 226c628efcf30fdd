"""Rank metrics with the reference's ``metrics.py`` interface.

The device produces integer rank vectors (hmmc_rank_count); the five scalars are then
formed on the host with the reference's exact expressions, including its quirks:
``MR = np.median(rank) + 1`` and python-float percentages in compute_metrics
(metrics.py:32-37); float32 percentages, the *lower* median and the extra ``Std_Rank`` /
``MedianR`` keys in tensor_text_to_video_metrics (metrics.py:71-75).
Parity with the reference is defined on tie-free matrices (its sort/where form emits
duplicate entries under ties).
"""
import logging

import numpy as np
import torch

from . import ops

logger = logging.getLogger(__name__)


def _to_device(x):
    t = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.cuda()
    return t


def metrics_from_ranks(ind):
    """metrics.py:31-38 given the rank vector ``ind`` (0 = retrieved first)."""
    ind = np.asarray(ind)
    metrics = {}
    metrics['R1'] = float(np.sum(ind == 0)) * 100 / len(ind)
    metrics['R5'] = float(np.sum(ind < 5)) * 100 / len(ind)
    metrics['R10'] = float(np.sum(ind < 10)) * 100 / len(ind)
    metrics['MR'] = np.median(ind) + 1
    metrics["MeanR"] = np.mean(ind) + 1
    return metrics


def t2v_metrics_from_ranks(valid_ranks, top_k=(1, 5, 10)):
    """metrics.py:71-76 given the valid caption ranks."""
    vr = torch.as_tensor(np.asarray(valid_ranks)).to(torch.int64).cpu()
    results = {f"R{k}": float(torch.sum(vr < k) * 100 / len(vr)) for k in top_k}
    results["MedianR"] = float(torch.median(vr + 1))
    results["MeanR"] = float(np.mean(vr.numpy() + 1))
    results["Std_Rank"] = float(np.std(vr.numpy() + 1))
    results['MR'] = results["MedianR"]
    return results


def compute_metrics(x, log=False):
    """metrics.py:12-39.  ``x`` is a square [N,N] similarity matrix (numpy or tensor);
    row i's ground truth is column i."""
    sim = _to_device(x)
    if sim.dim() != 2 or sim.shape[0] != sim.shape[1]:
        # the reference fails in `sx - d` for non-square input
        raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,1) "
                         % (sim.shape[0], sim.shape[1], sim.shape[0]))
    t2v, _ = ops.rank_count(sim, want_v2t=False)
    ind = t2v.cpu().numpy()
    if log:
        logger.info("correct:{}".format([int(i) for i in np.nonzero(ind == 0)[0]]))
    return metrics_from_ranks(ind)


def print_computed_metrics(metrics):
    r1 = metrics['R1']
    r5 = metrics['R5']
    r10 = metrics['R10']
    mr = metrics['MR']
    print('R@1: {:.4f} - R@5: {:.4f} - R@10: {:.4f} - Median R: {}'.format(r1, r5, r10, mr))


def tensor_text_to_video_metrics(sim_tensor, top_k=[1, 5, 10]):
    """metrics.py:49-76.  sim_tensor is the padded cube [V, maxlen, Nv] (-inf rows = padding);
    caption (v, l) belongs to video v."""
    cube = _to_device(sim_tensor)
    V, L, Nv = cube.shape
    flat = cube.reshape(V * L, Nv)
    gt = torch.arange(V, device=flat.device, dtype=torch.int32).repeat_interleave(L)
    gs = torch.arange(0, V * L + 1, L, device=flat.device, dtype=torch.int32)
    t2v, _ = ops.rank_count(flat, gt, gs, want_v2t=False)
    diag = flat[torch.arange(V * L, device=flat.device), gt.long()]
    mask = ~(torch.isinf(diag) | torch.isnan(diag))
    # the reference flattens [maxlen, V]; the scalars do not depend on the order
    valid = t2v.reshape(V, L).t().reshape(-1)[mask.reshape(V, L).t().reshape(-1)]
    return t2v_metrics_from_ranks(valid.cpu().numpy(), top_k)


def tensor_video_to_text_sim(sim_tensor):
    """metrics.py:79-86: NaN -> -inf, max over the caption axis, transpose -> [Nv, V]."""
    cube = _to_device(sim_tensor)
    V, L, Nv = cube.shape
    gs = torch.arange(0, V * L + 1, L, device=cube.device, dtype=torch.int32)
    return ops.group_max(cube.reshape(V * L, Nv), gs)


def multi_sentence_ranks(sim_matrix, cut_off_points_):
    """Integer ranks of the multi-sentence layout straight from the [Nt, Nv] matrix (no
    padded cube): returns (t2v ranks [Nt], v2t ranks [Nv]) as numpy int arrays."""
    sim = _to_device(sim_matrix)
    ends = np.asarray([c + 1 for c in cut_off_points_], dtype=np.int64)
    starts = np.concatenate([[0], ends[:-1]])
    gs = torch.as_tensor(np.concatenate([starts, ends[-1:]]).astype(np.int32))
    gt = torch.as_tensor(np.repeat(np.arange(len(ends)), ends - starts).astype(np.int32))
    t2v, v2t = ops.rank_count(sim, gt, gs)
    return t2v.cpu().numpy(), v2t.cpu().numpy()


def logging_rank_from_ranks(t2v, v2t, multi_sentence_, logger, shape=None):
    """The logging half of logging_rank for ranks that were counted without a matrix (fused / sharded eval):
    same scalars, same log lines; returns tv_metrics."""
    t2v, v2t = np.asarray(t2v), np.asarray(v2t)
    tv_metrics = t2v_metrics_from_ranks(t2v) if multi_sentence_ else metrics_from_ranks(t2v)
    vt_metrics = metrics_from_ranks(v2t)
    if shape is not None and not multi_sentence_:
        logger.info('\t Length-T: {}, Length-V:{}'.format(shape[0], shape[1]))
    _log_both(tv_metrics, vt_metrics, logger)
    return tv_metrics


def _log_both(tv_metrics, vt_metrics, logger):
    logger.info("Text-to-Video:")
    logger.info('\t>>>  R@1: {:.1f} - R@5: {:.1f} - R@10: {:.1f} - Median R: {:.1f} - Mean R: {:.1f}'.
                format(tv_metrics['R1'], tv_metrics['R5'], tv_metrics['R10'], tv_metrics['MR'], tv_metrics['MeanR']))
    logger.info("Video-to-Text:")
    logger.info(
        '\t>>>  V2T$R@1: {:.1f} - V2T$R@5: {:.1f} - V2T$R@10: {:.1f} - V2T$Median R: {:.1f} - V2T$Mean R: {:.1f}'.format(
            vt_metrics['R1'], vt_metrics['R5'], vt_metrics['R10'], vt_metrics['MR'], vt_metrics['MeanR']))


def logging_rank(sim_matrix, multi_sentence_, cut_off_points_, logger):
    """run similarity in one single gpu (metrics.py:89-143); returns tv_metrics."""
    if multi_sentence_:
        logger.info("before reshape, sim matrix size: {} x {}".format(sim_matrix.shape[0], sim_matrix.shape[1]))
        t2v, v2t = multi_sentence_ranks(sim_matrix, cut_off_points_)
        tv_metrics = t2v_metrics_from_ranks(t2v)
        vt_metrics = metrics_from_ranks(v2t)
    else:
        logger.info("sim matrix size: {}, {}".format(sim_matrix.shape[0], sim_matrix.shape[1]))
        sim = _to_device(sim_matrix)
        if sim.shape[0] != sim.shape[1]:
            raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,1) "
                             % (sim.shape[0], sim.shape[1], sim.shape[0]))
        t2v, v2t = ops.rank_count(sim)
        tv_metrics = metrics_from_ranks(t2v.cpu().numpy())
        vt_metrics = metrics_from_ranks(v2t.cpu().numpy())
        logger.info('\t Length-T: {}, Length-V:{}'.format(len(sim_matrix), len(sim_matrix[0])))

    logger.info("Text-to-Video:")
    logger.info('\t>>>  R@1: {:.1f} - R@5: {:.1f} - R@10: {:.1f} - Median R: {:.1f} - Mean R: {:.1f}'.
                format(tv_metrics['R1'], tv_metrics['R5'], tv_metrics['R10'], tv_metrics['MR'], tv_metrics['MeanR']))
    logger.info("Video-to-Text:")
    logger.info(
        '\t>>>  V2T$R@1: {:.1f} - V2T$R@5: {:.1f} - V2T$R@10: {:.1f} - V2T$Median R: {:.1f} - V2T$Mean R: {:.1f}'.format(
            vt_metrics['R1'], vt_metrics['R5'], vt_metrics['R10'], vt_metrics['MR'], vt_metrics['MeanR']))
    return tv_metrics
