"""Eval-time similarity with the reference's ``_run_on_single_gpu`` interface
(main_task_retrieval.py:321-357, duplicate main_pretrain.py:307-338) plus the
whole-set drivers used by eval_epoch (main_task_retrieval.py:490-524).
"""
import numpy as np
import torch

from . import metrics as M
from . import ops, parallel


def _scale_of(model):
    return model._logit_scale() if hasattr(model, "_logit_scale") else 100.0


def _run_on_single_gpu(model, batch_query_output_list, batch_visual_output_list, batch_title_output_list,
                       batch_frame_output_list):
    """Same contract as the reference: three lists (sim, title sim, frame sim) of numpy
    [bt, Nv] row blocks, one per text tile.  The gallery tiles are concatenated once and
    each text tile makes ONE device call and ONE device->host copy per output (the
    reference does 3 blocking copies per tile pair)."""
    scale = _scale_of(model)
    prec = getattr(model, "head_precision", None)
    top_k = model.top_frames
    squeeze = lambda t: t.reshape(-1, t.shape[-1]) if t.dim() != 2 else t
    video = torch.cat([squeeze(v.squeeze()) for v in batch_visual_output_list], dim=0)
    frames = torch.cat([f.reshape(-1, f.shape[-2], f.shape[-1]) for f in batch_frame_output_list], dim=0)
    titles = [squeeze(t.squeeze()) for t in batch_title_output_list]
    title = torch.cat(titles, dim=0) if titles else None
    sim_matrix, sim_matrix_title, sim_matrix_frame = [], [], []
    for query_output in batch_query_output_list:
        q = squeeze(query_output.squeeze())
        sim, fsim = ops.sim_topk(q, video, frames, scale, top_k, prec)
        # title block: an all-zero title gives 0/0 = NaN exactly like the reference
        tsim = ops.loose_similarity_raw(ops._f32c(q, "q"), ops._f32c(title, "title"), scale, ops.PREC_FP32) \
            if title is not None else None
        sim_matrix.append(sim.cpu().numpy())
        sim_matrix_frame.append(fsim.cpu().numpy())
        sim_matrix_title.append(tsim.cpu().numpy() if tsim is not None else None)
    return sim_matrix, sim_matrix_title, sim_matrix_frame


def similarity_matrix(model, text, video, frames, use_frame_fea=True, text_tile=4096):
    """sim (+ sim_frame when --use_frame_fea) for the whole set, kept on the device
    (main_task_retrieval.py:490-513 without the host round trips)."""
    scale = _scale_of(model)
    prec = getattr(model, "head_precision", None)
    out = []
    for i in range(0, text.shape[0], text_tile):
        sim, fsim = ops.sim_topk(text[i:i + text_tile], video, frames if use_frame_fea else None, scale,
                                 model.top_frames, prec)
        out.append(sim + fsim if fsim is not None else sim)
    return torch.cat(out, dim=0)


def eval_metrics(model, text, video, frames, multi_sentence_=False, cut_off_points_=None, use_frame_fea=True):
    """Similarity + ranking of one eval set on one GPU: returns (tv_metrics, vt_metrics)."""
    sim = similarity_matrix(model, text, video, frames, use_frame_fea)
    if multi_sentence_:
        t2v, v2t = M.multi_sentence_ranks(sim, cut_off_points_)
        return M.t2v_metrics_from_ranks(t2v), M.metrics_from_ranks(v2t)
    t2v, v2t = ops.rank_count(sim)
    return M.metrics_from_ranks(t2v.cpu().numpy()), M.metrics_from_ranks(v2t.cpu().numpy())
