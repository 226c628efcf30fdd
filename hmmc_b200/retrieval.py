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
        sim, _ = ops.sim_topk(text[i:i + text_tile], video, frames if use_frame_fea else None, scale,
                              model.top_frames, prec, combine=True)
        out.append(sim)
    return out[0] if len(out) == 1 else torch.cat(out, dim=0)


def eval_metrics(model, text, video, frames, multi_sentence_=False, cut_off_points_=None, use_frame_fea=True):
    """Similarity + ranking of one eval set on one GPU: returns (tv_metrics, vt_metrics)."""
    sim = similarity_matrix(model, text, video, frames, use_frame_fea)
    if multi_sentence_:
        t2v, v2t = M.multi_sentence_ranks(sim, cut_off_points_)
        return M.t2v_metrics_from_ranks(t2v), M.metrics_from_ranks(v2t)
    t2v, v2t = ops.rank_count(sim)
    return M.metrics_from_ranks(t2v.cpu().numpy()), M.metrics_from_ranks(v2t.cpu().numpy())


# ---------------------------------------------------------------------------------------------
# Large-gallery eval (BASELINE config 5): fused similarity + top-k + rank counting, gallery sharded
# over the ranks of the process group.  The [Nt, Nv] matrix never exists.
# ---------------------------------------------------------------------------------------------

def pack_caption_groups(per_video, tile=128):
    """Lay the captions out group-aligned: the captions of one video are contiguous and never
    straddle a ``tile``-row boundary.  ``per_video[j]`` = number of captions of video j (captions
    are ordered by video, as the reference's multi-sentence loaders produce them).
    Returns (src_row [Nt_pad] int32 with -1 padding, grp [Nt_pad] int32, group_start [Nv] int32)."""
    per = np.asarray(per_video, dtype=np.int64)
    if per.size and per.max() > tile:
        raise ValueError("the fused eval path needs <= %d captions per video (got %d)" % (tile, per.max()))
    Nv = per.size
    starts = np.empty(Nv, dtype=np.int64)
    if Nv and (per == per[0]).all():
        g = int(per[0])
        per_tile = tile // g
        j = np.arange(Nv, dtype=np.int64)
        starts = (j // per_tile) * tile + (j % per_tile) * g
    else:
        pos = 0
        for j in range(Nv):
            c = int(per[j])
            if (pos % tile) + c > tile:
                pos = (pos // tile + 1) * tile
            starts[j] = pos
            pos += c
    total = int(starts[-1] + per[-1]) if Nv else 0
    Nt_pad = max(tile, (total + tile - 1) // tile * tile)
    src_row = np.full(Nt_pad, -1, dtype=np.int32)
    grp = np.full(Nt_pad, -1, dtype=np.int32)
    first = np.concatenate([[0], np.cumsum(per)[:-1]]) if Nv else np.zeros(0, np.int64)
    idx = np.repeat(starts - first, per) + np.arange(int(per.sum()), dtype=np.int64)   # packed position of caption s
    src_row[idx] = np.arange(int(per.sum()), dtype=np.int32)
    grp[idx] = np.repeat(np.arange(Nv, dtype=np.int32), per)
    return src_row, grp, starts.astype(np.int32)


def fused_eval_ranks(text, video_local, frames_local, per_video, scale=100.0, top_k=2, precision="bf16",
                     video_range=None):
    """Integer ranks of a (possibly sharded) retrieval eval without materialising the matrix.

    text [Nt, D] (replicated on every rank, captions ordered by video), video_local [Nv_loc, D] /
    frames_local [Nv_loc, F, D] = this rank's gallery shard = videos ``video_range`` (default:
    parallel.shard_range(Nv)), per_video [Nv] caption counts of ALL videos.
    Returns (t2v ranks [Nt] int32 tensor, v2t ranks [Nv] int32 tensor), identical on every rank.
    """
    from . import _lib
    lib = _lib.load()
    prec = ops.resolve_precision(precision)
    if prec == ops.PREC_FP32:
        raise ops.HmmcError("fused eval runs on tensor cores: precision must be bf16 or bf16x3")
    per = np.asarray(per_video, dtype=np.int64)
    Nv = per.size
    W, rank = parallel.world()
    lo, hi = video_range if video_range is not None else parallel.shard_range(Nv, W, rank)
    text = ops._f32c(text, "text")
    video_local = ops._f32c(video_local, "video")
    frames_local = ops._f32c(frames_local, "frames")
    Nt, D = text.shape
    F = frames_local.shape[1]
    Nv_loc = hi - lo
    assert video_local.shape[0] == Nv_loc and int(per.sum()) == Nt
    if not lib.hmmc_eval_fused_supported(F, D, int(top_k)):
        raise ops.HmmcError("fused eval supports F=12, D %% 64 == 0, top_k <= 4 (got F=%d D=%d k=%d)" % (F, D, top_k))
    dev = text.device
    planes = 2 if prec == ops.PREC_BF16X3 else 1
    st = ops._stream
    plan = _eval_plan(per, lo, hi, dev)
    Nt_pad, d_src, d_grp = plan["Nt_pad"], plan["d_src"], plan["d_grp"]
    tp = torch.empty(Nt_pad, planes * D, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.hmmc_eval_pack_text(ops._p(text), ops._p(d_src), Nt_pad, D, prec, ops._p(tp), st()), "eval_pack_text")
    n_blk = (Nv_loc + 15) // 16
    gp = torch.empty(max(n_blk, 1) * 16 * (1 + F), planes * D, dtype=torch.bfloat16, device=dev)
    gt_score = torch.zeros(Nt_pad, dtype=torch.float32, device=dev)
    t2v = torch.zeros(Nt_pad, dtype=torch.int32, device=dev)
    v2t_loc = torch.zeros(max(Nv_loc, 1), dtype=torch.int32, device=dev)
    if Nv_loc > 0:
        _lib.check(lib.hmmc_eval_pack_gallery(ops._p(video_local), ops._p(frames_local), Nv_loc, F, D, prec, ops._p(gp),
                                              st()), "eval_pack_gallery")
        _lib.check(lib.hmmc_eval_gt_scores(ops._p(tp), ops._p(gp), Nt_pad, Nv_loc, D, prec, float(scale), int(top_k),
                                           int(lo), ops._p(d_grp), ops._p(plan["d_pairs"]), plan["n_pairs"],
                                           ops._p(gt_score), st()), "eval_gt_scores")
    parallel.all_reduce_sum_(gt_score)          # each caption's score is produced by exactly one shard
    if Nv_loc > 0:
        theta = torch.empty(Nv_loc, dtype=torch.float32, device=dev)
        _lib.check(lib.hmmc_eval_theta(ops._p(gt_score), ops._p(plan["d_gs"]), ops._p(plan["d_gc"]), Nv_loc,
                                       ops._p(theta), st()),
                   "eval_theta")
        _lib.check(lib.hmmc_eval_fused_rank(ops._p(tp), ops._p(gp), Nt_pad, Nv_loc, D, prec, float(scale), int(top_k),
                                            int(lo), ops._p(d_grp), ops._p(gt_score), ops._p(theta), ops._p(t2v),
                                            ops._p(v2t_loc), st()), "eval_fused_rank")
    parallel.all_reduce_sum_(t2v)
    counts = [parallel.shard_range(Nv, W, r) for r in range(W)] if video_range is None else None
    if W > 1 and counts is not None:
        v2t = parallel.all_gather_varlen(v2t_loc[:Nv_loc], [b - a for a, b in counts])
    else:
        v2t = v2t_loc[:Nv_loc]
    t2v_out = torch.empty(Nt, dtype=torch.int32, device=dev)
    t2v_out[plan["d_order"]] = t2v[plan["d_valid"]]
    return t2v_out, v2t


_eval_plans = {}


def _eval_plan(per, lo, hi, dev):
    """Index arrays of one (caption layout, gallery shard): built once per eval set, like the
    reference builds cut_off_points once per dataset."""
    key = (hash(per.tobytes()), per.size, lo, hi, str(dev))
    plan = _eval_plans.get(key)
    if plan is not None:
        return plan
    src_row, grp, gstart = pack_caption_groups(per)
    plan = {"Nt_pad": int(src_row.size),
            "d_src": torch.from_numpy(src_row).to(dev), "d_grp": torch.from_numpy(grp).to(dev),
            "d_valid": torch.from_numpy(np.nonzero(src_row >= 0)[0]).to(dev),
            "d_order": torch.from_numpy(src_row[src_row >= 0].astype(np.int64)).to(dev),
            "n_pairs": 0, "d_pairs": None, "d_gs": None, "d_gc": None}
    if hi > lo:
        # diagonal tiles: (caption tile m, gallery tile n) pairs that contain a ground-truth pair of this shard
        m_of = (gstart[lo:hi].astype(np.int64)) // 128
        n_of = (np.arange(lo, hi, dtype=np.int64) - lo) // 16
        pairs = np.unique(np.stack([m_of, n_of], axis=1), axis=0).astype(np.int32)
        plan["d_pairs"] = torch.from_numpy(np.ascontiguousarray(pairs)).to(dev)
        plan["n_pairs"] = int(pairs.shape[0])
        plan["d_gs"] = torch.from_numpy(np.ascontiguousarray(gstart[lo:hi])).to(dev)
        plan["d_gc"] = torch.from_numpy(per[lo:hi].astype(np.int32)).to(dev)
    if len(_eval_plans) > 16:
        _eval_plans.clear()
    _eval_plans[key] = plan
    return plan


# ---------------------------------------------------------------------------------------------
# eval_epoch (main_task_retrieval.py:358-524)
# ---------------------------------------------------------------------------------------------
class _NullLogger:
    def info(self, *a, **k):
        pass


def cache_eval_features(args, model, test_dataloader, device, multi_sentence_=False, cut_off_points_=()):
    """Step 1 of eval_epoch (main_task_retrieval.py:391-440): run the encoders batch by batch and keep
    the features.  Returns ONE tensor per stream on `device` (texts [Nt,D], videos [Nv,D], frames
    [Nv,F,D], titles [Nv,D] or None) instead of the reference's four Python lists of batch tensors.
    In the multi-sentence layout a video is encoded only at its cut-off caption, like the reference."""
    texts, videos, frames, titles = [], [], [], []
    total_video_num = 0
    cut = set(cut_off_points_)
    task = getattr(args, "task", "retrieval")
    for batch in test_dataloader:
        batch = tuple(t.to(device) for t in batch)
        if task == "retrieval_VT":
            query_ids, query_mask, video, video_frame, title_ids, title_mask = batch
        elif task == "retrieval":
            query_ids, query_mask, video, video_frame = batch
        else:
            raise ValueError("wrong task type:{}".format(task))
        texts.append(model.text_encoder(query_ids, query_mask))
        if multi_sentence_:
            b = video.shape[0]
            s_, e_ = total_video_num, total_video_num + b
            filter_inds = [i - s_ for i in range(s_, e_) if i in cut]
            if len(filter_inds) > 0:
                visual_output, frame_output = model.visual_encoder(video[filter_inds, ...], video_frame)
                videos.append(visual_output)
                frames.append(frame_output)
            total_video_num += b
        else:
            visual_output, frame_output = model.visual_encoder(video, video_frame)
            videos.append(visual_output)
            frames.append(frame_output)
            if task == "retrieval_VT":
                titles.append(model.text_encoder(title_ids, title_mask))
    flat = lambda xs: torch.cat([x.reshape(-1, x.shape[-1]) for x in xs], dim=0).float()
    fr = torch.cat([f.reshape(-1, f.shape[-2], f.shape[-1]) for f in frames], dim=0).float()
    return flat(texts), flat(videos), fr, (flat(titles) if titles else None)


# an [Nt, Nv] fp32 matrix above this size is not materialised: the ranks are counted tile by tile instead
MATRIX_BYTES_LIMIT = 2 << 30


def choose_eval_path(Nt, Nv, F, D, top_k, per_video, precision, task="retrieval", sharded=False):
    """Which similarity + ranking path an eval set takes:
      "matrix"  - materialise sim (+ sim_frame) [Nt, Nv] on this GPU, rank it (small sets; any shape);
      "fused"   - fused similarity + top-k + rank counting on this GPU, no matrix;
      "sharded" - the same with the gallery split over the ranks of the process group (every rank must call).
    The fused tiles need 12 frames, top_frames <= 4, D % 64 == 0, <= 128 captions per video, a tensor-core
    precision and the plain retrieval task."""
    from . import _lib
    prec = ops.resolve_precision(precision)
    per = np.asarray(per_video)
    fused_ok = (task == "retrieval" and prec != ops.PREC_FP32 and per.size == Nv and (per.size == 0 or per.max() <= 128)
                and bool(_lib.load().hmmc_eval_fused_supported(int(F), int(D), int(top_k))))
    if sharded and parallel.world()[0] > 1:
        if not fused_ok:
            raise ops.HmmcError("sharded eval needs the fused tiles (12 frames, top_frames <= 4, D %% 64 == 0, <= 128 "
                                "captions per video, bf16 / bf16x3, task retrieval)")
        return "sharded"
    if fused_ok and 4 * int(Nt) * int(Nv) > MATRIX_BYTES_LIMIT:
        return "fused"
    return "matrix"


def eval_epoch(args, model, test_dataloader, device, n_gpu, logger=None):
    """Drop-in for main_task_retrieval.py:358-524: cache the features, build the similarity matrix
    (sim + sim_frame when --use_frame_fea, + weight_title * sim_title for retrieval_VT), log and return
    the text-to-video metrics of `logging_rank`.

    The reference fans text tiles out over `n_gpu` devices of ONE process with threads and peer copies of the
    whole gallery (:447-488).  Here a process owns one GPU (`n_gpu` is accepted for signature compatibility)
    and the path is picked from the size (`choose_eval_path`): small sets materialise the matrix on this GPU;
    sets whose matrix would exceed MATRIX_BYTES_LIMIT count the ranks tile by tile without one; and with
    ``args.eval_sharded`` set - every rank of the process group then has to call eval_epoch, not rank 0 alone
    as main_task_retrieval.py:620-622 does - the gallery is split over the ranks (two small
    all-reduces and one all-gather of integer ranks)."""
    logger = logger or _NullLogger()
    if hasattr(model, 'module'):
        model = model.module
    model = model.to(device)
    model.eval()
    logger.info("args.task:{}".format(getattr(args, "task", "retrieval")))
    multi_sentence_ = False
    cut_off_points_ = []
    ds = getattr(test_dataloader, "dataset", None)
    if ds is not None and getattr(ds, 'multi_sentence_per_video', False):
        multi_sentence_ = True
        cut_off_points_ = [itm - 1 for itm in ds.cut_off_points]
    logger.info("multi_sentence_:{}".format(multi_sentence_))
    with torch.no_grad():
        text, video, frames, title = cache_eval_features(args, model, test_dataloader, device, multi_sentence_,
                                                         cut_off_points_)
        use_frame = bool(getattr(args, "use_frame_fea", True))
        task = getattr(args, "task", "retrieval")
        Nt, Nv = text.shape[0], video.shape[0]
        if multi_sentence_:
            ends = np.asarray([c + 1 for c in cut_off_points_], dtype=np.int64)
            per = np.diff(np.concatenate([[0], ends]))
        else:
            per = np.ones(Nv, dtype=np.int64) if Nt == Nv else np.zeros(0, dtype=np.int64)
        path = choose_eval_path(Nt, Nv, frames.shape[1], text.shape[1], model.top_frames, per,
                                getattr(model, "head_precision", None), task if use_frame else "no-frames",
                                sharded=bool(getattr(args, "eval_sharded", False)))
        logger.info("eval path: {}".format(path))
        if path != "matrix":
            W, rank = parallel.world() if path == "sharded" else (1, 0)
            lo, hi = parallel.shard_range(Nv, W, rank)
            prec = getattr(model, "head_precision", None) or ops.DEFAULT_PRECISION
            t2v, v2t = fused_eval_ranks(text, video[lo:hi], frames[lo:hi], per, _scale_of(model), model.top_frames, prec,
                                        video_range=None if path == "sharded" else (0, Nv))
            logger.info("sim matrix size:  {}".format((Nt, Nv)))
            return M.logging_rank_from_ranks(t2v.cpu().numpy(), v2t.cpu().numpy(), multi_sentence_, logger, (Nt, Nv))
        sim = similarity_matrix(model, text, video, frames, use_frame_fea=use_frame)
        if getattr(args, "task", "retrieval") == "retrieval_VT":
            tsim = ops.loose_similarity_raw(ops._f32c(text, "text"), ops._f32c(title, "title"), _scale_of(model),
                                            ops.PREC_FP32)
            sim = sim + model.weight_title * tsim
    logger.info("sim matrix size:  {}".format(tuple(sim.shape)))
    return M.logging_rank(sim, multi_sentence_, cut_off_points_, logger)
