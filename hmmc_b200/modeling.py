"""Host side of the contrastive head, mirroring the reference's ``modules/modeling.py``.

Names, positional arguments, return types and buffer names follow the reference so that
the methods here can be mixed into (or swapped for) the reference's classes:

    dist_collect                      modules/modeling.py:25-36
    CrossEn                           modules/until_module.py:196-205
    ContrastiveHeadMixin
        .loose_similarity             modules/modeling.py:207-229
        .copy_params / ._momentum_update          :231-242
        ._dequeue_and_enqueue         :244-284
        .contrastive_loss             :286-313
        .frame_self_loss / .frame_cross_loss      :315-332
        .frame_loss                   :665-680
    BirdPreTrainedModel.forward       :334-436   (head part; encoders are injected)
    BirdModel.forward                 :682-722

All arithmetic runs in libhmmc_head.so (see ops.py); there is no PyTorch fallback.
"""
import logging
import os
from types import SimpleNamespace

import torch
from torch import nn

from . import ops, parallel

logger = logging.getLogger(__name__)

# modules/cross-base/cross_config.json (head-relevant keys)
DEFAULT_CROSS_CONFIG = dict(temporal_hidden_size=512, weight_FAM=0.05, weight_VTM=0.45, weight_FTM=0.45,
                            weight_MLM=0.05, weight_VTM_finetune=0.85, weight_FTM_finetune=0.15)


# SMs left to the key all-gather while it overlaps the loss (tuned on 2 and 8 B200)
GATHER_RESERVED_SMS = int(os.environ.get("HMMC_GATHER_RESERVED_SMS", "0"))


# run the enqueue next to the tail of the loss on a side stream (0: strictly after it)
ENQUEUE_OVERLAP = os.environ.get("HMMC_ENQUEUE_OVERLAP", "1") != "0"
_side_streams = {}


# split the head around the momentum update / key encoders (head_loss_begin / head_loss_end); 0: one fused call
LOSS_OVERLAP = os.environ.get("HMMC_LOSS_OVERLAP", "1") != "0"


def use_split_schedule():
    """The two-half schedule pays on a single rank (0.408 -> 0.399 ms per step).  With several ranks the key
    all-gather has to hide behind the loss GEMMs, which the split moves beside the EMA: measured 0.589 ms
    against 0.537 ms at 8 ranks, so multi-rank runs keep the one-call head."""
    return LOSS_OVERLAP and parallel.world()[0] == 1
# SMs kept out of the loss GEMM grids while they run beside the EMA (tools/overlap_probe.py)
LOSS_GEMM_RESERVED = int(os.environ.get("HMMC_LOSS_GEMM_RESERVED", "120"))


def _side_stream(device, name="enqueue", priority=0):
    key = (device.index if device.index is not None else torch.cuda.current_device(), name)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device, priority=priority)
    return _side_streams[key]


def default_cross_config(**over):
    cfg = dict(DEFAULT_CROSS_CONFIG)
    cfg.update(over)
    return SimpleNamespace(**cfg)


def dist_collect(x):
    """collect all tensor from all GPUs (modules/modeling.py:25-36): differentiable
    all-gather, concatenated on dim 0 in rank order; backward = SUM reduce-scatter."""
    return parallel.all_gather_cat(x)


class CrossEn(nn.Module):
    """modules/until_module.py:196-205."""

    def forward(self, sim_matrix):
        return ops.cross_en(sim_matrix)


class ContrastiveHeadMixin:
    """The reference's head methods, backed by the sm_100a kernels.

    Attributes read (same as the reference): contrast_temperature, contrast_momentum,
    contrast_num_negative, top_frames, text_encoder.logit_scale, model_pairs, the six
    queue buffers.  ``head_precision`` ("fp32" | "bf16" | "bf16x3") selects how the
    contractions are carried out (default: env HMMC_PRECISION or "bf16x3").
    """

    head_precision = None

    # ---------------------------------------------------------------- fine-tune pieces
    def _logit_scale(self):
        # modules/modeling.py:216-217: exp() then clamp(max=100).  logit_scale is a plain
        # tensor attribute copied from CLIP's state dict, cached as a float.
        ls = self.text_encoder.logit_scale
        key = (id(ls), getattr(ls, "_version", 0))
        cache = getattr(self, "_hmmc_scale_cache", None)
        if cache is None or cache[0] != key:
            val = float(torch.clamp(torch.as_tensor(ls, dtype=torch.float32).detach().exp(), max=100).item())
            cache = (key, val)
            self._hmmc_scale_cache = cache
        return cache[1]

    def loose_similarity(self, sequence_output, visual_output):
        sequence_output, visual_output = sequence_output.contiguous(), visual_output.contiguous()
        visual_output = visual_output.squeeze()
        sequence_output = sequence_output.squeeze()
        if sequence_output.dim() != 2 or visual_output.dim() not in (2, 3):
            raise ValueError("loose_similarity: expected [Bt,D] and [Bv,D] or [Bv,F,D] after squeeze(), got %s and %s"
                             % (tuple(sequence_output.shape), tuple(visual_output.shape)))
        return ops.loose_similarity(sequence_output, visual_output, self._logit_scale(), self.head_precision)

    def frame_loss(self, query_output, frame_output):
        # sum_i (CE(S_i) + CE(S_i^T)) / F over the F text x frame_i matrices (:665-673)
        return ops.sym_ce(query_output, None, frame_output, self._logit_scale(), 0.0, 1.0, self.head_precision)

    def finetune_head_loss(self, query_output, visual_output, frame_output):
        """weight_FTM_finetune * frame_loss + weight_VTM_finetune * (CE(S)+CE(S^T)) in one
        fused pass (modules/modeling.py:702-709)."""
        use_frames = bool(getattr(self.task_config, "use_frame_fea", True))
        return ops.sym_ce(query_output, visual_output, frame_output if use_frames else None, self._logit_scale(),
                          self.weight_VTM_finetune, self.weight_FTM_finetune if use_frames else 0.0,
                          self.head_precision)

    # ---------------------------------------------------------------- pre-train pieces
    def contrastive_loss(self, q, k, queue):
        q = q.squeeze()
        k = k.squeeze()
        if q.dim() != 2 or k.shape != q.shape:
            raise ValueError("contrastive_loss: q and k must be [b, D] after squeeze(), got %s and %s"
                             % (tuple(q.shape), tuple(k.shape)))
        return ops.infonce(q, k, queue, ops.POS_PAIR, q.shape[0], 1, 1, self.contrast_temperature, 1.0,
                           self.head_precision)

    def frame_self_loss(self, frame_fea, frame_fea_k, queue_frame_ng):
        b, F = frame_fea.shape[0], frame_fea.shape[1]
        if F < 2:
            raise ZeroDivisionError("float division by zero")      # loss / (F - 1) in the reference
        return ops.infonce(frame_fea, frame_fea_k, queue_frame_ng, ops.POS_FRAME_NEIGHBOUR, b, F, F,
                           self.contrast_temperature, 1.0 / (F - 1), self.head_precision)

    def frame_cross_loss(self, frame_fea, frame_fea_k, queue_frame_ng, text_fea, text_fea_k, queue_text_ng):
        b, F = frame_fea.shape[0], frame_fea.shape[1]
        T = self.contrast_temperature
        a = ops.infonce(text_fea, frame_fea_k, queue_frame_ng, ops.POS_ONE_TO_FRAMES, b, 1, F, T, 1.0 / F,
                        self.head_precision)
        c = ops.infonce(frame_fea, text_fea_k, queue_text_ng, ops.POS_FRAMES_TO_ONE, b, F, 1, T, 1.0 / F,
                        self.head_precision)
        return a + c

    @torch.no_grad()
    def copy_params(self):
        for model_pair in self.model_pairs:
            for param, param_k in zip(model_pair[0].parameters(), model_pair[1].parameters()):
                param_k.data.copy_(param.data)  # initialize
                param_k.requires_grad = False  # not update by gradient
        self._hmmc_ema = None

    def _ema_pairs(self):
        return [(p, pk) for pair in self.model_pairs for p, pk in zip(pair[0].parameters(), pair[1].parameters())]

    @torch.no_grad()
    def _momentum_update(self):
        # p_k <- p_k*m + p*(1-m) for every parameter pair, one launch (modules/modeling.py:238-242).
        # The pointer table is built once; a few pointers are probed per call to catch re-allocation
        # (.to(), .half(), load_state_dict with assign); set self._hmmc_ema = None to force a rebuild.
        tab = getattr(self, "_hmmc_ema", None)
        if tab is not None and not tab.still_valid():
            tab = None
        if tab is None:
            tab = ops.EmaTable(self._ema_pairs())
            self._hmmc_ema = tab
        tab.run(self.contrast_momentum)

    def _queue_buffers(self):
        return [self.queue_v_cross_ng, self.queue_tag_cross_ng, self.queue_title_cross_ng,
                self.queue_frame_cross_ng, self.queue_frame_proj_ng]

    @torch.no_grad()
    def _gather_keys_async(self, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        """Start the exchange of _dequeue_and_enqueue (modules/modeling.py:249-258): ONE packed
        all-gather instead of five.  Returns a handle for _enqueue_gathered."""
        b = v_fea_k.shape[0]
        D = v_fea_k.shape[-1]
        frame_fea_k = frame_fea_k.reshape(b, -1, D)
        frame_proj_k = frame_proj_k.reshape(b, -1, D)
        F = frame_fea_k.shape[1]
        W, _ = parallel.world()
        keys = [v_fea_k.reshape(b, D), tag_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k, frame_proj_k]
        if W == 1:
            return (W, b, F, D, [ops._f32c(t, "key") for t in keys], None, lambda: None)
        send = ops.pack_rows(keys)
        gathered, wait = parallel.all_gather_rows_async(send)
        # the collective's CTAs share the SMs with the loss kernels issued until _enqueue_gathered:
        # keep some SMs out of the persistent GEMM grids meanwhile
        ops.set_reserved_sms(GATHER_RESERVED_SMS)
        return (W, b, F, D, None, gathered, wait)

    @torch.no_grad()
    def _enqueue_gathered(self, handle):
        W, b, F, D, direct, gathered, wait = handle
        K = self.contrast_num_negative
        wait()
        if W > 1:
            ops.set_reserved_sms(0)
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: the pointer is read and advanced on the device at every replay
            if K % (W * b) != 0:
                raise ValueError("graph capture of the enqueue needs K %% (world*batch) == 0 (K=%d, B=%d)" % (K, W * b))
            ops.enqueue(gathered, W, b, F, D, self._queue_buffers(), self.queue_ptr, -1, K,
                        ops.resolve_precision(self.head_precision), direct=direct)
            self._hmmc_ptr = None          # host copy is stale after replays: re-read on the next eager call
            return
        ver = self.queue_ptr._version
        if getattr(self, "_hmmc_ptr", None) is None or self._hmmc_ptr[1] != ver:
            self._hmmc_ptr = (int(self.queue_ptr), ver)        # one sync, then tracked on the host
        ptr = self._hmmc_ptr[0]
        ops.enqueue(gathered, W, b, F, D, self._queue_buffers(), self.queue_ptr, ptr, K,
                    ops.resolve_precision(self.head_precision), direct=direct)
        self._hmmc_ptr = ((ptr + W * b) % K, self.queue_ptr._version)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        self._enqueue_gathered(self._gather_keys_async(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k))


def patch_reference_classes(*classes):
    """Install the B200 head on the reference's own model classes: every method
    of ContrastiveHeadMixin replaces the class's method of the same name; classes that own the
    negative queues also get the fused ``head_loss``."""
    for cls in classes:
        for name, fn in vars(ContrastiveHeadMixin).items():
            if callable(fn) and not name.startswith("__"):
                setattr(cls, name, fn)
        if hasattr(cls, "_dequeue_and_enqueue") and cls.__name__ == "BirdPreTrainedModel":
            cls.head_loss = BirdPreTrainedModel.head_loss
            cls.head_loss_begin = BirdPreTrainedModel.head_loss_begin
            cls.head_loss_end = BirdPreTrainedModel.head_loss_end
    return classes


def _event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _register_queues(mod, D, K, F):
    """Queue buffers exactly as modules/modeling.py:138-151 creates them."""
    shapes = [("queue_v_cross_ng", K), ("queue_frame_proj_ng", K * F), ("queue_frame_cross_ng", K * F),
              ("queue_title_cross_ng", K), ("queue_tag_cross_ng", K)]
    for name, cols in shapes:
        mod.register_buffer(name, torch.nn.functional.normalize(torch.randn(D, cols), dim=0))
    mod.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))


class BirdPreTrainedModel(ContrastiveHeadMixin, nn.Module):
    """Pre-train model with the reference's head; the encoders / projectors / MLM head are
    injected (any callables with the reference's signatures), because they are out of
    scope here."""

    def __init__(self, cross_config, task_config, visual_encoder=None, text_encoder=None, visual_encoder_k=None,
                 text_encoder_k=None, v_projector=None, v_projector_k=None, v_predictor=None, get_mlm_loss=None):
        super().__init__()
        self.task_config = task_config
        self.rank = getattr(task_config, "local_rank", 0)
        self.top_frames = getattr(task_config, "top_frames", 3)
        self.weight_FAM = cross_config.weight_FAM
        self.weight_VTM = cross_config.weight_VTM
        self.weight_FTM = cross_config.weight_FTM
        self.weight_MLM = cross_config.weight_MLM
        self.contrast_momentum = task_config.contrast_momentum
        self.contrast_temperature = task_config.contrast_temperature
        self.contrast_num_negative = task_config.contrast_num_negative
        self.head_precision = getattr(task_config, "head_precision", None)
        self.visual_encoder, self.text_encoder = visual_encoder, text_encoder
        self.visual_encoder_k, self.text_encoder_k = visual_encoder_k, text_encoder_k
        self.v_projector, self.v_projector_k, self.v_predictor = v_projector, v_projector_k, v_predictor
        if get_mlm_loss is not None:
            self.get_mlm_loss = get_mlm_loss
        self.model_pairs = [[a, b] for a, b in ((visual_encoder, visual_encoder_k), (text_encoder, text_encoder_k),
                                                (v_projector, v_projector_k))
                            if isinstance(a, nn.Module) and isinstance(b, nn.Module)]
        if self.model_pairs:
            self.copy_params()
        _register_queues(self, cross_config.temporal_hidden_size, self.contrast_num_negative,
                         task_config.max_frames)
        self.loss_fct = CrossEn()

    def get_mlm_loss(self, input_ids, input_mask):       # out of scope: MLM head (modules/modeling.py:153-205)
        return torch.zeros((), device=self.queue_ptr.device)

    def head_loss(self, v_fea, frame_fea, title_fea, frame_pred, v_fea_k, frame_fea_k, title_fea_k, tag_fea_k,
                  frame_proj_k, loss_MLM=None):
        """modules/modeling.py:385-424 for dataset != "bird" (the only reachable branch)."""
        b = v_fea.shape[0]
        D = v_fea.shape[-1]
        # the key exchange does not depend on the loss: start it first, enqueue after the loss kernels
        marks = getattr(self, "_hmmc_marks", None)          # optional CUDA-event breakdown (bench.py)
        mark = (lambda: marks.append(_event())) if marks is not None else (lambda: None)
        mark()
        pending = self._gather_keys_async(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        mark()
        # the queues are free once the two GEMMs have read them: the enqueue then runs on a side stream next to
        # the rest of the loss (positives, gradient projection, reductions) and is joined before returning
        released = torch.cuda.Event() if ENQUEUE_OVERLAP and marks is None else None
        total, parts = ops.pretrain_head(v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                         v_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k, frame_proj_k,
                                         self.queue_v_cross_ng, self.queue_title_cross_ng, self.queue_frame_proj_ng,
                                         self.queue_frame_cross_ng, self.contrast_temperature, self.weight_FAM,
                                         self.weight_VTM, self.weight_FTM, self.task_config.use_frame_fea,
                                         self.head_precision, release_event=released)
        self.last_loss_parts = parts            # [FAM, VTM, FTM], device tensor (the reference logs them)
        mark()
        if released is None:
            self._enqueue_gathered(pending)
        else:
            main = torch.cuda.current_stream()
            side = _side_stream(main.device)
            side.wait_event(released)
            with torch.cuda.stream(side):
                self._enqueue_gathered(pending)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
        mark()
        if loss_MLM is None:
            return total
        return total + self.weight_MLM * loss_MLM

    def head_loss_begin(self, v_fea, frame_fea, title_fea, frame_pred):
        """Query-side half of head_loss: normalisation and both GEMM passes against the queues, issued on a
        high-priority side stream so that they run beside `_momentum_update()` and the key encoders that the
        reference's forward executes next (modules/modeling.py:364-377).  The GEMM grids are kept on a few SMs
        (LOSS_GEMM_RESERVED): the EMA is HBM-bound and needs the others to saturate the memory system.
        Returns the state for head_loss_end."""
        b, D = v_fea.shape[0], v_fea.shape[-1]
        side = _side_stream(torch.cuda.current_stream().device, "loss", priority=-1)
        ops.set_reserved_sms(LOSS_GEMM_RESERVED)
        try:
            state = ops.pretrain_head_begin(v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                            self.queue_v_cross_ng, self.queue_title_cross_ng, self.queue_frame_proj_ng,
                                            self.queue_frame_cross_ng, self.contrast_temperature, self.weight_FAM,
                                            self.weight_VTM, self.weight_FTM, self.task_config.use_frame_fea,
                                            self.head_precision, stream=side)
        finally:
            ops.set_reserved_sms(0)
        state_done = torch.cuda.Event()
        state_done.record(side)
        return (state, state_done, (v_fea, frame_fea, title_fea, frame_pred))

    def head_loss_end(self, begun, v_fea_k, frame_fea_k, title_fea_k, tag_fea_k, frame_proj_k, loss_MLM=None):
        """Key-side half: the enqueue (and the wait for the key all-gather) starts at once on a side stream —
        the queues are no longer read — while positives, losses and gradients run on the current stream."""
        state, gemm_done, (v_fea, frame_fea, title_fea, frame_pred) = begun
        b, D = v_fea.shape[0], v_fea.shape[-1]
        main = torch.cuda.current_stream()
        main.wait_event(gemm_done)
        pending = self._gather_keys_async(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        keys_ready = torch.cuda.Event()
        keys_ready.record(main)
        total, parts = ops.pretrain_head_end(state, v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                             v_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k, frame_proj_k)
        self.last_loss_parts = parts
        side = _side_stream(main.device)
        side.wait_event(keys_ready)
        with torch.cuda.stream(side):
            self._enqueue_gathered(pending)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)
        if loss_MLM is None:
            return total
        return total + self.weight_MLM * loss_MLM

    def forward(self, video_data, video_frame, tag_ids, tag_mask, title_ids, title_mask, global_step):
        tag_ids = tag_ids.view(-1, tag_ids.shape[-1])
        tag_mask = tag_mask.view(-1, tag_mask.shape[-1])
        title_ids = title_ids.view(-1, title_ids.shape[-1])
        title_mask = title_mask.view(-1, title_mask.shape[-1])
        video = torch.as_tensor(video_data)
        if not self.training:
            return None
        v_fea, frame_fea = self.visual_encoder(video, video_frame)
        title_fea = self.text_encoder(title_ids, title_mask)
        bs, frame, hidden = frame_fea.shape
        frame_fea = frame_fea.view(-1, hidden)
        frame_proj = self.v_projector(frame_fea)
        frame_pred = self.v_predictor(frame_proj)
        frame_fea = frame_fea.view(bs, frame, hidden)
        frame_pred = frame_pred.view(bs, frame, hidden)
        # the queries exist: their half of the loss runs beside the momentum update and the key encoders
        begun = self.head_loss_begin(v_fea, frame_fea, title_fea, frame_pred) if use_split_schedule() else None
        with torch.no_grad():  # no gradient to keys
            self._momentum_update()  # update the key encoder
            tag_fea_k = self.text_encoder_k(tag_ids, tag_mask)
            title_fea_k = self.text_encoder_k(title_ids, title_mask)
            v_fea_k, frame_fea_k = self.visual_encoder_k(video, video_frame)
            frame_fea_k = frame_fea_k.view(-1, hidden)
            frame_proj_k = self.v_projector_k(frame_fea_k)
            frame_fea_k = frame_fea_k.view(bs, frame, hidden)
            frame_proj_k = frame_proj_k.view(bs, frame, hidden)
        loss_MLM = self.get_mlm_loss(title_ids, title_mask)
        if begun is not None:
            return self.head_loss_end(begun, v_fea_k, frame_fea_k, title_fea_k, tag_fea_k, frame_proj_k, loss_MLM)
        return self.head_loss(v_fea, frame_fea, title_fea, frame_pred, v_fea_k, frame_fea_k, title_fea_k,
                              tag_fea_k, frame_proj_k, loss_MLM)


class BirdModel(ContrastiveHeadMixin, nn.Module):
    """Fine-tune / eval model with the reference's head (modules/modeling.py:648-722)."""

    def __init__(self, cross_config, task_config, text_encoder=None, visual_encoder=None, logit_scale=4.6052):
        super().__init__()
        self.task_config = task_config
        self.rank = getattr(task_config, "local_rank", 0)
        self.weight_VTM_finetune = cross_config.weight_VTM_finetune
        self.weight_FTM_finetune = cross_config.weight_FTM_finetune
        self.top_frames = task_config.top_frames
        self.head_precision = getattr(task_config, "head_precision", None)
        if text_encoder is None:
            text_encoder = SimpleNamespace()
        if not hasattr(text_encoder, "logit_scale"):
            text_encoder.logit_scale = torch.tensor(logit_scale, dtype=torch.float32)
        self.text_encoder = text_encoder
        self.visual_encoder = visual_encoder
        self.loss_fct = CrossEn()

    def head_loss(self, query_output, visual_output, frame_output):
        """modules/modeling.py:698-709: gather, then the fused hierarchical-matching loss.
        Every rank evaluates the global loss, as in the reference."""
        b = query_output.shape[0]
        F = frame_output.shape[1]
        D = query_output.shape[-1]
        # one packed exchange instead of the reference's three dist_collect calls
        packed = torch.cat([query_output.reshape(b, D), visual_output.reshape(b, D),
                            frame_output.reshape(b, F * D)], dim=1)
        # every rank evaluates the same global loss on the same gathered rows: the gather's backward needs no
        # exchange (parallel.all_gather_cat_replicated); dist_collect keeps the general reduce-scatter
        full = parallel.all_gather_cat_replicated(packed)
        if not bool(getattr(self.task_config, "use_frame_fea", True)):
            return self.finetune_head_loss(full[:, :D], full[:, D:2 * D], None)
        # the fused head reads the gathered rows in place and returns the gradient in the same layout
        return ops.sym_ce_packed(full, F, D, self._logit_scale(), self.weight_VTM_finetune,
                                 self.weight_FTM_finetune, self.head_precision)

    def forward(self, query_ids, query_mask, video_data, video_frame, idx, global_step):
        query_ids = query_ids.view(-1, query_ids.shape[-1])
        query_mask = query_mask.view(-1, query_mask.shape[-1])
        video = torch.as_tensor(video_data)
        if not self.training:
            return None
        query_output = self.text_encoder(query_ids, query_mask)
        visual_output, frame_output = self.visual_encoder(video, video_frame)
        return self.head_loss(query_output, visual_output, frame_output)
