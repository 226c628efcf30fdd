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
from types import SimpleNamespace

import torch
from torch import nn

from . import ops, parallel

logger = logging.getLogger(__name__)

# modules/cross-base/cross_config.json (head-relevant keys)
DEFAULT_CROSS_CONFIG = dict(temporal_hidden_size=512, weight_FAM=0.05, weight_VTM=0.45, weight_FTM=0.45,
                            weight_MLM=0.05, weight_VTM_finetune=0.85, weight_FTM_finetune=0.15)


# Schedule constants (measured on B200; these were environment knobs while being tuned):
# * the enqueue of a single-rank step runs on a side stream next to the tail of the loss;
# * the query-side GEMMs of the loss CAN run beside the momentum update (head_loss_begin / head_loss_end) with
#   their persistent grids kept on 148 - LOSS_GEMM_RESERVED SMs (the EMA is HBM-bound and needs the other SMs
#   to saturate the memory system, tools/overlap_probe.py).  With round 2's loss kernels the plain sequence is
#   faster (0.388 ms vs 0.400 ms per step on one B200), so forward() no longer splits by default;
# * with several ranks the key exchange and the enqueue of step i are deferred to run beside step i+1's momentum
#   update (nothing reads the queues before step i+1's loss), see ContrastiveHeadMixin.start_pending_exchange.
LOSS_GEMM_RESERVED = 120
_side_streams = {}


SPLIT_SCHEDULE = False


def use_split_schedule():
    """Whether forward() issues the query-side half of the loss beside the momentum update.  Off: measured
    slower than the plain sequence since the loss kernels shrank (see above); with several ranks the side
    stream beside the EMA carries the deferred key exchange instead."""
    return SPLIT_SCHEDULE and parallel.world()[0] == 1


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
        # The device pointer table is built once and rebuilt when any parameter has moved (.to(), .half(),
        # load_state_dict with assign, `.data =`): all pointers are compared on the host every call.
        tab = getattr(self, "_hmmc_ema", None)
        pairs = self._ema_pairs()
        if tab is not None and not tab.still_valid(pairs):
            tab = None
        if tab is None:
            tab = ops.EmaTable(pairs)
            self._hmmc_ema = tab
        tab.run(self.contrast_momentum)

    def _queue_buffers(self):
        return [self.queue_v_cross_ng, self.queue_tag_cross_ng, self.queue_title_cross_ng,
                self.queue_frame_cross_ng, self.queue_frame_proj_ng]

    # The exchange of _dequeue_and_enqueue (modules/modeling.py:249-258) is ONE packed all-gather of the five
    # key tensors instead of five; the enqueue kernel reads the gathered rows in place.
    @staticmethod
    def _key_blocks(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        b = v_fea_k.shape[0]
        D = v_fea_k.shape[-1]
        frame_fea_k = frame_fea_k.reshape(b, -1, D)
        frame_proj_k = frame_proj_k.reshape(b, -1, D)
        F = frame_fea_k.shape[1]
        return b, F, D, [v_fea_k.reshape(b, D), tag_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k,
                         frame_proj_k]

    def _enqueue_ptr_mode(self, B):
        """-1 = the kernels read and advance queue_ptr on the device (no host copy that a CUDA-graph replay or a
        deferred enqueue could leave stale); needs K % B == 0, the reference's own operating condition
        (modules/modeling.py:273-280 has no wrap-around).  Otherwise the host value, one sync per call."""
        K = self.contrast_num_negative
        if K % B == 0:
            return -1
        if torch.cuda.is_current_stream_capturing():
            raise ValueError("graph capture of the enqueue needs K %% (world*batch) == 0 (K=%d, B=%d)" % (K, B))
        return int(self.queue_ptr)

    @torch.no_grad()
    def _enqueue_rows(self, W, b, F, D, gathered=None, direct=None, staged=None, slot=None, prenormalised=False):
        ops.enqueue(gathered, W, b, F, D, self._queue_buffers(), self.queue_ptr, self._enqueue_ptr_mode(W * b),
                    self.contrast_num_negative, ops.resolve_precision(self.head_precision), direct=direct,
                    staged=staged, slot=slot, prenormalised=prenormalised)

    def _use_peer_exchange(self):
        """Exchange the keys over peer memory (parallel.PeerExchange) rather than with an NCCL all-gather:
        default whenever the ranks can map each other's memory; task_config.peer_exchange = False forces NCCL."""
        v = getattr(self.task_config, "peer_exchange", None)
        return True if v is None else bool(v)

    def _defer_enqueue(self):
        v = getattr(self.task_config, "defer_enqueue", None)
        return parallel.world()[0] > 1 if v is None else bool(v)

    @torch.no_grad()
    def _stage_keys(self, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        """Deferred schedule, end of step i: copy the step's keys into the persistent send buffer.  Their
        exchange and enqueue are issued by start_pending_exchange() of step i+1 (or flush_pending_enqueue()).
        A device-side mark travels with the buffer: the packing kernel sets it, the enqueue consumes it, so an
        eager flush followed by the replay of a captured step (whose graph carries the same enqueue) writes the
        keys once."""
        self._join_pending()
        b, F, D, keys = self._key_blocks(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        W, _ = parallel.world()
        if self.contrast_num_negative % (W * b) != 0:
            raise ValueError("the deferred enqueue needs K %% (world*batch) == 0 (K=%d, B=%d): the queue pointer "
                             "lives on the device" % (self.contrast_num_negative, W * b))
        dev = keys[0].device
        width = (3 + 2 * F) * D
        bufs = getattr(self, "_hmmc_xchg", None)
        if bufs is None or bufs[0].shape != (b, width) or bufs[3] != W or bufs[0].device != dev:
            # (collective when W > 1: every rank builds its buffers at the same first deferred step, outside a
            # graph capture)
            send = torch.empty(b, width, dtype=torch.float32, device=dev)
            peer = parallel.make_peer_exchange(b, width, dev) if (W > 1 and self._use_peer_exchange()) else None
            if W == 1:
                gathered = send
            elif peer is not None:
                gathered = peer.recv
            else:
                gathered = torch.empty(W * b, width, dtype=torch.float32, device=dev)
            bufs = (send, gathered, torch.zeros(1, dtype=torch.int32, device=dev), W, peer)
            self._hmmc_xchg = bufs
        # normalised at the source: each rank normalises its own b keys, the enqueue reads the W*b received rows once
        ops.pack_rows(keys, out=bufs[0], staged=bufs[2], norm_dim=D)
        self._hmmc_pending = {"dims": (W, b, F, D), "done": None}
        if not getattr(self, "_hmmc_hooked", False) and isinstance(self, nn.Module):
            # checkpoints must see the queues with every staged key in place
            self.register_state_dict_pre_hook(lambda m, *a, **k: m.flush_pending_enqueue())
            self._register_load_state_dict_pre_hook(lambda *a, **k: self.flush_pending_enqueue())
            self._hmmc_hooked = True

    @torch.no_grad()
    def start_pending_exchange(self, force=False):
        """Deferred schedule, start of step i+1: issue the all-gather and the enqueue of step i's keys on a side
        stream and return at once.  Call it before `_momentum_update()` (forward() does): the exchange then
        runs beside the HBM-bound parameter update and is joined right before the loss reads the queues.
        Safe to call at any time; does nothing when no keys are staged (the enqueue itself is guarded by the
        device-side mark, so even a redundant issue writes nothing)."""
        p = getattr(self, "_hmmc_pending", None)
        if getattr(self, "_hmmc_xchg", None) is None or p is None:
            return
        if p["done"] is not None and not force:
            return
        W, b, F, D = p["dims"]
        send, gathered, staged, _, peer = self._hmmc_xchg
        main = torch.cuda.current_stream()
        side = _side_stream(main.device, "exchange", priority=-1)
        fork = torch.cuda.Event()
        fork.record(main)                  # after step i's kernels (they read the queues) and the staging copy
        side.wait_event(fork)
        with torch.cuda.stream(side):
            if W > 1 and peer is not None:
                peer.exchange(send)
            elif W > 1:
                parallel.all_gather_rows_into(gathered, send, group=parallel.overlap_group())
            self._enqueue_rows(W, b, F, D, gathered=gathered, staged=staged, prenormalised=True,
                               slot=(peer.epoch, peer.slot_stride) if peer is not None else None)
            done = torch.cuda.Event()
            done.record(side)
        p["done"] = done

    def _join_pending(self):
        p = getattr(self, "_hmmc_pending", None)
        if p is None or p.get("joined"):
            return                  # nothing staged, or already ordered before the current stream's work
        if p["done"] is None:
            self.start_pending_exchange()
        torch.cuda.current_stream().wait_event(p["done"])
        p["joined"] = True          # (a capture starting later inherits this order; it must not wait on this event)

    def flush_pending_enqueue(self):
        """Make the queue buffers current (stream-ordered on the current stream): call before reading
        queue_*_ng / queue_ptr directly, and after replaying a captured step before issuing eager ones (a replay
        stages keys without this object noticing).  state_dict() and load_state_dict() do it themselves.
        Always issues the exchange: the device-side mark decides whether anything is written."""
        if getattr(self, "_hmmc_pending", None) is None:
            return
        p = self._hmmc_pending
        p["joined"] = False
        self.start_pending_exchange(force=True)
        self._join_pending()

    @torch.no_grad()
    def _dequeue_and_enqueue(self, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        """modules/modeling.py:244-284, immediately (the reference's call): gather, normalise, write columns."""
        self._join_pending()
        b, F, D, keys = self._key_blocks(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        W, _ = parallel.world()
        if W == 1:
            self._enqueue_rows(1, b, F, D, direct=[ops._f32c(t, "key") for t in keys])
        else:
            self._enqueue_rows(W, b, F, D, gathered=parallel.all_gather_rows(ops.pack_rows(keys)))


def patch_reference_classes(*classes):
    """Install the B200 head on the reference's own model classes: every method
    of ContrastiveHeadMixin replaces the class's method of the same name; classes that own the
    negative queues also get the fused ``head_loss``."""
    for cls in classes:
        for name, fn in vars(ContrastiveHeadMixin).items():
            if callable(fn) and not name.startswith("__"):
                setattr(cls, name, fn)
        if hasattr(cls, "_dequeue_and_enqueue") and cls.__name__ == "BirdPreTrainedModel":
            for name in ("head_loss", "head_loss_begin", "head_loss_end", "_enqueue_beside"):
                setattr(cls, name, getattr(BirdPreTrainedModel, name))
    return classes


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
        """modules/modeling.py:385-424 for dataset != "bird" (the only reachable branch): the three
        losses forward and backward, then `_dequeue_and_enqueue` of the step's keys.

        Single rank: the enqueue runs on a side stream next to the tail of the loss and is joined before
        returning.  Several ranks (or task_config.defer_enqueue): nothing reads the queues again before the
        NEXT step's loss, so the keys are only staged here; their all-gather and enqueue run beside the next
        step's momentum update (start_pending_exchange) and are joined right before that step's loss."""
        b = v_fea.shape[0]
        D = v_fea.shape[-1]
        defer = self._defer_enqueue()
        self._join_pending()                 # the previous step's keys must be in the queues this loss reads
        released = None if defer else torch.cuda.Event()
        total, parts = ops.pretrain_head(v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                         v_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k, frame_proj_k,
                                         self.queue_v_cross_ng, self.queue_title_cross_ng, self.queue_frame_proj_ng,
                                         self.queue_frame_cross_ng, self.contrast_temperature, self.weight_FAM,
                                         self.weight_VTM, self.weight_FTM, self.task_config.use_frame_fea,
                                         self.head_precision, release_event=released)
        self.last_loss_parts = parts            # [FAM, VTM, FTM], device tensor (the reference logs them)
        if defer:
            self._stage_keys(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        else:
            # the queues are free once the two GEMMs have read them
            self._enqueue_beside(released, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        if loss_MLM is None:
            return total
        return total + self.weight_MLM * loss_MLM

    def _enqueue_beside(self, after, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k):
        """Immediate enqueue on the side stream, started once `after` (an event of the current stream) has
        happened and joined into the current stream."""
        main = torch.cuda.current_stream()
        side = _side_stream(main.device)
        side.wait_event(after)
        with torch.cuda.stream(side):
            self._dequeue_and_enqueue(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)

    def head_loss_begin(self, v_fea, frame_fea, title_fea, frame_pred):
        """Query-side half of head_loss: normalisation and both GEMM passes against the queues, issued on a
        high-priority side stream so that they run beside `_momentum_update()` and the key encoders that the
        reference's forward executes next (modules/modeling.py:364-377).  The GEMM grids are kept on a few SMs
        (LOSS_GEMM_RESERVED): the EMA is HBM-bound and needs the others to saturate the memory system.
        Returns the state for head_loss_end."""
        b, D = v_fea.shape[0], v_fea.shape[-1]
        self._join_pending()
        side = _side_stream(torch.cuda.current_stream().device, "loss", priority=-1)
        state = ops.pretrain_head_begin(v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                        self.queue_v_cross_ng, self.queue_title_cross_ng, self.queue_frame_proj_ng,
                                        self.queue_frame_cross_ng, self.contrast_temperature, self.weight_FAM,
                                        self.weight_VTM, self.weight_FTM, self.task_config.use_frame_fea,
                                        self.head_precision, stream=side, reserved_sms=LOSS_GEMM_RESERVED)
        state_done = torch.cuda.Event()
        state_done.record(side)
        return (state, state_done, (v_fea, frame_fea, title_fea, frame_pred))

    def head_loss_end(self, begun, v_fea_k, frame_fea_k, title_fea_k, tag_fea_k, frame_proj_k, loss_MLM=None):
        """Key-side half: the enqueue starts at once on a side stream - the queues are no longer read - while
        positives, losses and gradients run on the current stream."""
        state, gemm_done, (v_fea, frame_fea, title_fea, frame_pred) = begun
        b, D = v_fea.shape[0], v_fea.shape[-1]
        main = torch.cuda.current_stream()
        main.wait_event(gemm_done)
        keys_ready = torch.cuda.Event()
        keys_ready.record(main)
        total, parts = ops.pretrain_head_end(state, v_fea.reshape(b, D), title_fea.reshape(b, D), frame_fea, frame_pred,
                                             v_fea_k.reshape(b, D), title_fea_k.reshape(b, D), frame_fea_k, frame_proj_k)
        self.last_loss_parts = parts
        if self._defer_enqueue():
            self._stage_keys(v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
        else:
            self._enqueue_beside(keys_ready, v_fea_k, tag_fea_k, title_fea_k, frame_fea_k, frame_proj_k)
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
        # deferred schedule: the previous step's key exchange and enqueue run beside the momentum update
        self.start_pending_exchange()
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
        if parallel.world()[0] == 1:
            # nothing to gather: the fused head reads the three tensors where they are
            return self.finetune_head_loss(query_output.reshape(b, D), visual_output.reshape(b, D), frame_output)
        # one packed exchange instead of the reference's three dist_collect calls (one launch to pack, one to
        # unpack the gradient)
        packed = ops.pack_rows_autograd([query_output.reshape(b, D), visual_output.reshape(b, D),
                                         frame_output.reshape(b, F * D)])
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
