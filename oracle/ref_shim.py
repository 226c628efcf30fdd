"""Load the *unmodified* reference head (cheetah003/HMMC) on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hmmc_b200/`` may import this file.
It only works where ``/root/reference`` exists (the build container); the GPU
box has no copy of the reference, so everything that travels is the golden
fixtures written by ``oracle/gen_golden.py`` under ``tests/golden/``.

What is shimmed (all *outside* the reference files):
  * absent third-party modules (diffdist, boto3, botocore, ftfy, tensorboardX,
    thop, lmdb) get empty stand-ins so ``modules/modeling.py`` imports;
  * ``torch.Tensor.cuda`` becomes the identity (modules/modeling.py:311
    hard-codes ``.cuda()`` for the labels);
  * ``modules.modeling.dist_collect`` is replaced by a single-process stand-in
    (identity for W=1).
The reference methods are then bound, unmodified, onto a SimpleNamespace
"self" whose encoders are stubs that return the synthetic embeddings.
"""
import os
import sys
import types
from types import SimpleNamespace as NS

import torch

REFERENCE_ROOT = os.environ.get("HMMC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "modeling.py"))


_loaded = {}


def load():
    """Import the reference's modeling / metrics / main_task_retrieval modules."""
    if _loaded:
        return _loaded["M"], _loaded["metrics"], _loaded["R"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in ["diffdist", "diffdist.functional", "boto3", "botocore",
                 "botocore.exceptions", "ftfy", "tensorboardX", "thop", "lmdb"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["diffdist"].functional = sys.modules["diffdist.functional"]
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["thop"].profile = None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self          # modeling.py:311
    torch.distributed.init_process_group = lambda *a, **k: None  # main_task_retrieval.py:28
    import modules.modeling as M
    import metrics
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        import main_task_retrieval as R
    finally:
        sys.argv = argv
    M.dist_collect = lambda x: x.contiguous()                    # W = 1
    _loaded.update(M=M, metrics=metrics, R=R)
    return M, metrics, R


def finetune_self(logit_scale=4.6052, top_frames=2):
    """A stand-in ``self`` carrying the reference's fine-tune head methods."""
    M, _, _ = load()
    s = NS()
    s.training = True
    s.rank = 1
    s.top_frames = top_frames
    s.task_config = NS(n_display=10 ** 9, use_frame_fea=True, logdir=None, local_rank=1)
    s.text_encoder = NS(logit_scale=torch.tensor(logit_scale, dtype=torch.float32))
    s.loss_fct = M.CrossEn()
    s.weight_VTM_finetune = 0.85    # modules/cross-base/cross_config.json
    s.weight_FTM_finetune = 0.15
    s.loose_similarity = types.MethodType(M.BirdPreTrainedModel.loose_similarity, s)
    s.frame_loss = types.MethodType(M.BirdModel.frame_loss, s)
    return s


def finetune_forward(t, v, fr, logit_scale=4.6052):
    """Run the reference's BirdModel.forward (modules/modeling.py:682-722) with
    stub encoders that return (t, v, fr).  Returns the scalar loss tensor."""
    M, _, _ = load()
    s = finetune_self(logit_scale)
    ls = torch.tensor(logit_scale, dtype=torch.float32)

    class _Txt:
        def __init__(self):
            self.logit_scale = ls

        def __call__(self, ids, mask):
            return t

    s.text_encoder = _Txt()
    s.visual_encoder = lambda video, video_frame: (v, fr)
    ids = torch.zeros(t.shape[0], 4, dtype=torch.long)
    return M.BirdModel.forward(s, ids, ids, torch.zeros(1), fr.shape[1], None, 1)


def pretrain_self(queues, K, F, T=0.07, momentum=0.99):
    M, _, _ = load()
    s = NS()
    s.training = True
    s.rank = 1
    s.task_config = NS(n_display=10 ** 9, dataset="chvtt", use_frame_fea=True,
                       logdir=None, max_frames=F)
    s.contrast_temperature = T
    s.contrast_momentum = momentum
    s.contrast_num_negative = K
    s.weight_FAM, s.weight_VTM, s.weight_FTM, s.weight_MLM = 0.05, 0.45, 0.45, 0.05
    for name, q in queues.items():
        setattr(s, name, q)
    s.queue_ptr = torch.zeros(1, dtype=torch.long)
    for name in ["contrastive_loss", "frame_self_loss", "frame_cross_loss",
                 "_dequeue_and_enqueue", "loose_similarity"]:
        setattr(s, name, types.MethodType(getattr(M.BirdPreTrainedModel, name), s))
    return s


def pretrain_forward(inp, queues, K, F, T=0.07):
    """Run the reference's BirdPreTrainedModel.forward (modules/modeling.py:334-436)
    with stub encoders.  ``inp`` holds v_fea, frame_fea, title_fea, frame_proj,
    frame_pred (q side) and v_fea_k, frame_fea_k, title_fea_k, tag_fea_k,
    frame_proj_k (key side).  MLM is stubbed to 0.  Returns (loss, self)."""
    M, _, _ = load()
    s = pretrain_self(queues, K, F, T)
    b = inp["v_fea"].shape[0]
    D = inp["v_fea"].shape[1]
    TAG = torch.zeros(b, 4, dtype=torch.long)
    TITLE = torch.ones(b, 4, dtype=torch.long)
    s.visual_encoder = lambda video, vf: (inp["v_fea"], inp["frame_fea"])
    s.text_encoder = lambda ids, mask: inp["title_fea"]
    s.v_projector = lambda x: inp["frame_proj"].reshape(-1, D)
    s.v_predictor = lambda x: inp["frame_pred"].reshape(-1, D)
    s._momentum_update = lambda: None
    s.visual_encoder_k = lambda video, vf: (inp["v_fea_k"], inp["frame_fea_k"])
    s.text_encoder_k = lambda ids, mask: inp["tag_fea_k"] if int(ids[0, 0]) == 0 else inp["title_fea_k"]
    s.v_projector_k = lambda x: inp["frame_proj_k"].reshape(-1, D)
    s.get_mlm_loss = lambda ids, mask: torch.zeros(())
    loss = M.BirdPreTrainedModel.forward(s, torch.zeros(1), F, TAG, TAG, TITLE, TITLE, 1)
    return loss, s
