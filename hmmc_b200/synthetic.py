"""Seeded synthetic embeddings for tests, golden fixtures and bench.py.

Everything is drawn with ``numpy.random.RandomState`` (the legacy MT19937
stream is bit-stable across numpy versions and machines), so the container that
generates ``tests/golden`` and the GPU box that checks against it see the same
bytes.  Shapes follow the reference's batch layouts:
``[b, D]`` for text / video embeddings and ``[b, F, D]`` for frame embeddings.
"""
import numpy as np


def _randn(rs, *shape):
    return rs.standard_normal(shape).astype(np.float32)


def normalize_cols(x, eps=1e-12):
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=0, keepdims=True))
    return (x / np.maximum(n, eps)).astype(np.float32)


def finetune_inputs(B, F=12, D=512, seed=1, corr=0.0):
    """text [B,D], video [B,D], frames [B,F,D] (modules/modeling.py:691-692)."""
    rs = np.random.RandomState(seed)
    t = _randn(rs, B, D)
    v = _randn(rs, B, D)
    fr = _randn(rs, B, F, D)
    if corr:
        v += corr * t
        fr += (0.66 * corr) * t[:, None, :]
    return t, v, fr


PRETRAIN_Q_NAMES = ["v_fea", "frame_fea", "title_fea", "frame_pred", "frame_proj"]
PRETRAIN_K_NAMES = ["v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k", "frame_proj_k"]
QUEUE_NAMES = ["queue_v_cross_ng", "queue_frame_proj_ng", "queue_frame_cross_ng",
               "queue_title_cross_ng", "queue_tag_cross_ng"]


def pretrain_inputs(b, F=12, D=512, seed=2, corr=0.5):
    """Query- and key-side embeddings of one pre-train step
    (modules/modeling.py:347-378).  Keys are correlated with queries so the
    positives are not lost in the noise."""
    rs = np.random.RandomState(seed)
    out = {}
    out["v_fea"] = _randn(rs, b, D)
    out["frame_fea"] = _randn(rs, b, F, D)
    out["title_fea"] = _randn(rs, b, D) + corr * out["v_fea"]
    out["frame_pred"] = _randn(rs, b, F, D)
    out["frame_proj"] = _randn(rs, b, F, D)
    out["v_fea_k"] = _randn(rs, b, D) + corr * out["v_fea"]
    out["frame_fea_k"] = _randn(rs, b, F, D) + corr * out["frame_fea"]
    out["title_fea_k"] = _randn(rs, b, D) + corr * out["title_fea"]
    out["tag_fea_k"] = _randn(rs, b, D)
    out["frame_proj_k"] = _randn(rs, b, F, D) + corr * out["frame_pred"]
    return out


def queues(K, F=12, D=512, seed=3):
    """The five negative queues in the state-dict layout ``[D, Kq]`` with
    unit-norm columns (modules/modeling.py:138-149)."""
    rs = np.random.RandomState(seed)
    shapes = {"queue_v_cross_ng": K, "queue_frame_proj_ng": K * F,
              "queue_frame_cross_ng": K * F, "queue_title_cross_ng": K,
              "queue_tag_cross_ng": K}
    return {n: normalize_cols(_randn(rs, D, shapes[n])) for n in QUEUE_NAMES}


def eval_inputs(Nt, Nv, F=12, D=512, seed=4, per_video=None, corr=0.15):
    """Text / video / frame embeddings of a retrieval eval set.

    Square (``per_video is None``): text i matches video i (MSR-VTT 1k-A,
    dataloaders/dataloader_msrvtt_retrieval.py:155-164).  Multi-sentence:
    ``per_video`` captions per video, text s matches video ``gt[s]`` (VATEX,
    dataloaders/dataloader_vatex_retrieval.py:77-89).  Returns
    (T, V, Fr, gt, cut_off_points) with cut_off_points already shifted by -1 as
    main_task_retrieval.py:380 does.
    """
    rs = np.random.RandomState(seed)
    if per_video is None:
        assert Nt == Nv
        gt = np.arange(Nt, dtype=np.int64)
        cut = None
    else:
        per = np.asarray(per_video, dtype=np.int64)
        assert per.shape[0] == Nv and int(per.sum()) == Nt
        gt = np.repeat(np.arange(Nv, dtype=np.int64), per)
        cut = (np.cumsum(per) - 1).tolist()
    T = _randn(rs, Nt, D)
    V = _randn(rs, Nv, D)
    Fr = _randn(rs, Nv, F, D)
    # make the ground-truth pair stand out: the video side carries a copy of
    # the mean of its captions
    acc = np.zeros((Nv, D), dtype=np.float32)
    np.add.at(acc, gt, T)
    cnt = np.maximum(np.bincount(gt, minlength=Nv), 1).astype(np.float32)[:, None]
    V += corr * acc / np.sqrt(cnt)
    Fr += (0.66 * corr) * (acc / np.sqrt(cnt))[:, None, :]
    return T, V, Fr, gt, cut


def ema_tensors(seed=5, sizes=((513,), (64, 33), (7,), (1024, 96), (3, 5, 7), (1,)),
                dtypes=("float32", "float16", "float32", "float16", "float32", "float32")):
    """(param, param_k) pairs in mixed fp32 / fp16 like the CLIP weights after
    convert_weights (modules/module_clip.py:506-527)."""
    rs = np.random.RandomState(seed)
    ps, pks = [], []
    for shp, dt in zip(sizes, dtypes):
        ps.append(_randn(rs, *shp).astype(dt))
        pks.append(_randn(rs, *shp).astype(dt))
    return ps, pks


def ema_param_numels(total=172325632, count=362):
    """Element counts of the 362 parameter tensors the reference's _momentum_update walks
    (ViT-B/32 87.85 M, temporal transformer 12.63 M, CLIP text 63.43 M, two MLPs; SURVEY.md
    §8 a10).  The shapes follow the transformer block pattern (qkv, out-proj, two MLP matrices,
    biases and LayerNorms); a final filler tensor makes the total exact."""
    sizes = []

    def block(width, mlp):
        sizes.extend([3 * width * width, 3 * width, width * width, width, width, width,
                      width * mlp, mlp, mlp * width, width, width, width])
    sizes.extend([768 * 3 * 32 * 32, 768, 50 * 768, 768, 768])        # ViT stem
    for _ in range(12):
        block(768, 3072)
    sizes.extend([768, 768, 768 * 512])
    sizes.extend([49408 * 512, 77 * 512])                               # CLIP text stem
    for _ in range(12):
        block(512, 2048)
    sizes.extend([512, 512, 512 * 512])
    sizes.extend([48 * 512])                                            # temporal transformer
    for _ in range(4):
        block(512, 2048)
    for _ in range(2):                                                  # projector MLPs
        sizes.extend([512 * 4096, 4096, 4096, 4096, 4096 * 512, 512])
    sizes = sizes[:count - 1]
    rest = total - sum(sizes)
    assert rest > 0, rest
    sizes.append(rest)
    return sizes


OPTIM_SIZES = ((37, 5), (1000,), (3,), (90, 100), (8192,), (8193,), (1,), (64, 129))
OPTIM_GROUP_OF = (0, 1, 1, 0, 0, 1, 1, 0)


def optim_groups(case="pretrain"):
    """Parameter-group hyper-parameters.  'pretrain': the shape prep_optimizer builds
    (main_pretrain.py:168-202: decay / no-decay groups with their own lr, warmup_cosine, b2=0.98,
    per-parameter max_grad_norm 1.0).  'plain': constant lr (t_total=-1), no per-parameter clip.
    'linear': warmup_linear with the constructor's default betas."""
    if case == "pretrain":
        common = dict(schedule='warmup_cosine', warmup=0.1, t_total=40, b1=0.9, b2=0.98, e=1e-6, max_grad_norm=1.0)
        return [dict(common, lr=1e-4, weight_decay=0.2), dict(common, lr=5e-5, weight_decay=0.0)]
    if case == "plain":
        common = dict(schedule='warmup_linear', warmup=-1, t_total=-1, b1=0.9, b2=0.999, e=1e-6, max_grad_norm=-1)
        return [dict(common, lr=3e-4, weight_decay=0.01), dict(common, lr=3e-4, weight_decay=0.0)]
    if case == "linear":
        common = dict(schedule='warmup_linear', warmup=0.25, t_total=8, b1=0.9, b2=0.999, e=1e-6, max_grad_norm=1.0)
        return [dict(common, lr=2e-4, weight_decay=0.01), dict(common, lr=1e-4, weight_decay=0.0)]
    raise ValueError(case)


def optim_tensors(seed=11, sizes=OPTIM_SIZES):
    """Initial fp32 parameters of the optimizer parity cases."""
    rs = np.random.RandomState(seed)
    return [(_randn(rs, *shp) * 0.05).astype(np.float32) for shp in sizes]


def optim_grads(step, seed=12, sizes=OPTIM_SIZES, scales=(0.02, 0.002, 0.05, 0.0005, 0.004, 0.03)):
    """Gradients of step `step`; the scales alternate between clipped (norm > 1) and unclipped steps."""
    rs = np.random.RandomState(seed + 977 * step)
    s = scales[step % len(scales)]
    return [(_randn(rs, *shp) * s).astype(np.float32) for shp in sizes]


def eval_epoch_case(multi, batch):
    """Features and a fake dataloader layout for the eval_epoch parity cases: batches of index tensors
    (query_ids = caption index, video = video index of that caption); the stub encoders of the test
    and of oracle/gen_golden.py look the features up by these indices."""
    if multi:
        rs = np.random.RandomState(21)
        per = rs.randint(1, 8, size=40)
        T, V, Fr, gt, cut = eval_inputs(int(per.sum()), 40, seed=22, per_video=per)
        vid_of_caption = np.repeat(np.arange(40), per)
        cut_1based = [c + 1 for c in cut]            # the dataset stores them 1-based (eval_epoch subtracts 1)
    else:
        T, V, Fr, gt, _ = eval_inputs(200, 200, seed=23, corr=0.07)
        vid_of_caption = np.arange(200)
        cut_1based = None
    n = T.shape[0]
    batches = [(np.arange(i, min(i + batch, n)), vid_of_caption[i:i + batch]) for i in range(0, n, batch)]
    return T, V, Fr, batches, cut_1based


def mlp_case(M=64, Din=64, Dh=128, Dout=64, seed=31):
    """Inputs, parameters (non-trivial BatchNorm affine and running statistics) and the upstream
    gradient of the MLP parity cases."""
    rs = np.random.RandomState(seed)
    f = lambda *shp, s=1.0: (_randn(rs, *shp) * s).astype(np.float32)
    return dict(x=f(M, Din), W1=f(Dh, Din, s=Din ** -0.5), b1=f(Dh, s=0.1), gamma=(1.0 + f(Dh, s=0.2)).astype(np.float32),
                beta=f(Dh, s=0.3), W2=f(Dout, Dh, s=Dh ** -0.5), b2=f(Dout, s=0.1),
                rm=f(Dh, s=0.1), rv=(1.0 + np.abs(f(Dh, s=0.2))).astype(np.float32), dy=f(M, Dout, s=0.05))
