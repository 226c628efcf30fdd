import types

import numpy as np
import pytest
import torch

from hmmc_b200 import modeling, ops
from hmmc_b200 import synthetic as syn
from oracle import head_oracle as O
from gpu_util import TOL, cu, rel

pytestmark = pytest.mark.gpu


def _model(K, F, D, prec, T=0.07):
    task = types.SimpleNamespace(local_rank=0, top_frames=3, contrast_momentum=0.99, contrast_temperature=T,
                                 contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision=prec)
    m = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).cuda()
    return m


def _load_queues(m, qs):
    with torch.no_grad():
        for n, x in qs.items():
            getattr(m, n).copy_(torch.from_numpy(x))


def test_contrastive_small_fp32(golden):
    g = golden("contrastive_small")
    m = _model(40, 1, 32, "fp32")
    q = cu(g["q"], True)
    loss = m.contrastive_loss(q, cu(g["k"]), cu(g["queue"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    assert rel(q.grad.cpu().numpy(), g["dq"]) < 2e-5


@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16"])
def test_pretrain_head_b32(golden, prec):
    """config 1: b=32, F=12, D=512, K=1024, T=0.07 against the reference's own outputs."""
    g = golden("pretrain_b32")
    b, F, D, K, T = int(g["b"]), int(g["F"]), int(g["D"]), int(g["K"]), float(g["T"])
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = syn.queues(K, F=F, D=D, seed=3)
    m = _model(K, F, D, prec, T)
    _load_queues(m, qs)
    ltol, gtol = TOL[prec]
    t = {n: cu(x, n in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for n, x in inp.items()}
    with torch.no_grad():
        fam = float(m.frame_self_loss(t["frame_pred"], t["frame_proj_k"], m.queue_frame_proj_ng))
        vtm = float(m.contrastive_loss(t["v_fea"], t["title_fea_k"], m.queue_title_cross_ng)
                    + m.contrastive_loss(t["title_fea"], t["v_fea_k"], m.queue_v_cross_ng))
        ftm = float(m.frame_cross_loss(t["frame_fea"], t["frame_fea_k"], m.queue_frame_cross_ng, t["title_fea"],
                                       t["title_fea_k"], m.queue_title_cross_ng))
    for got, name in ((fam, "fam"), (vtm, "vtm"), (ftm, "ftm")):
        assert abs(got - float(g[name])) / float(g[name]) < ltol, (name, got, float(g[name]))
    loss = m.head_loss(t["v_fea"], t["frame_fea"], t["title_fea"], t["frame_pred"], t["v_fea_k"], t["frame_fea_k"],
                       t["title_fea_k"], t["tag_fea_k"], t["frame_proj_k"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < ltol
    parts = m.last_loss_parts.cpu().numpy()                   # fused head reports FAM, VTM, FTM too
    for got, name in zip(parts, ("fam", "vtm", "ftm")):
        assert abs(got - float(g[name])) / float(g[name]) < ltol, (name, got)
    _, ref = O.pretrain_loss_and_grads(inp, qs, T)            # fp64 oracle (pinned to the golden grads)
    for n in ("v_fea", "title_fea", "frame_fea", "frame_pred"):
        got = t[n].grad.cpu().numpy()
        assert rel(got, ref[n]) < gtol, (n, rel(got, ref[n]))
        gg = g["d_" + n]
        assert rel(got[:gg.shape[0]], gg) < max(gtol, 1e-4), n
    # enqueue happened inside head_loss: pointer and the first written columns
    assert int(m.queue_ptr) == int(g["ptr"])
    for n in syn.QUEUE_NAMES:
        after = getattr(m, n).cpu().numpy()
        refq = g["after_" + n]
        np.testing.assert_allclose(after[:, :refq.shape[1]], refq, rtol=0, atol=2e-7)
        np.testing.assert_allclose(after.astype(np.float64).sum(1), g["sum_" + n], rtol=0, atol=2e-5)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_enqueue_wrap_and_packed_copies(golden, prec):
    g = golden("enqueue_wrap")
    K, F, D = 16, 4, 64
    m = _model(K, F, D, prec)
    _load_queues(m, syn.queues(K, F=F, D=D, seed=3))
    ptrs = []
    for step in range(3):
        k = syn.pretrain_inputs(8, F=F, D=D, seed=20 + step)
        m._dequeue_and_enqueue(cu(k["v_fea_k"]), cu(k["tag_fea_k"]), cu(k["title_fea_k"]), cu(k["frame_fea_k"]),
                               cu(k["frame_proj_k"]))
        ptrs.append(int(m.queue_ptr))
    assert ptrs == list(g["ptrs"])
    for n in syn.QUEUE_NAMES:
        np.testing.assert_allclose(getattr(m, n).cpu().numpy(), g["after_" + n], rtol=0, atol=2e-7)
    if prec != "fp32":
        # the packed operand copies written by the enqueue kernel equal a fresh re-pack
        for buf in m._queue_buffers():
            st, planes = ops.queue_state(buf), (2 if prec == "bf16x3" else 1)
            kd, dk = [x.clone() for x in st.pack(buf, planes)[:2]]      # as the enqueue kernel left them
            kd2, dk2 = st.repack(buf, planes)
            assert torch.equal(kd, kd2) and torch.equal(dk, dk2)
    # a batch that does not fit raises like the reference's slice assignment
    k = syn.pretrain_inputs(12, F=F, D=D, seed=1)
    with pytest.raises(Exception):
        m._dequeue_and_enqueue(cu(k["v_fea_k"]), cu(k["tag_fea_k"]), cu(k["title_fea_k"]), cu(k["frame_fea_k"]),
                               cu(k["frame_proj_k"]))


@pytest.mark.parametrize("ptr,b", [(3, 5), (4, 5), (0, 7), (6, 16)])
def test_enqueue_ragged_and_odd_pointer(ptr, b):
    """Shapes the vectorised scatter must guard: odd first column (scalar stores), odd batch (last
    column pair half empty), D not a multiple of the 64-wide tile.  Checked against numpy."""
    K, F, D = 22, 3, 70
    m = _model(K, F, D, "bf16x3")
    q0 = syn.queues(K, F=F, D=D, seed=3)
    _load_queues(m, q0)
    m.queue_ptr.fill_(ptr)
    k = syn.pretrain_inputs(b, F=F, D=D, seed=31)
    m._dequeue_and_enqueue(cu(k["v_fea_k"]), cu(k["tag_fea_k"]), cu(k["title_fea_k"]), cu(k["frame_fea_k"]),
                           cu(k["frame_proj_k"]))
    assert int(m.queue_ptr) == (ptr + b) % K
    src = {"queue_v_cross_ng": (k["v_fea_k"], 1), "queue_tag_cross_ng": (k["tag_fea_k"], 1),
           "queue_title_cross_ng": (k["title_fea_k"], 1), "queue_frame_cross_ng": (k["frame_fea_k"], F),
           "queue_frame_proj_ng": (k["frame_proj_k"], F)}
    for n, (x, mult) in src.items():
        x = x.reshape(-1, D).astype(np.float64)
        xn = (x / np.maximum(np.sqrt((x * x).sum(1, keepdims=True)), 1e-12)).T
        want = q0[n].copy()
        want[:, ptr * mult:(ptr + b) * mult] = xn
        got = getattr(m, n).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-7, err_msg=n)
        # untouched columns are bit-identical
        keep = np.ones(want.shape[1], bool)
        keep[ptr * mult:(ptr + b) * mult] = False
        assert np.array_equal(got[:, keep], q0[n][:, keep]), n
    for buf in m._queue_buffers():
        st = ops.queue_state(buf)
        kd, dk = [x.clone() for x in st.pack(buf, 2)[:2]]
        kd2, dk2 = st.repack(buf, 2)
        assert torch.equal(kd, kd2) and torch.equal(dk, dk2)


def test_ema_bit_exact(golden):
    g = golden("ema")
    ps, pks = syn.ema_tensors()

    class Mod(torch.nn.Module):
        def __init__(self, xs):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(cu(x), requires_grad=False) for x in xs])
    m = _model(16, 4, 64, "fp32")
    a, b = Mod(ps), Mod(pks)
    m.model_pairs = [[a, b]]
    for _ in range(int(g["steps"])):
        m._momentum_update()
    for i, pk in enumerate(b.ps):
        got = pk.detach().cpu().numpy()
        ref = g["out%d" % i]
        assert got.dtype == ref.dtype and np.array_equal(got.view(np.uint8), ref.view(np.uint8)), i
    m.copy_params()
    assert all(torch.equal(x, y) for x, y in zip(a.ps, b.ps))


def test_ema_large_matches_oracle():
    rs = np.random.RandomState(0)
    n = 3 * 8192 + 77
    p = rs.randn(n).astype(np.float32)
    pk = rs.randn(n).astype(np.float32)
    tp, tpk = cu(p), cu(pk)
    tab = ops.EmaTable([(tp[1:], tpk[1:])])       # misaligned views take the scalar path
    tab.run(0.99)
    ref = O.momentum_update([p[1:]], [pk[1:]], 0.99)[0]
    assert np.array_equal(tpk[1:].cpu().numpy().view(np.uint8), ref.view(np.uint8))


def test_unsupported_temperature_is_loud():
    m = _model(128, 2, 64, "fp32", T=0.001)
    q = cu(np.random.RandomState(0).randn(4, 64).astype(np.float32))
    with pytest.raises(Exception):
        m.contrastive_loss(q, q, m.queue_v_cross_ng)


def test_pretrain_head_small_fp32(golden):
    g = golden("pretrain_small")
    b, F, D, K, T = int(g["b"]), int(g["F"]), int(g["D"]), int(g["K"]), float(g["T"])
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    m = _model(K, F, D, "fp32", T)
    _load_queues(m, syn.queues(K, F=F, D=D, seed=3))
    t = {n: cu(x, n in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for n, x in inp.items()}
    loss = m.head_loss(t["v_fea"], t["frame_fea"], t["title_fea"], t["frame_pred"], t["v_fea_k"], t["frame_fea_k"],
                       t["title_fea_k"], t["tag_fea_k"], t["frame_proj_k"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-5
    for n in ("v_fea", "title_fea", "frame_fea", "frame_pred"):
        assert rel(t[n].grad.cpu().numpy(), g["d_" + n]) < 1e-4, n
    for n in syn.QUEUE_NAMES:
        np.testing.assert_allclose(getattr(m, n).cpu().numpy(), g["after_" + n], rtol=0, atol=2e-7)


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_fused_head_equals_granular_calls(prec):
    """The one-call head equals the reference-style composition of frame_self_loss /
    contrastive_loss / frame_cross_loss (modules/modeling.py:385-400), values and gradients;
    also without --use_frame_fea."""
    b, F, D, K = 24, 5, 128, 256
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=31)
    qs = syn.queues(K, F=F, D=D, seed=32)
    names = ("v_fea", "title_fea", "frame_fea", "frame_pred")
    for use_frames in (True, False):
        m = _model(K, F, D, prec)
        m.task_config.use_frame_fea = use_frames
        _load_queues(m, qs)
        a = {n: cu(x, n in names) for n, x in inp.items()}
        fam = m.frame_self_loss(a["frame_pred"], a["frame_proj_k"], m.queue_frame_proj_ng)
        vtm = m.contrastive_loss(a["v_fea"], a["title_fea_k"], m.queue_title_cross_ng) \
            + m.contrastive_loss(a["title_fea"], a["v_fea_k"], m.queue_v_cross_ng)
        ftm = m.frame_cross_loss(a["frame_fea"], a["frame_fea_k"], m.queue_frame_cross_ng, a["title_fea"],
                                 a["title_fea_k"], m.queue_title_cross_ng) if use_frames else 0.0
        l1 = 0.05 * fam + 0.45 * vtm + 0.45 * ftm
        l1.backward()
        c = {n: cu(x, n in names) for n, x in inp.items()}
        l2 = m.head_loss(c["v_fea"], c["frame_fea"], c["title_fea"], c["frame_pred"], c["v_fea_k"], c["frame_fea_k"],
                         c["title_fea_k"], c["tag_fea_k"], c["frame_proj_k"])
        l2.backward()
        assert abs(float(l1) - float(l2)) / float(l1) < 3e-6
        for n in names:
            if n == "frame_fea" and not use_frames:
                assert a[n].grad is None and float(c[n].grad.abs().max()) == 0.0
                continue
            assert rel(c[n].grad.cpu().numpy(), a[n].grad.cpu().numpy()) < 1e-5, n


def test_graphed_step_matches_eager():
    """Replaying the captured step (hmmc_b200/graphs.py) leaves the same queues, pointer, momentum
    parameters, loss and gradients as issuing it from Python."""
    from hmmc_b200.graphs import GraphedStep
    b, F, D, K = 16, 12, 128, 128
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=41)
    qs = syn.queues(K, F=F, D=D, seed=42)
    names = ("v_fea", "title_fea", "frame_fea", "frame_pred")
    order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
             "frame_proj_k"]

    class Enc(torch.nn.Module):
        def __init__(self, seed):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.w = torch.nn.Parameter(torch.randn(1000, generator=g).cuda(), requires_grad=False)

    def make():
        m = _model(K, F, D, "bf16x3")
        _load_queues(m, qs)
        m.model_pairs = [[Enc(1), Enc(2)]]
        t = {n: cu(x, n in names) for n, x in inp.items()}

        def step():
            for n in names:
                t[n].grad = None
            with torch.no_grad():
                m._momentum_update()
            loss = m.head_loss(*[t[n] for n in order])
            loss.backward()
            return loss
        return m, t, step

    m1, t1, step1 = make()
    m2, t2, step2 = make()
    n_warm, n_run = 3, 5          # GraphedStep runs 3 warm-up steps + 1 capture pass (capture does not execute)
    g = GraphedStep(step2, warmup=n_warm)
    for _ in range(n_run):
        l2 = g.replay()
    for _ in range(n_warm + n_run):
        l1 = step1()
    torch.cuda.synchronize()
    assert int(m1.queue_ptr) == int(m2.queue_ptr) == ((n_warm + n_run) * b) % K
    assert float(l1) == float(l2)
    for n in syn.QUEUE_NAMES:
        assert torch.equal(getattr(m1, n), getattr(m2, n)), n
    assert torch.equal(m1.model_pairs[0][1].w, m2.model_pairs[0][1].w)
    for n in names:
        assert torch.equal(t1[n].grad, t2[n].grad), n


@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_checkpoint_round_trip_rebuilds_operand_copies(prec, tmp_path):
    """N4: a state_dict written after some enqueues, loaded into a fresh model through init_preweight,
    gives the same loss / gradients / next enqueue bit for bit (the bf16 operand copies are re-derived)."""
    import io
    from hmmc_b200 import checkpoint as C
    K, F, D, b = 128, 4, 128, 16
    a = _model(K, F, D, prec)
    _load_queues(a, syn.queues(K, F=F, D=D, seed=3))
    for step in range(3):
        k = syn.pretrain_inputs(b, F=F, D=D, seed=40 + step)
        a._dequeue_and_enqueue(cu(k["v_fea_k"]), cu(k["tag_fea_k"]), cu(k["title_fea_k"]), cu(k["frame_fea_k"]),
                               cu(k["frame_proj_k"]))
    buf = io.BytesIO()
    torch.save(a.state_dict(), buf)
    buf.seek(0)
    sd = torch.load(buf, map_location="cpu")
    assert set(C.QUEUE_KEYS) <= set(sd) and sd["queue_frame_cross_ng"].shape == (D, K * F)
    bm = _model(K, F, D, prec)
    # touch the fresh model's operand copies first so that a stale pack would be noticed
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=50)
    order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
             "frame_proj_k"]

    def run(m):
        t = {n: cu(inp[n], grad=n in ("v_fea", "frame_fea", "title_fea", "frame_pred")) for n in order}
        loss = m.head_loss(*[t[n] for n in order])
        loss.backward()
        return loss.detach(), t["frame_fea"].grad
    run(bm)
    C.init_preweight(bm, sd)
    assert bm._hmmc_load_report == ([], [], [])
    assert int(bm.queue_ptr) == int(a.queue_ptr) == 48
    la, ga = run(a)
    lb, gb = run(bm)
    assert torch.equal(la, lb) and torch.equal(ga, gb)
    for n in syn.QUEUE_NAMES:
        assert torch.equal(getattr(a, n), getattr(bm, n)), n


@pytest.mark.parametrize("prec", ["bf16", "bf16x3", "fp32"])
def test_split_head_equals_fused_call(prec):
    """head_loss_begin (query side, on a side stream with a reduced GEMM grid) + head_loss_end (key side) give
    the loss, the gradients and the queues of the one-call head_loss bit for bit, also with the momentum
    update issued in between as the reference's forward does."""
    K, F, D, b = 128, 4, 128, 16
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=61)
    qs = syn.queues(K, F=F, D=D, seed=62)
    order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
             "frame_proj_k"]
    res = []
    for split in (False, True):
        m = _model(K, F, D, prec)
        _load_queues(m, qs)
        t = {n: cu(inp[n], grad=n in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for n in order}
        for _ in range(2):                       # two steps: the second one sees the enqueued keys
            for n in order[:4]:
                t[n].grad = None
            if split:
                begun = m.head_loss_begin(*[t[n] for n in order[:4]])
                junk = torch.randn(1 << 20, device="cuda").sum()       # unrelated work on the main stream
                loss = m.head_loss_end(begun, *[t[n] for n in order[4:]])
            else:
                loss = m.head_loss(*[t[n] for n in order])
            loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach().clone(), [t[n].grad.clone() for n in order[:4]],
                    [getattr(m, n).clone() for n in syn.QUEUE_NAMES], int(m.queue_ptr)))
    assert torch.equal(res[0][0], res[1][0])
    for a, c in zip(res[0][1], res[1][1]):
        assert torch.equal(a, c)
    for a, c in zip(res[0][2], res[1][2]):
        assert torch.equal(a, c)
    assert res[0][3] == res[1][3] == 32


def test_head_accepts_fp16_autocast_embeddings():
    """--enable_amp (main_pretrain.py:82,258): whatever the autocast region hands the head in fp16 is widened once,
    the arithmetic stays the fp32-parity path, and the gradients come back in the inputs' dtype.  Checked against the
    float64 oracle evaluated on the SAME fp16-rounded embeddings (1e-5 on the loss; the fp16 rounding of the returned
    gradients bounds their error at ~1e-3)."""
    b, F, D, K = 32, 12, 512, 1024
    inp = {n: x.astype(np.float16).astype(np.float32) for n, x in syn.pretrain_inputs(b, F=F, D=D, seed=2).items()}
    qs = syn.queues(K, F=F, D=D, seed=3)
    ref_loss, ref_g = O.pretrain_loss_and_grads(inp, qs, 0.07)
    m = _model(K, F, D, "bf16x3")
    _load_queues(m, qs)
    names = ("v_fea", "title_fea", "frame_fea", "frame_pred")
    t = {n: torch.from_numpy(x).cuda().half().requires_grad_(n in names) for n, x in inp.items()}
    with torch.autocast("cuda", dtype=torch.float16):
        loss = m.head_loss(t["v_fea"], t["frame_fea"], t["title_fea"], t["frame_pred"], t["v_fea_k"], t["frame_fea_k"],
                           t["title_fea_k"], t["tag_fea_k"], t["frame_proj_k"])
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    scaler.scale(loss).backward()
    assert loss.dtype == torch.float32 and abs(float(loss.detach()) - ref_loss) / ref_loss < 1e-5
    for n in names:
        assert t[n].grad.dtype == torch.float16
        assert rel(t[n].grad.float().cpu().numpy() / 1024.0, ref_g[n]) < 2e-3, n
