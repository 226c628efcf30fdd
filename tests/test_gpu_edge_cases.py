"""Edge cases and size-independent properties: degenerate shapes, maximum group sizes, BASELINE's
full config-4 size against the fp64 oracle, linearity in the upstream gradient, idempotence of the
derived state."""
import types

import numpy as np
import pytest
import torch

from hmmc_b200 import modeling, ops, retrieval
from hmmc_b200 import synthetic as syn
from oracle import head_oracle as O
from gpu_util import cu, rel

pytestmark = pytest.mark.gpu
QN = ("v_fea", "title_fea", "frame_fea", "frame_pred")
ORDER = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
         "frame_proj_k"]


def _pre(K, F, D, prec):
    task = types.SimpleNamespace(local_rank=0, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                 contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision=prec)
    return modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).cuda()


def _load(m, qs):
    with torch.no_grad():
        for n, x in qs.items():
            getattr(m, n).copy_(torch.from_numpy(x))


@pytest.mark.parametrize("b", [128, 256])
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_full_size_config4_against_fp64_oracle(prec, b):
    """BASELINE config 4 at its real size (b=128, F=12, D=512, K=1024) and the north-star shape (b=256):
    loss and all four gradients against the float64 oracle; tolerances 1e-5 / 5e-5 (fp32-parity mode)
    and 1e-3 (bf16)."""
    F, D, K = 12, 512, 1024
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = syn.queues(K, F=F, D=D, seed=3)
    ref_loss, ref_g = O.pretrain_loss_and_grads(inp, qs, 0.07)
    m = _pre(K, F, D, prec)
    _load(m, qs)
    t = {n: cu(inp[n], grad=n in QN) for n in ORDER}
    loss = m.head_loss(*[t[n] for n in ORDER])
    loss.backward()
    ltol, gtol = {"bf16x3": (1e-5, 5e-5), "bf16": (1e-3, 1e-3)}[prec]
    assert abs(float(loss.detach()) - ref_loss) / ref_loss < ltol
    for n in QN:
        assert rel(t[n].grad.cpu().numpy(), ref_g[n]) < gtol, n
    assert int(m.queue_ptr) == b


def test_backward_is_linear_in_the_upstream_gradient():
    b, F, D, K = 16, 4, 128, 128
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=5)
    qs = syn.queues(K, F=F, D=D, seed=6)
    grads = []
    for factor in (1.0, -2.5):
        m = _pre(K, F, D, "bf16x3")
        _load(m, qs)
        t = {n: cu(inp[n], grad=n in QN) for n in ORDER}
        (m.head_loss(*[t[n] for n in ORDER]) * factor).backward()
        grads.append([t[n].grad.clone() for n in QN])
    for a, c in zip(*grads):
        assert torch.equal(a * np.float32(-2.5), c)


def test_enqueue_replaces_the_whole_queue_when_batch_equals_K():
    K, F, D = 16, 3, 64
    m = _pre(K, F, D, "bf16x3")
    _load(m, syn.queues(K, F=F, D=D, seed=3))
    k = syn.pretrain_inputs(K, F=F, D=D, seed=9)
    m._dequeue_and_enqueue(cu(k["v_fea_k"]), cu(k["tag_fea_k"]), cu(k["title_fea_k"]), cu(k["frame_fea_k"]),
                           cu(k["frame_proj_k"]))
    assert int(m.queue_ptr) == 0
    x = k["frame_fea_k"].reshape(-1, D).astype(np.float64)
    want = (x / np.sqrt((x * x).sum(1, keepdims=True))).T
    np.testing.assert_allclose(m.queue_frame_cross_ng.cpu().numpy(), want, rtol=0, atol=2e-7)


def test_ema_limits_are_exact():
    """m = 1 leaves the momentum parameters untouched, m = 0 copies the online ones (both bit for bit)."""
    ps, pks = syn.ema_tensors()
    for mom in (1.0, 0.0):
        a = [torch.nn.Parameter(cu(x), requires_grad=False) for x in ps]
        bb = [torch.nn.Parameter(cu(x), requires_grad=False) for x in pks]
        ops.EmaTable(list(zip(a, bb))).run(mom)
        for i in range(len(ps)):
            want = pks[i] if mom == 1.0 else ps[i]
            assert np.array_equal(bb[i].detach().cpu().numpy(), want), (mom, i)
            assert np.array_equal(a[i].detach().cpu().numpy(), ps[i])


def test_derived_queue_copies_are_idempotent():
    m = _pre(32, 4, 64, "bf16x3")
    buf = m.queue_frame_proj_ng
    st = ops.queue_state(buf)
    kd, dk = [x.clone() for x in st.pack(buf, 2)[:2]]
    st.repack(buf, 2)
    kd2, dk2 = st.repack(buf, 2)
    assert torch.equal(kd, kd2) and torch.equal(dk, dk2)


def test_rank_count_known_answers():
    n = 257
    eye = torch.eye(n, device="cuda") * 3.0 + torch.rand(n, n, device="cuda")
    t2v, v2t = ops.rank_count(eye)
    assert int(t2v.abs().sum()) == 0 and int(v2t.abs().sum()) == 0
    # row i: strictly decreasing in (j - i) mod n  ->  the ground truth column i is the best of its row,
    # column i of the matrix sees row i ranked first as well
    j = torch.arange(n, device="cuda")
    circ = -(((j[None, :] - j[:, None]) % n).float())
    t2v, v2t = ops.rank_count(circ)
    assert int(t2v.abs().sum()) == 0
    # reversed: ground truth is the worst entry of every row
    t2v, _ = ops.rank_count(-circ)
    assert bool((t2v == n - 1).all())
    one, _ = ops.rank_count(torch.ones(1, 1, device="cuda"))
    assert one.tolist() == [0]


def test_eval_degenerate_shapes():
    task = types.SimpleNamespace(local_rank=0, top_frames=2, use_frame_fea=True, head_precision="bf16x3")
    m = modeling.BirdModel(modeling.default_cross_config(), task)
    T, V, Fr, gt, _ = syn.eval_inputs(1, 1, seed=3)
    sim = retrieval.similarity_matrix(m, cu(T), cu(V), cu(Fr))
    ref = O.eval_scores(T, V, Fr, 2)
    assert sim.shape == (1, 1) and abs(float(sim) - float(ref)) < 6e-4
    # 3 texts x 130 videos: neither a tile multiple nor square
    per = np.zeros(130, dtype=np.int64)
    per[5], per[77] = 1, 2
    T, V, Fr, gt, _ = syn.eval_inputs(3, 130, seed=4, per_video=per)
    sim = retrieval.similarity_matrix(m, cu(T), cu(V), cu(Fr)).cpu().numpy()
    assert np.abs(sim - O.eval_scores(T, V, Fr, 2)).max() < 6e-4


def test_fused_eval_group_limits():
    """Videos without captions, a video with the maximum 128 captions, and the loud failure above it."""
    per = np.array([0, 128, 1, 0, 5, 0], dtype=np.int64)
    Nt, Nv = int(per.sum()), per.size
    T, V, Fr, gtv, cut = syn.eval_inputs(Nt, Nv, seed=8, per_video=per)
    t2v, v2t = retrieval.fused_eval_ranks(cu(T), cu(V), cu(Fr), per, 100.0, 2, "bf16x3")
    ref = O.eval_scores(T, V, Fr, 2)
    want_t = (ref > ref[np.arange(Nt), gtv][:, None]).sum(1)
    assert int((t2v.cpu().numpy() != want_t).sum()) <= 1          # exact near-ties only
    assert t2v.shape == (Nt,) and v2t.shape == (Nv,)
    with pytest.raises(ValueError):
        retrieval.pack_caption_groups(np.array([129]))


def test_finetune_single_sample_and_frames_only():
    """B = 1 (every softmax is over one entry: loss 0, zero gradients) and frame_loss alone."""
    task = types.SimpleNamespace(local_rank=0, top_frames=2, use_frame_fea=True, head_precision="bf16x3")
    m = modeling.BirdModel(modeling.default_cross_config(), task)
    t, v, fr = syn.finetune_inputs(1, seed=2)
    a = [cu(t, True), cu(v, True), cu(fr, True)]
    loss = m.head_loss(*a)
    loss.backward()
    assert abs(float(loss.detach())) < 1e-6 and float(a[0].grad.abs().max()) < 1e-6
    t, v, fr = syn.finetune_inputs(64, seed=3)
    tt, tf = cu(t, True), cu(fr, True)
    fl = m.frame_loss(tt, tf)
    fl.backward()
    ref = O.frame_loss(t, fr, dtype=np.float64)
    assert abs(float(fl.detach()) - float(ref)) / abs(float(ref)) < 1e-5
    assert tt.grad is not None and tf.grad is not None and bool(torch.isfinite(tf.grad).all())


def test_reserved_sms_do_not_change_results():
    """hmmc_head_schedule.reserved_sms only shrinks the persistent GEMM grids of that call: same units, same
    bits (and no library state is left behind: the plain call afterwards matches too)."""
    b, F, D, K = 32, 12, 512, 1024
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = {n: cu(x) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
    qb = [qs[n] for n in ("queue_v_cross_ng", "queue_title_cross_ng", "queue_frame_proj_ng", "queue_frame_cross_ng")]
    outs = []
    for reserved in (0, 40, 140, None):
        t = {n: cu(inp[n], grad=n in QN) for n in ORDER}
        q4 = [t["v_fea"], t["title_fea"], t["frame_fea"], t["frame_pred"]]
        k4 = [t["v_fea_k"], t["title_fea_k"], t["frame_fea_k"], t["frame_proj_k"]]
        if reserved is None:
            loss, _ = ops.pretrain_head(*q4, *k4, *qb, 0.07, 0.05, 0.45, 0.45, True, "bf16")
        else:
            st = ops.pretrain_head_begin(*q4, *qb, 0.07, 0.05, 0.45, 0.45, True, "bf16", reserved_sms=reserved)
            loss, _ = ops.pretrain_head_end(st, *q4, *k4)
        loss.backward()
        outs.append((loss.detach().clone(), t["frame_pred"].grad.clone(), t["title_fea"].grad.clone()))
    for l, g, g2 in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(g, outs[0][1]) and torch.equal(g2, outs[0][2])
