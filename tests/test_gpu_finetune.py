import types

import numpy as np
import pytest
import torch

from hmmc_b200 import modeling, ops
from hmmc_b200 import synthetic as syn
from oracle import head_oracle as O
from gpu_util import TOL, cu, rel

pytestmark = pytest.mark.gpu


def _model(prec, top_frames=2):
    task = types.SimpleNamespace(local_rank=0, top_frames=top_frames, use_frame_fea=True, head_precision=prec)
    return modeling.BirdModel(modeling.default_cross_config(), task)


@pytest.mark.parametrize("prec", ["fp32"])
def test_loose_similarity_and_cross_en_golden(golden, prec):
    g = golden("similarity")
    m = _model(prec)
    s2 = m.loose_similarity(cu(g["q"]), cu(g["v"]))
    s3 = m.loose_similarity(cu(g["q"]), cu(g["fr"]))
    assert tuple(s2.shape) == g["s2"].shape and tuple(s3.shape) == g["s3"].shape
    np.testing.assert_allclose(s2.cpu().numpy(), g["s2"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(s3.cpu().numpy(), g["s3"], rtol=0, atol=3e-5)
    ce = m.loss_fct(cu(g["s2"][:7, :7].copy()))
    assert abs(float(ce) - float(g["ce"])) < 1e-5


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_loose_similarity_tensor_core(prec):
    rs = np.random.RandomState(3)
    q = rs.randn(130, 512).astype(np.float32)
    fr = rs.randn(40, 12, 512).astype(np.float32)
    m = _model(prec)
    s3 = m.loose_similarity(cu(q), cu(fr)).cpu().numpy()
    ref = O.loose_similarity(q, fr, dtype=np.float64)
    assert s3.shape == ref.shape == (130, 40, 12)
    assert np.abs(s3 - ref).max() < (2e-4 if prec == "bf16x3" else 0.12)


@pytest.mark.parametrize("B", [32, 256])
@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "bf16"])
def test_finetune_head(golden, B, prec):
    g = golden("finetune_B%d" % B)
    t, v, fr = syn.finetune_inputs(B, seed=int(g["seed"]))
    m = _model(prec)
    tt, tv, tf = cu(t, True), cu(v, True), cu(fr, True)
    loss = m.head_loss(tt, tv, tf)
    loss.backward()
    ltol, gtol = TOL[prec]
    if prec == "bf16" and B % 64 == 0:
        gtol = 2e-2      # logits carry scale 100: a single bf16 plane moves the softmax weights by ~1 %
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < ltol
    _, dt, dv, dfr = O.finetune_loss_and_grads(t, v, fr)
    assert rel(tt.grad.cpu().numpy(), dt) < gtol
    assert rel(tv.grad.cpu().numpy(), dv) < gtol
    assert rel(tf.grad.cpu().numpy(), dfr) < gtol
    gd = g["dt"]
    assert rel(tt.grad.cpu().numpy()[:gd.shape[0]], gd) < max(1e-4, gtol)


def test_granular_path_equals_fused():
    """frame_loss + loose_similarity + loss_fct composed as BirdModel.forward does
    (modules/modeling.py:702-709) agree with the fused head, values and gradients."""
    B = 48
    t, v, fr = syn.finetune_inputs(B, seed=5)
    m = _model("fp32")
    a = [cu(t, True), cu(v, True), cu(fr, True)]
    loss = 0.15 * m.frame_loss(a[0], a[2])
    sim = m.loose_similarity(a[0], a[1])
    loss = loss + 0.85 * (m.loss_fct(sim) + m.loss_fct(sim.T))
    loss.backward()
    b = [cu(t, True), cu(v, True), cu(fr, True)]
    fused = m.head_loss(*b)
    fused.backward()
    assert abs(float(loss) - float(fused)) / float(fused) < 2e-6
    for x, y in zip(a, b):
        assert rel(x.grad.cpu().numpy(), y.grad.cpu().numpy()) < 1e-5


def test_single_row_and_no_grad():
    m = _model("fp32")
    t, v, fr = syn.finetune_inputs(2, seed=9)
    with torch.no_grad():
        l = m.head_loss(cu(t), cu(v), cu(fr))
    assert abs(float(l) - float(O.finetune_loss(t, v, fr))) < 1e-4


@pytest.mark.parametrize("B,prec", [(256, "bf16x3"), (64, "bf16"), (32, "bf16x3"), (96, "bf16x3"), (32, "bf16"), (32, "fp32")])
def test_packed_layout_equals_separate_tensors(B, prec):
    """hmmc_sym_ce_packed_fwd_bwd (rows [text | video | frames], the all-gather's layout) gives the
    loss and gradients of hmmc_sym_ce_fwd_bwd on the three separate tensors, bit for bit."""
    F, D = 12, 512
    t, v, fr = [cu(x) for x in syn.finetune_inputs(B, seed=7)]
    p = ops.resolve_precision(prec)
    loss, dt, dv, df = ops.sym_ce_raw(t, v, fr, 100.0, 0.85, 0.15, p, True)
    packed = torch.cat([t, v, fr.reshape(B, F * D)], dim=1).contiguous()
    loss_p, dp = ops.sym_ce_packed_raw(packed, F, D, 100.0, 0.85, 0.15, p, True)
    assert torch.equal(loss, loss_p)
    assert torch.equal(dp[:, :D], dt) and torch.equal(dp[:, D:2 * D], dv)
    assert torch.equal(dp[:, 2 * D:].reshape(B, F, D), df)
    # forward only
    loss_f, none = ops.sym_ce_packed_raw(packed, F, D, 100.0, 0.85, 0.15, p, False)
    assert none is None and torch.equal(loss_f, loss)
