"""Parity of the tensor-core MLP (modules/modeling.py:788-807) against the float64
oracle and the golden vectors of the reference's own class."""
import numpy as np
import pytest
import torch

from hmmc_b200 import synthetic as syn
from hmmc_b200.mlp import MLP
from oracle import mlp_oracle as MO
from gpu_util import cu, rel

pytestmark = pytest.mark.gpu
# (output, gradient) relative-L2 tolerances: fp32-parity split / single bf16 plane
TOL = {"bf16x3": (1e-5, 5e-5), "bf16": (1e-2, 3e-2)}


def _build(c, prec, Din, Dh, Dout):
    m = MLP(in_dim=Din, inner_dim=Dh, out_dim=Dout, num_layers=2, precision=prec).cuda()
    lin1, bn = m.linear_hidden[1], m.linear_hidden[2]
    with torch.no_grad():
        lin1.weight.copy_(cu(c["W1"])); lin1.bias.copy_(cu(c["b1"]))
        bn.weight.copy_(cu(c["gamma"])); bn.bias.copy_(cu(c["beta"]))
        bn.running_mean.copy_(cu(c["rm"])); bn.running_var.copy_(cu(c["rv"]))
        m.linear_out.weight.copy_(cu(c["W2"])); m.linear_out.bias.copy_(cu(c["b2"]))
    return m, lin1, bn


def _grads(m, lin1, bn, x):
    return dict(dx=x.grad, dW1=lin1.weight.grad, db1=lin1.bias.grad, dgamma=bn.weight.grad, dbeta=bn.bias.grad,
                dW2=m.linear_out.weight.grad, db2=m.linear_out.bias.grad)


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_mlp_small_vs_golden_and_oracle(golden, prec):
    g = golden("mlp")
    c = syn.mlp_case()
    m, lin1, bn = _build(c, prec, 64, 128, 64)
    m.train()
    x = cu(c["x"], True)
    y = m(x)
    y.backward(cu(c["dy"]))
    otol, gtol = TOL[prec]
    assert rel(y.detach().cpu().numpy(), g["y"]) < otol
    got = _grads(m, lin1, bn, x)
    y64, cache = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    ref = MO.backward(c["dy"], cache, c["W1"], c["gamma"], c["W2"])
    for k in ("dx", "dW1", "dgamma", "dbeta", "dW2", "db2"):
        assert rel(got[k].cpu().numpy(), ref[k]) < gtol, k
        assert rel(got[k].cpu().numpy(), g[k]) < max(gtol, 1e-5), k
    assert float(got["db1"].abs().max()) < 1e-5
    assert rel(bn.running_mean.cpu().numpy(), g["rm"]) < max(otol, 1e-5)
    assert rel(bn.running_var.cpu().numpy(), g["rv"]) < max(otol, 1e-5)
    assert int(bn.num_batches_tracked) == int(g["nbt"])
    m.eval()
    with torch.no_grad():
        ye = m(cu(c["x"]))
    assert rel(ye.cpu().numpy(), g["y_eval"]) < otol


def test_mlp_reference_size_vs_oracle():
    """The size the pre-train step runs: 128 samples x 12 frames = 1536 rows, 512 -> 4096 -> 512."""
    c = syn.mlp_case(M=1536, Din=512, Dh=4096, Dout=512, seed=33)
    m, lin1, bn = _build(c, "bf16x3", 512, 4096, 512)
    m.train()
    x = cu(c["x"], True)
    y = m(x)
    y.backward(cu(c["dy"]))
    y64, cache = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    ref = MO.backward(c["dy"], cache, c["W1"], c["gamma"], c["W2"])
    assert rel(y.detach().cpu().numpy(), y64) < 1e-5
    got = _grads(m, lin1, bn, x)
    # 6.3 M pre-activations: a few dozen lie within the fp32-level error of the GEMM around zero and may
    # take the other ReLU branch (as they would between any two fp32 implementations); the bound of what
    # those elements can move is granted on top of the 5e-5 relative tolerance
    slack = MO.relu_flip_slack(cache, ref, c["W1"], c["gamma"], width=2e-5)
    assert slack["count"] < 400
    for k in ("dx", "dW1", "dgamma", "dbeta", "dW2", "db2"):
        err = np.linalg.norm(got[k].cpu().numpy().astype(np.float64) - ref[k])
        assert err < 5e-5 * np.linalg.norm(ref[k]) + slack.get(k, 0.0), (k, err, slack)
    rm, rv = MO.running_stats(cache, c["rm"], c["rv"])
    assert rel(bn.running_mean.cpu().numpy(), rm) < 1e-5 and rel(bn.running_var.cpu().numpy(), rv) < 1e-5


def test_mlp_no_grad_training_mode_updates_running_stats():
    """The key-side projector runs under no_grad in training mode (modules/modeling.py:369-377)."""
    c = syn.mlp_case()
    m, lin1, bn = _build(c, "bf16x3", 64, 128, 64)
    m.train()
    with torch.no_grad():
        y = m(cu(c["x"]))
    y64, cache = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    assert rel(y.cpu().numpy(), y64) < 1e-5 and not y.requires_grad
    rm, _ = MO.running_stats(cache, c["rm"], c["rv"])
    assert rel(bn.running_mean.cpu().numpy(), rm) < 1e-5


def test_mlp_rows_not_multiple_of_64_forward_only_and_loud_backward():
    from hmmc_b200.ops import HmmcError
    c = syn.mlp_case(M=40)
    m, lin1, bn = _build(c, "bf16x3", 64, 128, 64)
    m.train()
    with torch.no_grad():
        y = m(cu(c["x"]))
    y64, _ = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    assert rel(y.cpu().numpy(), y64) < 1e-5
    with pytest.raises(HmmcError):
        m(cu(c["x"], True))


@pytest.mark.parametrize("use_temp", [True, False])
def test_visual_tail_vs_reference(golden, use_temp):
    """Residual + per-frame L2-normalise + mean over frames (modules/module_cross.py:207-213) against the
    reference's own VisualEncoder.forward (stub encoders) and the float64 oracle."""
    from hmmc_b200.mlp import visual_tail
    g = golden("visual_tail")
    orig = cu(g["orig"].reshape(6, 12, 64), True)
    temp = cu(np.ascontiguousarray(g["temp"].transpose(1, 0, 2)), True) if use_temp else None
    out, frame_output = visual_tail(temp, orig)
    out.backward(cu(g["g"]))
    tag = "temp" if use_temp else "notemp"
    assert frame_output is orig
    assert rel(out.detach().cpu().numpy(), g["out_" + tag]) < 1e-6
    assert rel(orig.grad.cpu().numpy().reshape(72, 64), g["dorig_" + tag]) < 2e-6
    if use_temp:
        assert rel(temp.grad.cpu().numpy().transpose(1, 0, 2), g["dtemp"]) < 2e-6
    # the pre-train size: 128 x 12 x 512
    rs = np.random.RandomState(5)
    o, t = rs.randn(128, 12, 512).astype(np.float32), rs.randn(128, 12, 512).astype(np.float32)
    gg = rs.randn(128, 512).astype(np.float32)
    oo, tt = cu(o, True), (cu(t, True) if use_temp else None)
    out, _ = visual_tail(tt, oo)
    out.backward(cu(gg))
    ref, h, n = MO.visual_tail(t if use_temp else None, o)
    assert rel(out.detach().cpu().numpy(), ref) < 1e-6
    assert rel(oo.grad.cpu().numpy(), MO.visual_tail_backward(gg, h, n)) < 2e-6
