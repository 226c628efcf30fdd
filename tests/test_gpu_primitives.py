import numpy as np
import pytest
import torch

from hmmc_b200 import ops
from gpu_util import cu, rel

pytestmark = pytest.mark.gpu


def _bf16_split(x):
    t = torch.from_numpy(x)
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi.float().numpy(), lo.float().numpy()


def test_device_is_b200():
    ops.device_check()
    assert torch.cuda.get_device_capability()[0] == 10


@pytest.mark.parametrize("eps", [1e-12, 0.0])
def test_rownorm_pack(eps):
    rs = np.random.RandomState(0)
    x = rs.randn(37, 192).astype(np.float32)
    x[5] = 0.0
    xhat, inv, packed = ops.rownorm_pack(cu(x), eps, 2, want_xhat=True)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True))
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = x / (np.maximum(n, eps) if eps > 0 else n)
    got = xhat.cpu().numpy()
    keep = np.arange(37) != 5
    np.testing.assert_allclose(got[keep], ref[keep], rtol=0, atol=2e-7)
    if eps > 0:
        assert (got[5] == 0).all()
    else:
        assert np.isnan(got[5]).all()          # 0/0 like the reference's loose_similarity
    hi, lo = _bf16_split(got[keep])
    p = packed.float().cpu().numpy()
    assert np.array_equal(p[keep][:, :192], hi) and np.array_equal(p[keep][:, 192:], lo)


def test_gemm_f32_strided():
    rs = np.random.RandomState(1)
    A = rs.randn(70, 130).astype(np.float32)
    B = rs.randn(45, 130).astype(np.float32)
    C = ops.gemm_f32(cu(A), cu(B), 0.5).cpu().numpy()
    assert rel(C, 0.5 * A.astype(np.float64) @ B.T.astype(np.float64)) < 1e-6
    # transposed views (non-unit inner stride)
    At = cu(np.ascontiguousarray(A.T)).t()
    Bt = cu(np.ascontiguousarray(B.T)).t()
    C2 = ops.gemm_f32(At, Bt, 1.0).cpu().numpy()
    assert rel(C2, A.astype(np.float64) @ B.T.astype(np.float64)) < 1e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (300, 200, 512), (1, 16, 64),
                                   (257, 1040, 192), (1536, 512, 1024)])
@pytest.mark.parametrize("planes", [1, 2])
def test_umma_gemm_nt(M, N, K, planes):
    """tcgen05 GEMM against an exact product of the bf16-rounded planes."""
    rs = np.random.RandomState(M + N + K)
    A = (rs.randn(M, K) / np.sqrt(K)).astype(np.float32)
    B = (rs.randn(N, K) / np.sqrt(K)).astype(np.float32)
    ah, al = _bf16_split(A)
    bh, bl = _bf16_split(B)
    Ap = torch.from_numpy(np.concatenate([ah, al], 1)[:, :planes * K]).to(torch.bfloat16).cuda().contiguous()
    Bp = torch.from_numpy(np.concatenate([bh, bl], 1)[:, :planes * K]).to(torch.bfloat16).cuda().contiguous()
    C = ops.umma_gemm_nt(Ap, Bp, K, planes, 2.0).cpu().numpy()
    f = lambda x: x.astype(np.float64)
    if planes == 1:
        ref = 2.0 * f(ah) @ f(bh).T
    else:
        ref = 2.0 * (f(ah) @ f(bh).T + f(ah) @ f(bl).T + f(al) @ f(bh).T)
    assert C.shape == (M, N)
    assert rel(C, ref) < 1e-5, rel(C, ref)   # tcgen05 accumulates in fp32 with truncation: ~5e-6 at K~1k
    if planes == 2:      # and the split product is fp32-grade w.r.t. the unsplit operands
        assert rel(C, 2.0 * f(A) @ f(B).T) < 2e-5


def test_pack_rows_autograd_is_cat_with_slice_gradients():
    """ops.pack_rows_autograd (send buffer of the fine-tune gather): forward == torch.cat, backward == the slices."""
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(24, 128, device="cuda", generator=g, requires_grad=True)
    b = torch.randn(24, 128, device="cuda", generator=g, requires_grad=True)
    c = torch.randn(24, 12, 128, device="cuda", generator=g, requires_grad=True)
    packed = ops.pack_rows_autograd([a, b, c.reshape(24, -1)])
    want = torch.cat([a, b, c.reshape(24, -1)], dim=1)
    assert torch.equal(packed, want)
    up = torch.randn_like(packed)
    packed.backward(up)
    assert torch.equal(a.grad, up[:, :128]) and torch.equal(b.grad, up[:, 128:256])
    assert torch.equal(c.grad, up[:, 256:].reshape(24, 12, 128))


@pytest.mark.parametrize("D", [128, 512])
def test_pack_rows_normalised_matches_the_enqueue_arithmetic(D):
    """pack_rows(norm_dim=D) (keys normalised at the source, before the exchange): every D-vector equals
    x * (1 / max(||x||, 1e-12)) with the sum of squares taken lane-strided then by the shuffle tree -- compared with
    F.normalize to 2 ulp, zero vectors stay zero, and the staged mark is set."""
    g = torch.Generator(device="cuda").manual_seed(5)
    v = torch.randn(16, D, device="cuda", generator=g)
    f = torch.randn(16, 12, D, device="cuda", generator=g) * 3.0
    v[3].zero_()
    staged = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = ops.pack_rows([v, f], staged=staged, norm_dim=D)
    assert int(staged) == 1 and out.shape == (16, 13 * D)
    want = torch.cat([torch.nn.functional.normalize(v, dim=-1, eps=1e-12),
                      torch.nn.functional.normalize(f, dim=-1, eps=1e-12).reshape(16, -1)], dim=1)
    assert float((out - want).abs().max()) <= 2.5e-7
    assert float(out[3, :D].abs().max()) == 0.0
    norms = out.reshape(16, 13, D).norm(dim=-1)
    norms[3, 0] = 1.0
    assert float((norms - 1).abs().max()) < 1e-6
