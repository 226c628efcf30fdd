"""Host-side logic that needs no GPU."""
import numpy as np

from hmmc_b200 import retrieval


def test_pack_caption_groups_never_straddles():
    rs = np.random.RandomState(0)
    per = rs.randint(1, 40, size=500)
    src, grp, starts = retrieval.pack_caption_groups(per)
    assert src.size % 128 == 0 and (np.sort(src[src >= 0]) == np.arange(per.sum())).all()
    for j in range(500):
        s, c = int(starts[j]), int(per[j])
        assert s // 128 == (s + c - 1) // 128 and (grp[s:s + c] == j).all()
    src2, grp2, st2 = retrieval.pack_caption_groups(np.full(30, 10))
    assert (st2[:13] == np.array([0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 110, 128])).all()


def test_bert_adam_mirror_validation_and_schedules():
    """Constructor checks, schedule functions and get_lr of hmmc_b200.optimization follow
    modules/optimization.py:26-101 (no GPU needed; step() itself has no CPU path)."""
    import math
    import os
    import sys
    import pytest
    import torch
    from hmmc_b200 import optimization as opt
    from hmmc_b200.ops import HmmcError
    p = torch.nn.Parameter(torch.zeros(4))
    for bad in (dict(lr=-1.0), dict(lr=1e-3, schedule="nope"), dict(lr=1e-3, warmup=1.5), dict(lr=1e-3, b1=1.0),
                dict(lr=1e-3, b2=-0.1), dict(lr=1e-3, e=-1e-6)):
        with pytest.raises(ValueError):
            opt.BertAdam([p], **bad)
    o = opt.BertAdam([p], lr=1e-3, warmup=0.1, t_total=100, schedule="warmup_cosine")
    assert o.defaults["b2"] == 0.999 and o.defaults["max_grad_norm"] == 1.0 and o.defaults["weight_decay"] == 0.01
    assert o.get_lr() == []
    p.grad = torch.ones(4)
    assert o.get_lr() == [0]
    with pytest.raises(HmmcError):
        o.step()
    assert opt.warmup_cosine(0.05, 0.1) == 0.5 and opt.warmup_cosine(0.5, 0.1) == 0.5 * (1.0 + math.cos(math.pi * 0.5))
    assert opt.warmup_constant(0.5, 0.1) == 1.0 and opt.warmup_linear(1.5, 0.1) == 0
    if os.path.isdir("/root/reference/modules"):
        sys.path.insert(0, "/root/reference")
        try:
            import modules.optimization as ref
        finally:
            sys.path.pop(0)
        for name in ("warmup_cosine", "warmup_constant", "warmup_linear"):
            for x in (0.0, 0.001, 0.05, 0.1, 0.37, 0.99, 1.0, 1.2):
                for w in (0.002, 0.1, 0.25):
                    assert getattr(opt, name)(x, w) == getattr(ref, name)(x, w)
