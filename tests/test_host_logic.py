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
