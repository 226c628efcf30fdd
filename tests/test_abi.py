"""The C-ABI library loads and exports every symbol include/hmmc_head.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hmmc_head.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmmc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from hmmc_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    from hmmc_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_string(lib):
    lib.hmmc_last_error.restype = ctypes.c_char_p
    assert lib.hmmc_version() >= 100
    assert isinstance(lib.hmmc_last_error(), bytes)
    assert lib.hmmc_ema_block_elems() > 0


def test_product_path_has_no_cpu_fallback():
    """ops refuse CPU tensors loudly instead of falling back."""
    import torch
    from hmmc_b200 import ops
    from hmmc_b200._lib import HmmcError
    with pytest.raises(HmmcError):
        ops.cross_en(torch.zeros(4, 4))
    with pytest.raises(HmmcError):
        ops.sim_topk(torch.zeros(4, 64), torch.zeros(4, 64), torch.zeros(4, 2, 64), 100.0, 1)


def test_product_code_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hmmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_flags_match_reference_names():
    from hmmc_b200.flags import get_head_args
    a = get_head_args(["--top_frames", "2", "--contrast_num_negative", "1024", "--use_frame_fea"])
    assert a.top_frames == 2 and a.contrast_num_negative == 1024 and a.use_frame_fea
    assert a.contrast_momentum == 0.99 and a.contrast_temperature == 0.07 and a.max_frames == 12
