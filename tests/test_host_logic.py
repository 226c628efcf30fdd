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


def _tiny_pretrain_model():
    import types
    from hmmc_b200 import modeling
    task = types.SimpleNamespace(local_rank=0, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                 contrast_num_negative=8, max_frames=2, use_frame_fea=True, head_precision="fp32")
    return modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=16), task)


def test_checkpoint_keys_and_load_semantics(tmp_path):
    """state_dict carries the reference's queue buffers ([D,Kq] fp32 + queue_ptr) and nothing derived;
    init_preweight renames gamma/beta, applies the prefix, reports missing / unexpected keys and bumps
    the buffers' version counters (which is what triggers the re-pack of the bf16 operand copies)."""
    import types
    import torch
    from hmmc_b200 import checkpoint as C
    m = _tiny_pretrain_model()
    sd = m.state_dict()
    assert set(C.QUEUE_KEYS) <= set(sd.keys())
    assert all("pack" not in k for k in sd)
    assert sd["queue_frame_proj_ng"].shape == (16, 16) and sd["queue_ptr"].dtype == torch.long
    m.add_module("ln", torch.nn.LayerNorm(16))
    src = {k: torch.randn_like(v) if v.is_floating_point() else torch.full_like(v, 4) for k, v in sd.items()}
    src["ln.gamma"] = torch.full((16,), 2.0)
    src["ln.beta"] = torch.full((16,), -1.0)
    src["stray.weight"] = torch.zeros(1)
    before = m.queue_v_cross_ng._version
    C.init_preweight(m, dict(src))
    missing, unexpected, errors = m._hmmc_load_report
    assert torch.equal(m.ln.weight.data, src["ln.gamma"]) and torch.equal(m.ln.bias.data, src["ln.beta"])
    assert torch.equal(m.queue_v_cross_ng, src["queue_v_cross_ng"]) and int(m.queue_ptr) == 4
    assert m.queue_v_cross_ng._version > before
    assert missing == [] and errors == []
    # with a prefix every key moves under it: nothing matches a module without that prefix
    m2 = _tiny_pretrain_model()
    C.init_preweight(m2, dict(sd), prefix="module.")
    assert set(m2._hmmc_load_report[0]) >= set(C.QUEUE_KEYS)
    # save_model writes the reference's file name and torch.load reads it back
    args = types.SimpleNamespace(output_dir=str(tmp_path))
    f = C.save_model(3, args, m, type_name="pretrain")
    assert f.endswith("pytorch_model.bin.pretrain.3")
    m3 = C.load_head_state(_tiny_pretrain_model(), f)
    assert torch.equal(m3.queue_tag_cross_ng, m.queue_tag_cross_ng)


def test_init_preweight_matches_reference():
    import pytest
    import torch
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.load()
    from modules.until_module import PreTrainedModel
    from hmmc_b200 import checkpoint as C
    a, b = _tiny_pretrain_model(), _tiny_pretrain_model()
    for m in (a, b):
        m.add_module("ln", torch.nn.LayerNorm(16))
    sd = {k: torch.randn_like(v) if v.is_floating_point() else torch.full_like(v, 2) for k, v in a.state_dict().items()}
    sd["ln.gamma"] = sd.pop("ln.weight")
    sd["extra.thing"] = torch.zeros(2)
    del sd["queue_tag_cross_ng"]
    PreTrainedModel.init_preweight(a, dict(sd), task_config=None)
    C.init_preweight(b, dict(sd))
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and (ka == "queue_tag_cross_ng" or torch.equal(va, vb)), ka
    assert b._hmmc_load_report[0] == ["queue_tag_cross_ng"]


def test_cache_eval_features_filters_videos_at_cut_off_points():
    """eval_epoch step 1 (main_task_retrieval.py:391-440): in the multi-sentence layout a video is
    encoded once, at its last caption; features come back as one tensor per stream."""
    import types
    import torch
    from hmmc_b200 import synthetic as syn
    T, V, Fr, batches, cut = syn.eval_epoch_case(True, 32)
    Tt, Vt, Ft = torch.from_numpy(T), torch.from_numpy(V), torch.from_numpy(Fr)
    seen = []

    def visual(video, video_frame):
        seen.extend(int(i) for i in video[:, 0])
        return Vt[video[:, 0]], Ft[video[:, 0]]
    model = types.SimpleNamespace(text_encoder=lambda ids, mask: Tt[ids[:, 0]], visual_encoder=visual)
    dl = [(torch.from_numpy(ci)[:, None], torch.ones(len(ci), 1, dtype=torch.long), torch.from_numpy(vi)[:, None],
           torch.full((len(ci),), 12)) for ci, vi in batches]
    args = types.SimpleNamespace(task="retrieval")
    text, video, frames, title = retrieval.cache_eval_features(args, model, dl, torch.device("cpu"), True,
                                                               [c - 1 for c in cut])
    assert seen == list(range(V.shape[0])) and title is None
    assert torch.equal(text, Tt) and torch.equal(video, Vt) and torch.equal(frames, Ft)
    import pytest
    with pytest.raises(ValueError):
        retrieval.cache_eval_features(types.SimpleNamespace(task="caption"), model, dl, torch.device("cpu"))


def test_fine_tune_head_warns_when_the_batch_leaves_the_tensor_path():
    """batches the tensor-core tiling does not take must not be silently slow."""
    import warnings
    from hmmc_b200 import ops
    ops._warned_slow_symce.clear()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ops._warn_symce_path(256, 512, ops.PREC_BF16X3)      # tensor path: silent
        ops._warn_symce_path(64, 512, ops.PREC_BF16)
        ops._warn_symce_path(40, 512, ops.PREC_FP32)         # fp32 is the CUDA-core path by choice: silent
        assert not w
        ops._warn_symce_path(40, 512, ops.PREC_BF16X3)
        ops._warn_symce_path(40, 512, ops.PREC_BF16X3)       # once per shape
        ops._warn_symce_path(96, 512, ops.PREC_BF16)         # 96 % 64 != 0
        ops._warn_symce_path(64, 1536, ops.PREC_BF16X3)      # D > 1024
    assert len(w) == 3 and all(issubclass(x.category, RuntimeWarning) for x in w)
    assert "CUDA-core path" in str(w[0].message)
