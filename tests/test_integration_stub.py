"""The reference-side stub of INTEGRATION.md applied to the reference's own classes (only where
the reference tree exists: the build container)."""
import inspect

import pytest

from oracle import ref_shim


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_patch_reference_classes():
    M, _, _ = ref_shim.load()
    import hmmc_b200.modeling as B
    names = ["loose_similarity", "contrastive_loss", "frame_self_loss", "frame_cross_loss", "_momentum_update",
             "copy_params", "_dequeue_and_enqueue", "frame_loss"]
    # same positional signatures as the reference's methods
    for n in names:
        ref = getattr(M.BirdPreTrainedModel, n, None) or getattr(M.BirdModel, n)
        ours = getattr(B.ContrastiveHeadMixin, n)
        assert list(inspect.signature(ref).parameters) == list(inspect.signature(ours).parameters), n
    saved = {c: dict(vars(c)) for c in (M.BirdPreTrainedModel, M.BirdModel)}
    try:
        B.patch_reference_classes(M.BirdPreTrainedModel, M.BirdModel)
        for n in names:
            assert getattr(M.BirdModel, n) is getattr(B.ContrastiveHeadMixin, n)
        assert M.BirdPreTrainedModel.head_loss is B.BirdPreTrainedModel.head_loss
        assert M.BirdPreTrainedModel.head_loss_begin is B.BirdPreTrainedModel.head_loss_begin
        assert M.BirdPreTrainedModel.head_loss_end is B.BirdPreTrainedModel.head_loss_end
    finally:
        for c, d in saved.items():
            for k in list(vars(c)):
                if k not in d:
                    delattr(c, k)
            for k, v in d.items():
                if k not in ("__dict__", "__weakref__", "__doc__", "__module__"):
                    try:
                        setattr(c, k, v)
                    except (AttributeError, TypeError):
                        pass


def test_reference_metric_signatures():
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    _, metrics, R = ref_shim.load()
    import hmmc_b200.metrics as GM
    import hmmc_b200.retrieval as GR
    for n in ("compute_metrics", "tensor_text_to_video_metrics", "tensor_video_to_text_sim", "logging_rank"):
        assert list(inspect.signature(getattr(metrics, n)).parameters) == \
            list(inspect.signature(getattr(GM, n)).parameters), n
    assert list(inspect.signature(R._run_on_single_gpu).parameters) == \
        list(inspect.signature(GR._run_on_single_gpu).parameters)
    ref_ev = list(inspect.signature(R.eval_epoch).parameters)
    assert list(inspect.signature(GR.eval_epoch).parameters)[:len(ref_ev)] == ref_ev
    import modules.optimization as RO
    import hmmc_b200.optimization as GO
    assert list(inspect.signature(RO.BertAdam.__init__).parameters) == list(inspect.signature(GO.BertAdam.__init__).parameters)
    assert set(RO.SCHEDULES) == set(GO.SCHEDULES)
