"""Checkpoint I/O boundary of the head state.

The reference saves `model.state_dict()` with torch.save (main_pretrain.py:212-219) and loads it
through `PreTrainedModel.init_preweight` (modules/until_module.py:104-160): `gamma`/`beta` keys are
renamed to `weight`/`bias`, an optional prefix is prepended, then every sub-module pulls its own
entries with `_load_from_state_dict`.  The head's state in such a file is the five `queue_*_ng`
buffers in the reference's `[D, Kq]` fp32 layout plus `queue_ptr` — exactly the tensors this package
keeps authoritative, so files travel in both directions unchanged.  The K-major bf16 operand copies
are derived state: they are never written, and a load bumps the buffers' version counters so the
next head call re-packs them (hmmc_b200/ops.py `queue_state`).
"""
import os

import torch

QUEUE_KEYS = ("queue_v_cross_ng", "queue_frame_proj_ng", "queue_frame_cross_ng", "queue_title_cross_ng",
              "queue_tag_cross_ng", "queue_ptr")


def init_preweight(model, state_dict, prefix=None, task_config=None, logger=None):
    """modules/until_module.py:104-160 for any nn.Module: returns `model` after loading; attaches the
    lists the reference only logs as `model._hmmc_load_report = (missing, unexpected, errors)`."""
    renamed = {}
    for key, value in state_dict.items():
        new_key = key
        if 'gamma' in new_key:
            new_key = new_key.replace('gamma', 'weight')
        if 'beta' in new_key:
            new_key = new_key.replace('beta', 'bias')
        renamed[new_key] = value
    if prefix is not None:
        renamed = {prefix + k: v for k, v in renamed.items()}
    missing, unexpected, errors = [], [], []
    metadata = getattr(state_dict, '_metadata', None)

    def load(module, pfx=''):
        local_metadata = {} if metadata is None else metadata.get(pfx[:-1], {})
        module._load_from_state_dict(renamed, pfx, local_metadata, True, missing, unexpected, errors)
        for name, child in module._modules.items():
            if child is not None:
                load(child, pfx + name + '.')

    load(model)
    model._hmmc_load_report = (missing, unexpected, errors)
    if logger is not None and prefix is None and (task_config is None or getattr(task_config, "local_rank", 0) == 0):
        if missing:
            logger.info("Weights of %s not initialized from pretrained model: %s", model.__class__.__name__, missing)
        if unexpected:
            logger.info("Weights from pretrained model not used in %s: %s", model.__class__.__name__, unexpected)
        if errors:
            logger.error("Weights from pretrained model cause errors in %s: %s", model.__class__.__name__, errors)
    return model


def save_model(epoch, args, model, type_name="", logger=None):
    """main_pretrain.py:212-219 / main_task_retrieval.py:222-229: only the model itself, file name
    `pytorch_model.bin.<type_name.><epoch>` under args.output_dir."""
    model_to_save = model.module if hasattr(model, 'module') else model
    output_model_file = os.path.join(
        args.output_dir, "pytorch_model.bin.{}{}".format("" if type_name == "" else type_name + ".", epoch))
    torch.save(model_to_save.state_dict(), output_model_file)
    if logger is not None:
        logger.info("Model saved to %s", output_model_file)
    return output_model_file


def load_head_state(model, model_file, map_location='cpu'):
    """The part of load_model (main_pretrain.py:222-240) that concerns this package: read the file and
    load whatever entries the model owns (queues, pointer, any injected sub-modules)."""
    return init_preweight(model, torch.load(model_file, map_location=map_location))
