"""The command-line flags that parameterise the head, with the reference's names,
types and defaults (main_pretrain.py:33-100, main_task_retrieval.py:33-97).  Only the
head-relevant subset is declared; ``add_head_flags`` can be applied to the reference's
own parser without clashes when the flags already exist.
"""
import argparse

HEAD_FLAGS = [
    # name, kwargs (defaults of main_pretrain.py; main_task_retrieval.py differs only in batch sizes)
    ("--use_frame_fea", dict(action="store_true", help="whether use frame feature matching text")),
    ("--batch_size", dict(type=int, default=256, help="batch size (GLOBAL: per-GPU = batch_size // n_gpu)")),
    ("--batch_size_val", dict(type=int, default=3500, help="batch size eval")),
    ("--max_frames", dict(type=int, default=12, help="")),
    ("--top_frames", dict(type=int, default=3, help="")),
    ("--contrast_num_negative", dict(type=int, default=4096, help="Num of negative sample in queue")),
    ("--contrast_momentum", dict(type=float, default=0.99, help="momentum")),
    ("--contrast_temperature", dict(type=float, default=0.07, help="temperature")),
    ("--enable_amp", dict(action="store_true", help="whether to use pytorch amp")),
    ("--n_gpu", dict(type=int, default=1, help="Changed in the execute process.")),
    ("--local_rank", dict(default=0, type=int, help="distribted training")),
]
# not in the reference: how the contractions are carried out on the B200
EXTRA_FLAGS = [
    ("--head_precision", dict(type=str, default=None, choices=["fp32", "bf16", "bf16x3"],
                              help="fp32 = CUDA cores, bf16 = tcgen05, bf16x3 = tcgen05 split (fp32 parity)")),
]


def add_head_flags(parser, extra=True):
    have = {s for a in parser._actions for s in a.option_strings}
    for name, kw in HEAD_FLAGS + (EXTRA_FLAGS if extra else []):
        if name not in have:
            parser.add_argument(name, **kw)
    return parser


def get_head_args(argv=None):
    return add_head_flags(argparse.ArgumentParser(description="HMMC contrastive head")).parse_args(argv)
