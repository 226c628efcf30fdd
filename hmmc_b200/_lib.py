"""ctypes binding of libhmmc_head.so (the C ABI declared in include/hmmc_head.h).

There is no fallback: if the shared library is missing or a call fails, this
module raises.  The library is built in-tree by ``python -m hmmc_b200.build``.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint64, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhmmc_head.so")

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}
POS_PAIR, POS_FRAME_NEIGHBOUR, POS_ONE_TO_FRAMES, POS_FRAMES_TO_ONE = 0, 1, 2, 3


class HmmcError(RuntimeError):
    pass


class hmmc_queue(Structure):
    _fields_ = [("dk", c_void_p), ("pack_kd", c_void_p), ("pack_dk", c_void_p),
                ("D", c_int32), ("Kq", c_int32), ("planes", c_int32), ("reserved", c_int32)]


class hmmc_pretrain_io(Structure):
    _fields_ = [(n, c_void_p) for n in ("v_fea", "title_fea", "frame_fea", "frame_pred", "v_fea_k", "title_fea_k",
                                        "frame_fea_k", "frame_proj_k", "d_v_fea", "d_title_fea", "d_frame_fea",
                                        "d_frame_pred")]


class hmmc_head_schedule(Structure):
    _fields_ = [("phase", c_int32), ("reserved_sms", c_int32), ("queues_released", c_void_p)]


class hmmc_mlp_params(Structure):
    _fields_ = [(n, c_void_p) for n in ("W1", "b1", "gamma", "beta", "W2", "b2", "running_mean", "running_var")]


# name -> (restype, argtypes); must list every symbol of include/hmmc_head.h
SIGNATURES = {
    "hmmc_last_error": (c_char_p, []),
    "hmmc_version": (c_int, []),
    "hmmc_launch_count": (ctypes.c_ulonglong, []),
    "hmmc_device_check": (c_int, []),
    "hmmc_rownorm_pack": (c_int, [c_void_p, c_int64, c_int, c_int64, c_float, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p]),
    "hmmc_gemm_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                              c_int, c_int, c_int, c_float, c_void_p]),
    "hmmc_umma_gemm_nt": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                  c_int, c_int, c_float, c_void_p]),
    "hmmc_umma_gemm_nt_tiled": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                        c_int, c_int, c_float, c_int, c_void_p]),
    "hmmc_queue_pack": (c_int, [POINTER(hmmc_queue), c_void_p]),
    "hmmc_infonce_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "hmmc_infonce_queue_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           POINTER(hmmc_queue), c_float, c_float, c_int, c_void_p, c_void_p,
                                           c_void_p, c_size_t, c_void_p]),
    "hmmc_pretrain_head_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "hmmc_pretrain_head_fwd_bwd": (c_int, [POINTER(hmmc_pretrain_io), c_int, c_int, c_int, POINTER(hmmc_queue),
                                           POINTER(hmmc_queue), POINTER(hmmc_queue), POINTER(hmmc_queue), c_float,
                                           c_float, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
    "hmmc_pretrain_head_fwd_bwd_sched": (c_int, [POINTER(hmmc_pretrain_io), c_int, c_int, c_int, POINTER(hmmc_queue),
                                                 POINTER(hmmc_queue), POINTER(hmmc_queue), POINTER(hmmc_queue), c_float,
                                                 c_float, c_float, c_float, c_int, c_int, c_void_p,
                                                 POINTER(hmmc_head_schedule), c_void_p, c_size_t, c_void_p]),
    "hmmc_enqueue_norm_direct": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         POINTER(hmmc_queue), c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "hmmc_ema_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float,
                               c_float, c_void_p]),
    "hmmc_ema_block_elems": (c_int, []),
    "hmmc_visual_tail_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hmmc_visual_tail_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hmmc_mlp_ctx_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "hmmc_mlp_fwd_a": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(hmmc_mlp_params), c_int, c_int, c_void_p,
                               c_size_t, POINTER(c_void_p), c_void_p]),
    "hmmc_mlp_fwd_b": (c_int, [c_int, c_int, c_int, c_int, POINTER(hmmc_mlp_params), c_float, c_float, ctypes.c_double,
                               c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "hmmc_mlp_bwd_a": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(hmmc_mlp_params), c_int, c_void_p, c_size_t,
                               c_void_p, c_void_p, POINTER(c_void_p), c_void_p]),
    "hmmc_mlp_bwd_b": (c_int, [c_int, c_int, c_int, c_int, POINTER(hmmc_mlp_params), ctypes.c_double, c_int, c_void_p,
                               c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hmmc_bert_adam_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "hmmc_bert_adam_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int64, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "hmmc_clip_grad_norm_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p,
                                          c_size_t, c_void_p]),
    "hmmc_enqueue_norm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(hmmc_queue), c_void_p, c_int64,
                                  c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "hmmc_peer_push_rows": (c_int, [c_void_p, c_int64, POINTER(c_uint64), POINTER(c_uint64), c_int, c_int, c_int64,
                                    c_void_p, c_void_p, c_void_p]),
    "hmmc_peer_wait": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "hmmc_scale_tensors": (c_int, [POINTER(c_uint64), POINTER(c_int64), c_int, c_void_p, c_void_p]),
    "hmmc_pack_rows": (c_int, [POINTER(c_uint64), POINTER(c_int32), c_int, c_int64, c_void_p, c_void_p, c_int,
                               c_void_p]),
    "hmmc_unpack_rows": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_int32), c_int, c_int64, c_void_p]),
    "hmmc_similarity_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "hmmc_loose_similarity_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_int,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "hmmc_loose_similarity_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hmmc_cross_en_fwd_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "hmmc_sym_ce_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "hmmc_sym_ce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                                    c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hmmc_sym_ce_packed_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "hmmc_sym_ce_packed_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_void_p,
                                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "hmmc_sim_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "hmmc_sim_topk_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_int,
                                  c_int, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "hmmc_rank_count": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "hmmc_eval_fused_supported": (c_int, [c_int, c_int, c_int]),
    "hmmc_eval_pack_text": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "hmmc_eval_pack_gallery": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hmmc_eval_gt_scores": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_int, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "hmmc_eval_theta": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "hmmc_eval_fused_rank": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hmmc_group_max": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HmmcError("%s not found: build it with `python -m hmmc_b200.build` "
                        "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().hmmc_last_error()
        raise HmmcError("%s failed (%d): %s" % (what or "libhmmc_head call", rc,
                                               msg.decode() if msg else "?"))
