"""ctypes binding of the C-ABI library (``include/eavqa_b200.h``).

The library is the product: if it cannot be loaded this module raises -- there is no
PyTorch/CPU fallback for any of the operators.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeavqa_b200.so")

c_void_p, c_int32, c_int64, c_char_p, c_size_t = C.c_void_p, C.c_int32, C.c_int64, C.c_char_p, C.c_size_t
c_float_p = C.c_void_p   # device pointers travel as integers


class EavqaConfig(C.Structure):
    _fields_ = [(n, c_int32) for n in ("n_layer", "n_head", "d_model", "vocab", "n_positions", "prefix_length",
                                       "clip_length", "clip_dim", "mapper_type", "mapper_layers")]


MAPPER_MLP, MAPPER_TRANSFORMER = 0, 1
F32, BF16 = 0, 1
ACT = {"none": 0, "gelu_new": 1, "tanh": 2, "relu": 3}

# name -> (restype, argtypes); every symbol include/eavqa_b200.h declares
PROTOTYPES = {
    "eavqa_last_error": (c_char_p, []),
    "eavqa_abi_version": (C.c_int, []),
    "eavqa_create": (C.c_int, [C.POINTER(EavqaConfig), C.POINTER(c_void_p)]),
    "eavqa_destroy": (C.c_int, [c_void_p]),
    "eavqa_load_lm_weight": (C.c_int, [c_void_p, c_char_p, c_void_p, c_int32, c_int64, c_void_p]),
    "eavqa_finalize_lm": (C.c_int, [c_void_p, c_void_p]),
    "eavqa_mapper_param_count": (c_int64, [c_void_p]),
    "eavqa_mapper_num_tensors": (c_int32, [c_void_p]),
    "eavqa_mapper_tensor_info": (C.c_int, [c_void_p, c_int32, c_char_p, c_size_t, C.POINTER(c_int64), C.POINTER(c_int64),
                                           C.POINTER(c_int64)]),
    "eavqa_train_step": (C.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "eavqa_forward_logits": (C.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                       c_void_p]),
    "eavqa_generate": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                 c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_void_p, C.POINTER(c_int32),
                                 c_void_p]),
    "eavqa_grad_bucket_count": (c_int32, [c_void_p]),
    "eavqa_grad_bucket_range": (C.c_int, [c_void_p, c_int32, C.POINTER(c_int64), C.POINTER(c_int64)]),
    "eavqa_set_grad_events": (C.c_int, [c_void_p, C.POINTER(c_void_p), c_int32]),
    "eavqa_build_caption_labels": (C.c_int, [c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "eavqa_ensemble_select": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "eavqa_rices_search": (C.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "eavqa_rices_rerank": (C.c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "eavqa_scale_grads": (C.c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "eavqa_splice": (C.c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int64,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "eavqa_launch_count": (c_int64, []),
    "eavqa_adamw_step": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_float, c_int32, C.c_float, c_void_p]),
    "eavqa_sharded_adamw_range": (C.c_int, [c_int64, c_int32, c_int32, C.POINTER(c_int64), C.POINTER(c_int64)]),
    "eavqa_sharded_adamw_step": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_uint32, c_int32, c_int32, c_void_p,
                                           c_void_p, c_int64, c_int64, c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                           c_int32, C.c_float, c_void_p]),
    "eavqa_profile_begin": (C.c_int, []),
    "eavqa_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_int64), c_char_p, c_size_t]),
    "eavqa_op_gemm": (C.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32,
                                c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32,
                                c_void_p]),
    "eavqa_op_gemm_wgrad": (C.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32,
                                      c_void_p]),
    "eavqa_op_lmhead_ce": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "eavqa_op_layernorm_fwd": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "eavqa_op_layernorm_bwd": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_int32, c_int32, c_void_p]),
    "eavqa_op_lm_attention_fwd": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "eavqa_op_lm_attention_bwd": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                            c_int32, c_int32, c_void_p]),
    "eavqa_op_mapper_attention_fwd": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "eavqa_op_mapper_attention_bwd": (C.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "eavqa_op_convert_transpose": (C.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
}

_lib: Optional[C.CDLL] = None


class EavqaError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load ``libeavqa_b200.so`` (building it with nvcc first when absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise EavqaError("CUDA library %s is missing; run `python __graft_entry__.py build`" % LIB_PATH)
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.eavqa_abi_version() != 1:
        raise EavqaError("ABI version mismatch between lib.py and %s" % LIB_PATH)
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().eavqa_last_error()
        raise EavqaError(msg.decode("utf-8", "replace") if msg else "eavqa_b200 call failed (status %d)" % status)


def ptr(t) -> Optional[int]:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
