"""ctypes binding of libcsn_b200.so (the C-ABI declared in include/csn_b200.h).

The library is the product: there is NO CPU or PyTorch fallback.  If the shared object is missing, or a call
fails, this module raises -- loudly -- instead of routing anywhere else.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsn_b200.so")

F32, BF16 = 0, 1
LAYOUT_BCT, LAYOUT_BTC, LAYOUT_TBC = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
DINO_SINGLE, DINO_MULTICROP_REF, DINO_MULTICROP_CANONICAL = 0, 1, 2

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> argtypes; every function returns int.  Mirrors include/csn_b200.h one to one
# (tests/test_abi.py checks the header against this table and against the built library).
SIGNATURES = {
    "csn_device_info": [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "csn_sosfilt_f32": [_vp, _vp, C.POINTER(C.c_double), _i, _i, _i, _i, _i, _i, _i, _vp],
    "csn_btc_to_tbc": [_vp, _vp, _i, _i, _i, _i, _vp],
    "csn_cast": [_vp, _i, _vp, _i, _sz, _vp],
    "csn_gemm_f32": [_i, _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _f, _vp, _i, _vp, _i, _vp],
    "csn_gemm_f32_rowsum": [_i, _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _f, _vp, _i, _vp, _i, _vp, _i, _vp],
    "csn_colsum_f32": [_vp, _vp, _i, _i, _i, _i, _vp],
    "csn_scale_f32": [_vp, _sz, _vp, _f, _vp],
    "csn_act_fwd": [_vp, _vp, _sz, _i, _vp],
    "csn_act_bwd": [_vp, _vp, _vp, _sz, _i, _vp],
    "csn_l2norm_fwd": [_vp, _vp, _vp, _i, _i, _vp],
    "csn_l2norm_bwd": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "csn_batchnorm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _vp],
    "csn_batchnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "csn_weight_norm_fwd": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "csn_weight_norm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "csn_gemm_bf16_tc": [_i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp],
    "csn_lstm_layer_bytes": [_i, _i, _i, _i, _i, C.POINTER(_sz), C.POINTER(_sz)],
    "csn_lstm_layer_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "csn_lstm_layer_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "csn_lstm_set_cta_budget": [_i],
    "csn_loss_workspace_bytes": [_i, C.POINTER(_sz)],
    "csn_dino_loss_fwd_bwd": [_vp, _vp, _vp, _i, _f, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp],
    "csn_center_ema": [_vp, _vp, _sz, _f, _f, _vp],
    "csn_adam_step": [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _i, _f, _vp],
    "csn_adam_step_graph": [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _vp, _vp, _f, _vp],
    "csn_dp_last_timeout": [C.POINTER(_i)],
    "csn_dp_allreduce_twoshot": [_vp, _vp, _i, _i, _sz, _sz, _vp, _i, _vp, _vp],
    "csn_dp_wait_done": [_vp, _i, _vp, _i, _vp],
    "csn_dp_wait_done_zero": [_vp, _i, _vp, _vp, _sz, _vp],
    "csn_dp_adam_step_peer": [_vp, _vp, _vp, _sz, _vp, _vp, _i, _i, _vp, _sz, _f, _f, _vp, _vp, _f, _f, _f, _f, _f, _i, _f, _vp],
    "csn_feature_dist_loss_fwd_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp],
    "csn_cosine_loss_fwd_bwd": [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp],
    "csn_kd_loss_fwd_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _f, _vp, _vp],
    "csn_head_dino_supported": [_i, _i, _i],
    "csn_head_dino_fwd_bwd": [_vp, _i, _vp, _vp, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp],
    "csn_sosfilt_gather_f32": [_vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, C.POINTER(C.c_double), _i, _i, _i, _i, _vp],
    "csn_select_crop_zscore": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "csn_gather_trials": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _vp],
    "csn_topk_workspace_bytes": [_i, _i, _i, C.POINTER(_sz)],
    "csn_topk_search": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "csn_topk_tc_workspace_bytes": [_i, _i, _i, _i, C.POINTER(_sz)],
    "csn_topk_search_tc": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp],
    "csn_ema_update": [_vp, _vp, _sz, _f, _vp],
    "csn_fused_optim_workspace_bytes": [_i, C.c_longlong, C.POINTER(_sz)],
    "csn_fused_optim_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, C.c_longlong, _vp, _i, _f, _f, _f, _i, _f, _f,
                             _vp, _vp, _vp],
    "csn_clip_grad_segments": [_vp, _vp, _i, C.c_longlong, _vp, _f, _vp],
}
# bring-up / self-test hooks (include/csn_b200_debug.h): not part of the product ABI, bound for tests and scripts only
DEBUG_SIGNATURES = {
    "csn_dbg_umma_tile": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "csn_dbg_lstm_profile_buffer": [_vp],
    "csn_dbg_umma_bench": [_vp, _i, _i, _i, _i, _i, _vp],
    "csn_dbg_store_bw": [_vp, _vp, _sz, _i, _i, _i, _vp],
}
EXTRA_SYMBOLS = ["csn_version", "csn_last_error", "csn_launch_count"]


class CsnError(RuntimeError):
    pass


_lib = None


def load():
    """Load libcsn_b200.so (once).  Raises CsnError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CsnError(
            "libcsn_b200.so not found at %s -- build it with `python -m cerebralsignalnetworks_b200.build` "
            "(there is no CPU / PyTorch fallback for the CUDA path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.csn_version.restype = _i
    lib.csn_version.argtypes = []
    lib.csn_last_error.restype = C.c_char_p
    lib.csn_last_error.argtypes = []
    lib.csn_launch_count.restype = C.c_ulonglong
    lib.csn_launch_count.argtypes = []
    for name, argtypes in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(lib, name)
        fn.restype = _i
        fn.argtypes = argtypes
    _lib = lib
    return lib


def call(name: str, *args):
    """Invoke an entry point; raise CsnError with the library's message on a non-zero return."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CsnError("%s failed (%d): %s" % (name, rc, lib.csn_last_error().decode(errors="replace")))


def launch_count() -> int:
    return int(load().csn_launch_count())


def require_gpu():
    import torch

    if not torch.cuda.is_available():
        raise CsnError("cerebralsignalnetworks_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
