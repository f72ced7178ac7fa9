"""ctypes binding of libmmf_b200.so (include/mmf_b200.h).  There is NO fallback: if the
library is missing, or there is no sm_100 device when a handle is created, this raises."""
from __future__ import annotations

import ctypes as C
import os

OK, ERR_BAD_ARG, ERR_CUDA, ERR_NOT_LOADED, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NOMEM, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6, -7
SHARD_ID_BYTES = 128
F32, F16, BF16, F64 = 0, 1, 2, 3
VAULT_FP32, VAULT_BF16 = 0, 1
ALGO_AUTO, ALGO_STREAM, ALGO_MMA = 0, 1, 2
MAX_TOP_K = 256
FUSION_PARAMS = 2530

_p, _i, _l, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/mmf_b200.h one to one
SIGNATURES = {
    "mmf_version": (C.c_char_p, []),
    "mmf_arch": (_i, []),
    "mmf_status_string": (C.c_char_p, [_i]),
    "mmf_create": (_i, [_i, C.POINTER(_p)]),
    "mmf_destroy": (_i, [_p]),
    "mmf_last_error": (C.c_char_p, [_p]),
    "mmf_cosine_pairs": (_i, [_p, _p, _p, _l, _i, _d, _p, _p, _p]),
    "mmf_vault_load": (_i, [_p, _p, _i, _l, _i, _i, _i, _l]),
    "mmf_vault_unload": (_i, [_p]),
    "mmf_vault_info": (_i, [_p, C.POINTER(_l), C.POINTER(_i), C.POINTER(_i), C.POINTER(_l)]),
    "mmf_vault_search": (_i, [_p, _p, _l, _i, _d, _i, _p, _p, _p, _p]),
    "mmf_vault_search_host": (_i, [_p, _p, _l, _i, _d, _i, _p, _p, _p]),
    "mmf_vault_search_candidates": (_i, [_p, _p, _l, _i, _i, _p, _p]),
    "mmf_topk_merge": (_i, [_p, _p, _i, _l, _i, _i, _d, _p, _p, _p, _p]),
    "mmf_shard_unique_id": (_i, [_p]),
    "mmf_shard_init": (_i, [_p, _i, _i, _p]),
    "mmf_shard_finalize": (_i, [_p]),
    "mmf_shard_info": (_i, [_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "mmf_vault_search_sharded": (_i, [_p, _p, _l, _i, _d, _i, _p, _p, _p, _p]),
    "mmf_shard_all_gather": (_i, [_p, _p, _l, _p, _p]),
    "mmf_exchange_layout": (_i, [_i, _l, _i, C.POINTER(_l), C.POINTER(_l)]),
    "mmf_exchange_attach": (_i, [_p, _i, _i, C.POINTER(C.c_uint64), _l]),
    "mmf_exchange_detach": (_i, [_p]),
    "mmf_vault_search_push": (_i, [_p, _p, _l, _i, _i, _p]),
    "mmf_vault_exchange_merge": (_i, [_p, _i, _d, _p, _p, _p, _p]),
    "mmf_vault_search_exchange": (_i, [_p, _p, _l, _i, _i, _d, _i, _p, _p, _p, _p]),
    "mmf_fusion_load": (_i, [_p, _p]),
    "mmf_fusion_forward": (_i, [_p, _p, _l, _p, _p, _p, _p]),
    "mmf_verdict_batch": (_i, [_p, _p, _p, _l, _p, _p, _p, _p]),
    "mmf_score_batch_host": (_i, [_p, _p, _p, _p, _p, _l, _i, _d, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mmf_verdict_assemble": (_i, [_p, _p, _p, _l, _p, _p, _p, _p, _p, _p, _p]),
    "mmf_score_batch": (_i, [_p, _p, _p, _p, _p, _l, _i, _d, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mmf_score_batch_submit": (_i, [_p, _i, _p, _p, _p, _p, _l, _i, _d, _i]),
    "mmf_score_batch_collect": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mmf_set_option": (_i, [_p, C.c_char_p, _i]),
    "mmf_get_option": (_i, [_p, C.c_char_p, C.POINTER(_i)]),
    "mmf_mma_plan_check": (_i, [_l, _l, _i, C.POINTER(_l), C.POINTER(_i), C.POINTER(_i)]),
    "mmf_mma_hist_bound": (_i, [_p, _l, _i, C.POINTER(C.c_float)]),
    "mmf_mma_screen_eps": (_d, []),
    "mmf_launch_count": (_l, [_p]),
    "mmf_collective_count": (_l, [_p]),
}


class MMFError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libmmf_b200: {message} (status {status})")
        self.status = status


def library_path() -> str:
    return os.environ.get("MMF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmmf_b200.so")


_LIB = None


def load() -> C.CDLL:
    """dlopen the library and bind every entry point of the header."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise MMFError(ERR_NO_DEVICE, f"{path} not found -- run ./build.sh (or __graft_entry__.build()); "
                                          "there is no CPU fallback")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB
