"""ctypes binding of libssdhead.so (declared in include/ssdhead.h).

There is no fallback: if the library is missing or an entry point fails, a RuntimeError
is raised.  The library is loaded lazily on first use so that importing the package (and
the drop-in ``Util`` / ``Losses`` modules) works in CPU-only processes such as DataLoader
workers, which never reach the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libssdhead.so")

ABI_VERSION = 4
WS_MATCH, WS_LOSS, WS_DETECT, WS_NMS, WS_ROWS = 0, 1, 2, 3, 4
E_BADARG, E_UNSUPPORTED, E_WORKSPACE, E_ALIGN, E_STATE = -1, -2, -3, -4, -5

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/ssdhead.h one to one
SIGNATURES = {
    "ssdhead_abi_version": (_i, []),
    "ssdhead_error_string": (C.c_char_p, [_i]),
    "ssdhead_launch_count": (C.c_uint64, []),
    "ssdhead_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ssdhead_cxcywh_to_xyxy": (_i, [_vp, _vp, _i, _vp]),
    "ssdhead_xyxy_to_cxcywh": (_i, [_vp, _vp, _i, _vp]),
    "ssdhead_encode": (_i, [_vp, _vp, _vp, _i, _vp]),
    "ssdhead_decode": (_i, [_vp, _vp, _vp, _i, _vp]),
    "ssdhead_iou_matrix": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "ssdhead_intersection_matrix": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "ssdhead_match_from_iou": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "ssdhead_match": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_multibox_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_ce_stream": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_ce_match_stream": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _sz, _vp, _sz, _i, _vp]),
    "ssdhead_multibox_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "ssdhead_multibox_step_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _sz, _vp, _sz, _i, _i, C.c_uint, _vp, _vp, _vp, _vp]),
    "ssdhead_multibox_step_resident": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _i, _i, C.c_uint, _vp, _vp, _vp, _vp]),
    "ssdhead_xchg_bytes": (_sz, []),
    "ssdhead_ctx_xchg_export": (_i, [_vp, _vp]),
    "ssdhead_ctx_xchg_import": (_i, [_vp, _vp, _i, _i]),
    "ssdhead_ctx_xchg_error": (_i, [_vp]),
    "ssdhead_mine": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                          _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_mine_sparse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _i, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_finish_loss": (_i, [_vp, _vp, _vp, _vp]),
    "ssdhead_scale_grads": (_i, [_vp, _sz, _vp, _sz, _vp, _vp]),
    "ssdhead_detect": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_detect_fallbacks": (_i, [_vp, _sz, _i, _i, _i, _i, _vp, _vp]),
    "ssdhead_detect_from_scores": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_voc_ap_workspace_bytes": (_sz, [_i, _i, _i]),
    "ssdhead_voc_ap": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "ssdhead_ctx_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp]),
    "ssdhead_ctx_destroy": (None, [_vp]),
    "ssdhead_host_alloc": (_vp, [_sz]),
    "ssdhead_host_free": (None, [_vp]),
    "ssdhead_ctx_multibox_loss_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "ssdhead_ctx_multibox_loss_dev_resident": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "ssdhead_ctx_multibox_loss_begin": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, C.POINTER(_vp), _vp]),
    "ssdhead_ctx_multibox_loss_end": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ssdhead_ctx_multibox_loss_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "ssdhead_ctx_multibox_loss_host_sparse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _i, _vp, _vp, _vp, _vp]),
    "ssdhead_ctx_detect_host": (_i, [_vp, _vp, _vp, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "ssdhead_pack_gt": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "ssdhead_multibox_step_levels": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _sz, _vp, _sz, _vp]),
    "ssdhead_multibox_step_levels_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                                  _vp, _sz, _vp, _sz, _i, _i, C.c_uint, _vp, _vp, _vp, _vp]),
    "ssdhead_ctx_multibox_loss_levels_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    "ssdhead_detect_levels": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

MAX_LEVELS = 8


class Levels(C.Structure):
    """``ssdhead_levels`` of include/ssdhead.h: per-pyramid-level head tensors."""
    _fields_ = [("num_levels", C.c_int32), ("count", C.c_int32 * MAX_LEVELS),
                ("conf", C.c_void_p * MAX_LEVELS), ("loc", C.c_void_p * MAX_LEVELS),
                ("grad_conf", C.c_void_p * MAX_LEVELS), ("grad_loc", C.c_void_p * MAX_LEVELS)]

_lib = None


def load() -> C.CDLL:
    """Load libssdhead.so; raise loudly when it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found. Build it with `python -m objectdetection_ssd_b200.build` "
            "(needs nvcc; targets sm_100a). The SSD head path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.ssdhead_abi_version() != ABI_VERSION:
        raise RuntimeError("libssdhead.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().ssdhead_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def launch_count() -> int:
    return int(load().ssdhead_launch_count())
