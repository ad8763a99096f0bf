"""ctypes binding of libofstab.so (the C ABI declared in include/ofstab.h).

Loading never falls back: a missing library is built with nvcc if a toolchain is present and
otherwise raises.  Every call that touches the device raises ``OfstabError`` on a non-zero
status, carrying the library's thread-local message.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OFS_OK, OFS_EINVAL, OFS_ECUDA, OFS_ENOTSM100, OFS_ENOMEM, OFS_ESTATE = range(6)
PREC_BF16, PREC_FP16 = 0, 1

_STATUS_NAMES = {1: "OFS_EINVAL", 2: "OFS_ECUDA", 3: "OFS_ENOTSM100", 4: "OFS_ENOMEM", 5: "OFS_ESTATE"}


class OfstabError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{_STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class NamedArray(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


_i, _f, _p, _ll = C.c_int, C.c_float, C.c_void_p, C.c_longlong

# name -> (restype, argtypes); mirrors include/ofstab.h one to one
PROTOTYPES = {
    "ofs_version": (_i, []),
    "ofs_last_error": (C.c_char_p, []),
    "ofs_device_check": (_i, [_i]),
    "ofs_launch_count": (C.c_uint64, []),
    "ofs_tf_warp": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "ofs_set_warp_variant": (_i, [_i]),
    "ofs_flow_resize": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ofs_flow_resize_warp": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ofs_flow_resize_ex": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _p]),
    "ofs_warp_perspective_u8": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ofs_tf1_resize_bilinear": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ofs_cv_resize_linear_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "ofs_flow_box_blur_ema": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    "ofs_medfilt_nd3": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "ofs_grid_sample_affine": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ofs_grid_sample_projective": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ofs_vec2mtrx": (_i, [_p, _p, _i, _i, _i, _p]),
    "ofs_lie_warp": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ofs_net_create": (_i, [C.POINTER(_p), _i, _i, _i]),
    "ofs_net_destroy": (_i, [_p]),
    "ofs_net_load_weights": (_i, [_p, C.POINTER(NamedArray), _i]),
    "ofs_net_forward": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "ofs_net_stabilize": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "ofs_net_stabilize_host": (_i, [_p, _p, _p, _p, _i, _i, _i]),
    "ofs_net_host_h2d_bytes": (_ll, [_p, _i, _i, _i]),
    "ofs_net_get_activation": (_i, [_p, C.c_char_p, _i, _p, C.c_int64, C.POINTER(_i), _p]),
    "ofs_net_profile": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, C.POINTER(_i), _p]),
    "ofs_net_time_kernels": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "ofs_net_launches_per_forward": (_i, [_p]),
    "ofs_net_graph_stats": (_ll, [_p, _i]),
    "ofs_chain_trace_read": (_i, [_p, _i]),
    "ofs_clips_create": (_i, [_p, _p, _i, _i, _i]),
    "ofs_clips_destroy": (_i, [_p]),
    "ofs_clips_reset": (_i, [_p]),
    "ofs_clips_frame_index": (C.c_longlong, [_p]),
    "ofs_clips_step_host": (_i, [_p, _p, _p, _p]),
    "ofs_clips_submit_host": (_i, [_p, _p, _p, _p]),
    "ofs_clips_wait": (_i, [_p]),
    "ofs_clips_in_flight": (_i, [_p]),
    "ofs_clips_depth": (_i, []),
    "ofs_conv2d_nhwc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ofs_conv2d_nhwc_ex": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
}
# host-only introspection used by the CPU test-suite (not part of the product API)
DEBUG_PROTOTYPES = {
    "ofs_debug_conv_plan": (_i, [_i] * 11 + [_p, _p, _p, _p, _p, _ll, _p]),
    "ofs_debug_conv_plan_ex": (_i, [_i] * 12 + [_p, _p, _p, _p, _p, _p, _p, _ll, _p]),
    "ofs_debug_conv_schedule": (_i, [_i] * 12 + [_p]),
    "ofs_debug_cvt16": (C.c_uint, [_f, _i]),
    "ofs_debug_host_pack_bf16": (_i, [_p, _p, _ll, _i]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing or stale).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not _build.is_fresh():
        if os.path.exists(_build.LIB_PATH) and os.environ.get("OFSTAB_NO_REBUILD") == "1":
            pass
        else:
            _build.build_library()
    lib = C.CDLL(_build.LIB_PATH, mode=C.RTLD_GLOBAL)
    for table in (PROTOTYPES, DEBUG_PROTOTYPES):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)  # AttributeError = header and library disagree: fail loudly
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != OFS_OK:
        raise OfstabError(status, load().ofs_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device/host pointer of a contiguous float32 torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream_ptr(device=None):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
