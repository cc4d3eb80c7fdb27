"""ctypes binding of libsuperslam_b200.so (the C-ABI declared in include/superslam_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises,
and every wrapper raises SsbError on a non-zero status from the library.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsuperslam_b200.so")


class SsbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"superslam_b200 status {status}: {message}")
        self.status = status


SSB_OK, SSB_ERR_INVALID, SSB_ERR_CUDA, SSB_ERR_IO, SSB_ERR_EXHAUSTED, SSB_ERR_NODEVICE = range(6)

# Every symbol include/superslam_b200.h declares: (name, restype, argtypes)
_u8pp = C.POINTER(C.POINTER(C.c_uint8))
_fpp = C.POINTER(C.POINTER(C.c_float))
_vp = C.c_void_p
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_dp = C.POINTER(C.c_double)

SYMBOLS = [
    ("ssb_last_error", C.c_char_p, []),
    ("ssb_version", C.c_int, []),
    ("ssb_device_count", C.c_int, []),
    ("ssb_sp_create", C.c_int, [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_sp_destroy", None, [_vp]),
    ("ssb_sp_extract", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fpp, _fpp, _ip,
                                 C.POINTER(_vp), _ip]),
    ("ssb_sp_slot_retain", C.c_int, [_vp, C.c_int]),
    ("ssb_sp_slot_release", C.c_int, [_vp, C.c_int]),
    ("ssb_sp_slots_in_use", C.c_int, [_vp]),
    ("ssb_sp_max_keypoints", C.c_int, [_vp]),
    ("ssb_sp_debug_read", C.c_int, [_vp, C.c_char_p, _vp, C.c_size_t]),
    ("ssb_lg_create", C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_lg_clone_context", C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_lg_destroy", None, [_vp]),
    ("ssb_lg_match_device", C.c_int, [_vp, _fp, C.c_int, _vp, _fp, C.c_int, _vp, _i32p, _fp]),
    ("ssb_lg_match_host", C.c_int, [_vp, _fp, C.c_int, _fp, _fp, C.c_int, _fp, _i32p, _fp]),
    ("ssb_desc_to_host_f32", C.c_int, [C.c_int, _vp, C.c_int, C.c_int, _fp]),
    ("ssb_lg_debug_read", C.c_int, [_vp, C.c_char_p, _vp, C.c_size_t]),
    ("ssb_fe_create", C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_fe_destroy", None, [_vp]),
    ("ssb_fe_process", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp, _i32p, _fp, _fp,
                                 _u8p]),
    ("ssb_fe_enqueue_device", C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    ("ssb_fe_fetch", C.c_int, [_vp, C.c_int, _ip, _fp, _fp, _i32p, _fp, _fp, _u8p]),
    ("ssb_fe_submit", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("ssb_fe_collect", C.c_int, [_vp, _ip, _ip, _fp, _fp, _i32p, _fp, _fp, _u8p]),
    ("ssb_fe_sync", C.c_int, [_vp]),
    ("ssb_fe_set_extract_only", C.c_int, [_vp, C.c_int]),
    ("ssb_fe_enable_tracking", C.c_int, [_vp, C.c_int]),
    ("ssb_fe_reset_tracking", C.c_int, [_vp]),
    ("ssb_fe_promote_keyframes", C.c_int, [_vp, _u8p, C.c_int]),
    ("ssb_fe_tracking_results", C.c_int, [_vp, C.c_int, _ip, _i32p, _fp, _u8p]),
    ("ssb_mg_create", C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_int, _ip, C.c_int, C.POINTER(_vp)]),
    ("ssb_mg_destroy", None, [_vp]),
    ("ssb_mg_device_count", C.c_int, [_vp]),
    ("ssb_mg_device_of_pair", C.c_int, [_vp, C.c_int]),
    ("ssb_mg_last_error", C.c_char_p, [_vp]),
    ("ssb_mg_process", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp, _i32p, _fp, _fp, _u8p]),
    ("ssb_fe_event_record", C.c_int, [_vp, C.c_int]),
    ("ssb_fe_event_elapsed_ms", C.c_int, [_vp, C.c_int, C.c_int, _fp]),
    ("ssb_fe_upload_images", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_fe_kernel_launches_per_call", C.c_int, [_vp, C.c_int]),
    ("ssb_fe_superpoint", _vp, [_vp]),
    ("ssb_fe_lightglue", _vp, [_vp]),
    ("ssb_ep_create", C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_ep_destroy", None, [_vp]),
    ("ssb_ep_descriptor_dim", C.c_int, [_vp]),
    ("ssb_ep_compute", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    ("ssb_ep_add", C.c_int, [_vp, C.c_uint64, _fp, C.c_int]),
    ("ssb_ep_query", C.c_int, [_vp, _fp, C.c_int, C.c_uint64, C.c_int, C.c_float, C.POINTER(C.c_uint64), _fp,
                               C.c_int, _ip]),
    ("ssb_ep_index_size", C.c_int, [_vp]),
    ("ssb_ep_debug_read", C.c_int, [_vp, C.c_char_p, _vp, C.c_size_t]),
    ("ssb_rect_create", C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_rect_destroy", None, [_vp]),
    ("ssb_rect_convert_maps", C.c_int, [_fp, _fp, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)]),
    ("ssb_rect_remap", C.c_int, [_vp, _u8pp, C.c_int, C.c_int, _u8pp]),
    ("ssb_rect_remap_device", C.c_int, [_vp, _vp, C.c_int, _vp]),
    ("ssb_fe_set_rectifiers", C.c_int, [_vp, _vp, _vp]),
    ("ssb_rgbd_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    ("ssb_rgbd_destroy", None, [_vp]),
    ("ssb_rgbd_process", C.c_int, [_vp, _fp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                   C.c_double, C.c_double, C.c_double, _fp, _dp, _u8p]),
    ("ssb_kernel_launch_count", C.c_longlong, []),
    ("ssb_profile_enable", None, [C.c_int]),
    ("ssb_profile_collect", None, []),
    ("ssb_profile_report", C.c_int, [C.c_char_p, C.c_size_t]),
]

_lib = None


def load():
    """Load the shared library (once).  Raises OSError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C superslam_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != SSB_OK:
        msg = load().ssb_last_error()
        raise SsbError(status, msg.decode() if msg else "")
