"""ctypes binding of libgplum_b200.so (the C ABI in include/gplum_b200.h).

The shared library is the product; this module only loads it.  There is no fallback: if the
library is missing or no sm_100 device is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPLUM_B200_LIB", os.path.join(_HERE, "libgplum_b200.so"))

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> (restype, argtypes); every symbol include/gplum_b200.h declares
SYMBOLS = {
    "gplum_b200_abi_version": (_i, []),
    "gplum_b200_last_error": (C.c_char_p, []),
    "gplum_b200_init": (_i, [_i, C.c_size_t, C.c_size_t]),
    "gplum_b200_finalize": (_i, []),
    "gplum_b200_set_params": (_i, [_f, _i, _i]),
    "gplum_b200_epep": (_i, [_vp, _i, _vp, _i, _vp, _f]),
    "gplum_b200_epsp": (_i, [_vp, _i, _vp, _i, _vp, _f, _i]),
    "gplum_b200_dispatch": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i]),
    "gplum_b200_retrieve": (_i, [_i, _i, _vp, _vp]),
    "gplum_b200_calc_walks": (_i, [_i] + [_vp] * 10 + [_i, _vp, _i, _vp, _i]),
    "gplum_b200_walks_upload": (_i, [_i] + [_vp] * 10 + [_i, _vp, _i]),
    "gplum_b200_walks_select": (_i, [_i]),
    "gplum_b200_set_tile_cap": (_i, [_i]),
    "gplum_b200_walks_run": (_i, [_i]),
    "gplum_b200_walks_pack": (_i, []),
    "gplum_b200_walks_download": (_i, [_vp]),
    "gplum_b200_walks_time": (_i, [_i, _i, C.POINTER(_f)]),
    "gplum_b200_walks_set_packed_dev": (_i, [_vp, _i, _vp, _i]),
    "gplum_b200_pack_epj_dev": (_i, [_vp, _i, _vp]),
    "gplum_b200_pack_spj_dev": (_i, [_vp, _i, _vp]),
    "gplum_b200_gather_epj_packed_dev": (_i, [_vp, _vp, _i, _vp]),
    "gplum_b200_peer_setup": (_i, [_i, _i, _i, _vp]),
    "gplum_b200_peer_open": (_i, [_vp]),
    "gplum_b200_peer_pack": (_i, [_vp, _i]),
    "gplum_b200_peer_wait": (_i, []),
    "gplum_b200_peer_close": (_i, []),
    "gplum_b200_peer_free": (_i, []),
    "gplum_b200_packed_sizes": (None, [C.POINTER(_i), C.POINTER(_i)]),
    "gplum_b200_set_stream": (_i, [_vp]),
    "gplum_b200_synchronize": (_i, []),
    "gplum_b200_counters": (None, [C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_ll), _i]),
    "gplum_b200_tree_build_gpu": (_i, [_i, _vp, _vp, _vp, _vp, C.c_double, _i, _i, _i, _vp]),
    "gplum_b200_pinned_alloc": (_vp, [C.c_size_t]),
    "gplum_b200_pinned_free": (None, [_vp]),
    "gplum_b200_tree_set_motion": (_i, [_i, _vp, _vp]),
    "gplum_b200_tree_set_motion_sparse": (_i, [_i, _vp, _vp, _vp, _vp]),
    "gplum_b200_tree_set_motion_gather": (_i, [_i, _vp, _vp, _vp, _vp]),
    "gplum_b200_tree_download_compact": (_i, [_vp, _vp, _vp, _i, C.POINTER(_i)]),
    "gplum_b200_tree_build_gpu_vel": (_i, [_i, _vp, _vp, _vp, _vp, _vp, C.c_double, _i, _i, _i, _vp]),
    "gplum_b200_tree_build_gpu_epj": (_i, [_i, _vp, _i, C.c_double, _i, _i, _vp]),
    "gplum_b200_tree_copy_gpu": (_i, [_vp] * 12),
    "gplum_b200_tree_gpu_times": (_i, [C.POINTER(_f)]),
    "gplum_b200_tree_download_original": (_i, [_vp]),
    "gplum_b200_tree_build_gpu_part": (_i, [_i, _vp, C.c_double, _i, _i, _i, _i, _vp]),
    "gplum_b200_tree_build_gpu_part_rec48": (_i, [_i, _vp, C.c_double, _i, _i, _i, _i, _vp]),
    "gplum_b200_walks_download_range": (_i, [_vp, _ll, _ll]),
    "gplum_b200_tree_gpu_stamps": (_i, [_vp, _i]),
    "gplum_b200_state_upload": (_i, [_i, _vp, _vp, _vp]),
    "gplum_b200_state_download": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "gplum_b200_state_tree_build": (_i, [C.c_double, _i, _i, _vp]),
    "gplum_b200_state_kick": (_i, [_i, _i, C.c_double]),
    "gplum_b200_state_drift": (_i, [_vp, C.c_double, C.c_double, _i, _vp, _vp]),
    "gplum_b200_state_pull_unhandled": (_i, [_vp, _vp, _i, C.POINTER(_i)]),
    "gplum_b200_state_push": (_i, [_vp, _vp, _i]),
    "gplum_b200_debug_trace": (_i, [_i, _vp, _i, C.POINTER(_i)]),
    "gplum_b200_debug_build_items": (_i, [_i, _vp, _vp, _vp, _ll, _i, _i, _i, _vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _vp, _i, C.POINTER(_i)]),
    "gplum_b200_fp32_peak": (_i, [_i, C.POINTER(_f), C.POINTER(_f)]),
    "gplum_b200_soft_corr_enable": (_i, [_i, _ll]),
    "gplum_b200_correct_long_run": (_i, [_i, _vp, _i]),
    "gplum_b200_correct_long_download": (_i, [_i, _vp, _vp, _vp, _ll, C.POINTER(_ll), C.POINTER(_ll)]),
    "gplum_b200_correct_long_download_compact": (_i, [_i, _vp, _ll, C.POINTER(_ll), _vp, _ll, C.POINTER(_ll), C.POINTER(_ll)]),
    "gplum_b200_correct_long_time": (_i, [_i, _vp, _i, _i, C.POINTER(_f)]),
}

_lib = None


class GplumB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GplumB200Error("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise GplumB200Error("libgplum_b200 error %d: %s" % (rc, lib().gplum_b200_last_error().decode()))
