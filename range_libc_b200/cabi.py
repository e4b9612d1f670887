"""ctypes binding of librangelib_b200.so -- one Python function per symbol of
include/rangelib_b200.h, nothing else.  Fails loudly when the library is missing: there is no
CPU fallback in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RL_B200_LIB: development aid -- load an alternative build of the same library (tools/trace_fused.py, A/B builds)
LIB_PATH = os.environ.get("RL_B200_LIB") or os.path.join(_HERE, "librangelib_b200.so")

RL_BL, RL_RM, RL_CDDT, RL_PCDDT, RL_GLT = 0, 1, 2, 3, 4
RL_OK, RL_E_INVALID, RL_E_CUDA, RL_E_NO_DEVICE, RL_E_STATE, RL_E_MIXED = 0, -1, -2, -3, -4, -5

_vp = C.c_void_p
_i = C.c_int
_f = C.c_float

# symbol -> (restype, argtypes); mirrors include/rangelib_b200.h one to one
SIGNATURES = {
    "rl_last_error": (C.c_char_p, []),
    "rl_stat_kernel_launches": (C.c_uint64, []),
    "rl_map_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "rl_map_set_world": (_i, [_vp, _f, _f, _f, _f, _f, _f]),
    "rl_map_width": (_i, [_vp]),
    "rl_map_height": (_i, [_vp]),
    "rl_map_is_occupied": (_i, [_vp, _i, _i]),
    "rl_map_get": (_i, [_vp, _vp]),
    "rl_map_update": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "rl_map_destroy": (None, [_vp]),
    "rl_method_create": (_i, [_i, _vp, _f, C.c_uint, _i, C.POINTER(_vp)]),
    "rl_method_destroy": (None, [_vp]),
    "rl_method_prune": (_i, [_vp, _f]),
    "rl_method_save_cddt": (_i, [_vp, C.c_char_p]),
    "rl_method_get_params": (_i, [_vp, C.POINTER(_f), C.POINTER(C.c_uint), C.POINTER(_i)]),
    "rl_method_create_from_cddt": (_i, [_vp, C.c_char_p, _i, C.POINTER(_vp)]),
    "rl_pf_normalize_weights": (_i, [_vp, _vp, _i, C.c_double, C.POINTER(C.c_double)]),
    "rl_pf_resample": (_i, [_vp, _vp, _vp, _vp, _i, C.c_double]),
    "rl_pf_motion_update": (_i, [_vp, _vp, _i, _f, _f, _f, _vp]),
    "rl_method_set_stream": (_i, [_vp, _vp]),
    "rl_method_use_own_stream": (_i, [_vp]),
    "rl_method_synchronize": (_i, [_vp]),
    "rl_method_update_map": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "rl_method_update_map_batch": (_i, [_vp, _vp, _vp, _i]),
    "rl_method_memory": (C.c_int64, [_vp]),
    "rl_calc_range": (_i, [_vp, _f, _f, _f, C.POINTER(_f)]),
    "rl_calc_range_many": (_i, [_vp, _vp, _vp, _i]),
    "rl_numpy_calc_range": (_i, [_vp, _vp, _vp, _i]),
    "rl_numpy_calc_range_angles": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "rl_set_sensor_model": (_i, [_vp, _vp, _i]),
    "rl_eval_sensor_model": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "rl_calc_range_repeat_angles_eval_sensor_model": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "rl_method_set_map_occupancy_grid": (_i, [_vp, _vp, _i, _i]),
    "rl_method_set_map_rgba": (_i, [_vp, _vp, _i, _i, C.c_float]),
    "rl_debug_get_occ": (_i, [_vp, _vp]),
    "rl_calc_range_many_radial_optimized": (_i, [_vp, _vp, _vp, _i, _i, C.c_float, C.c_float]),
    "rl_calc_range_repeat_angles_eval_sensor_model_peers": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_vp), _i, C.c_int64, _i, _i]),
    "rl_method_peers_init": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i, _i]),
    "rl_calc_range_repeat_angles_eval_sensor_model_signalled": (_i, [_vp, _vp, _vp, _vp, C.c_int64, _i, _i, C.POINTER(_i)]),
    "rl_method_peers_wait": (_i, [_vp]),
    "rl_calc_range_repeat_angles_eval_sensor_model_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, C.c_int64, _i, _i, C.c_int64]),
    "rl_debug_get_dt": (_i, [_vp, _vp]),
    "rl_debug_cddt_dims": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp, _vp]),
    "rl_debug_cddt_dump": (_i, [_vp, _vp, _vp]),
    "rl_debug_glt_dump": (_i, [_vp, _vp]),
    "rl_debug_sincosf": (_i, [_vp, _vp, _vp, _i]),
    "rl_debug_set_coop_threshold": (_i, [_vp, _i]),
    "rl_debug_set_persistent": (_i, [_vp, _i]),
    "rl_debug_set_spatial_sort": (_i, [_vp, _i]),
}

_lib = None


class RangeLibError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rangelib_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m range_libc_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RangeLibError(rc, (lib().rl_last_error() or b"").decode("utf-8", "replace"))
    return rc
