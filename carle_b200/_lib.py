"""ctypes binding of libcarle_b200.so (C ABI: include/carle_b200.h).

There is no CPU implementation behind this module: if the shared library is
missing or cannot be loaded, importing the package's compute classes raises —
loudly — instead of degrading to torch ops.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CARLE_B200_LIB") or os.path.join(_HERE, "lib", "libcarle_b200.so")

# mirrors of the enums in include/carle_b200.h
CARLE_OK, CARLE_EINVAL, CARLE_ECUDA, CARLE_ENODEV, CARLE_ERULE = 0, -1, -2, -3, -4
F32, U8, PACKED = 0, 1, 2
CNT_STEP_NUMBER, CNT_STEPS_SINCE_ACTION, CNT_RESETS, CNT_GENERATIONS = 0, 1, 2, 3
CNT_LAST_NOT_ALL_ONES, CNT_LAST_ANY_TOGGLE, CNT_LAST_RESET_COND = 4, 5, 6
RED_LIVE, RED_SH, RED_SW, RED_WINDOW_LIVE = 0, 1, 2, 3
RLE_KEEP_TAIL = 1

_c = ctypes
_vp, _i64, _i32, _u32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_uint32



class StepArgs(_c.Structure):
    """``carle_step_args`` of include/carle_b200.h (argument block of ``carle_step_ex``)."""
    _fields_ = [("struct_size", _c.c_uint32), ("action_dtype", _c.c_int32),
                ("state_in", _vp), ("state_out", _vp), ("action", _vp),
                ("action_batch", _i64), ("counters", _vp), ("reductions", _vp),
                ("reward_zero", _vp), ("obs", _vp), ("obs_dtype", _c.c_int32),
                ("defer_reset", _c.c_int32), ("speed_com_prev", _vp), ("speed_com_next", _vp),
                ("speed_velocity", _vp), ("speed_out", _vp), ("speed_sumsq", _vp),
                ("speed_primed", _vp)]


#: every symbol declared in include/carle_b200.h -> (restype, argtypes)
PROTOTYPES = {
    "carle_version": (_i32, []),
    "carle_last_error": (_c.c_char_p, []),
    "carle_create": (_i32, [_c.POINTER(_vp), _i32, _i64, _i32, _i32, _i32, _i32]),
    "carle_destroy": (_i32, [_vp]),
    "carle_geometry": (_i32, [_vp, _c.POINTER(_c.c_int32 * 8)]),
    "carle_set_rule": (_i32, [_vp, _u32, _u32]),
    "carle_pack_state": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "carle_unpack_state": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "carle_pack_action": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp]),
    "carle_pack_action_host": (_i32, [_i32, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32]),
    "carle_pack_action_host_copy": (_i32, [_i32, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _i32, _vp]),
    "carle_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "carle_step_many": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "carle_step_action": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "carle_step_ex": (_i32, [_vp, _c.POINTER(StepArgs), _vp]),
    "carle_apply_reset": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "carle_band_create": (_i32, [_c.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "carle_band_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "carle_band_push_halos": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "carle_dev_alloc": (_i32, [_i32, _c.c_uint64, _c.POINTER(_vp)]),
    "carle_dev_free": (_i32, [_i32, _vp]),
    "carle_ipc_export": (_i32, [_vp, _vp]),
    "carle_ipc_open": (_i32, [_vp, _c.POINTER(_vp)]),
    "carle_ipc_close": (_i32, [_vp]),
    "carle_random_action": (_i32, [_vp, _c.c_uint64, _u32, _c.c_double, _i64, _vp, _vp]),
    "carle_step_random": (_i32, [_vp, _vp, _vp, _c.c_uint64, _u32, _c.c_double, _i64, _vp, _vp, _vp, _vp]),
    "carle_unpack_action": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "carle_apply_action": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "carle_reduce": (_i32, [_vp, _vp, _vp, _vp]),
    "carle_masked_count": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "carle_action_count": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "carle_speed_tail": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "carle_puffer_tail": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "carle_morpho_match": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "carle_rle_encode_host": (_i64, [_vp, _i32, _i32, _i32, _vp, _i64]),
    "carle_rle_decode_host": (_i32, [_c.c_char_p, _i64, _i32, _i32, _vp]),
    "carle_jit_probe": (_i32, [_i32, _u32, _u32, _c.POINTER(_c.c_int64)]),
    "carle_jit_loaded": (_i32, []),
}


class CarleLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CarleLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m carle_b200.build` "
            "(needs nvcc; sm_100a).  carle_b200 has no CPU or torch fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover
        raise CarleLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    msg = load().carle_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    """Map C error codes onto the exception types the reference raises."""
    if rc == CARLE_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == CARLE_ERULE:
        raise TypeError(msg)              # reference: reduce() of empty sequence
    if rc == CARLE_EINVAL:
        raise ValueError(msg)
    raise CarleLibraryError(msg)
