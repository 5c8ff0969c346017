"""ctypes binding of ``libpydvs_b200.so`` (``include/pydvs_b200.h``).

There is no CPU path and no fallback: if the shared library has not been built, importing this module
raises.  Build it with ``python pydvs_b200/build.py`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYDVS_B200_LIB selects another build of the same library (kernel tuning experiments: tools/build_variant.py)
LIB_PATH = os.environ.get("PYDVS_B200_LIB") or os.path.join(_HERE, "libpydvs_b200.so")

# error codes / enums of the header
OK = 0
ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_NO_DEVICE = 1001, 1002, 1003
STATUS_LUT_INDEX, STATUS_SLOT_RANGE, STATUS_EVENT_OVERFLOW, STATUS_INDEX = 1, 2, 4, 8
POL_UP, POL_DOWN, POL_MERGED, POL_RECTIFIED = 0, 1, 2, 3
UPD_RATE, UPD_RATE_ADPT, UPD_TIME_BIN_RAW, UPD_TIME_BIN_THR, UPD_NONE = 0, 1, 2, 3, 4
ENC_RATE, ENC_TIME, ENC_TIME_BIN, ENC_TIME_BIN_THR, ENC_NONE = 0, 1, 2, 3, 4
FRAME_I16, FRAME_U8 = 0, 1
STATE_I16, STATE_U8 = 0, 1
ABI_VERSION = 4
PATH_GENERIC, PATH_FAST, PATH_FAST_4X4 = 0, 1, 2


class Limits(C.Structure):
    _fields_ = [("max_tile_px", C.c_int32), ("max_slots", C.c_int32), ("max_width", C.c_int32),
                ("sm_count", C.c_int32)]


class EncParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("uncoded", C.c_int32), ("threshold", C.c_int32),
                ("n_slots", C.c_int32), ("u16_vals", C.c_int32), ("lut_len", C.c_int32),
                ("lut", C.c_void_p)]


class Tiling(C.Structure):
    _fields_ = [("S", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("tile_rows", C.c_int32)]


class StepParams(C.Structure):
    _fields_ = [("tiling", Tiling), ("frame_dtype", C.c_int32), ("adaptive", C.c_int32),
                ("update_mode", C.c_int32), ("uncoded", C.c_int32), ("threshold", C.c_int32),
                ("min_thr", C.c_int32), ("max_thr", C.c_int32), ("thr_down", C.c_int32),
                ("thr_up", C.c_int32), ("max_time_ms", C.c_int32), ("inh_k", C.c_int32),
                ("polarity", C.c_int32), ("force_generic", C.c_int32), ("drop_silent", C.c_int32),
                ("record_stride", C.c_int32), ("state_dtype", C.c_int32),
                ("history_weight", C.c_double), ("enc", EncParams),
                ("curr", C.c_void_p), ("ref", C.c_void_p), ("thr", C.c_void_p),
                ("diff_out", C.c_void_p), ("abs_diff_out", C.c_void_p), ("spikes_out", C.c_void_p),
                ("records", C.c_void_p), ("tile_counts", C.c_void_p), ("status", C.c_void_p),
                ("redo", C.c_void_p)]


class NvsParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved0", C.c_int32), ("n_px", C.c_int64),
                ("frame", C.c_void_p), ("ref", C.c_void_p), ("thr", C.c_void_p), ("spikes", C.c_void_p),
                ("active", C.c_void_p), ("leak_mask", C.c_void_p), ("ref_off", C.c_void_p),
                ("thr_off", C.c_void_p), ("spikes_off", C.c_void_p), ("leak_mask_off", C.c_void_p),
                ("reference_leak", C.c_double), ("threshold_base", C.c_double),
                ("threshold_mult_incr", C.c_double), ("threshold_leak", C.c_double),
                ("target_off", C.c_double)]


class KeyList(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("stream_offsets", C.c_void_p), ("n", C.c_int64),
                ("S", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("flag_shift", C.c_int32),
                ("data_shift", C.c_int32), ("data_mask", C.c_int32)]


NVS_0, NVS_LEAKY, NVS_LEAKY_ADAPTIVE, NVS_NOISY_LEAKY_ADAPTIVE, NVS_BI_NOISY_LEAKY_ADAPTIVE = 0, 1, 2, 3, 4
DECODE_KEY_NEGATIVE, DECODE_INDEX, DECODE_VALUE_NAN, DECODE_VALUE_RANGE = 0, 1, 2, 3
DECODE_NO_ERROR = 0x7f7f7f7f
ANIM_TRAVERSE, ANIM_FADE = 0, 1

_P, _I32, _I64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

_SIGNATURES = {
    "dvs_abi_version": (C.c_int, []),
    "dvs_last_step_path": (C.c_int, []),
    "dvs_last_error_string": (C.c_char_p, []),
    "dvs_get_limits": (C.c_int, [C.POINTER(Limits)]),
    "dvs_choose_tile_rows": (C.c_int, [_I32, _I32, _I32, _I32]),
    "dvs_host_narrow_i16_u8": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "dvs_thresholded_difference_i16": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _I64, _P]),
    "dvs_local_inhibition_i16": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "dvs_update_reference_i16": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32,
                                           _I32, _I32, _D, _P, _I32, _I64, _P, _P]),
    "dvs_update_reference_noise_i16": (C.c_int, [_I32, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _D, _P,
                                                 _I32, _P, _I64, _P, _P, _P]),
    "dvs_step_f32": (C.c_int, [_P, _P, _P, _P, _F, _F, _F, _I64, _P]),
    "dvs_step_fused": (C.c_int, [C.POINTER(StepParams), _P]),
    "dvs_split_tiles": (C.c_int, [C.POINTER(Tiling), _P, _P, _I32, C.POINTER(EncParams), _P, _P, _P, _P,
                                  _P]),
    "dvs_lists_to_tiles": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I32, C.POINTER(EncParams), _P, _P,
                                     _P, _P]),
    "dvs_scan_tiles": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "dvs_scan_tiles_ticket": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "dvs_aer_expand": (C.c_int, [_I32, _I32, _I32, C.POINTER(EncParams), _P, _P, _P, _P, _P, _I64, _P,
                                 _P]),
    "dvs_gather_split": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _I64, _P, _I64, _P]),
    "dvs_keys_from_events": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "dvs_pad_rows": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _P]),
    "dvs_keys_from_events_devn": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "dvs_move_mask_i16": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _P]),
    "dvs_attention_targets_i16": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P]),
    "dvs_animate_i16": (C.c_int, [_I32, _P, _I32, _P, _P, _D, _I32, _I32, _P, _I32, _I32, _I32, _P]),
    "dvs_nvs_step_f32": (C.c_int, [C.POINTER(NvsParams), _P]),
    "dvs_decay_i16": (C.c_int, [_P, _P, _D, _I64, _P]),
    "dvs_decode_rate_i16": (C.c_int, [C.POINTER(KeyList), _P, _I32, _P, _P]),
    "dvs_decode_time_round_i16": (C.c_int, [C.POINTER(KeyList), _I32, _P, _P, _P, _P, _I32, _D, _I32,
                                            C.c_uint32, _I32, _P, _P, _P, _P, _P]),
}

EXPORTS = tuple(_SIGNATURES)


class DVSLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pydvs_b200: %s is missing. This package has no CPU fallback; build the CUDA library with "
            "`python pydvs_b200/build.py` (needs nvcc)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:  # the .so does not export what the header declares
            raise ImportError("pydvs_b200: %s does not export %s (stale build): run `python pydvs_b200/build.py`"
                              % (LIB_PATH, name))
        fn.restype, fn.argtypes = res, args
    if lib.dvs_abi_version() != ABI_VERSION:
        raise ImportError("pydvs_b200: ABI version mismatch, rebuild the library")
    return lib


lib = _load()


def check(rc):
    """Raise for a non-zero return code of any dvs_* entry point."""
    if rc != OK:
        msg = lib.dvs_last_error_string().decode("utf-8", "replace")
        if rc == ERR_INVALID_ARGUMENT:
            raise ValueError("pydvs_b200: " + msg)
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError("pydvs_b200: " + msg)
        raise DVSLibraryError("pydvs_b200: CUDA error %d: %s" % (rc, msg))


def limits():
    out = Limits()
    check(lib.dvs_get_limits(C.byref(out)))
    return out
