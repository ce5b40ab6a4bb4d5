"""ctypes binding of libkws.so (include/kws.h).  There is no fallback: if the
library is missing or no B200 is present, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KWS_LIBKWS") or os.path.join(HERE, "libkws.so")   # KWS_LIBKWS: e.g. the profiling build libkws_prof.so

FEAT_RAW, FEAT_SPEC, FEAT_LOGMEL, FEAT_MFCC = -1, 0, 1, 2
PREC_FP32, PREC_TC = 0, 1
MAX_VIEWS = 16


class KwsError(RuntimeError):
    pass


class TensorH(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_int64)]


_vp, _i, _f, _d, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64

# name -> (restype, argtypes); mirrors include/kws.h one to one
SIGNATURES = {
    "kws_abi_version": (_i, []),
    "kws_create": (_i, [C.POINTER(_vp), _i, _i]),
    "kws_destroy": (None, [_vp]),
    "kws_last_error": (C.c_char_p, [_vp]),
    "kws_set_precision": (_i, [_vp, _i]),
    "kws_set_fusion": (_i, [_vp, _i]),
    "kws_launch_count": (_i64, [_vp]),
    "kws_timing_enable": (_i, [_vp, _i]),
    "kws_timing_read": (_i, [_vp, C.POINTER(_d), C.POINTER(_i64), _i]),
    "kws_set_noise_bank": (_i, [_vp, _vp, C.POINTER(_i64), _i]),
    "kws_augment": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "kws_augment_pcm16": (_i, [_vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "kws_time_stretch_pcm16": (_i, [_vp, _vp, _i, _d, _vp, _vp]),
    "kws_time_stretch_host_pcm16": (_i, [_vp, _vp, _i, _d, _vp]),
    "kws_frontend_config": (_i, [_vp, _i, _i, _i, _i, _f, _f, _i]),
    "kws_frontend_config_contrib": (_i, [_vp, _i, _i, _i, _f, _f, _i, _i]),
    "kws_frontend_frames": (_i, [_vp]),
    "kws_features": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "kws_model_load": (_i, [_vp, _i, _i, C.POINTER(TensorH), _i]),
    "kws_model_classes": (_i, [_vp, _i]),
    "kws_forward": (_i, [_vp, _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(_f), _i, _vp, _vp, _vp]),
    "kws_debug_activation": (_i, [_vp, _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(_f), _i, _i, _vp, _vp]),
    "kws_convert_classes": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_int32), _i, _vp, _vp, _vp]),
    "kws_select": (_i, [_vp, _vp, _i, _i, _d, _vp, _vp, _vp]),
    "kws_vote": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "kws_predict_host": (_i, [_vp, _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(_f), _i, _vp, _vp]),
    "kws_get_data_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "kws_set_host_staging": (_i, [_vp, _i]),
    "kws_predict_host_pcm16": (_i, [_vp, _i, _vp, _f, _i, C.POINTER(C.c_int32), C.POINTER(_f), _i, _vp, _vp]),
    "kws_pipeline_host_pcm16": (_i, [_vp, _i, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, C.POINTER(C.c_int32),
                                     C.POINTER(_f), _i, _vp, _vp, _vp]),
    "kws_pipeline_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.POINTER(C.c_int32),
                               C.POINTER(_f), _i, _vp, _vp, _vp]),
}

_lib = None


def load():
    """dlopen libkws.so and bind every symbol of include/kws.h (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KwsError(f"{LIB_PATH} not found: build it with `python -m speech_recognition_b200.build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
