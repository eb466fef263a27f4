"""ctypes binding of libctrlv_b200.so (the C ABI declared in include/ctrlv_b200.h).

There is no fallback: if the shared library is missing or a call returns an error, a
CtrlvError is raised.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CTRLV_B200_LIB: developer override for A/B runs of two builds on one box (same exported symbols)
LIB_PATH = os.environ.get("CTRLV_B200_LIB") or os.path.join(HERE, "libctrlv_b200.so")

CTRLV_MAX_SRC = 4
CTRLV_MAX_SEG = 20


class CtrlvError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p),
        ("ld_rowbias", C.c_int32),
        ("rb_mode", C.c_int32),
        ("rb_div", C.c_int32),
        ("rb_mod", C.c_int32),
        ("rb_B", C.c_int32),
        ("geglu", C.c_int32),
        ("s_acc", C.c_float),
        ("res1", C.c_void_p),
        ("ld_res1", C.c_int32),
        ("s_res1", C.c_float),
        ("res2", C.c_void_p),
        ("ld_res2", C.c_int32),
        ("s_res2", C.c_float),
        ("out", C.c_void_p),
        ("ld_out", C.c_int32),
        ("out_f32", C.c_void_p),
        ("ld_out_f32", C.c_int32),
        ("n_store", C.c_int32),
        ("gn_sums", C.c_void_p),
        ("gn_rows_per_unit", C.c_int32),
        ("gn_cg", C.c_int32),
        ("gn_c_off", C.c_int32),
        ("gn_units", C.c_int32),
        ("gn_rep", C.c_int32),
        ("rb_off", C.c_int32),
        ("splitk_ws", C.c_void_p),
        ("splitk_bytes", C.c_int64),
    ]


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("sx", C.c_int64), ("sy", C.c_int64),
                ("sz", C.c_int64)]


class Seg(C.Structure):
    _fields_ = [("src", C.c_int32), ("c0", C.c_int32), ("nchunk", C.c_int32), ("dx", C.c_int32),
                ("dy", C.c_int32), ("dz", C.c_int32)]


class IgemmDesc(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int32),
        ("src", Src * CTRLV_MAX_SRC),
        ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32),
        ("nseg", C.c_int32),
        ("seg", Seg * CTRLV_MAX_SEG),
        ("W", C.c_void_p),
        ("N", C.c_int32), ("K", C.c_int32),
        ("bn", C.c_int32),
        ("out_mul_x", C.c_int32), ("out_mul_y", C.c_int32), ("out_off_x", C.c_int32), ("out_off_y", C.c_int32),
        ("out_X", C.c_int32), ("out_Y", C.c_int32),
        ("ep", Epilogue),
    ]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_F = C.c_float

# name -> (restype, argtypes); every symbol include/ctrlv_b200.h declares
SIGNATURES = {
    "ctrlv_last_error": (C.c_char_p, []),
    "ctrlv_version": (C.c_char_p, []),
    "ctrlv_device_check": (_I, []),
    "ctrlv_launch_count": (_L, []),
    "ctrlv_plan_create": (_I, [_P, C.POINTER(_P)]),
    "ctrlv_plan_fork": (_I, []),
    "ctrlv_plan_join": (_I, []),
    "ctrlv_plan_finish": (_I, [_P]),
    "ctrlv_plan_size": (_L, [_P]),
    "ctrlv_plan_run": (_I, [_P, _P]),
    "ctrlv_plan_destroy": (_I, [_P]),
    "ctrlv_memset_zero": (_I, [_P, _L, _P]),
    "ctrlv_igemm": (_I, [C.POINTER(IgemmDesc), _P]),
    "ctrlv_igemm_plan": (_I, [C.POINTER(IgemmDesc), _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ctrlv_igemm_override": (_I, [_I, _I, _I]),
    "ctrlv_igemm_streamk": (_I, [_I]),
    "ctrlv_linear": (_I, [_P, _L, _I, _I, _P, _I, C.POINTER(Epilogue), _P]),
    "ctrlv_feedforward": (_I, [_P, _L, _I, _I, _P, _P, _P, C.POINTER(Epilogue), _P]),
    "ctrlv_feedforward_ln": (_I, [_P, _L, _I, _I, _F, _P, _I, _I, _I, _P, _P, _P, C.POINTER(Epilogue), _P]),
    "ctrlv_linear_ln": (_I, [_P, _L, _I, _I, _F, _P, _I, _I, _I, _P, _I, C.POINTER(Epilogue), _P]),
    "ctrlv_feedforward_override": (_I, [_I]),
    "ctrlv_conv3x3": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I,
                           C.POINTER(Epilogue), _P]),
    "ctrlv_conv_t3": (_I, [_P, _I, _I, _I, _I, _P, _I, C.POINTER(Epilogue), _P]),
    "ctrlv_groupnorm_workspace": (_L, [_I]),
    "ctrlv_groupnorm": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _F, _I, _P, _P, _P]),
    "ctrlv_groupnorm_apply": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _F, _I, _P, _P, _I, _P]),
    "ctrlv_axpby_gn": (_I, [_P, _P, _F, _F, _L, _I, _P, _P, _I, _I, _I, _I, _P]),
    "ctrlv_layernorm": (_I, [_P, _L, _I, _I, _P, _P, _F, _P, _I, _I, _I, _P, _P]),
    "ctrlv_attn_spatial": (_I, [_P, _I, _I, _I, _F, _P, _P]),
    "ctrlv_attn_temporal": (_I, [_P, _I, _I, _I, _I, _F, _P, _P]),
    "ctrlv_cross_attn": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _F, _I, _I, _I, _I, _P, _P]),
    "ctrlv_small_linear": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "ctrlv_sinusoid": (_I, [_P, _I, _I, _I, _P, _P]),
    "ctrlv_prep_input": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "ctrlv_cfg_euler": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "ctrlv_upsample2x": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "ctrlv_axpby": (_I, [_P, _P, _F, _F, _L, _P, _P]),
    "ctrlv_nchw_to_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ctrlv_nhwc_to_nchw": (_I, [_P, _I, _L, _I, _I, _I, _I, _P, _P]),
    "ctrlv_conv3x3_s2_pad01": (_I, [_P, _I, _I, _I, _I, _P, _I, C.POINTER(Epilogue), _P]),
    "ctrlv_upsample2x_conv3x3": (_I, [_P, _I, _I, _I, _I, _P, _I, C.POINTER(Epilogue), _P]),
    "ctrlv_softmax_rows": (_I, [_P, _L, _I, _I, _F, _P, _L, _P]),
    "ctrlv_time_conv_out": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "ctrlv_blur1d_reflect": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "ctrlv_resize_bicubic_ac": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "ctrlv_clip_patchify": (_I, [_P, _I, _I, _I, _I, _I, _F, _F, _I, _P, _P, _I, _P, _P]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the shared library (building it in-tree first if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise CtrlvError(f"{LIB_PATH} not built (run python ctrl-v_b200/build.py)")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ctrlv_build", os.path.join(HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ctrlv_last_error().decode("utf-8", "replace")
        raise CtrlvError(f"{what} failed (code {rc}): {msg}")
