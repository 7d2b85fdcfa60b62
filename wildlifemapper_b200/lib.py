"""ctypes binding of libwm_b200.so (C ABI declared in include/wm_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WM_LIB_NAME selects an alternative in-tree build (e.g. the -DWM_DEBUG_WAIT diagnostics build)
LIB_PATH = os.path.join(_HERE, os.environ.get("WM_LIB_NAME", "libwm_b200.so"))

WM_ACT = {"none": 0, "gelu": 1, "relu": 2, "sigmoid": 3}

_p, _i64, _i, _f, _d = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double

# name -> argtypes; must list every symbol include/wm_b200.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "wm_version": [],
    "wm_last_error": [],
    "wm_device_check": [],
    "wm_set_flash_version": [_i],
    "wm_set_option": [C.c_char_p, _i],
    "wm_debug_flash_trace": [_p],
    "wm_debug_window_trace": [_p],
    "wm_gemm_bf16": [_p, _i64, _p, _i64, _p, _p, _i64, _i, _p, _i64, _p, _i64, _i, _i, _i, _i, _i, _p],
    "wm_num_sms": [],
    "wm_conv3x3_nhwc_bf16": [_p, _p, _p, _p, _i, _i, _i, _p],
    "wm_layernorm": [_p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _f, _p],
    "wm_patchify": [_p, _p, _p, _i, _i, _i, _p],
    "wm_transpose_split": [_p, _p, _i, _i, _i, _p],
    "wm_transpose": [_p, _p, _i, _i, _i, _i, _p],
    "wm_hfc_finalize": [_p, _p, _p, _p, _i, _p],
    "wm_add_cast": [_p, _p, _i, _p, _i, _i, _p],
    "wm_attn_flash": [_p, _i64, _i64, _i64, _i, _p, _i64, _i64, _i64, _i, _p, _i64, _i64, _i64, _i, _p, _p, _i64,
                      _i, _i, _i, _i, _i, _f, _p],
    "wm_attn_window": [_p, _p, _p, _i, _i, _i, _f, _p],
    "wm_attn_small": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i, _i, _i, _i, _i, _f, _p],
    "wm_postprocess": [_p, _p, _p, _f, _i, _p, _p, _p, _p, _i, _i, _i, _p],
    "wm_sigmoid_topk": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "wm_nms": [_p, _p, _p, _i, _d, _p, _p, _p, _p, _p],
    "wm_nms_batched": [_p, _p, _i, _i, _f, _d, _i, _p, _p, _p],
    "wm_tiles_from_u8": [_p, _i, _i, _i64, _p, _i, _i, _i, _p, _p, _p, _p],
    "wm_resize_tiles_u8": [_p, _i, _i, _i64, _p, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _p, _i, _p],
    "wm_merge_detections": [_p, _p, _p, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p],
    "wm_match_cost": [_p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _p, _p],
    "wm_set_criterion": [_p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _f, _p, _p, _p],
    "wm_pack_coco": [_p, _p, _p, _p, _i, _p, _p, _p],
}

_lib = None


class WmError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WmError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or wildlifemapper_b200/csrc/build.sh). wildlifemapper_b200 has no CPU or PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "wm_last_error" else C.c_int
    _lib = lib
    for env, opt in (("WM_FLASH_VERSION", b"flash_version"), ("WM_GEMM_PAIRS", b"gemm_pairs")):
        if os.environ.get(env):  # measurement knobs, see include/wm_b200.h
            lib.wm_set_option(opt, int(os.environ[env]))
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().wm_last_error()
        raise WmError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
