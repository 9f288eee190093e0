"""ctypes binding of libb200vc.so (include/b200vc.h).

The library is the product path: if it is missing or fails to load there is no
CPU fallback — importing callers get a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

MAX_TAPS = 128

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU, ACT_TANH, ACT_SIGMOID, ACT_EXP = range(7)
BACKEND_SIMT, BACKEND_TC, BACKEND_TC_V1, BACKEND_TC_TILE, BACKEND_TC_WS = 0, 1, 2, 3, 4   # 2 = experimental persistent kernel


class Tap(C.Structure):
    _fields_ = [
        ("c_off", C.c_int32),
        ("dw", C.c_int16),
        ("dh", C.c_int16),
        ("dp", C.c_int16),
        ("widx", C.c_int16),
    ]


class TapGemmParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p),
        ("a_dim", C.c_int32 * 5),
        ("a_stride", C.c_int64 * 5),
        ("Wt", C.c_void_p),
        ("ldw", C.c_int64),
        ("wstride", C.c_int64),
        ("w_batch_step", C.c_int32),
        ("Kc", C.c_int32),
        ("N", C.c_int32),
        ("ntaps", C.c_int32),
        ("OW", C.c_int32),
        ("OH", C.c_int32),
        ("OB", C.c_int32),
        ("BW", C.c_int32),
        ("BH", C.c_int32),
        ("osh", C.c_int32),
        ("osw", C.c_int32),
        ("ooh", C.c_int32),
        ("oow", C.c_int32),
        ("o_fh", C.c_int32),
        ("o_fw", C.c_int32),
        ("o_sb", C.c_int64),
        ("o_sh", C.c_int64),
        ("o_sw", C.c_int64),
        ("r_sb", C.c_int64),
        ("r_sh", C.c_int64),
        ("r_sw", C.c_int64),
        ("o_sn", C.c_int64),
        ("r_sn", C.c_int64),
        ("o2_sb", C.c_int64),
        ("o2_sh", C.c_int64),
        ("o2_sw", C.c_int64),
        ("o2_sn", C.c_int64),
        ("out2_own", C.c_int32),
        ("row_scale_pre", C.c_void_p),
        ("bias", C.c_void_p),
        ("bias_per_row", C.c_int32),
        ("act_pre", C.c_int32),
        ("act_pre_p", C.c_float),
        ("row_scale", C.c_void_p),
        ("res", C.c_void_p),
        ("res_op", C.c_int32),
        ("scale", C.c_float),
        ("res2", C.c_void_p),
        ("act_post", C.c_int32),
        ("act_post_p", C.c_float),
        ("out", C.c_void_p),
        ("out2", C.c_void_p),
        ("act2", C.c_int32),
        ("act2_p", C.c_float),
        ("vec4", C.c_int32),
        ("round_tf32", C.c_int32),
        ("dtype", C.c_int32),
        ("split", C.c_int32),
        ("o_split", C.c_int64),
        ("r_split", C.c_int64),
        ("acc_in", C.c_void_p),
        ("ai_sb", C.c_int64),
        ("ai_sh", C.c_int64),
        ("ai_sw", C.c_int64),
        ("taps", Tap * MAX_TAPS),
    ]


_LIB = None
_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libb200vc.so"


def lib_path() -> Path:
    return _LIB_PATH


def _declare(lib):
    lib.b200vc_version.restype = C.c_char_p
    lib.b200vc_last_error.restype = C.c_char_p
    lib.b200vc_launch_count.restype = C.c_int64
    lib.b200vc_sizeof_tapgemm_params.restype = C.c_int64
    lib.b200vc_count_launches.argtypes = [C.c_int64]
    lib.b200vc_count_launches.restype = None
    lib.b200vc_plan_begin.argtypes = [C.POINTER(C.c_void_p)]
    lib.b200vc_plan_end.argtypes = []
    lib.b200vc_plan_size.argtypes = [C.c_void_p]
    lib.b200vc_plan_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200vc_plan_destroy.argtypes = [C.c_void_p]
    for fn in (lib.b200vc_plan_begin, lib.b200vc_plan_end, lib.b200vc_plan_size, lib.b200vc_plan_run, lib.b200vc_plan_destroy):
        fn.restype = C.c_int
    if lib.b200vc_sizeof_tapgemm_params() != C.sizeof(TapGemmParams):
        raise RuntimeError(
            f"ABI mismatch: sizeof(b200vc_tapgemm_params)={lib.b200vc_sizeof_tapgemm_params()} "
            f"but ctypes mirror is {C.sizeof(TapGemmParams)}"
        )
    lib.b200vc_tapgemm.argtypes = [C.POINTER(TapGemmParams), C.c_int, C.c_void_p]
    lib.b200vc_tapgemm.restype = C.c_int
    lib.b200vc_tapgemm_tc_supported.argtypes = [C.POINTER(TapGemmParams)]
    lib.b200vc_tapgemm_tc_supported.restype = C.c_int
    lib.b200vc_tapgemm_ws_applicable.argtypes = [C.POINTER(TapGemmParams)]
    lib.b200vc_tapgemm_ws_applicable.restype = C.c_int
    lib.b200vc_tapgemm_set_rows256.argtypes = [C.c_int]
    lib.b200vc_tapgemm_set_rows256.restype = C.c_int
    # later-declared entry points register themselves via declare_optional()
    for name, (argtypes, restype) in _EXTRA.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype


_EXTRA: dict = {}


def declare(name: str, argtypes, restype=C.c_int):
    """Register the signature of an entry point (called at import time by the op modules)."""
    _EXTRA[name] = (argtypes, restype)
    if _LIB is not None:
        fn = getattr(_LIB, name)
        fn.argtypes = argtypes
        fn.restype = restype


def lib():
    """Load (once) and return the ctypes handle. Raises RuntimeError if the CUDA library is absent."""
    global _LIB
    if _LIB is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} not built: run `python -m aicovergen_b200.build` "
                "(there is no CPU fallback for the b200vc hot path)"
            )
        try:
            handle = C.CDLL(str(_LIB_PATH))
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"cannot load {_LIB_PATH}: {e}") from e
        _declare(handle)
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "b200vc"):
    if rc != 0:
        msg = lib().b200vc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def on_device(fn):
    """Decorator for the public methods of the operator objects: run with the object's own CUDA device current.
    Every C-ABI launch goes to `torch.cuda.current_stream()` of the CURRENT device and the kernels / TMA descriptors are
    created on the current device, so an operator built for cuda:k must not run while cuda:0 is current."""
    import functools

    import torch

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = torch.device(self.device)
        if dev.type != "cuda" or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapper


def launch_count() -> int:
    return int(lib().b200vc_launch_count())
