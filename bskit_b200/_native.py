"""ctypes binding of the C-ABI library ``libbskit_b200.so`` (include/bskit_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call
fails, the product path raises.  A CPU substitute would void every parity
claim (the float64 oracle lives in ``oracle/`` and is test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbskit_b200.so")

F32, F64 = 0, 1
KIND_DATA, KIND_UNIT, KIND_KPOW = 0, 1, 2
MAX_CHUNK = 64

#: every symbol include/bskit_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = (
    "bsk_version", "bsk_last_error", "bsk_plan_create", "bsk_plan_destroy", "bsk_plan_info",
    "bsk_set_compensation", "bsk_plan_set_stream", "bsk_fold_even", "bsk_forward_local", "bsk_forward_finish", "bsk_modes_per_bin", "bsk_shells_x", "bsk_shells_yz",
    "bsk_shells", "bsk_shells_prepare", "bsk_cplan_create", "bsk_cplan_destroy", "bsk_cplan_info", "bsk_cplan_set_path",
    "bsk_cplan_path", "bsk_tc_schedule_info", "bsk_tc_schedule_eval", "bsk_contract",
    "bsk_reduce_list", "bsk_paint_cic", "bsk_launch_count",
)


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nmesh", "neval", "ncrop", "precision", "world", "rank", "max_shells", "fft_precision", "no_prune", "transposed")]


class Info(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("kx", "ky", "kz", "nx0", "nxl", "mx0", "mxl", "fwd_batch", "fwd_work_complex",
                 "planes_local_complex", "planes_all_complex", "cube_complex",
                 "xcols_complex_per_shell", "planes2d_complex_per_shell", "field_real_per_shell",
                 "fft_work_bytes", "kyl", "ky0", "xplanes_complex_per_shell", "pruned")]


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C bskit_b200/csrc`).  bskit_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    L.bsk_version.restype = C.c_int
    L.bsk_last_error.restype = C.c_char_p
    L.bsk_launch_count.restype = C.c_int64
    L.bsk_plan_create.argtypes = [C.POINTER(vp), C.POINTER(Geometry), dp, dp, dp, vp]
    L.bsk_plan_destroy.argtypes = [vp]
    L.bsk_plan_info.argtypes = [vp, C.POINTER(Info)]
    L.bsk_set_compensation.argtypes = [vp, dp, dp, dp]
    L.bsk_plan_set_stream.argtypes = [vp, vp]
    L.bsk_fold_even.argtypes = [vp, ip, ip, ip, ip, ip, vp, C.c_int64, vp]
    L.bsk_forward_local.argtypes = [vp, vp, ip, vp, vp, vp]
    L.bsk_forward_finish.argtypes = [vp, vp, vp]
    L.bsk_modes_per_bin.argtypes = [vp, ip, dp, dp, C.POINTER(C.c_int64)]
    L.bsk_shells_prepare.argtypes = [vp, ip]
    L.bsk_shells.argtypes = [vp, vp, ip, C.c_double, ip, dp, dp, vp, vp, vp]
    L.bsk_shells_x.argtypes = [vp, vp, ip, C.c_double, ip, dp, dp, vp]
    L.bsk_shells_yz.argtypes = [vp, ip, vp, vp, vp]
    L.bsk_cplan_create.argtypes = [C.POINTER(vp), ip, C.POINTER(C.c_int32), ip, ip]
    L.bsk_cplan_destroy.argtypes = [vp]
    L.bsk_cplan_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.bsk_cplan_set_path.argtypes = [vp, ip]
    L.bsk_cplan_path.argtypes = [vp, C.POINTER(C.c_int64)]
    L.bsk_tc_schedule_info.argtypes = [ip, C.POINTER(C.c_int32), ip, C.POINTER(C.c_int64)]
    L.bsk_tc_schedule_eval.argtypes = [ip, C.POINTER(C.c_int32), ip, C.c_int64, dp, dp]
    L.bsk_contract.argtypes = [vp, C.POINTER(vp), ip, ip, C.c_int64, ip, C.POINTER(C.c_int32), dp, vp]
    L.bsk_reduce_list.argtypes = [C.POINTER(vp), ip, ip, C.c_int64, ip, C.POINTER(C.c_int32), dp, vp]
    L.bsk_paint_cic.argtypes = [vp, ip, C.c_int64, ip, dp, vp, vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().bsk_last_error()
        raise NativeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def dptr(arr):
    """ctypes pointer to a contiguous float64 numpy array (or None)."""
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(C.c_double))


def launch_count():
    return int(lib().bsk_launch_count())
