"""ctypes binding of the C ABI in include/qvnt_b200.h (libqvnt_b200.so).

There is no fallback: if the shared library is missing the import fails
loudly, and if no CUDA device is present every register call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

from .optypes import QvntOp

_HERE = os.path.dirname(os.path.abspath(__file__))
# QVNT_B200_LIB selects an experimental build variant of the same library (see csrc/Makefile)
LIB_PATH = os.environ.get("QVNT_B200_LIB") or os.path.join(_HERE, "libqvnt_b200.so")

STATS_CLASSES = 5
IPC_BLOB_BYTES = 256
STATUS_NAMES = {0: "OK", 1: "INVALID", 2: "BAD_MASK", 3: "OOM", 4: "CUDA", 5: "COMM", 6: "UNSUPPORTED"}


class QvntStats(ctypes.Structure):
    _fields_ = [
        ("launches", c_uint64 * STATS_CLASSES),
        ("ms", c_double * STATS_CLASSES),
        ("alg_bytes", c_uint64 * STATS_CLASSES),
        ("ops_applied", c_uint64),
        ("passes", c_uint64),
        ("h2d_bytes", c_uint64),
        ("d2h_bytes", c_uint64),
        ("peer_bytes", c_uint64),
    ]


class QvntError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"qvnt_b200: {STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


# every symbol include/qvnt_b200.h declares: name -> (restype, argtypes)
_REG = c_void_p
PROTOTYPES = {
    "qvnt_version": (c_int, []),
    "qvnt_last_error": (c_char_p, []),
    "qvnt_device_count": (c_int, [POINTER(c_int)]),
    "qvnt_reg_create": (c_int, [c_uint32, c_uint64, POINTER(_REG)]),
    "qvnt_reg_create_sharded": (c_int, [c_uint32, c_uint64, c_uint32, c_uint32, c_int, POINTER(_REG)]),
    "qvnt_reg_create_multi": (c_int, [c_uint32, c_uint64, c_uint32, POINTER(_REG)]),
    "qvnt_reg_set_gpus": (c_int, [_REG, c_uint32, POINTER(_REG)]),
    "qvnt_reg_combine": (c_int, [_REG, _REG, POINTER(_REG)]),
    "qvnt_reg_combine_unitary": (c_int, [_REG, _REG, c_void_p, POINTER(_REG)]),
    "qvnt_reg_linear_composition": (c_int, [_REG, _REG, c_double, c_double, c_double, c_double]),
    "qvnt_reg_sample_all": (c_int, [_REG, c_uint64, c_uint64, c_void_p]),
    "qvnt_reg_export_ipc": (c_int, [_REG, c_void_p]),
    "qvnt_reg_attach_peers": (c_int, [_REG, c_void_p]),
    "qvnt_reg_clone": (c_int, [_REG, POINTER(_REG)]),
    "qvnt_reg_destroy": (c_int, [_REG]),
    "qvnt_reg_q_num": (c_int, [_REG, POINTER(c_uint32)]),
    "qvnt_reg_apply": (c_int, [_REG, POINTER(QvntOp), c_size_t]),
    "qvnt_plan_describe": (c_int, [c_uint32, c_uint32, c_uint32, c_int, c_int, c_int, c_int, POINTER(QvntOp),
                                   c_size_t, c_char_p, c_size_t, POINTER(c_size_t)]),
    "qvnt_reg_norm_sqr": (c_int, [_REG, POINTER(c_double)]),
    "qvnt_reg_probabilities": (c_int, [_REG, c_uint64, c_uint64, c_void_p]),
    "qvnt_reg_polar": (c_int, [_REG, c_uint64, c_uint64, c_void_p]),
    "qvnt_reg_measure_mask": (c_int, [_REG, c_uint64, c_double, POINTER(c_uint64), POINTER(c_uint64)]),
    "qvnt_reg_measure_mask_rng": (c_int, [_REG, c_uint64, POINTER(c_uint64)]),
    "qvnt_reg_collapse": (c_int, [_REG, c_uint64, c_uint64]),
    "qvnt_reg_normalize": (c_int, [_REG]),
    "qvnt_reg_reset": (c_int, [_REG, c_uint64]),
    "qvnt_reg_reset_by_mask": (c_int, [_REG, c_uint64]),
    "qvnt_reg_read": (c_int, [_REG, c_uint64, c_uint64, c_void_p]),
    "qvnt_reg_write": (c_int, [_REG, c_uint64, c_uint64, c_void_p]),
    "qvnt_reg_tensor_prod": (c_int, [_REG, _REG, POINTER(_REG)]),
    "qvnt_reg_sync": (c_int, [_REG]),
    "qvnt_reg_set_option": (c_int, [_REG, c_char_p, c_int64]),
    "qvnt_reg_stats": (c_int, [_REG, POINTER(QvntStats)]),
    "qvnt_reg_stats_reset": (c_int, [_REG]),
    "qvnt_reg_mark": (c_int, [_REG, c_int]),
    "qvnt_reg_elapsed_ms": (c_int, [_REG, c_int, c_int, POINTER(c_double)]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libqvnt_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C qvnt_b200/csrc`. qvnt_b200 has no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)        # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().qvnt_last_error()
        raise QvntError(status, msg.decode() if msg else "")


def device_count() -> int:
    n = c_int(0)
    check(lib().qvnt_device_count(byref(n)))
    return n.value
