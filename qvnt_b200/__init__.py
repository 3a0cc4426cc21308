"""qvnt_b200 -- B200-native state-vector engine behind QVNT's `qvnt::prelude` API.

    from qvnt_b200 import op, QReg, CReg, VReg, MultiOp, SingleOp

    reg = QReg.with_state(20, 0)
    reg.apply(op.qft(0xFFFFF))
    c = reg.measure_mask(0b100)

Everything numerical runs in libqvnt_b200.so (hand-written sm_100a CUDA, C ABI in
include/qvnt_b200.h).  This package is the thin host mirror of the reference's
operator/register interface; importing it does not require a GPU, creating a
register does (there is no CPU fallback).
"""
from . import op                                   # noqa: F401
from .op import MultiOp, SingleOp                  # noqa: F401
from .optypes import QvntOp                        # noqa: F401
from .register import CReg, QReg, VReg             # noqa: F401
from ._ffi import QvntError, device_count, lib     # noqa: F401

__all__ = ["op", "MultiOp", "SingleOp", "QReg", "CReg", "VReg", "QvntOp", "QvntError", "device_count", "lib"]
