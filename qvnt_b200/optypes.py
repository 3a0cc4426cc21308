"""POD operator descriptor of the C ABI (`qvnt_op_t`, include/qvnt_b200.h).

`kind` follows the variant order of the reference's `AtomicOpDispatch` enum
(reference: src/operator/atomic/dispatch.rs:82-105).
"""
import ctypes

KIND_NAMES = [
    "Id", "X", "RX", "RXX", "Y", "RY", "RYY", "Z", "S", "T", "RZ", "RZZ",
    "U1", "U2", "H1", "H2", "Swap", "ISwap", "SqrtSwap", "SqrtISwap",
]
(K_ID, K_X, K_RX, K_RXX, K_Y, K_RY, K_RYY, K_Z, K_S, K_T, K_RZ, K_RZZ,
 K_U1, K_U2, K_H1, K_H2, K_SWAP, K_ISWAP, K_SQRTSWAP, K_SQRTISWAP) = range(20)


class QvntOp(ctypes.Structure):
    """Mirror of `qvnt_op_t` (304 bytes)."""
    _fields_ = [
        ("kind", ctypes.c_uint32),
        ("dagger", ctypes.c_uint32),
        ("a_mask", ctypes.c_uint64),
        ("b_mask", ctypes.c_uint64),
        ("ctrl", ctypes.c_uint64),
        ("phase_re", ctypes.c_double),
        ("phase_im", ctypes.c_double),
        ("matrix", ctypes.c_double * 32),
    ]


assert ctypes.sizeof(QvntOp) == 304
