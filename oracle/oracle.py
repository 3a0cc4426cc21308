"""ctypes wrapper of the CPU oracle (oracle/qvnt_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package
(qvnt_b200/).  `OracleReg` exposes the same methods as `qvnt_b200.QReg` so parity
tests can drive both with one circuit.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_double, c_int, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqvnt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qvnt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libqvnt_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = ctypes.CDLL(LIB_PATH)
        l.qo_reg_create.restype = c_void_p
        l.qo_reg_create.argtypes = [c_uint32, c_uint64, c_int]
        l.qo_reg_destroy.argtypes = [c_void_p]
        l.qo_reg_clone.restype = c_void_p
        l.qo_reg_clone.argtypes = [c_void_p]
        l.qo_reg_len.restype = c_uint64
        l.qo_reg_len.argtypes = [c_void_p]
        l.qo_reg_data.restype = c_void_p
        l.qo_reg_data.argtypes = [c_void_p]
        l.qo_reg_set_threads.argtypes = [c_void_p, c_int]
        l.qo_reg_reset.argtypes = [c_void_p, c_uint64]
        l.qo_reg_apply.restype = c_int
        l.qo_reg_apply.argtypes = [c_void_p, c_void_p, c_uint64]
        l.qo_reg_norm_sqr.restype = c_double
        l.qo_reg_norm_sqr.argtypes = [c_void_p]
        l.qo_reg_probabilities.argtypes = [c_void_p, c_void_p]
        l.qo_reg_normalize.argtypes = [c_void_p]
        l.qo_reg_reset_by_mask.argtypes = [c_void_p, c_uint64]
        l.qo_reg_collapse.argtypes = [c_void_p, c_uint64, c_uint64]
        l.qo_weighted_index.restype = c_uint64
        l.qo_weighted_index.argtypes = [c_void_p, c_uint64, c_double]
        l.qo_reg_measure_mask.restype = c_uint64
        l.qo_reg_measure_mask.argtypes = [c_void_p, c_uint64, c_double, POINTER(c_uint64)]
        l.qo_reg_tensor_prod.restype = c_void_p
        l.qo_reg_tensor_prod.argtypes = [c_void_p, c_void_p]
        l.qo_reg_read.argtypes = [c_void_p, c_uint64, c_uint64, c_void_p]
        l.qo_reg_write.argtypes = [c_void_p, c_uint64, c_uint64, c_void_p]
        l.qo_sweep.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_int]
        l.qo_max_threads.restype = c_int
        l.qo_sizeof_op.restype = c_uint64
        _lib = l
    return _lib


def max_threads() -> int:
    return int(lib().qo_max_threads())


class OracleReg:
    """CPU register following register/quant.rs (out-of-place sweeps)."""

    def __init__(self, q_num: int, state: int = 0, threads: int = 1, _handle=None):
        self._h = _handle if _handle is not None else lib().qo_reg_create(q_num, state, threads)
        if not self._h:
            raise MemoryError("oracle register allocation failed")
        self.q_num = q_num
        self.q_mask = (1 << q_num) - 1
        self.threads = threads

    new = classmethod(lambda cls, q, threads=1: cls(q, 0, threads))
    with_state = classmethod(lambda cls, q, s, threads=1: cls(q, s, threads))

    def close(self):
        if self._h:
            lib().qo_reg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self):
        return OracleReg(self.q_num, threads=self.threads, _handle=lib().qo_reg_clone(self._h))

    def num(self):
        return self.q_num

    def apply(self, op):
        from qvnt_b200.op import SingleOp
        from qvnt_b200.optypes import QvntOp
        if isinstance(op, SingleOp):
            arr, n = (QvntOp * 1)(op.to_c()), 1
        else:
            arr, n = op.to_c_array()
        if n and lib().qo_reg_apply(self._h, arr, n) != 0:
            raise MemoryError("oracle apply: allocation failed")

    def apply_raw(self, arr, n):
        if lib().qo_reg_apply(self._h, arr, n) != 0:
            raise MemoryError("oracle apply: allocation failed")

    def amplitudes(self, off=0, cnt=None):
        n = 1 << self.q_num if cnt is None else cnt
        out = np.empty(n, dtype=np.complex128)
        lib().qo_reg_read(self._h, off, n, out.ctypes.data)
        return out

    def buffer_len(self):
        return int(lib().qo_reg_len(self._h))

    def write_amplitudes(self, data, off=0):
        data = np.ascontiguousarray(data, dtype=np.complex128)
        lib().qo_reg_write(self._h, off, data.size, data.ctypes.data)

    def get_absolute(self):
        return float(lib().qo_reg_norm_sqr(self._h))

    def get_probabilities(self):
        out = np.empty(1 << self.q_num, dtype=np.float64)
        lib().qo_reg_probabilities(self._h, out.ctypes.data)
        return out

    def normalize(self):
        lib().qo_reg_normalize(self._h)
        return self

    def reset(self, state):
        lib().qo_reg_reset(self._h, state)

    def reset_by_mask(self, mask):
        lib().qo_reg_reset_by_mask(self._h, mask)

    def collapse_mask(self, idy, mask):
        lib().qo_reg_collapse(self._h, idy, mask)

    def measure_mask_full(self, mask, u):
        s = c_uint64(0)
        out = lib().qo_reg_measure_mask(self._h, mask, float(u), byref(s))
        return int(out), int(s.value)

    def measure_mask(self, mask, u):
        from qvnt_b200.register import CReg
        return CReg.with_state(self.q_num, self.measure_mask_full(mask, u)[0])

    def __mul__(self, other):
        h = lib().qo_reg_tensor_prod(self._h, other._h)
        return OracleReg(self.q_num + other.q_num, threads=max(self.threads, other.threads), _handle=h)


def weighted_index(weights: np.ndarray, u: float) -> int:
    w = np.ascontiguousarray(weights, dtype=np.float64)
    return int(lib().qo_weighted_index(w.ctypes.data, w.size, float(u)))


# -- crate-private register helpers of the reference, restated in numpy (small, elementwise) -------
def combine(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """quant.rs:245-271: the two state vectors as lower / upper half of a register one qubit larger."""
    return np.concatenate([a, b])


def combine_with_unitary(a: np.ndarray, b: np.ndarray, c) -> np.ndarray:
    """quant.rs:274-307: out[idx] = c[0]*a + c[1]*b where the new top bit is clear, c[2]*a + c[3]*b where set."""
    c = [complex(z) for z in c]
    return np.concatenate([c[0] * a + c[1] * b, c[2] * a + c[3] * b])


def linear_composition(self_psi: np.ndarray, psi: np.ndarray, c) -> np.ndarray:
    """quant.rs:310-328: self[i] = self[i]*c.0 + psi[i]*c.1."""
    return self_psi * complex(c[0]) + psi * complex(c[1])


def sample_all_moments(p: np.ndarray, count: int):
    """quant.rs:513-594: mean and standard deviation of every histogram bin before rounding:
    counts_i = c p_i + sqrt(c) (n_i - p_i sum_j n_j), n_i = sqrt(p_i) g_i, g ~ N(0, 1) i.i.d."""
    c = float(count)
    mean = c * p
    var = c * (p * (1 - p) ** 2 + p ** 2 * (1 - p))          # = c p (1 - p)
    return mean, np.sqrt(var)
