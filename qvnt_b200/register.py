"""Host-side mirror of `qvnt::register::{QReg, CReg, VReg}`.

`QReg` owns an opaque device handle; every method is one C-ABI call
(include/qvnt_b200.h).  `CReg`/`VReg` are host-only integers and masks, kept
bit-exact with the reference.

Mirrors (reference file:line, relative to /root/reference/src):
  QReg                     register/quant.rs:103-636
  CReg                     register/class.rs:9-121
  VReg                     register/virtl.rs:11-87
"""
from __future__ import annotations

import ctypes
import math
import os
from ctypes import byref, c_double, c_uint32, c_uint64, c_void_p
from typing import Callable, Iterable, List, Optional, Union

import numpy as np

from . import _ffi
from .op import MultiOp, SingleOp, _rust_complex_debug

U64 = (1 << 64) - 1
MAX_LEN_TO_DISPLAY = 8      # quant.rs:16


def _bits_iter(mask: int) -> List[int]:
    """math/bits_iter.rs:1-27: set bits low -> high as single-bit masks."""
    return [1 << i for i in range(64) if (mask >> i) & 1]


class CReg:
    """Classical register (register/class.rs)."""

    def __init__(self, q_num: int, state: int = 0):
        self.q_num = q_num
        self.q_mask = ((1 << q_num) - 1) & U64 if q_num < 64 else U64
        self.value = state          # with_state stores `state` unmasked (class.rs:24)

    new = classmethod(lambda cls, q_num: cls(q_num, 0))
    with_state = classmethod(lambda cls, q_num, state: cls(q_num, state))

    def num(self) -> int:
        return self.q_num

    def reset(self, i_state: int):
        self.value = i_state & self.q_mask

    def set(self, bit: bool, mask: int):
        self.value = (self.value | mask) if bit else (self.value & ~mask & U64)

    def xor(self, bit: bool, mask: int):
        if bit:
            self.value ^= mask

    def get(self) -> int:
        return self.value

    def get_by_mask(self, mask: int) -> int:
        out = 0
        for idx, val in enumerate(_bits_iter(mask & self.q_mask)):
            if self.value & val:
                out |= 1 << idx
        return out

    def __mul__(self, other: "CReg") -> "CReg":
        return CReg(self.q_num + other.q_num, self.value | (other.value << self.q_num))

    def __eq__(self, o):
        return isinstance(o, CReg) and (self.value, self.q_num) == (o.value, o.q_num)

    __hash__ = None

    def __repr__(self):
        s = ""
        for b in _bits_iter(self.q_mask):
            s = ("1" if b & self.value else "0") + s
        return f"({s})"


class VReg:
    """Mask-building sugar (register/virtl.rs): v[3], v[[0,7]], v[:], v[callable]."""

    def __init__(self, num: Optional[int] = None, mask: Optional[int] = None):
        if mask is None:
            mask = ((1 << num) - 1) & U64
        self.bits = _bits_iter(mask)

    @classmethod
    def new_with_mask(cls, mask: int) -> "VReg":
        return cls(mask=mask)

    def __getitem__(self, key: Union[int, slice, Iterable[int], Callable[[int], bool]]) -> int:
        if isinstance(key, int):
            return self.bits[key]
        if isinstance(key, slice):
            if key == slice(None):
                return self[lambda _i: True]
            raise TypeError("only v[:] is supported (RangeFull)")
        if callable(key):
            out = 0
            for i, b in enumerate(self.bits):
                if key(i):
                    out |= b
            return out
        idx = list(key)
        return self[lambda i: i in idx]


class QReg:
    """Quantum register resident in B200 HBM (mirror of register/quant.rs `Reg`)."""

    def __init__(self, q_num: int, state: int = 0, *, _handle=None):
        self._h = c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            _ffi.check(_ffi.lib().qvnt_reg_create(q_num, state & U64, byref(self._h)))
        self.q_num = q_num
        self.q_mask = ((1 << q_num) - 1) & U64
        self.rank, self.world = 0, 1
        self.group = False          # True: one handle over several GPUs (QReg.multi / num_threads)

    # -- constructors -------------------------------------------------------
    @classmethod
    def new(cls, q_num: int) -> "QReg":                       # quant.rs:113
        return cls(q_num, 0)

    @classmethod
    def with_state(cls, q_num: int, state: int) -> "QReg":    # quant.rs:129
        return cls(q_num, state)

    @classmethod
    def sharded(cls, q_num: int, state: int, rank: int, world: int, device: int = -1) -> "QReg":
        """One shard of a register split by its top log2(world) qubits (one per GPU)."""
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_create_sharded(q_num, state & U64, rank, world, device, byref(h)))
        r = cls(q_num, _handle=h)
        r.rank, r.world = rank, world
        return r

    def export_ipc(self) -> bytes:
        buf = ctypes.create_string_buffer(_ffi.IPC_BLOB_BYTES)
        _ffi.check(_ffi.lib().qvnt_reg_export_ipc(self._h, buf))
        return buf.raw

    def attach_peers(self, blobs: List[bytes]):
        raw = b"".join(blobs)
        assert len(raw) == self.world * _ffi.IPC_BLOB_BYTES
        buf = ctypes.create_string_buffer(raw, len(raw))
        _ffi.check(_ffi.lib().qvnt_reg_attach_peers(self._h, buf))

    @classmethod
    def multi(cls, q_num: int, state: int = 0, n_gpus: int = 1) -> "QReg":
        """The register sharded by its top log2(n_gpus) qubits over n_gpus GPUs of this box, driven
        from this one process through one handle (qvnt_reg_create_multi)."""
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_create_multi(q_num, state & U64, n_gpus, byref(h)))
        r = cls(q_num, _handle=h)
        r.world = max(1, n_gpus)
        r.group = n_gpus > 1
        return r

    def num_threads(self, num_threads: int) -> Optional["QReg"]:
        """quant.rs:186-200 with GPUs for threads: the register continues on `num_threads` GPUs
        (1, 2, 4 or 8 of this box: sharded by its top qubits, state kept).  Like the reference it
        consumes `self` and answers None for 0 or for more than the box has."""
        n = int(num_threads)
        avail = 8 if os.environ.get("QVNT_MULTI_SHARE_DEVICES") else _ffi.device_count()
        if n == 0 or n & (n - 1) or n > 8 or n > avail:
            return None
        if n == self.world and self.rank == 0:
            return self
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_set_gpus(self._h, n, byref(h)))
        r = QReg(self.q_num, _handle=h)
        r.world = n
        r.group = n > 1
        self.close()
        return r

    def clone(self) -> "QReg":
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_clone(self._h, byref(h)))
        return QReg(self.q_num, _handle=h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _ffi.lib().qvnt_reg_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- hot path -------------------------------------------------------------
    def num(self) -> int:
        return self.q_num

    def apply(self, op: Union[MultiOp, SingleOp]):                # quant.rs:376
        if isinstance(op, SingleOp):
            arr = (_ffi.QvntOp * 1)(op.to_c())
            n = 1
        else:
            arr, n = op.to_c_array()
        if n:
            _ffi.check(_ffi.lib().qvnt_reg_apply(self._h, arr, n))

    def apply_raw(self, arr, n: int):
        """Apply a pre-lowered qvnt_op_t array (bench: lowering outside the timed region)."""
        _ffi.check(_ffi.lib().qvnt_reg_apply(self._h, arr, n))

    def measure_mask(self, mask: int, u: Optional[float] = None) -> CReg:   # quant.rs:490
        out = c_uint64(0)
        if u is None:
            _ffi.check(_ffi.lib().qvnt_reg_measure_mask_rng(self._h, mask & U64, byref(out)))
        else:
            _ffi.check(_ffi.lib().qvnt_reg_measure_mask(self._h, mask & U64, float(u), byref(out), None))
        return CReg.with_state(self.q_num, out.value)

    def measure_mask_full(self, mask: int, u: float):
        """(outcome, sampled full index) -- harness helper."""
        out, smp = c_uint64(0), c_uint64(0)
        _ffi.check(_ffi.lib().qvnt_reg_measure_mask(self._h, mask & U64, float(u), byref(out), byref(smp)))
        return out.value, smp.value

    def measure(self, u: Optional[float] = None) -> CReg:         # quant.rs:505
        return self.measure_mask(self.q_mask, u)

    def collapse_mask(self, idy: int, mask: int):                 # quant.rs:468
        _ffi.check(_ffi.lib().qvnt_reg_collapse(self._h, idy & U64, mask & U64))

    def normalize(self) -> "QReg":                                # quant.rs:397
        _ffi.check(_ffi.lib().qvnt_reg_normalize(self._h))
        return self

    def reset(self, i_state: int):                                # quant.rs:202
        _ffi.check(_ffi.lib().qvnt_reg_reset(self._h, i_state & U64))

    def reset_by_mask(self, mask: int):                           # quant.rs:207
        _ffi.check(_ffi.lib().qvnt_reg_reset_by_mask(self._h, mask & U64))

    def get_absolute(self) -> float:                              # quant.rs:458
        out = c_double(0.0)
        _ffi.check(_ffi.lib().qvnt_reg_norm_sqr(self._h, byref(out)))
        return out.value

    def _local_range(self):
        if self.group:
            return 0, 1 << self.q_num
        n_local = self.q_num - (self.world.bit_length() - 1)
        return self.rank << n_local, 1 << n_local

    def get_probabilities(self) -> np.ndarray:                    # quant.rs:434
        off, cnt = self._local_range()
        out = np.empty(cnt, dtype=np.float64)
        _ffi.check(_ffi.lib().qvnt_reg_probabilities(self._h, off, cnt, out.ctypes.data))
        return out

    def get_polar(self) -> np.ndarray:                            # quant.rs:417
        off, cnt = self._local_range()
        out = np.empty((cnt, 2), dtype=np.float64)
        _ffi.check(_ffi.lib().qvnt_reg_polar(self._h, off, cnt, out.ctypes.data))
        return out

    def amplitudes(self, off: Optional[int] = None, cnt: Optional[int] = None) -> np.ndarray:
        """Copy amplitudes (this rank's shard by default) to the host as complex128."""
        lo, n = self._local_range()
        off = lo if off is None else off
        cnt = n if cnt is None else cnt
        out = np.empty(cnt, dtype=np.complex128)
        _ffi.check(_ffi.lib().qvnt_reg_read(self._h, off, cnt, out.ctypes.data))
        return out

    def write_amplitudes(self, data: np.ndarray, off: Optional[int] = None):
        lo, _ = self._local_range()
        data = np.ascontiguousarray(data, dtype=np.complex128)
        _ffi.check(_ffi.lib().qvnt_reg_write(self._h, lo if off is None else off, data.size, data.ctypes.data))

    def sample_all(self, count: int, seed: Optional[int] = None) -> np.ndarray:
        """quant.rs:513-594: the histogram of `count` shots in the reference's Gaussian approximation
        (no collapse), computed on the device (qvnt_reg_sample_all); statistical, not bit, parity --
        the reference draws its normals from thread_rng."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        out = np.empty(1 << self.q_num, dtype=np.uint64)
        _ffi.check(_ffi.lib().qvnt_reg_sample_all(self._h, int(count), seed & U64, out.ctypes.data))
        return out

    # -- crate-private helpers of the reference (quant.rs:245-328) --------------------------------
    @staticmethod
    def combine(a: "QReg", b: "QReg") -> Optional["QReg"]:
        if a.q_num != b.q_num:
            return None
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_combine(a._h, b._h, byref(h)))
        return QReg(a.q_num + 1, _handle=h)

    @staticmethod
    def combine_with_unitary(a: "QReg", b: "QReg", c) -> Optional["QReg"]:
        if a.q_num != b.q_num:
            return None
        m = np.ascontiguousarray(np.asarray(c, dtype=np.complex128).reshape(4))
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_combine_unitary(a._h, b._h, m.ctypes.data, byref(h)))
        return QReg(a.q_num + 1, _handle=h)

    def linear_composition(self, other: "QReg", c) -> None:
        c0, c1 = complex(c[0]), complex(c[1])
        _ffi.check(_ffi.lib().qvnt_reg_linear_composition(self._h, other._h, c0.real, c0.imag, c1.real, c1.imag))

    def __mul__(self, other: "QReg") -> "QReg":                   # tensor_prod quant.rs:330
        h = c_void_p()
        _ffi.check(_ffi.lib().qvnt_reg_tensor_prod(self._h, other._h, byref(h)))
        return QReg(self.q_num + other.q_num, _handle=h)

    # -- tuning / instrumentation -------------------------------------------
    def set_option(self, key: str, value: int):
        _ffi.check(_ffi.lib().qvnt_reg_set_option(self._h, key.encode(), int(value)))

    def stats(self) -> dict:
        s = _ffi.QvntStats()
        _ffi.check(_ffi.lib().qvnt_reg_stats(self._h, byref(s)))
        return {"launches": list(s.launches), "ms": list(s.ms), "alg_bytes": list(s.alg_bytes), "ops_applied": s.ops_applied,
                "passes": s.passes, "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes,
                "peer_bytes": s.peer_bytes}

    def stats_reset(self):
        _ffi.check(_ffi.lib().qvnt_reg_stats_reset(self._h))

    def sync(self):
        _ffi.check(_ffi.lib().qvnt_reg_sync(self._h))

    def mark(self, slot: int):
        _ffi.check(_ffi.lib().qvnt_reg_mark(self._h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = c_double(0.0)
        _ffi.check(_ffi.lib().qvnt_reg_elapsed_ms(self._h, a, b, byref(ms)))
        return ms.value

    def __repr__(self):
        """Rust `{:?}` of the register (quant.rs:603-623)."""
        n = min(1 << self.q_num, MAX_LEN_TO_DISPLAY)
        a = self.amplitudes(0, n)
        body = ", ".join(f"{i}: {_rust_complex_debug(float(z.real), float(z.imag))}" for i, z in enumerate(a))
        tail = "" if (1 << self.q_num) <= MAX_LEN_TO_DISPLAY else ", .."
        return f"QReg {{ {body}{tail} }}"
