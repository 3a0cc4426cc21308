"""Host-side mirror of the reference's operator algebra (`qvnt::prelude::op`).

Pure metadata, exactly like the reference ("gates are lazy",
src/operator/mod.rs:7-9): nothing here touches the device.  A `MultiOp` is
lowered to an array of `qvnt_op_t` and handed to `qvnt_reg_apply` in ONE call so
the device-side planner can fuse adjacent gates.

Mirrors (reference file:line):
  SingleOp {act, ctrl, func}            src/operator/single/mod.rs:43-119
  MultiOp (VecDeque<SingleOp>)          src/operator/multi/mod.rs:63-191
  h(mask) -> H2/H1 list                 src/operator/multi/h.rs:14-45
  qft / qft_swapped                     src/operator/multi/qft.rs:4-52
  op::* constructors, arg order         src/operator/mod.rs:113-522
  constructor validation (Option)       src/operator/single/{pauli,rotate,swap}.rs
  dgr of rotations = negated phase      src/operator/atomic/rx.rs:42-47 (and ry/rz/rxx/ryy/rzz)
  u1/u2 dgr = conjugate transpose       src/math/matrix.rs:32-35,76-96
"""
from __future__ import annotations

import math
import struct
from collections import deque
from typing import Iterable, List, Optional, Sequence

from .optypes import (K_H1, K_H2, K_ID, K_ISWAP, K_RX, K_RXX, K_RY, K_RYY, K_RZ, K_RZZ, K_S,
                      K_SQRTISWAP, K_SQRTSWAP, K_SWAP, K_T, K_U1, K_U2, K_X, K_Y, K_Z, QvntOp)

PI = math.pi
FRAC_PI_2 = math.pi / 2
U64 = (1 << 64) - 1


def _popcount(x: int) -> int:
    return bin(x & U64).count("1")


def _rust_f64(x: float) -> str:
    """Rust `{}`/`{:?}` for f64 in the range the gate names use."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == int(x) and abs(x) < 1e16:
        return ("-" if math.copysign(1.0, x) < 0 else "") + f"{abs(int(x))}.0"
    r = repr(x)
    if "e" in r:  # Rust Display never uses exponent notation
        r = format(x, "f").rstrip("0")
        # re-expand to a round-tripping decimal
        for prec in range(1, 340):
            s = f"{x:.{prec}f}"
            if float(s) == x:
                r = s
                break
    return r


def _rust_complex_debug(re: float, im: float) -> str:
    return f"Complex {{ re: {_rust_f64(re)}, im: {_rust_f64(im)} }}"


class SingleOp:
    """One atomic gate + control mask (reference: src/operator/single/mod.rs:43-47)."""

    __slots__ = ("kind", "dagger", "a_mask", "b_mask", "ctrl", "phase", "matrix", "act")

    def __init__(self, kind, a_mask=0, b_mask=0, phase=(0.0, 0.0), matrix=None, dagger=False,
                 ctrl=0, act=None):
        self.kind = kind
        self.dagger = bool(dagger)
        self.a_mask = a_mask & U64
        self.b_mask = b_mask & U64
        self.ctrl = ctrl & U64
        self.phase = (float(phase[0]), float(phase[1]))
        self.matrix = None if matrix is None else [complex(z) for z in matrix]
        # acts_on(): u2 reports only a_mask (reference quirk, atomic/u2.rs:70-72)
        if act is None:
            if kind == K_H2:
                act = self.a_mask | self.b_mask
            elif kind == K_ID:
                act = 0
            else:
                act = self.a_mask
        self.act = act

    # -- Applicable ---------------------------------------------------------
    def act_on(self) -> int:
        return self.act | self.ctrl

    def c(self, c_mask: int) -> Optional["SingleOp"]:
        """Add control qubits; None when the mask overlaps (single/mod.rs:109-118)."""
        if self.act_on() & c_mask:
            return None
        r = self.clone()
        r.ctrl = self.ctrl | c_mask
        return r

    def dgr(self) -> "SingleOp":
        r = self.clone()
        k = self.kind
        if k in (K_S, K_T, K_ISWAP, K_SQRTSWAP, K_SQRTISWAP):
            r.dagger = not self.dagger
        elif k in (K_RX, K_RXX, K_RY, K_RYY, K_RZ, K_RZZ):
            r.phase = (-self.phase[0], -self.phase[1])          # phase: -self.phase
        elif k == K_U1:
            u = self.matrix
            r.matrix = [u[0].conjugate(), u[2].conjugate(), u[1].conjugate(), u[3].conjugate()]
        elif k == K_U2:
            u = self.matrix
            r.matrix = [u[4 * c + rr].conjugate() for rr in range(4) for c in range(4)]
        return r

    def unwrap(self) -> "SingleOp":
        return self

    def clone(self) -> "SingleOp":
        return SingleOp(self.kind, self.a_mask, self.b_mask, self.phase, self.matrix, self.dagger,
                        self.ctrl, self.act)

    # -- naming (Debug) -----------------------------------------------------
    def _func_name(self) -> str:
        k, a = self.kind, self.a_mask
        if k == K_ID:
            return "Id"
        if k in (K_RX, K_RXX, K_RY, K_RYY, K_RZ, K_RZZ):
            ang = 2.0 * math.atan2(self.phase[1], self.phase[0])
            nm = {K_RX: "RX", K_RXX: "RXX", K_RY: "RY", K_RYY: "RYY", K_RZ: "RZ", K_RZZ: "RZZ"}[k]
            return f"{nm}{a}({_rust_f64(ang)})"
        if k == K_U1:
            m = [_rust_complex_debug(z.real, z.imag) for z in self.matrix]
            return f"U{a}[[{m[0]}, {m[1]}], [{m[2]}, {m[3]}]]"
        if k == K_U2:
            m = [_rust_complex_debug(z.real, z.imag) for z in self.matrix]
            rows = ", ".join("[" + ", ".join(m[4 * r:4 * r + 4]) + "]" for r in range(4))
            return f"U{a | self.b_mask}[{rows}]"
        if k == K_H2:
            return f"H{a | self.b_mask}"
        simple = {K_X: "X", K_Y: "Y", K_Z: "Z", K_S: "S", K_T: "T", K_H1: "H", K_SWAP: "SWAP",
                  K_ISWAP: "iSWAP"}
        if k in simple:
            return f"{simple[k]}{a}"
        if k == K_SQRTSWAP:
            return f"sqrt(SWAP{a})"
        if k == K_SQRTISWAP:
            return f"sqrt(iSWAP{a})"
        raise ValueError(k)

    def name(self) -> str:
        """`(C{ctrl}_)?{gate}{mask}` (single/mod.rs:74-80)."""
        return (f"C{self.ctrl}_" if self.ctrl else "") + self._func_name()

    __repr__ = name

    def __eq__(self, o):
        return (isinstance(o, SingleOp) and self.kind == o.kind and self.dagger == o.dagger
                and self.a_mask == o.a_mask and self.b_mask == o.b_mask and self.ctrl == o.ctrl
                and self.phase == o.phase and self.matrix == o.matrix and self.act == o.act)

    __hash__ = None

    # -- lowering -----------------------------------------------------------
    def to_c(self) -> QvntOp:
        o = QvntOp()
        o.kind = self.kind
        o.dagger = 1 if self.dagger else 0
        o.a_mask = self.a_mask
        o.b_mask = self.b_mask
        o.ctrl = self.ctrl
        o.phase_re, o.phase_im = self.phase
        if self.matrix is not None:
            for i, z in enumerate(self.matrix):
                o.matrix[2 * i] = z.real
                o.matrix[2 * i + 1] = z.imag
        return o

    def singles(self) -> List["SingleOp"]:
        return [self]

    def matrix_repr(self, size: int, reg_factory):
        return _matrix(self, size, reg_factory)


class MultiOp:
    """Queue of SingleOps applied front-first (multi/mod.rs:63,96-114)."""

    def __init__(self, ops: Iterable[SingleOp] = ()):
        self.ops = deque(ops)

    @staticmethod
    def from_single(s: SingleOp) -> "MultiOp":
        # From<SingleOp> drops "Id" (multi/mod.rs:135-146)
        return MultiOp([] if s.name() == "Id" else [s])

    def __len__(self):
        return len(self.ops)

    def __iter__(self):
        return iter(self.ops)

    def __getitem__(self, i):
        return self.ops[i]

    def append(self, other: "MultiOp"):
        self.ops.extend(other.ops)
        other.ops.clear()

    def push_back(self, s: SingleOp):
        self.ops.append(s)

    def __mul__(self, rhs):
        r = MultiOp(self.ops)
        r *= rhs
        return r

    def __imul__(self, rhs):
        if isinstance(rhs, SingleOp):
            rhs = MultiOp.from_single(rhs)
        self.ops.extend(rhs.ops)
        return self

    def act_on(self) -> int:
        m = 0
        for s in self.ops:
            m |= s.act_on()
        return m

    def dgr(self) -> "MultiOp":
        return MultiOp([s.dgr() for s in reversed(self.ops)])

    def c(self, c_mask: int) -> Optional["MultiOp"]:
        if self.act_on() & c_mask:
            return None
        return MultiOp([s.c(c_mask) for s in self.ops])

    def unwrap(self) -> "MultiOp":
        return self

    def clone(self) -> "MultiOp":
        return MultiOp([s.clone() for s in self.ops])

    def ends_with(self, suffix: "MultiOp") -> bool:
        return all(a == b for a, b in zip(reversed(self.ops), reversed(suffix.ops)))

    def singles(self) -> List[SingleOp]:
        return list(self.ops)

    def to_c_array(self):
        arr = (QvntOp * max(1, len(self.ops)))()
        for i, s in enumerate(self.ops):
            arr[i] = s.to_c()
        return arr, len(self.ops)

    def __eq__(self, o):
        return isinstance(o, MultiOp) and list(self.ops) == list(o.ops)

    __hash__ = None

    def __repr__(self):
        return "[" + ", ".join(s.name() for s in self.ops) + "]"

    def matrix_repr(self, size: int, reg_factory):
        return _matrix(self, size, reg_factory)


def _matrix(op, size: int, reg_factory):
    """`Applicable::matrix` (applicable.rs:17-46): apply to each basis vector,
    transpose.  `reg_factory(q_num, state)` returns a register exposing
    apply()/amplitudes(); used with the oracle and with the device register."""
    dim = 1 << size
    cols = []
    for idx in range(dim):
        reg = reg_factory(size, idx)
        reg.apply(op)
        cols.append(list(reg.amplitudes()[:dim]))
    return [[cols[j][i] for j in range(dim)] for i in range(dim)]


# ---------------------------------------------------------------------------
# single-gate constructors (Option-returning, reference single/*.rs)
# ---------------------------------------------------------------------------
def _half_phase(phase: float):
    h = phase / 2.0
    return (math.cos(h), math.sin(h))


def _checked(op: SingleOp, bits: int) -> Optional[SingleOp]:
    masks_ok = _popcount(op.a_mask) == bits
    return op if masks_ok else None


class single:
    """Namespace mirroring `operator::single::{pauli,rotate,swap}`."""

    @staticmethod
    def x(a): return SingleOp(K_X, a)
    @staticmethod
    def y(a): return SingleOp(K_Y, a)
    @staticmethod
    def z(a): return SingleOp(K_Z, a)
    @staticmethod
    def s(a): return SingleOp(K_S, a)
    @staticmethod
    def t(a): return SingleOp(K_T, a)
    @staticmethod
    def h1(a): return SingleOp(K_H1, a)
    @staticmethod
    def h2(a, b): return SingleOp(K_H2, a, b)

    @staticmethod
    def rx(a, phase): return _checked(SingleOp(K_RX, a, phase=_half_phase(phase)), 1)
    @staticmethod
    def ry(a, phase): return _checked(SingleOp(K_RY, a, phase=_half_phase(phase)), 1)
    @staticmethod
    def rz(a, phase): return _checked(SingleOp(K_RZ, a, phase=_half_phase(phase)), 1)
    @staticmethod
    def rxx(ab, phase): return _checked(SingleOp(K_RXX, ab, phase=(math.cos(phase * 0.5), math.sin(phase * 0.5))), 2)
    @staticmethod
    def ryy(ab, phase): return _checked(SingleOp(K_RYY, ab, phase=_half_phase(phase)), 2)
    @staticmethod
    def rzz(ab, phase): return _checked(SingleOp(K_RZZ, ab, phase=_half_phase(phase)), 2)
    @staticmethod
    def swap(ab): return _checked(SingleOp(K_SWAP, ab), 2)
    @staticmethod
    def sqrt_swap(ab): return _checked(SingleOp(K_SQRTSWAP, ab), 2)
    @staticmethod
    def i_swap(ab): return _checked(SingleOp(K_ISWAP, ab), 2)
    @staticmethod
    def sqrt_i_swap(ab): return _checked(SingleOp(K_SQRTISWAP, ab), 2)

    @staticmethod
    def u1(a, matrix: Sequence[complex]) -> Optional[SingleOp]:
        m = [complex(z) for z in matrix]
        if _popcount(a) != 1 or not is_unitary_m1(m):
            return None
        return SingleOp(K_U1, a, matrix=m)

    @staticmethod
    def u2(a, b, matrix: Sequence[complex]) -> Optional[SingleOp]:
        m = [complex(z) for z in matrix]
        if _popcount(a) != 1 or _popcount(b) != 1 or not is_unitary_m2(m):
            return None
        return SingleOp(K_U2, a, b, matrix=m)


# -- unitarity checks (math/matrix.rs:24-75, approx_cmp.rs: 2-ULP float compare)
def _ordered_bits(x: float) -> int:
    (i,) = struct.unpack("<q", struct.pack("<d", x))
    return i if i >= 0 else -(i & 0x7FFFFFFFFFFFFFFF)


def _approx_eq(a: float, b: float) -> bool:
    """float_cmp 0.8 `approx_eq!(f64, a, b, ulps = 2)`: equal, or |a-b| <= EPSILON,
    or at most 2 ULPs apart (math/approx_cmp.rs:5-10)."""
    if a == b:
        return True
    if abs(a - b) <= 2.220446049250313e-16:
        return True
    return abs(_ordered_bits(a) - _ordered_bits(b)) <= 2


def _nsq(z: complex) -> float:
    return z.real * z.real + z.imag * z.imag


def is_unitary_m1(u) -> bool:
    e00 = _nsq(u[0]) + _nsq(u[1])
    e11 = _nsq(u[2]) + _nsq(u[3])
    e01 = u[0] * u[2].conjugate() + u[1] * u[3].conjugate()
    return _approx_eq(e00, 1.0) and _approx_eq(e11, 1.0) and _approx_eq(e01.real + e01.imag, 0.0)


def is_unitary_m2(u) -> bool:
    def hm(i, j):
        i, j = (i << 2) & 0xF, (j << 2) & 0xF
        if i == j:
            return complex((_nsq(u[i]) + _nsq(u[1 | i])) + (_nsq(u[2 | i]) + _nsq(u[3 | i])), 0.0)
        return ((u[i] * u[j].conjugate() + u[1 | i] * u[1 | j].conjugate())
                + (u[2 | i] * u[2 | j].conjugate() + u[3 | i] * u[3 | j].conjugate()))
    for i in range(4):
        if not _approx_eq(hm(i, i).real, 1.0):
            return False
    for i in range(4):
        for j in range(i + 1, 4):
            e = hm(i, j)
            if not _approx_eq(e.real + e.imag, 0.0):
                return False
    return True


# ---------------------------------------------------------------------------
# public `op::*` (reference src/operator/mod.rs:113-522); argument order (phase, mask)
# ---------------------------------------------------------------------------
def _expect(s: Optional[SingleOp], msg: str) -> MultiOp:
    if s is None:
        raise ValueError(msg)      # reference: .expect(msg) panics
    return MultiOp.from_single(s)


def id() -> MultiOp:  # noqa: A001 - mirrors op::id()
    return MultiOp()


def x(a_mask): return MultiOp.from_single(single.x(a_mask))
def y(a_mask): return MultiOp.from_single(single.y(a_mask))
def z(a_mask): return MultiOp.from_single(single.z(a_mask))
def s(a_mask): return MultiOp.from_single(single.s(a_mask))
def t(a_mask): return MultiOp.from_single(single.t(a_mask))
def rx(phase, a_mask): return _expect(single.rx(a_mask, phase), "Mask should contain 1 bit!")
def ry(phase, a_mask): return _expect(single.ry(a_mask, phase), "Mask should contain 1 bit!")
def rz(phase, a_mask): return _expect(single.rz(a_mask, phase), "Mask should contain 1 bit!")
def rxx(phase, ab_mask): return _expect(single.rxx(ab_mask, phase), "Mask should contain 2 bit!")
def ryy(phase, ab_mask): return _expect(single.ryy(ab_mask, phase), "Mask should contain 2 bit!")
def rzz(phase, ab_mask): return _expect(single.rzz(ab_mask, phase), "Mask should contain 2 bit!")
def swap(ab_mask): return _expect(single.swap(ab_mask), "Mask should contain 2 bit!")
def sqrt_swap(ab_mask): return _expect(single.sqrt_swap(ab_mask), "Mask should contain 2 bit!")
def i_swap(ab_mask): return _expect(single.i_swap(ab_mask), "Mask should contain 2 bit!")
def sqrt_i_swap(ab_mask): return _expect(single.sqrt_i_swap(ab_mask), "Mask should contain 2 bit!")


def h(a_mask: int) -> MultiOp:
    """Pair set bits low->high into H2(hi, lo), trailing H1 (multi/h.rs:14-45)."""
    count = _popcount(a_mask)
    if count == 0:
        return MultiOp()
    if count == 1:
        return MultiOp.from_single(single.h1(a_mask))
    res = MultiOp()
    bit, first, is_first = 1, 0, True
    while bit <= a_mask:
        if bit & a_mask:
            if is_first:
                first, is_first = bit, False
            else:
                res.push_back(single.h2(bit, first))
                is_first = True
        bit <<= 1
    if not is_first:
        res.push_back(single.h1(first))
    return res


def u1(lam, a_mask): return rz(lam, a_mask)                                   # mod.rs:472
def u2(phi, lam, a_mask): return rz(lam, a_mask) * ry(FRAC_PI_2, a_mask) * rz(phi, a_mask)   # :480
def u3(the, phi, lam, a_mask): return rz(lam, a_mask) * ry(the, a_mask) * rz(phi, a_mask)    # :499


def qft(a_mask: int) -> MultiOp:
    """H1(v[i]) then RZ(v[i+j], pi*0.5^j).c(v[i]) for j=1.. (multi/qft.rs:4-33)."""
    count = _popcount(a_mask)
    if count == 0:
        return MultiOp()
    if count == 1:
        return h(a_mask)
    vec = [1 << i for i in range(64) if (a_mask >> i) & 1]
    res = MultiOp()
    for i in range(count - 1):
        res.append(h(vec[i]))
        for j in range(1, count - i):
            res.push_back(single.rz(vec[i + j], PI * (0.5 ** j)).c(vec[i]))
    res.append(h(vec[count - 1]))
    return res


def qft_swapped(a_mask: int) -> MultiOp:
    vec = [1 << i for i in range(64) if (a_mask >> i) & 1]
    swaps = MultiOp()
    n = len(vec)
    for i in range(n >> 1):
        swaps *= single.swap(vec[i] | vec[n - i - 1])
    return qft(a_mask) * swaps


def bench_circuit() -> MultiOp:
    """operator/mod.rs:524-535 (test-only helper in the reference)."""
    return (MultiOp() * h(0b111) * h(0b100).c(0b001) * x(0b001).c(0b110) * rx(1.2, 0b100)
            * rz(1.0, 0b010).c(0b001) * h(0b001).c(0b100) * z(0b010) * rxx(math.pi / 6, 0b101))
