"""Deterministic synthetic circuits for BASELINE.json's configs (host metadata only).

Built with the public `op::*` mirror, so the op lists are exactly what a user of the reference
would hand to `QReg::apply` (reference: src/operator/mod.rs:113-522).
"""
from __future__ import annotations

import math

from . import op
from .op import MultiOp

MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def angle(self) -> float:
        return 2.0 * math.pi * (self.next() >> 11) * (2.0 ** -53)


def random_layered(n: int, depth: int, seed: int = 0x51564E54) -> MultiOp:
    """configs[1] / configs[3]: per layer one of {h, rx, ry, rz} on every qubit, then a brick
    pattern of controlled-x: x(1<<(q+1)).c(1<<q) for q = l mod 2."""
    rng = SplitMix64(seed)
    circ = MultiOp()
    for layer in range(depth):
        for q in range(n):
            kind = rng.next() & 3
            theta = rng.angle()
            m = 1 << q
            if kind == 0:
                circ *= op.h(m)
            elif kind == 1:
                circ *= op.rx(theta, m)
            elif kind == 2:
                circ *= op.ry(theta, m)
            else:
                circ *= op.rz(theta, m)
        for q in range(layer & 1, n - 1, 2):
            circ *= op.x(1 << (q + 1)).c(1 << q)
    return circ


def qft_full(n: int) -> MultiOp:
    """configs[0] / configs[2]: op::qft over all n qubits."""
    return op.qft((1 << n) - 1)


def qft_plus_h(n: int) -> MultiOp:
    """configs[2]: full QFT followed by the Hadamard transform."""
    return op.qft((1 << n) - 1) * op.h((1 << n) - 1)


def mixed_all_kinds(n: int, layers: int, seed: int = 7) -> MultiOp:
    """Every one of the 20 atomic kinds with random masks/controls/daggers (parity tests)."""
    rng = SplitMix64(seed)
    circ = MultiOp()

    def bit():
        return 1 << (rng.next() % n)

    def two():
        a = bit()
        b = bit()
        while b == a:
            b = bit()
        return a, b

    def mask():
        m = rng.next() & ((1 << n) - 1)
        return m or 1

    for _ in range(layers):
        k = rng.next() % 20
        th = rng.angle()
        if k == 0:
            g = op.x(mask())
        elif k == 1:
            g = op.y(mask())
        elif k == 2:
            g = op.z(mask())
        elif k == 3:
            g = op.s(mask())
        elif k == 4:
            g = op.t(mask())
        elif k == 5:
            g = op.h(mask())
        elif k == 6:
            g = op.rx(th, bit())
        elif k == 7:
            g = op.ry(th, bit())
        elif k == 8:
            g = op.rz(th, bit())
        elif k == 9:
            a, b = two()
            g = op.rxx(th, a | b)
        elif k == 10:
            a, b = two()
            g = op.ryy(th, a | b)
        elif k == 11:
            a, b = two()
            g = op.rzz(th, a | b)
        elif k == 12:
            a, b = two()
            g = op.swap(a | b)
        elif k == 13:
            a, b = two()
            g = op.i_swap(a | b)
        elif k == 14:
            a, b = two()
            g = op.sqrt_swap(a | b)
        elif k == 15:
            a, b = two()
            g = op.sqrt_i_swap(a | b)
        elif k == 16:
            g = op.u3(th, rng.angle(), rng.angle(), bit())
        elif k == 17:
            c, s = math.cos(th), math.sin(th)
            ph = complex(math.cos(rng.angle()), 0)
            g = MultiOp.from_single(op.SingleOp(op.K_U1, bit(), matrix=[c * ph, -s * ph, s * ph, c * ph]))
        elif k == 18:
            a, b = two()
            c, s = math.cos(th), math.sin(th)
            g = MultiOp.from_single(op.SingleOp(op.K_U2, a, b, matrix=[
                c, 0, 0, -s, 0, c, -s, 0, 0, s, c, 0, s, 0, 0, c]))
        else:
            g = op.qft(mask() & mask()) if n <= 12 else op.h(bit())
        if rng.next() & 1:
            g = g.dgr()
        if rng.next() % 3 == 0 and len(g):
            free = ((1 << n) - 1) & ~g.act_on()
            # u2's act_on omits b_mask (reference quirk): keep controls off both targets
            for s_ in g:
                free &= ~(s_.a_mask | s_.b_mask)
            c = free & rng.next() & rng.next()
            if c:
                g = g.c(c)
        circ *= g
    return circ


def qasm_config5(n: int, layers: int, seed: int = 0x51415335) -> str:
    """configs[4]: OpenQASM 2.0 text, generated deterministically --
    `h q;` then layers of multi-controlled Toffolis (ccx / cccx chains reaching across the
    register, so the controls land on sharded qubits), rzz and i_swap pairs, a final
    `measure q -> c;`.  Gate names as lowered by qasm/int/gates.rs:111,115."""
    rng = SplitMix64(seed)
    out = ["OPENQASM 2.0;", f"qreg q[{n}];", f"creg c[{n}];", "h q;"]

    def pick(k):
        s = []
        while len(s) < k:
            v = rng.next() % n
            if v not in s:
                s.append(v)
        return s
    for layer in range(layers):
        for _ in range(max(1, n // 6)):
            a, b, c = pick(3)
            out.append(f"ccx q[{a}],q[{b}],q[{c}];")
        a, b, c, d = pick(4)
        out.append(f"cccx q[{a}],q[{b}],q[{c}],q[{d}];")
        for _ in range(max(1, n // 4)):
            a, b = pick(2)
            out.append(f"rzz({rng.angle():.17g}) q[{a}],q[{b}];")
        for _ in range(max(1, n // 6)):
            a, b = pick(2)
            out.append(f"i_swap q[{a}],q[{b}];")
        if layer % 2 == 1:
            out.append(f"rx(pi/{2 + layer % 5}) q[{pick(1)[0]}];")
    out.append("measure q -> c;")
    return "\n".join(out) + "\n"


def fast_mix(n: int, layers: int, seed: int) -> MultiOp:
    """Every kind the fast stage interpreter carries, with random controls (incl. multi-bit masks): x / cx
    leave register slots inverted, controlled diagonals and h follow on the same slots."""
    rng = SplitMix64(seed)
    circ = MultiOp()
    for _ in range(layers):
        k = rng.next() % 11
        b = 1 << (rng.next() % n)
        m = (rng.next() & ((1 << n) - 1)) or 1
        th = rng.angle()
        g = [op.x(m), op.y(m), op.z(m), op.s(m), op.t(m), op.h(m), op.rx(th, b), op.ry(th, b), op.rz(th, b),
             op.rzz(th, b | (1 << ((b.bit_length() + 2) % n)) if b != 1 << ((b.bit_length() + 2) % n) else b | (b << 1) % (1 << n) or 3),
             op.t(m).dgr()][k]
        if rng.next() % 3 == 0:
            free = ((1 << n) - 1) & ~g.act_on()
            for s_ in g:
                free &= ~(s_.a_mask | s_.b_mask)
            c = free & rng.next() & rng.next()
            if c:
                g = g.c(c)
        circ *= g
    return circ
