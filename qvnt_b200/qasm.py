"""OpenQASM 2.0 front-end for the device register (the caller of the hot path).

The reference parses with the un-vendored crates `qvnt-qasm 0.2.0` / `meval 0.2.0`; this module
restates the part that matters for the gate-application path -- how statements are LOWERED to op
lists and executed -- following (citations relative to /root/reference/src):

  gate names -> op::*          qasm/int/gates.rs:76-124   (`process`; leading c/C strips one control
                                                           register recursively: cx, ccx, cccx, ch, crz...;
                                                           quirk: sdg/tdg lower to plain s/t, :98,:100)
  registers -> masks           qasm/int/mod.rs:186-200,306-341 (qregs concatenated in declaration
                                                           order; q[i] = i-th set bit of the alias mask)
  measure / reset / if / gate  qasm/int/mod.rs:205-305, int/ext_op.rs:5-67 (ops are split by
                                                           separators Nop / Measure / IfBranch / Reset)
  execution                    qasm/sym.rs:40-87          (`Sym::finish`: apply each segment, then
                                                           measure_mask / reset_by_mask / conditional apply)

Host-side string processing only; the amplitudes never leave HBM.
"""
from __future__ import annotations

import ast as _pyast
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from . import op
from .op import MultiOp
from .register import CReg, QReg


class QasmError(Exception):
    def __init__(self, kind: str, *args):
        super().__init__(f"{kind}{args}")
        self.kind = kind
        self.args_ = args


def _count_bits(x: int) -> int:
    return bin(x).count("1")


# --------------------------------------------------------------------------- gates.rs:76-124
def process(name: str, regs: List[int], args: List[float]) -> MultiOp:
    """Lower one gate application: `regs` are qubit masks in argument order, `args` evaluated
    angles.  Raises QasmError with the reference's error kinds."""
    if name[:1] in ("c", "C"):
        if not regs:
            raise QasmError("WrongRegNumber", name, 0)
        ctrl, rest = regs[0], regs[1:]
        try:
            inner = process(name[1:], rest, args)
        except QasmError as e:
            if e.kind == "WrongRegNumber":
                raise QasmError("WrongRegNumber", name, 1 + e.args_[1])
            if e.kind == "WrongArgNumber":
                raise QasmError("WrongArgNumber", name, e.args_[1])
            if e.kind == "UnknownGate":
                raise QasmError("UnknownGate", name)
            raise
        act = inner.act_on()
        res = inner.c(ctrl)
        if res is None:
            raise QasmError("InvalidControlMask", ctrl, act)
        return res
    mask = 0
    for r in regs:
        mask |= r
    low = name.lower() if name in (name.lower(), name.upper()) else name
    any_gates = {"x": op.x, "y": op.y, "z": op.z, "s": op.s, "sdg": op.s, "t": op.t, "tdg": op.t,
                 "h": op.h, "qft": op.qft}
    if low in any_gates:
        if mask == 0:
            raise QasmError("WrongRegNumber", name, 0)
        if args:
            raise QasmError("WrongArgNumber", name, len(args))
        return any_gates[low](mask)
    rot = {"rx": (op.rx, 1), "ry": (op.ry, 1), "rz": (op.rz, 1), "rxx": (op.rxx, 2), "ryy": (op.ryy, 2),
           "rzz": (op.rzz, 2)}
    if low in rot:
        fn, bits = rot[low]
        if _count_bits(mask) != bits:
            raise QasmError("WrongRegNumber", name, _count_bits(mask))
        if len(args) != 1:
            raise QasmError("WrongArgNumber", name, len(args))
        return fn(args[0], mask)
    two = {"swap": op.swap, "sqrt_swap": op.sqrt_swap, "i_swap": op.i_swap, "sqrt_i_swap": op.sqrt_i_swap}
    if low in two:
        if _count_bits(mask) != 2:
            raise QasmError("WrongRegNumber", name, _count_bits(mask))
        if args:
            raise QasmError("WrongArgNumber", name, len(args))
        return two[low](mask)
    uni = {"u1": (op.u1, 1), "u2": (op.u2, 2), "u3": (op.u3, 3)}
    if low in uni:
        fn, n_args = uni[low]
        if _count_bits(mask) != 1:
            raise QasmError("WrongRegNumber", name, _count_bits(mask))
        if len(args) != n_args:
            raise QasmError("WrongArgNumber", name, len(args))
        return fn(*args, mask)
    raise QasmError("UnknownGate", name)


# --------------------------------------------------------------------------- expression evaluation
_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log,
          "sqrt": math.sqrt, "abs": abs, "asin": math.asin, "acos": math.acos, "atan": math.atan}
_CONSTS = {"pi": math.pi, "e": math.e}


def eval_extended(expr: str, env: Optional[Dict[str, float]] = None) -> float:
    """Arithmetic over numbers, pi, e, + - * / ^ and the usual functions (qasm/int/parse.rs:30-40)."""
    env = dict(_CONSTS, **(env or {}))
    try:
        tree = _pyast.parse(expr.replace("^", "**"), mode="eval")
    except SyntaxError as ex:
        raise QasmError("UnevaluatedArgument", expr, str(ex))

    def ev(n):
        if isinstance(n, _pyast.Expression):
            return ev(n.body)
        if isinstance(n, _pyast.Constant) and isinstance(n.value, (int, float)):
            return float(n.value)
        if isinstance(n, _pyast.Name) and n.id in env:
            return float(env[n.id])
        if isinstance(n, _pyast.UnaryOp) and isinstance(n.op, (_pyast.USub, _pyast.UAdd)):
            v = ev(n.operand)
            return -v if isinstance(n.op, _pyast.USub) else v
        if isinstance(n, _pyast.BinOp):
            a, b = ev(n.left), ev(n.right)
            if isinstance(n.op, _pyast.Add):
                return a + b
            if isinstance(n.op, _pyast.Sub):
                return a - b
            if isinstance(n.op, _pyast.Mult):
                return a * b
            if isinstance(n.op, _pyast.Div):
                return a / b
            if isinstance(n.op, _pyast.Pow):
                return a ** b
        if isinstance(n, _pyast.Call) and isinstance(n.func, _pyast.Name) and n.func.id in _FUNCS:
            return float(_FUNCS[n.func.id](*[ev(a) for a in n.args]))
        raise QasmError("UnevaluatedArgument", expr, "unsupported expression")
    return ev(tree)


# --------------------------------------------------------------------------- ExtOp (int/ext_op.rs)
@dataclass
class Sep:
    kind: str = "Nop"          # Nop | Measure | IfBranch | Reset
    a: int = 0
    b: int = 0


@dataclass
class ExtOp:
    segs: List[Tuple[MultiOp, Sep]] = field(default_factory=list)
    tail: MultiOp = field(default_factory=MultiOp)

    def push(self, other: MultiOp):                      # ext_op.rs:38-48
        if len(self.tail) == 0 and self.segs and self.segs[-1][1].kind == "Nop":
            self.segs[-1] = (self.segs[-1][0] * other, self.segs[-1][1])
        else:
            self.tail = self.tail * other

    def branch(self, sep: Sep):
        """Int::branch (int/mod.rs:374-379): close the pending ops with `sep` -- only if there are any."""
        if len(self.tail):
            self.branch_with_id(sep)

    def branch_with_id(self, sep: Sep):
        """Int::branch_with_id (int/mod.rs:381-384): always emits a segment (measure / reset)."""
        self.segs.append((self.tail, sep))
        self.tail = MultiOp()

    def n_ops(self) -> int:
        return sum(len(m) for m, _ in self.segs) + len(self.tail)


@dataclass
class _Macro:
    params: List[str]
    qargs: List[str]
    body: List[str]


class Int:
    """Interpreter state: registers, macros and the lowered program (qasm/int/mod.rs:33-41)."""

    _STMT = re.compile(r"^(?P<name>[A-Za-z_][A-Za-z0-9_]*)\s*(\((?P<args>.*)\))?\s*(?P<regs>.*)$", re.S)

    def __init__(self, source: Optional[str] = None, xor: bool = False):
        self.m_op = "Xor" if xor else "Set"
        self.q_reg: List[str] = []
        self.c_reg: List[str] = []
        self.q_ops = ExtOp()
        self.macros: Dict[str, _Macro] = {}
        if source is not None:
            self.add(source)

    # -- registers --------------------------------------------------------------------------
    @staticmethod
    def _mask(names: List[str], alias: str) -> int:
        m = 0
        for i, nm in enumerate(names):
            if nm == alias:
                m |= 1 << i
        return m

    def _idx(self, names: List[str], arg: str, kind: str) -> int:
        arg = arg.strip()
        m = re.fullmatch(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\s*(\d+)\s*\])?", arg)
        if not m:
            raise QasmError("BadArgument", arg)
        alias, idx = m.group(1), m.group(3)
        mask = self._mask(names, alias)
        if mask == 0:
            raise QasmError("NoQReg" if kind == "q" else "NoCReg", alias)
        if idx is None:
            return mask
        bits = [1 << i for i in range(mask.bit_length()) if (mask >> i) & 1]
        if int(idx) >= len(bits):
            raise QasmError("IdxOutOfRange", alias, int(idx))
        return bits[int(idx)]

    def q_idx(self, arg: str) -> int:
        return self._idx(self.q_reg, arg, "q")

    def c_idx(self, arg: str) -> int:
        return self._idx(self.c_reg, arg, "c")

    def _declare(self, names: List[str], alias: str, size: int):
        if len(alias.encode()) >= 32:
            raise QasmError("IdentIsTooLarge", alias, len(alias.encode()))
        if size >= 64:
            raise QasmError("RegisterIsTooLarge", alias, size)
        if alias in self.q_reg:
            raise QasmError("DupQReg", alias, self.q_reg.count(alias))
        if alias in self.c_reg:
            raise QasmError("DupCReg", alias, self.c_reg.count(alias))
        names.extend([alias] * size)

    # -- statements ---------------------------------------------------------------------------
    @staticmethod
    def _split(source: str) -> List[str]:
        src = re.sub(r"//[^\n]*", "", source)
        out, depth, cur = [], 0, ""
        for ch in src:
            if ch == "{":
                depth += 1
            if ch == "}":
                depth -= 1
                cur += ch
                if depth == 0:
                    out.append(cur.strip())
                    cur = ""
                continue
            if ch == ";" and depth == 0:
                if cur.strip():
                    out.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            out.append(cur.strip())
        return out

    def add(self, source: str) -> "Int":
        for st in self._split(source):
            self._statement(st)
        return self

    def _apply_gate(self, name: str, regs: List[int], args: List[float]) -> MultiOp:
        mac = self.macros.get(name)
        if mac is None:
            return process(name, regs, args)
        if len(regs) != len(mac.qargs):
            raise QasmError("WrongRegNumber", name, len(regs))
        if len(args) != len(mac.params):
            raise QasmError("WrongArgNumber", name, len(args))
        env = dict(zip(mac.params, args))
        qenv = dict(zip(mac.qargs, regs))
        res = MultiOp()
        for st in mac.body:
            m = self._STMT.match(st)
            nm = m.group("name")
            if nm == "barrier":
                continue
            a = [eval_extended(x, env) for x in self._csv(m.group("args"))]
            r = []
            for q in self._csv(m.group("regs")):
                if q not in qenv:
                    raise QasmError("UnknownReg", q)
                r.append(qenv[q])
            res = res * self._apply_gate(nm, r, a)
        return res

    @staticmethod
    def _csv(s: Optional[str]) -> List[str]:
        if s is None or not s.strip():
            return []
        out, depth, cur = [], 0, ""
        for ch in s:
            if ch == "(":
                depth += 1
            if ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                out.append(cur.strip())
                cur = ""
            else:
                cur += ch
        out.append(cur.strip())
        return out

    def _statement(self, st: str):
        if st.startswith("OPENQASM") or st.startswith("include"):
            return
        m = re.fullmatch(r"(qreg|creg)\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]", st)
        if m:
            self._declare(self.q_reg if m.group(1) == "qreg" else self.c_reg, m.group(2), int(m.group(3)))
            return
        if st.startswith("barrier") or st.startswith("opaque"):
            return
        m = re.fullmatch(r"gate\s+([A-Za-z_][A-Za-z0-9_]*)\s*(\(([^)]*)\))?\s*([^{]*)\{(.*)\}", st, re.S)
        if m:
            name = m.group(1)
            if name in self.macros:
                raise QasmError("MacroAlreadyDefined", name)
            self.macros[name] = _Macro(self._csv(m.group(3)), self._csv(m.group(4)),
                                       [b.strip() for b in m.group(5).split(";") if b.strip()])
            return
        m = re.fullmatch(r"reset\s+(.+)", st)
        if m:
            self.q_ops.branch_with_id(Sep("Reset", self.q_idx(m.group(1))))
            return
        m = re.fullmatch(r"measure\s+(.+?)\s*->\s*(.+)", st)
        if m:
            q, c = self.q_idx(m.group(1)), self.c_idx(m.group(2))
            if _count_bits(q) != _count_bits(c):
                raise QasmError("UnmatchedRegSize", _count_bits(q), _count_bits(c))
            self.q_ops.branch_with_id(Sep("Measure", q, c))
            return
        m = re.fullmatch(r"if\s*\(\s*([A-Za-z_][A-Za-z0-9_]*)\s*==\s*(\d+)\s*\)\s*(.+)", st, re.S)
        if m:
            # int/mod.rs:292-303.  Reference quirk kept for parity: when gates are pending before the
            # `if`, they are closed as a Nop segment and ExtOp::push (ext_op.rs:38-48) then MERGES the
            # conditional gate into that segment -- it runs unconditionally.  Directly after a
            # measure / reset (the usual pattern) the gate gets its own IfBranch segment.
            self.q_ops.branch(Sep("Nop"))
            val = self.c_idx(m.group(1))
            self._gate_statement(m.group(3))
            self.q_ops.branch(Sep("IfBranch", val, int(m.group(2))))
            return
        self._gate_statement(st)

    def _gate_statement(self, st: str):
        m = self._STMT.match(st)
        if not m:
            raise QasmError("Syntax", st)
        regs = [self.q_idx(r) for r in self._csv(m.group("regs"))]
        args = [eval_extended(a) for a in self._csv(m.group("args"))]
        self.q_ops.push(self._apply_gate(m.group("name"), regs, args))


class Sym:
    """Executor over a device register (qasm/sym.rs).  `us`: uniform variates to inject into the
    successive measurements (parity tests); None draws from the library's RNG."""

    def __init__(self, prog: Int, reg_factory=None):
        self.int = prog
        n = len(prog.q_reg)
        self.q_reg = (reg_factory or QReg.new)(n)
        self.c_reg = CReg.new(len(prog.c_reg))

    def reset(self):
        self.q_reg.reset(0)
        self.c_reg = CReg.new(len(self.int.c_reg))

    def _measure(self, q_arg: int, c_arg: int, u: Optional[float]):
        got = self.q_reg.measure_mask(q_arg, u).get()
        qs = [1 << i for i in range(q_arg.bit_length()) if (q_arg >> i) & 1]
        cs = [1 << i for i in range(c_arg.bit_length()) if (c_arg >> i) & 1]
        val = self.c_reg.get()
        for q, c in zip(qs, cs):
            bit = (got & q) != 0
            if self.int.m_op == "Set":
                val = (val | c) if bit else (val & ~c)
            elif bit:
                val ^= c
        self.c_reg = CReg.with_state(len(self.int.c_reg), val)

    def finish(self, us: Optional[List[float]] = None) -> "Sym":
        us = list(us) if us is not None else None
        for mop, sep in self.int.q_ops.segs:
            if sep.kind == "Nop":
                self.q_reg.apply(mop)
            elif sep.kind == "Measure":
                self.q_reg.apply(mop)
                self._measure(sep.a, sep.b, us.pop(0) if us else None)
            elif sep.kind == "IfBranch":
                if self.c_reg.get_by_mask(sep.a) == sep.b:
                    self.q_reg.apply(mop)
            elif sep.kind == "Reset":
                self.q_reg.apply(mop)
                self.q_reg.reset_by_mask(sep.a)
        self.q_reg.apply(self.int.q_ops.tail)
        return self

    def get_class(self) -> CReg:
        return self.c_reg

    def get_probabilities(self):
        return self.q_reg.get_probabilities()
