"""The QASM front-end subset (qvnt_b200/qasm.py): gate-name lowering pinned to the reference's
own tests (qasm/int/gates.rs:131-248), statement lowering / separators (int/mod.rs, ext_op.rs),
and `Sym::finish` execution (sym.rs:40-74) -- on the CPU against the oracle register, on the GPU
against the same executor over the oracle (configs[4] twin: ccx/cccx/rzz/i_swap + measure)."""
import math

import numpy as np
import pytest

from qvnt_b200 import op, qasm, workloads
from qvnt_b200.qasm import Int, QasmError, Sym, process


def err(fn):
    with pytest.raises(QasmError) as e:
        fn()
    return (e.value.kind,) + tuple(e.value.args_)


# ---- qasm/int/gates.rs:131-248, test by test --------------------------------------------------
def test_try_process_x():
    assert process("x", [0b111], []) == op.x(0b111)
    assert err(lambda: process("x", [0b111], [1.0])) == ("WrongArgNumber", "x", 1)


def test_try_process_cx():
    assert process("cx", [0b100, 0b010, 0b001], []) == op.x(0b011).c(0b100)
    assert err(lambda: process("cx", [0b100], [])) == ("WrongRegNumber", "cx", 1)
    assert err(lambda: process("cx", [0b100, 0b010, 0b001], [1.0])) == ("WrongArgNumber", "cx", 1)


def test_try_process_ccx():
    assert process("ccx", [0b100, 0b010, 0b001], []) == op.x(0b001).c(0b110)
    assert err(lambda: process("ccx", [0b100], [])) == ("WrongRegNumber", "ccx", 1)
    assert err(lambda: process("ccx", [0b100, 0b010, 0b001], [1.0])) == ("WrongArgNumber", "ccx", 1)


def test_try_process_rx_rxx_swap():
    assert process("rx", [0b100], [1.0]) == op.rx(1.0, 0b100)
    assert err(lambda: process("rx", [0b101], [1.0])) == ("WrongRegNumber", "rx", 2)
    assert err(lambda: process("rx", [0b100], [])) == ("WrongArgNumber", "rx", 0)
    assert process("rxx", [0b101], [1.0]) == op.rxx(1.0, 0b101)
    assert err(lambda: process("rxx", [0b100], [1.0])) == ("WrongRegNumber", "rxx", 1)
    assert err(lambda: process("rxx", [0b101], [2.0, 1.0])) == ("WrongArgNumber", "rxx", 2)
    assert process("swap", [0b101], []) == op.swap(0b101)
    assert err(lambda: process("swap", [0b111], [1.0])) == ("WrongRegNumber", "swap", 3)
    assert err(lambda: process("swap", [0b101], [1.0])) == ("WrongArgNumber", "swap", 1)


def test_try_process_unitary_and_any():
    assert process("u1", [0b001], [1.0]) == op.u1(1.0, 0b001)
    assert process("u2", [0b001], [1.0, 2.0]) == op.u2(1.0, 2.0, 0b001)
    assert process("u3", [0b001], [1.0, 2.0, 3.0]) == op.u3(1.0, 2.0, 3.0, 0b001)
    assert process("x", [0b001, 0b100], []) == op.x(0b101)
    assert process("y", [0b11], []) == op.y(0b11)
    assert process("ch", [0b100, 0b010, 0b001], []) == op.h(0b011).c(0b100)
    assert process("swap", [0b100, 0b010], []) == op.swap(0b110)
    assert err(lambda: process("swap", [0b001], [])) == ("WrongRegNumber", "swap", 1)
    assert err(lambda: process("foo", [1], [])) == ("UnknownGate", "foo")


def test_sdg_tdg_quirk_and_control_overlap():
    # gates.rs:98,100: the `dgr` macro arm never calls .dgr()
    assert process("sdg", [0b1], []) == op.s(0b1)
    assert process("tdg", [0b10], []) == op.t(0b10)
    assert err(lambda: process("cx", [0b1, 0b1], [])) == ("InvalidControlMask", 1, 1)


# ---- statements -> ExtOp ------------------------------------------------------------------------
SRC = """
OPENQASM 2.0;
include "qelib1.inc";
qreg a[2]; qreg b[3];
creg c[2]; creg d[3];
gate bell x, y { h x; cx x, y; }
h a;
bell b[0], b[2];
rz(pi/4 + 0.5*2) b[1];   // comment
measure a -> c;
if (c == 3) x b;
reset b[0];
cu1(pi/8) a[0], b[1];
"""


def test_statement_lowering():
    prog = Int(SRC)
    assert prog.q_reg == ["a", "a", "b", "b", "b"] and prog.c_reg == ["c", "c", "d", "d", "d"]
    assert prog.q_idx("b[2]") == 0b10000 and prog.q_idx("a") == 0b00011 and prog.c_idx("d") == 0b11100
    segs, tail = prog.q_ops.segs, prog.q_ops.tail
    assert [s.kind for _, s in segs] == ["Measure", "IfBranch", "Reset"]
    first = op.h(0b00011) * op.h(0b00100) * op.x(0b10000).c(0b00100) * op.rz(math.pi / 4 + 1.0, 0b01000)
    assert segs[0][0] == first and (segs[0][1].a, segs[0][1].b) == (0b00011, 0b00011)
    assert segs[1][0] == op.x(0b11100) and (segs[1][1].a, segs[1][1].b) == (0b00011, 3)
    assert len(segs[2][0]) == 0 and segs[2][1].a == 0b00100
    assert tail == op.u1(math.pi / 8, 0b01000).c(0b00001)
    with pytest.raises(QasmError):
        Int("qreg q[2]; creg c[3]; measure q -> c;")
    with pytest.raises(QasmError):
        Int("qreg q[2]; h r;")


def test_if_after_pending_gates_quirk():
    """ext_op.rs:38-48 + int/mod.rs:292-303: a conditional gate that follows pending gates is merged
    into their Nop segment (reference behaviour, reproduced for parity)."""
    prog = Int("qreg q[2]; creg c[1]; h q[0]; if (c == 1) x q[1];")
    assert [s.kind for _, s in prog.q_ops.segs] == ["Nop"]
    assert prog.q_ops.segs[0][0] == op.h(0b01) * op.x(0b10) and len(prog.q_ops.tail) == 0


def test_eval_extended():
    assert qasm.eval_extended("pi/2") == math.pi / 2
    assert qasm.eval_extended("2^3 - sin(0)") == 8.0
    assert qasm.eval_extended("-x*2", {"x": 1.5}) == -3.0
    with pytest.raises(QasmError):
        qasm.eval_extended("__import__('os')")


def test_sym_finish_on_oracle(oracle):
    prog = Int(SRC)
    sym = Sym(prog, reg_factory=lambda n: oracle.OracleReg.new(n))
    sym.finish(us=[0.3])
    c = sym.get_class().get()
    assert c & ~0b00011 == 0
    p = sym.get_probabilities()
    assert abs(p.sum() - 1.0) < 1e-12
    assert np.all(p[[i for i in range(32) if i & 0b00100]] == 0)     # b[0] was reset


def test_config5_generator_lowers():
    n = 12
    prog = Int(workloads.qasm_config5(n, 3))
    assert len(prog.q_reg) == n and prog.q_ops.segs[-1][1].kind == "Measure"
    kinds = {s.kind for m, _ in prog.q_ops.segs for s in m}
    assert {op.K_X, op.K_RZZ, op.K_ISWAP, op.K_H2}.issubset(kinds)
    assert any(bin(s.ctrl).count("1") == 3 for m, _ in prog.q_ops.segs for s in m)


# ---- GPU: configs[4] twin -------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,layers", [(12, 4), (20, 3)])
def test_config5_twin_device_vs_oracle(oracle, n, layers):
    src = workloads.qasm_config5(n, layers) + "if (c == 0) x q[0];\nreset q[1];\nh q[2];\n"
    us = [0.6180339887]
    dev = Sym(Int(src)).finish(us=us)
    ref = Sym(Int(src), reg_factory=lambda k: oracle.OracleReg.new(k, threads=oracle.max_threads())).finish(us=us)
    assert dev.get_class().get() == ref.get_class().get()
    assert np.abs(dev.q_reg.amplitudes() - ref.q_reg.amplitudes()).max() <= 1e-10
    assert np.abs(dev.get_probabilities() - ref.get_probabilities()).max() <= 1e-12


# ---- the reference's own smoke programs (qasm/mod.rs:15-39) ----------------------------------------
_EXAMPLES = "/root/reference/src/qasm/examples/source"
_NAMES = ["adder", "bigadder", "Deutsch_Algorithm", "inverseqft1", "inverseqft2", "ipea_3_pi_8", "qe_qft_3",
          "qe_qft_4", "qe_qft_5", "qec", "qft", "qpt", "rb", "teleport", "teleportv2", "W-state", "W3test"]


@pytest.mark.skipif(not __import__("os").path.isdir(_EXAMPLES),
                    reason="reference checkout not present (it never is on the GPU box)")
@pytest.mark.parametrize("name", _NAMES)
def test_reference_example_programs_run(oracle, name):
    """`run_qasm` of the reference: build the program, Sym::reset(); Sym::finish() -- read from the
    reference checkout where it lies (never copied); here additionally: probabilities sum to 1."""
    src = open(f"{_EXAMPLES}/{name}.qasm").read()
    prog = Int(src)
    sym = Sym(prog, reg_factory=lambda k: oracle.OracleReg.new(k))
    sym.reset()
    sym.finish(us=[0.37] * 64)
    assert abs(float(sym.get_probabilities().sum()) - 1.0) < 1e-9
    assert sym.get_class().get() >> max(1, len(prog.c_reg)) == 0
