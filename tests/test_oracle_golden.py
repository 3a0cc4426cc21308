"""Pins the CPU oracle (oracle/qvnt_oracle.c) against every known-answer vector the reference's
own tests hold for the gate-application path (tests/golden/reference_kat.py cites each one).
Exact f64 equality, like the reference's `assert_eq!`.  No GPU needed."""
import numpy as np
import pytest

from qvnt_b200 import op
from tests.golden import reference_kat as kat


def _factory(oracle):
    return lambda q, s: oracle.OracleReg(q, s)


@pytest.mark.parametrize("label,build,name,size,expected", kat.ATOMIC_KATS, ids=[k[0] for k in kat.ATOMIC_KATS])
def test_atomic_matrix_repr(oracle, label, build, name, size, expected):
    sop = build(op)
    assert sop is not None
    assert sop.name() == kat.expected_name(name)
    got = op._matrix(sop, size, _factory(oracle))
    exp = np.array(expected, dtype=np.complex128)
    got = np.array(got, dtype=np.complex128)
    assert got.shape == exp.shape
    # assert_eq! on Complex<f64>: exact (and -0.0 == 0.0)
    assert np.array_equal(got.real, exp.real) and np.array_equal(got.imag, exp.imag), (label, got, exp)


def test_quantum_reg_golden(oracle):
    g = kat.QUANTUM_REG
    operator = kat.quantum_reg_op(op)
    assert repr(operator) == g["op_debug"]
    reg = oracle.OracleReg.with_state(g["q_num"], g["state"])
    reg.apply(operator)
    psi = reg.amplitudes()
    assert np.array_equal(psi.real, np.array(g["psi"])) and not psi.imag.any()
    for u in (0.0, 0.3, 0.77, 0.999999):
        r2 = reg.clone()
        out, _ = r2.measure_mask_full(g["mask"], u)
        assert out & ~g["mask"] == 0


def test_bell_doctest_probabilities(oracle):
    # register/quant.rs:86-96: assert_eq!(prob, [0.5, 0.0, 0.0, 0.5])
    q = oracle.OracleReg.new(2)
    q.apply(op.h(0b01) * op.x(0b10).c(0b01))
    assert list(q.get_probabilities()) == [0.5, 0.0, 0.0, 0.5]


def test_tensor_golden(oracle):
    # register/quant.rs:680-711
    pend = op.h(0b01)
    r1, r2 = oracle.OracleReg.with_state(2, 0b01), oracle.OracleReg.with_state(1, 0b1)
    r1.apply(pend)
    r2.apply(pend)
    p = (r1 * r2).get_probabilities()
    assert np.all(np.abs(p - np.array(kat.TENSOR_PROB)) < kat.TENSOR_EPS)
    r3 = oracle.OracleReg.with_state(3, 0b101)
    r3.apply(op.h(0b101))
    assert np.all(np.abs(r3.get_probabilities() - np.array(kat.TENSOR_PROB)) < kat.TENSOR_EPS)


def test_qft_uniform(oracle):
    # analytic check: qft on a basis state gives |a|^2 = 2^-n
    n = 10
    r = oracle.OracleReg.with_state(n, 0x2A5)
    r.apply(op.qft((1 << n) - 1))
    p = r.get_probabilities()
    assert np.all(np.abs(p - 2.0 ** -n) < 1e-12)


def test_weighted_index_semantics(oracle):
    # rand 0.8.5 WeightedIndex: first i with cumsum_i > u*total; never past the last index
    w = np.array([0.25, 0.0, 0.25, 0.5])
    assert oracle.weighted_index(w, 0.0) == 0
    assert oracle.weighted_index(w, 0.2499) == 0
    assert oracle.weighted_index(w, 0.25) == 2      # cum[0] = .25 <= x -> skips the zero weight too
    assert oracle.weighted_index(w, 0.5) == 3
    assert oracle.weighted_index(w, 0.999999) == 3


def test_normalize_and_reset_by_mask(oracle):
    r = oracle.OracleReg.new(3)
    r.apply(op.h(0b111))
    r.reset_by_mask(0b001)                 # zero odd indices, renormalise (quant.rs:207-229)
    a = r.amplitudes()
    assert np.all(a[1::2] == 0)
    assert abs(r.get_absolute() - 1.0) < 1e-12
    r.reset_by_mask(0b111)                 # full mask -> reset(0)
    a = r.amplitudes()
    assert a[0] == 1 and not a[1:].any()


def test_op_struct_layout(oracle):
    import ctypes
    from qvnt_b200.optypes import QvntOp
    assert oracle.lib().qo_sizeof_op() == ctypes.sizeof(QvntOp) == 304
