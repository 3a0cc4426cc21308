"""Host-side mirror of the reference's operator algebra and registers (no GPU, no oracle).
Mirrors the reference's own host-logic tests: operator/single/mod.rs:142-161,
operator/multi/mod.rs:201-218, register/virtl.rs tests, math/bits_iter.rs:34."""
import math

import pytest

from qvnt_b200 import CReg, VReg, op
from qvnt_b200.op import single
from tests.golden import reference_kat as kat


def test_single_names():                       # single/mod.rs:142-151
    s = single.x(123)
    assert s.name() == "X123" and repr(s) == "X123"
    s = s.c(4)
    assert s.name() == "C4_X123" and repr(s) == "C4_X123"


def test_wrong_ctrl_mask():                    # single/mod.rs:153-161
    o = single.ryy(0b101, 1.35)
    assert o.act_on() & 0b001 != 0
    assert o.clone().c(0b001) is None
    assert o.act_on() & 0b010 == 0
    assert o.c(0b010) is not None


def test_multi_ops_len():                      # multi/mod.rs:201-208
    pend = op.id() * op.h(0b001).c(0b010) * op.x(0b011).c(0b100) * op.rz(5.0, 0b001)
    assert len(pend) == 3


def test_ends_with():                          # multi/mod.rs:211-218
    a = op.x(0x010101) * op.y(0x101010) * op.z(0x011011)
    b = op.y(0x101010) * op.z(0x011011)
    assert b.ends_with(a)


def test_validity_checks():                    # atomic/u1.rs:67-80, u2.rs:98-112, single/*.rs Option
    for a, m in kat.U1_INVALID:
        assert single.u1(a, m) is None
    for a, b, m in kat.U2_INVALID:
        assert single.u2(a, b, m) is None
    assert single.rx(0b11, 1.0) is None and single.rzz(0b1, 1.0) is None and single.swap(0b111) is None
    with pytest.raises(ValueError):
        op.rx(1.0, 0b11)


def test_h_lowering():                         # multi/h.rs:14-45
    assert repr(op.h(0b1111)) == "[H3, H12]"
    assert repr(op.h(0b10101)) == "[H5, H16]"
    assert repr(op.h(0)) == "[]"
    assert repr(op.h(0b100)) == "[H4]"


def test_qft_lowering():                       # multi/qft.rs:4-33
    q = op.qft(0b1011)
    names = [s.name() for s in q]
    assert names[0] == "H1" and names[1].startswith("C1_RZ2(") and names[2].startswith("C1_RZ8(")
    assert names[3] == "H2" and names[4].startswith("C2_RZ8(") and names[5] == "H8"
    assert len(op.qft(0xFFFFF)) == 210 and len(op.qft(0xFFFFFFFF)) == 528
    # angles pi * 0.5^j
    assert q[1].phase == (math.cos(math.pi * 0.5 / 2), math.sin(math.pi * 0.5 / 2))
    assert q[2].phase == (math.cos(math.pi * 0.25 / 2), math.sin(math.pi * 0.25 / 2))
    sw = op.qft_swapped(0b1011)
    assert [s.name() for s in sw][-1] == "SWAP9"


def test_u_aliases():                          # operator/mod.rs:472-501
    assert op.u1(0.3, 0b1) == op.rz(0.3, 0b1)
    assert op.u2(0.1, 0.2, 0b10) == op.rz(0.2, 0b10) * op.ry(math.pi / 2, 0b10) * op.rz(0.1, 0b10)
    assert op.u3(0.5, 0.1, 0.2, 0b10) == op.rz(0.2, 0b10) * op.ry(0.5, 0b10) * op.rz(0.1, 0b10)


def test_dgr_semantics():
    # rotations: phase negated (re AND im) -- atomic/rx.rs:42-47; multi: reversed order
    r = single.rx(0b1, 1.0)
    d = r.dgr()
    assert d.phase == (-r.phase[0], -r.phase[1])
    m = (op.x(1) * op.s(2) * op.rz(0.4, 4)).dgr()
    assert [s.kind for s in m] == [op.K_RZ, op.K_S, op.K_X] and m[1].dagger
    # c() on a MultiOp: None on overlap
    assert (op.x(1) * op.y(2)).c(2) is None
    assert repr((op.x(1) * op.y(2)).c(4)) == "[C4_X1, C4_Y2]"


def test_creg():
    c = CReg.with_state(4, 0b1010)
    assert c.get() == 0b1010 and repr(c) == "(1010)"
    assert c.get_by_mask(0b1010) == 0b11 and c.get_by_mask(0b0110) == 0b01
    assert (CReg.with_state(2, 0b01) * CReg.with_state(2, 0b10)).get() == 0b1001


def test_vreg():
    v = VReg(mask=0b101101)
    assert v[0] == 1 and v[1] == 0b100 and v[3] == 0b100000
    assert v[[0, 2]] == 0b1001 and v[:] == 0b101101
    assert v[lambda i: i % 2 == 1] == 0b100100
