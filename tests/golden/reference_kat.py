"""Known-answer vectors held by the REFERENCE's own tests for the gate-application path.

The reference is Rust and cannot be executed in this image (no rustc/cargo, crates un-vendored),
so these are literal transcriptions of the expected VALUES asserted by its unit tests -- not
outputs of our own code.  Every entry cites the reference test it comes from (paths relative to
/root/reference/src).  The reference asserts them with `assert_eq!` on f64, i.e. EXACT equality;
so do our tests (tests/test_oracle_golden.py on the CPU oracle, tests/test_gpu_parity.py on the
CUDA path).

`ATOMIC_KATS`: (label, build(op_module) -> SingleOp, expected name(), matrix size in qubits,
expected dense matrix as rows of complex) -- from every `operator/atomic/*.rs::matrix_repr`.
"""
import math

ANGLE = 1.23456                       # `const ANGLE: R = 1.23456;` in every rotation test
FRAC_1_SQRT_2 = 0.70710678118654752440084436210485
O = 0j
I = 1 + 0j
i = 1j
COS = complex(math.cos(0.5 * ANGLE), 0.0)
SIN = complex(math.sin(0.5 * ANGLE), 0.0)
I_SIN = complex(0.0, math.sin(0.5 * ANGLE))
EXP = complex(math.cos(0.5 * ANGLE), math.sin(0.5 * ANGLE))
SQ = complex(FRAC_1_SQRT_2, 0.0)
O_5 = 0.5 + 0j
SQRT_I = 0.5 + 0.5j
EXP_I_PI_4 = complex(FRAC_1_SQRT_2, FRAC_1_SQRT_2)


def _rust_cdbg(z):
    from qvnt_b200.op import _rust_complex_debug
    return _rust_complex_debug(z.real, z.imag)


ATOMIC_KATS = [
    # atomic/id.rs:30-40
    ("id", lambda s: s.SingleOp(0), "Id", 1, [[I, O], [O, I]]),
    # atomic/x.rs:38-54
    ("x1", lambda s: s.single.x(0b1), "X1", 1, [[O, I], [I, O]]),
    ("x1_in2", lambda s: s.single.x(0b01), "X1", 2,
     [[O, I, O, O], [I, O, O, O], [O, O, O, I], [O, O, I, O]]),
    # atomic/y.rs:44-61
    ("y1", lambda s: s.single.y(0b1), "Y1", 1, [[O, -i], [i, O]]),
    ("y3", lambda s: s.single.y(0b11), "Y3", 2,
     [[O, O, O, -I], [O, O, I, O], [O, I, O, O], [-I, O, O, O]]),
    # atomic/z.rs:42-52
    ("z1", lambda s: s.single.z(0b1), "Z1", 1, [[I, O], [O, -I]]),
    # atomic/s.rs:49-62  (tests S-dagger)
    ("s1_dgr", lambda s: s.single.s(0b1).dgr(), "S1", 1, [[I, O], [O, -i]]),
    # atomic/t.rs:59-70  (tests T-dagger)
    ("t1_dgr", lambda s: s.single.t(0b1).dgr(), "T1", 1, [[I, O], [O, EXP_I_PI_4.conjugate()]]),
    # atomic/rx.rs:52-70
    ("rx1", lambda s: s.single.rx(0b1, ANGLE), "RX1(1.23456)", 1, [[COS, -I_SIN], [-I_SIN, COS]]),
    # atomic/ry.rs:57-75
    ("ry1", lambda s: s.single.ry(0b1, ANGLE), "RY1(1.23456)", 1, [[COS, -SIN], [SIN, COS]]),
    # atomic/rz.rs:53-68
    ("rz1", lambda s: s.single.rz(0b1, ANGLE), "RZ1(1.23456)", 1, [[EXP.conjugate(), O], [O, EXP]]),
    # atomic/rxx.rs:53-80
    ("rxx3", lambda s: s.single.rxx(0b11, ANGLE), "RXX3(1.23456)", 2,
     [[COS, O, O, -I_SIN], [O, COS, -I_SIN, O], [O, -I_SIN, COS, O], [-I_SIN, O, O, COS]]),
    # atomic/ryy.rs:57-116 (three embeddings)
    ("ryy3", lambda s: s.single.ryy(0b11, ANGLE), "RYY3(1.23456)", 2,
     [[COS, O, O, I_SIN], [O, COS, -I_SIN, O], [O, -I_SIN, COS, O], [I_SIN, O, O, COS]]),
    ("ryy6", lambda s: s.single.ryy(0b110, ANGLE), "RYY6(1.23456)", 3,
     [[COS, O, O, O, O, O, I_SIN, O],
      [O, COS, O, O, O, O, O, I_SIN],
      [O, O, COS, O, -I_SIN, O, O, O],
      [O, O, O, COS, O, -I_SIN, O, O],
      [O, O, -I_SIN, O, COS, O, O, O],
      [O, O, O, -I_SIN, O, COS, O, O],
      [I_SIN, O, O, O, O, O, COS, O],
      [O, I_SIN, O, O, O, O, O, COS]]),
    ("ryy5", lambda s: s.single.ryy(0b101, ANGLE), "RYY5(1.23456)", 3,
     [[COS, O, O, O, O, I_SIN, O, O],
      [O, COS, O, O, -I_SIN, O, O, O],
      [O, O, COS, O, O, O, O, I_SIN],
      [O, O, O, COS, O, O, -I_SIN, O],
      [O, -I_SIN, O, O, COS, O, O, O],
      [I_SIN, O, O, O, O, COS, O, O],
      [O, O, O, -I_SIN, O, O, COS, O],
      [O, O, I_SIN, O, O, O, O, COS]]),
    # atomic/rzz.rs:53-77
    ("rzz3", lambda s: s.single.rzz(0b11, ANGLE), "RZZ3(1.23456)", 2,
     [[EXP.conjugate(), O, O, O], [O, EXP, O, O], [O, O, EXP, O], [O, O, O, EXP.conjugate()]]),
    # atomic/u1.rs:57-93
    ("u1_id", lambda s: s.single.u1(0b1, [I, O, O, I]),
     "U1[[{I}, {O}], [{O}, {I}]]", 1, [[I, O], [O, I]]),
    ("u1_h", lambda s: s.single.u1(0b1, [SQ, SQ, SQ, -SQ]),
     "U1[[{SQ}, {SQ}], [{SQ}, {NSQ}]]", 1, [[SQ, SQ], [SQ, -SQ]]),
    # atomic/u2.rs:88-131
    ("u2_id", lambda s: s.single.u2(0b01, 0b10, [I, O, O, O, O, I, O, O, O, O, I, O, O, O, O, I]),
     "U3[[{I}, {O}, {O}, {O}], [{O}, {I}, {O}, {O}], [{O}, {O}, {I}, {O}], [{O}, {O}, {O}, {I}]]", 2,
     [[I, O, O, O], [O, I, O, O], [O, O, I, O], [O, O, O, I]]),
    ("u2_hh", lambda s: s.single.u2(0b01, 0b10, [SQ, SQ, O, O, SQ, -SQ, O, O, O, O, -SQ, -SQ, O, O, -SQ, SQ]),
     "U3[[{SQ}, {SQ}, {O}, {O}], [{SQ}, {NSQ}, {O}, {O}], [{O}, {O}, {NSQ}, {NSQ}], [{O}, {O}, {NSQ}, {SQ}]]", 2,
     [[SQ, SQ, O, O], [SQ, -SQ, O, O], [O, O, -SQ, -SQ], [O, O, -SQ, SQ]]),
    # atomic/h1.rs:47-60
    ("h1", lambda s: s.single.h1(0b1), "H1", 1, [[SQ, SQ], [SQ, -SQ]]),
    # atomic/h2.rs:65-82
    ("h3", lambda s: s.single.h2(0b01, 0b10), "H3", 2,
     [[O_5, O_5, O_5, O_5], [O_5, -O_5, O_5, -O_5], [O_5, O_5, -O_5, -O_5], [O_5, -O_5, -O_5, O_5]]),
    # atomic/swap.rs:47-61
    ("swap3", lambda s: s.single.swap(0b11), "SWAP3", 2,
     [[I, O, O, O], [O, O, I, O], [O, I, O, O], [O, O, O, I]]),
    # atomic/i_swap.rs:65-81 (tests iSWAP-dagger)
    ("iswap3_dgr", lambda s: s.single.i_swap(0b11).dgr(), "iSWAP3", 2,
     [[I, O, O, O], [O, O, -i, O], [O, -i, O, O], [O, O, O, I]]),
    # atomic/sqrt_swap.rs:65-85
    ("sqrt_swap3", lambda s: s.single.sqrt_swap(0b11), "sqrt(SWAP3)", 2,
     [[I, O, O, O], [O, SQRT_I, SQRT_I.conjugate(), O], [O, SQRT_I.conjugate(), SQRT_I, O], [O, O, O, I]]),
    # atomic/sqrt_i_swap.rs:65-85
    ("sqrt_iswap3", lambda s: s.single.sqrt_i_swap(0b11), "sqrt(iSWAP3)", 2,
     [[I, O, O, O], [O, FRAC_1_SQRT_2 * I, FRAC_1_SQRT_2 * i, O],
      [O, FRAC_1_SQRT_2 * i, FRAC_1_SQRT_2 * I, O], [O, O, O, I]]),
]


def expected_name(template: str) -> str:
    """u1/u2 names embed Rust `{:?}` of Complex<f64> (atomic/u1.rs:69-86, u2.rs:101-121)."""
    return template.format(I=_rust_cdbg(I), O=_rust_cdbg(O), SQ=_rust_cdbg(SQ), NSQ=_rust_cdbg(-SQ))


# is_valid() assertions: atomic/u1.rs:67-80, atomic/u2.rs:98-112
U1_INVALID = [(0b1, [I, I, O, I]), (0b11, [I, O, O, I])]
U2_INVALID = [(0b01, 0b10, [I, O, O, O, I, I, O, O, O, O, I, O, O, O, O, I]),
              (0b11, 0b10, [I, O, O, O, O, I, O, O, O, O, I, O, O, O, O, I])]

# register/quant.rs:643-677 `quantum_reg`
QUANTUM_REG = {
    "q_num": 4, "state": 0b1100, "mask": 0b0110,
    "op_debug": "[H3, H12, C8_H3, C2_SWAP9]",
    "psi": [0.25, 0.25, 0.25, 0.0, -0.25, -0.25, -0.25, 0.0, -0.5, 0.0, 0.25, 0.0, 0.5, 0.0, -0.25, 0.0],
    "reg_debug": ("QReg { 0: Complex { re: 0.25, im: 0.0 }, 1: Complex { re: 0.25, im: 0.0 }, "
                  "2: Complex { re: 0.25, im: 0.0 }, 3: Complex { re: 0.0, im: 0.0 }, "
                  "4: Complex { re: -0.25, im: 0.0 }, 5: Complex { re: -0.25, im: 0.0 }, "
                  "6: Complex { re: -0.25, im: 0.0 }, 7: Complex { re: 0.0, im: 0.0 }, .. }"),
}


def quantum_reg_op(op):
    """`op::h(0b1111) * op::h(0b0011).c(0b1000).unwrap() * op::swap(0b1001).c(0b0010).unwrap()`"""
    return op.h(0b1111) * op.h(0b0011).c(0b1000) * op.swap(0b1001).c(0b0010)


# register/quant.rs:680-711 `tensor` (EPS = 1e-9) and the doctest at :86-96 (Bell pair)
TENSOR_PROB = [0.25, 0.25, 0.0, 0.0, 0.25, 0.25, 0.0, 0.0]
TENSOR_EPS = 1e-9
