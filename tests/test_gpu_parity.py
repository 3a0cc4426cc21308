"""Parity tests proper: the CUDA path, called through the C ABI (qvnt_b200.QReg -> ctypes ->
libqvnt_b200.so), against the CPU oracle and the reference's golden vectors.

Bars: masks / indices bit-exact, amplitudes <= 1e-10 abs, probabilities <= 1e-12.
With fusion off ("fuse"=0) every SingleOp is one in-place sweep using the reference's own
formulas without FMA contraction, and we additionally demand BIT-EXACT amplitudes.
"""
import math

import numpy as np
import pytest

from qvnt_b200 import QReg, QvntError, op, workloads
from qvnt_b200.op import MultiOp, SingleOp, single
from tests.golden import reference_kat as kat

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-10
PROB_TOL = 1e-12


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return v / np.linalg.norm(v)


def both(oracle, n, seed):
    v = rand_state(n, seed)
    g = QReg.new(n)
    g.write_amplitudes(v)
    o = oracle.OracleReg.new(n)
    o.write_amplitudes(v)
    return g, o


def assert_close(g, o, tol=AMP_TOL, exact=False):
    a, b = g.amplitudes(), o.amplitudes()
    if exact:
        assert np.array_equal(a.view(np.float64), b.view(np.float64)), np.abs(a - b).max()
    else:
        assert np.abs(a - b).max() <= tol, np.abs(a - b).max()


# ---------------------------------------------------------------- reference golden vectors
@pytest.mark.parametrize("label,build,name,size,expected", kat.ATOMIC_KATS, ids=[k[0] for k in kat.ATOMIC_KATS])
@pytest.mark.parametrize("fuse", [0, 1])
def test_atomic_matrix_repr(label, build, name, size, expected, fuse):
    def factory(q, s):
        r = QReg.with_state(q, s)
        r.set_option("fuse", fuse)
        return r
    got = np.array(op._matrix(build(op), size, factory), dtype=np.complex128)
    exp = np.array(expected, dtype=np.complex128)
    assert np.array_equal(got.real, exp.real) and np.array_equal(got.imag, exp.imag), (label, got)


@pytest.mark.parametrize("fuse", [0, 1])
def test_quantum_reg_golden(fuse):
    g = kat.QUANTUM_REG
    reg = QReg.with_state(g["q_num"], g["state"])
    reg.set_option("fuse", fuse)
    reg.apply(kat.quantum_reg_op(op))
    psi = reg.amplitudes()
    assert np.array_equal(psi.real, np.array(g["psi"])) and not psi.imag.any()
    assert repr(reg) == g["reg_debug"]
    assert reg.measure_mask(g["mask"]).get() & ~g["mask"] == 0


def test_bell_and_tensor_golden():
    q = QReg.new(2)
    q.apply(op.h(0b01) * op.x(0b10).c(0b01))
    assert list(q.get_probabilities()) == [0.5, 0.0, 0.0, 0.5]
    r1, r2 = QReg.with_state(2, 0b01), QReg.with_state(1, 0b1)
    r1.apply(op.h(0b01))
    r2.apply(op.h(0b01))
    p = (r1 * r2).get_probabilities()
    assert np.all(np.abs(p - np.array(kat.TENSOR_PROB)) < kat.TENSOR_EPS)
    r3 = QReg.with_state(3, 0b101)
    r3.apply(op.h(0b101))
    assert np.all(np.abs(r3.get_probabilities() - np.array(kat.TENSOR_PROB)) < kat.TENSOR_EPS)


# ---------------------------------------------------------------- every kind, every bit position
def _all_kind_ops(n, rng):
    def bit():
        return 1 << int(rng.integers(n))

    def two():
        a, b = rng.choice(n, 2, replace=False)
        return (1 << int(a)) | (1 << int(b)), 1 << int(a), 1 << int(b)

    def mask():
        return int(rng.integers(1, 1 << n))
    th = float(rng.uniform(0, 2 * math.pi))
    ab, a, b = two()
    c, s = math.cos(th), math.sin(th)
    e = complex(math.cos(0.3), math.sin(0.3))
    return [
        single.x(mask()), single.y(mask()), single.z(mask()), single.s(mask()), single.s(mask()).dgr(),
        single.t(mask()), single.t(mask()).dgr(), single.h1(bit()), single.h2(a, b),
        single.rx(bit(), th), single.ry(bit(), th), single.rz(bit(), th), single.rx(bit(), th).dgr(),
        single.rxx(ab, th), single.ryy(ab, th), single.rzz(ab, th), single.rzz(ab, th).dgr(),
        single.swap(ab), single.i_swap(ab), single.i_swap(ab).dgr(), single.sqrt_swap(ab),
        single.sqrt_swap(ab).dgr(), single.sqrt_i_swap(ab), single.sqrt_i_swap(ab).dgr(),
        SingleOp(op.K_U1, bit(), matrix=[c * e, -s, s, c * e.conjugate()]),
        SingleOp(op.K_U1, bit(), matrix=[c * e, -s, s, c * e.conjugate()]).dgr(),
        SingleOp(op.K_U2, a, b, matrix=[c, 0, 0, -s * e, 0, c, -s, 0, 0, s, c, 0, s * e.conjugate(), 0, 0, c]),
        SingleOp(op.K_U2, a, b, matrix=[c, 0, 0, -s * e, 0, c, -s, 0, 0, s, c, 0, s * e.conjugate(), 0, 0, c]).dgr(),
    ]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 14])
@pytest.mark.parametrize("fuse", [0, 1])
def test_every_kind_single_sweep(oracle, n, fuse):
    rng = np.random.default_rng(100 + n)
    for rep in range(3 if n > 1 else 1):
        ops = _all_kind_ops(n, rng) if n >= 2 else [
            single.x(1), single.y(1), single.z(1), single.h1(1), single.rx(1, 0.3), single.t(1)]
        for sop in ops:
            targets = sop.a_mask | sop.b_mask
            if targets >> n:
                continue
            variants = [sop]
            free = ((1 << n) - 1) & ~targets
            if free:
                cm = free & int(rng.integers(1, 1 << n))
                if cm:
                    variants.append(sop.c(cm))
            for v in variants:
                g, o = both(oracle, n, seed=rep)
                g.set_option("fuse", fuse)
                g.apply(v)
                o.apply(v)
                assert_close(g, o, exact=True)     # one op per pass: same arithmetic, no FMA


@pytest.mark.parametrize("n", [6, 11, 16])
@pytest.mark.parametrize("fuse", [0, 1, 3])
def test_mixed_circuit_all_kinds(oracle, n, fuse):
    circ = workloads.mixed_all_kinds(n, 150, seed=n)
    g, o = both(oracle, n, seed=n)
    g.set_option("fuse", fuse & 1)
    g.set_option("lower_two_bit", fuse >> 1)      # 3: swap / i_swap / rxx / ryy as products of fast kinds
    g.apply(circ)
    o.apply(circ)
    assert_close(g, o, exact=(fuse == 0))
    assert abs(g.get_absolute() - o.get_absolute()) <= PROB_TOL


@pytest.mark.parametrize("n,layers,seed", [(12, 500, 3), (16, 400, 11), (20, 250, 5)])
@pytest.mark.parametrize("opts", [{}, {"single_ctrl": 0, "butterfly": 0}, {"ptx_ops": 0}])
def test_fast_interpreter_forms_with_inverted_slots(oracle, n, layers, seed, opts):
    """x / cx leave register slots inverted (lazy); controlled phases, rz, h and rotations follow on the
    same slots: the single-control arms, the butterfly h, the generic masked arms and the C++ loop must
    all pick the exchanged coefficient block / predicates."""
    circ = workloads.fast_mix(n, layers, seed)
    g, o = both(oracle, n, seed=seed)
    for k, v in opts.items():
        g.set_option(k, v)
    g.apply(circ)
    o.apply(circ)
    assert_close(g, o)


@pytest.mark.parametrize("n,depth", [(12, 20), (18, 10), (22, 4)])
def test_random_layered_config2_scaled(oracle, n, depth):
    # configs[1] generator at oracle-feasible sizes
    circ = workloads.random_layered(n, depth)
    g, o = QReg.new(n), oracle.OracleReg.new(n, threads=oracle.max_threads())
    g.apply(circ)
    o.apply(circ)
    assert_close(g, o)
    pg, po = g.get_probabilities(), o.get_probabilities()
    assert np.abs(pg - po).max() <= PROB_TOL


# ---------------------------------------------------------------- config 1: 20-qubit QFT + measure
@pytest.mark.parametrize("state", [0, 0x5A5A5])
def test_config1_qft20_measure(oracle, state):
    n = 20
    circ = op.qft(0xFFFFF)
    assert len(circ) == 210
    g, o = QReg.with_state(n, state), oracle.OracleReg.with_state(n, state, threads=oracle.max_threads())
    g.apply(circ)
    o.apply(circ)
    assert_close(g, o)
    p = g.get_probabilities()
    assert np.abs(p - 2.0 ** -n).max() <= PROB_TOL
    po = o.get_probabilities()
    cum = np.cumsum(po)
    for mask in (0xFFFFF, 0b100):
        for u in (0.0, 0.25, 0.5, 0.999999):
            g2, o2 = g.clone(), o.clone()
            out_g, idx_g = g2.measure_mask_full(mask, u)
            out_o, idx_o = o2.measure_mask_full(mask, u)
            assert out_g == idx_g & mask
            if idx_g != idx_o:
                # only legal when u*total sits within rounding of a cumulative boundary
                x = u * cum[-1]
                lo = cum[idx_g - 1] if idx_g else 0.0
                assert lo <= x + 1e-12 and cum[idx_g] > x - 1e-12, (u, idx_g, idx_o)
                o2 = o.clone()
                o2.collapse_mask(idx_g, mask)
            assert_close(g2, o2)                    # collapsed, NOT renormalised (quant.rs:468-501)


def test_measure_matches_oracle_generic(oracle):
    n = 13
    g, o = both(oracle, n, seed=5)
    for u in (0.0, 0.1234, 0.5, 0.87, 0.999999):
        for mask in (0b1, 0b1010101, (1 << n) - 1, 1 << (n - 1)):
            g2, o2 = g.clone(), o.clone()
            rg, ro = g2.measure_mask_full(mask, u), o2.measure_mask_full(mask, u)
            assert rg == ro
            assert_close(g2, o2, exact=True)
    assert g.measure_mask(0).get() == 0             # quant.rs:491-494


def test_normalize_reset_collapse(oracle):
    n = 10
    g, o = both(oracle, n, seed=9)
    g.collapse_mask(0b1100110011, 0b0101010101)
    o.collapse_mask(0b1100110011, 0b0101010101)
    assert_close(g, o, exact=True)
    assert abs(g.get_absolute() - o.get_absolute()) <= PROB_TOL
    g.normalize()
    o.normalize()
    assert_close(g, o)
    g.reset_by_mask(0b11)
    o.reset_by_mask(0b11)
    assert_close(g, o)
    g.reset_by_mask((1 << n) - 1)
    a = g.amplitudes()
    assert a[0] == 1 and not a[1:].any()
    g.reset(77)
    a = g.amplitudes()
    assert a[77] == 1 and np.count_nonzero(a) == 1
    # norm <= 1e-15 -> reset(0)  (quant.rs:399-402)
    g.collapse_mask(0, 1 << 6)       # state 77 has bit 6 set -> everything zeroed
    g.normalize()
    a = g.amplitudes()
    assert a[0] == 1 and np.count_nonzero(a) == 1


def test_error_behaviour():
    r = QReg.new(5)
    with pytest.raises(QvntError) as e:
        r.apply(single.x(1 << 5))
    assert e.value.status == 2          # BAD_MASK: the reference would index out of bounds
    bad = SingleOp(op.K_RX, 0b11, phase=(1.0, 0.0))
    with pytest.raises(QvntError):
        r.apply(bad)
    overlap = SingleOp(op.K_X, 0b1, ctrl=0b1)
    with pytest.raises(QvntError):
        r.apply(overlap)
    r.apply(MultiOp())                  # empty op list is fine
    assert r.amplitudes()[0] == 1


def test_dgr_roundtrip_non_rotation():
    # s/t/iswap/sqrt-swap daggers invert; h/x/y/z/swap are involutions (rotations: see quirk 2)
    n = 12
    v = rand_state(n, 3)
    g = QReg.new(n)
    g.write_amplitudes(v)
    circ = (op.h(0xABC) * op.s(0x0F0) * op.t(0x111) * op.i_swap(0b11000) * op.sqrt_swap(0b101)
            * op.sqrt_i_swap(0b1000010) * op.x(0x333) * op.y(0x444) * op.z(0x888) * op.swap(0x801))
    g.apply(circ)
    g.apply(circ.dgr())
    assert np.abs(g.amplitudes() - v).max() <= AMP_TOL


# ---------------------------------------------------------------- full-size properties (no oracle)
@pytest.mark.parametrize("n", [26, 30])
def test_fullsize_qft_uniform(n):
    r = QReg.with_state(n, 1)
    r.apply(op.qft((1 << n) - 1))
    assert abs(r.get_absolute() - 1.0) <= 1e-9
    rng = np.random.default_rng(n)
    for off in rng.integers(0, (1 << n) - 4096, 8):
        off = int(off)
        a = r.amplitudes(off, 4096)
        assert np.abs(np.abs(a) ** 2 - 2.0 ** -n).max() <= 2.0 ** -n * 1e-9   # << 1e-12 absolute
    # hadamard transform of |0..0> is uniform with zero phase; H twice is the identity
    r.reset(0)
    r.apply(op.h((1 << n) - 1))
    a = r.amplitudes(12345, 1024)
    assert np.abs(a - 2.0 ** (-n / 2)).max() <= 1e-12
    r.apply(op.h((1 << n) - 1))
    a = r.amplitudes(0, 1024)
    assert abs(a[0] - 1) <= 1e-10 and np.abs(a[1:]).max() <= 1e-10
    out, idx = r.measure_mask_full((1 << n) - 1, 0.5)
    assert out == idx == 0


def test_fullsize_config2_fused_equals_unfused():
    # 28-qubit configs[1] circuit, first 2 layers: fused schedule vs one-sweep-per-gate
    n = 28
    circ = MultiOp(list(workloads.random_layered(n, 2)))
    a, b = QReg.new(n), QReg.new(n)
    b.set_option("fuse", 0)
    a.apply(circ)
    b.apply(circ)
    assert abs(a.get_absolute() - 1.0) <= 1e-9
    rng = np.random.default_rng(1)
    for off in rng.integers(0, (1 << n) - (1 << 16), 6):
        off = int(off)
        assert np.abs(a.amplitudes(off, 1 << 16) - b.amplitudes(off, 1 << 16)).max() <= AMP_TOL


def test_polar_clone_and_sample_all(oracle):
    """get_polar (quant.rs:417-431: Complex::to_polar = (norm, arg)), Clone (:102) and the shape
    of sample_all (quant.rs:513-594 is only pinned by `histogram`, :714-725: length and sum)."""
    n = 10
    g, o = both(oracle, n, seed=42)
    g.apply(workloads.mixed_all_kinds(n, 30, seed=5))
    o.apply(workloads.mixed_all_kinds(n, 30, seed=5))
    a = o.amplitudes()
    pol = g.get_polar()
    assert np.abs(pol[:, 0] - np.abs(a)).max() <= 1e-12
    big = np.abs(a) > 1e-6
    d = np.angle(np.exp(1j * (pol[big, 1] - np.angle(a[big]))))
    assert np.abs(d).max() <= 1e-9
    c = g.clone()
    g.apply(op.x(1))                                  # the clone owns its own buffer
    assert np.abs(c.amplitudes() - a).max() <= AMP_TOL
    hist = c.sample_all(2048, seed=1)
    assert len(hist) == 1 << n and int(hist.sum()) == 2048
    p = c.get_probabilities()
    assert np.abs(np.array(hist) / 2048.0 - p).max() < 0.05


def test_sample_all_on_device_statistics(oracle):
    """quant.rs:513-594 on the device: the histogram sums to `count` (the reference's +-delta
    correction), is reproducible per seed, and every bin stays within the Gaussian model's spread
    (mean c p_i, sigma sqrt(c p_i (1 - p_i))) -- statistical parity, the reference draws from
    thread_rng.  Also over a one-handle multi-GPU register."""
    import os
    n = 14
    circ = workloads.random_layered(n, 3)
    g = QReg.new(n)
    g.apply(circ)
    p = g.get_probabilities()
    mean, sd = oracle.sample_all_moments(p, 1 << 22)
    for count in (1 << 22, 1000, 3):
        h1 = g.sample_all(count, seed=7)
        h2 = g.sample_all(count, seed=7)
        h3 = g.sample_all(count, seed=8)
        assert h1.dtype == np.uint64 and h1.size == 1 << n
        assert int(h1.sum()) == count and int(h3.sum()) == count
        assert np.array_equal(h1, h2)
        if count > 1000:
            assert not np.array_equal(h1, h3)
            big = mean > 20.0                                 # bins where rounding to integers does not dominate
            assert big.sum() > 100
            z = (h1.astype(np.float64)[big] - mean[big]) / sd[big]
            assert np.abs(z).max() < 7.0                      # |z| > 7 has probability ~ 1e-12 per bin
            assert 0.8 < z.std() < 1.2                        # the spread is the model's, not zero
            assert np.abs(h1.astype(np.float64)[~big] - mean[~big]).max() < 7.0 * np.sqrt(20.0) + 1.0
    os.environ["QVNT_MULTI_SHARE_DEVICES"] = "1"
    try:
        m = QReg.multi(n, 0, 2)
        m.apply(circ)
        hm = m.sample_all(1 << 22, seed=7)
        m.close()
    finally:
        del os.environ["QVNT_MULTI_SHARE_DEVICES"]
    assert np.array_equal(hm, g.sample_all(1 << 22, seed=7))   # the generator is keyed by the global index
    g.close()


def test_combine_and_linear_composition(oracle):
    """The crate-private register helpers of quant.rs:245-328 (untested in the reference) against
    their numpy restatement."""
    n = 9
    a, oa = both(oracle, n, seed=3)
    b, ob = both(oracle, n, seed=4)
    va, vb = oa.amplitudes().copy(), ob.amplitudes().copy()
    c = QReg.combine(a, b)
    assert c.q_num == n + 1
    assert np.array_equal(c.amplitudes(), oracle.combine(va, vb))
    h = 0.5 ** 0.5
    u = [h, h, h, -h]
    cu = QReg.combine_with_unitary(a, b, u)
    assert np.abs(cu.amplitudes() - oracle.combine_with_unitary(va, vb, u)).max() <= 1e-15
    u2 = [0.6, 0.8j, 0.8j, 0.6]
    cu2 = QReg.combine_with_unitary(a, b, u2)
    assert np.abs(cu2.amplitudes() - oracle.combine_with_unitary(va, vb, u2)).max() <= 1e-15
    assert QReg.combine(a, QReg.new(n + 1)) is None
    a.linear_composition(b, (0.3 - 0.4j, 0.2 + 0.9j))
    assert np.abs(a.amplitudes() - oracle.linear_composition(va, vb, (0.3 - 0.4j, 0.2 + 0.9j))).max() <= 1e-15
    for r in (a, b, c, cu, cu2):
        r.close()
