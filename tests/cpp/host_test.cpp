// C++ host-mirror test: the reference's own register / operator tests, restated against
// include/qvnt.hpp (citations: /root/reference/src).  Built by tests/test_cpp_host.py with
// g++ -std=c++17 -Iinclude ... -lqvnt_b200; run on a GPU box.  `--host-only` runs the checks
// that need no device (operator algebra, lowering).
#include <cstdio>
#include <cstring>

#include "qvnt.hpp"

using namespace qvnt;

static int failures = 0;
#define CHECK(cond)                                                            \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
            ++failures;                                                        \
        }                                                                      \
    } while (0)

static void host_only() {
    // operator/multi/mod.rs:201-209: id() is dropped
    MultiOp m = op::id() * op::x(0b001) * op::h(0b110) * op::id();
    CHECK(m.len() == 2);
    // operator/single/mod.rs:142-161: names, control-mask overlap -> None
    CHECK(single::x(0b11).name() == "X3");
    CHECK(single::h2(0b100, 0b001).c(0b010)->name() == "C2_H5");
    CHECK(!single::x(0b011).c(0b001).has_value());
    CHECK(!op::h(0b0011).c(0b0010).has_value());
    // register/quant.rs:652: Debug of the op list
    MultiOp q = op::h(0b1111) * *op::h(0b0011).c(0b1000) * *op::swap(0b1001).c(0b0010);
    CHECK(q.debug() == "[H3, H12, C8_H3, C2_SWAP9]");
    // constructors validate their masks (single/mod.rs:4-11)
    CHECK(!single::rx(0b11, 1.0).has_value() && single::rx(0b10, 1.0).has_value());
    CHECK(!single::swap(0b111).has_value());
    // rotation .dgr() negates the whole phase (rx.rs:42-47)
    const SingleOp r = *single::rx(1, 1.23456), rd = r.dgr();
    CHECK(rd.d.phase_re == -r.d.phase_re && rd.d.phase_im == -r.d.phase_im);
    // qft lowering: n H1 + n(n-1)/2 controlled RZ (multi/qft.rs:4-33)
    CHECK(op::qft(0xFFFFF).len() == 20 + 190);
    CHECK(op::u3(0.1, 0.2, 0.3, 0b1).len() == 3);
    CHECK(sizeof(qvnt_op_t) == 304);
}

static void device() {
    // register/quant.rs:643-677 `quantum_reg`
    QReg reg = QReg::with_state(4, 0b1100);
    MultiOp q = op::h(0b1111) * *op::h(0b0011).c(0b1000) * *op::swap(0b1001).c(0b0010);
    reg.apply(q);
    const double golden[16] = {0.25, 0.25, 0.25, 0.0, -0.25, -0.25, -0.25, 0.0,
                               -0.5, 0.0,  0.25, 0.0, 0.5,   0.0,   -0.25, 0.0};
    const auto psi = reg.amplitudes();
    for (int i = 0; i < 16; ++i) CHECK(psi[i].real() == golden[i] && psi[i].imag() == 0.0);
    const N mask = 0b0110;
    CHECK((reg.measure_mask(mask).get() & ~mask) == 0);
    // doctest quant.rs:86-96: Bell pair
    QReg bell = QReg::new_(2);
    bell.apply(op::h(0b01) * *op::x(0b10).c(0b01));
    const auto p = bell.get_probabilities();
    CHECK(p[0] == 0.5 && p[1] == 0.0 && p[2] == 0.0 && p[3] == 0.5);
    // atomic/x.rs:38-50 matrix_repr: 2-qubit embedding of X on bit 0
    const auto mx = matrix(op::x(0b01), 2);
    const double x2[4][4] = {{0, 1, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 1}, {0, 0, 1, 0}};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) CHECK(mx[i][j] == C(x2[i][j], 0.0));
    // atomic/h1.rs:47-56
    const auto mh = matrix(MultiOp(single::h1(0b1)), 1);
    const double h = 0.70710678118654752440;
    CHECK(mh[0][0] == C(h, 0) && mh[0][1] == C(h, 0) && mh[1][0] == C(h, 0) && mh[1][1] == C(-h, 0));
    // tensor product (quant.rs:680-711)
    QReg r1 = QReg::with_state(2, 0b01), r2 = QReg::with_state(1, 0b1);
    r1.apply(op::h(0b01));
    r2.apply(op::h(0b01));
    QReg r3 = r1 * r2;
    const double tp[8] = {0.25, 0.25, 0.0, 0.0, 0.25, 0.25, 0.0, 0.0};
    const auto p3 = r3.get_probabilities();
    for (int i = 0; i < 8; ++i) CHECK(std::fabs(p3[i] - tp[i]) < 1e-9);
    // qft of a basis state: uniform probabilities
    QReg f = QReg::with_state(12, 0x5A5);
    f.apply(op::qft(0xFFF));
    double worst = 0;
    for (double v : f.get_probabilities()) worst = std::fmax(worst, std::fabs(v - 1.0 / 4096));
    CHECK(worst < 1e-12);
    // out-of-range mask: the C ABI reports an error, the wrapper throws (Rust: panic)
    bool threw = false;
    try { f.apply(op::x(N(1) << 20)); } catch (const Error &e) { threw = e.status == QVNT_ERR_BAD_MASK; }
    CHECK(threw);
}

int main(int argc, char **argv) {
    host_only();
    if (!(argc > 1 && !std::strcmp(argv[1], "--host-only"))) device();
    std::printf(failures ? "%d check(s) failed\n" : "cpp host mirror ok\n", failures);
    return failures ? 1 : 0;
}
