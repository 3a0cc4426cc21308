// qvnt.hpp -- header-only C++17 host mirror of `qvnt::prelude` over the C ABI (qvnt_b200.h).
//
// The reference's host language is Rust; no Rust toolchain exists in this image, so the host
// side above the C ABI is C++ (this header) and Python (qvnt_b200/).  Names, argument order
// `(phase, mask)`, `Option`-returning constructors (std::optional) and the reference's quirks
// are kept, so a test written against the reference reads the same here.
// Citations are relative to /root/reference/src.
//
//   prelude (lib.rs:16-24):  op, Applicable (= SingleOp / MultiOp), MultiOp, SingleOp, QReg, CReg, VReg
//
// Everything here is host metadata ("gates are lazy", operator/mod.rs:7-9); the amplitudes
// live in HBM behind the opaque qvnt_reg_t handle and only `QReg` talks to the device.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <deque>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "qvnt_b200.h"

namespace qvnt {

using N = uint64_t;                 // math/mod.rs:22 (usize)
using R = double;                   // :24
using C = std::complex<double>;     // :26

inline int count_ones(N x) { return __builtin_popcountll(x); }

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &what) : std::runtime_error(what), status(s) {}
};
inline void check(int rc) {
    if (rc != QVNT_OK) throw Error(rc, qvnt_last_error());   // the Rust wrapper panics here
}

// ---------------------------------------------------------------------------------------------
// SingleOp {act, ctrl, func} (operator/single/mod.rs:43-47); func flattened to the POD descriptor
// ---------------------------------------------------------------------------------------------
class SingleOp {
public:
    qvnt_op_t d{};      // what crosses the boundary
    N act = 0;          // acts_on() of the atomic op (u2 reports only a_mask: atomic/u2.rs:70-72)

    static SingleOp make(uint32_t kind, N a, N b = 0, C phase = C(0, 0), bool dagger = false) {
        SingleOp s;
        s.d.kind = kind;
        s.d.a_mask = a;
        s.d.b_mask = b;
        s.d.phase_re = phase.real();
        s.d.phase_im = phase.imag();
        s.d.dagger = dagger ? 1u : 0u;
        s.act = kind == QVNT_ID ? 0 : (kind == QVNT_H2 ? (a | b) : a);
        return s;
    }
    N act_on() const { return act | d.ctrl; }                     // single/mod.rs:95-97
    // `.c(mask)`: None when the mask overlaps the qubits acted on (single/mod.rs:109-118)
    std::optional<SingleOp> c(N c_mask) const {
        if (act_on() & c_mask) return std::nullopt;
        SingleOp r = *this;
        r.d.ctrl |= c_mask;
        return r;
    }
    // `.dgr()` (single/mod.rs:101-107 -> atomic/*.rs::dgr)
    SingleOp dgr() const {
        SingleOp r = *this;
        switch (d.kind) {
        case QVNT_S: case QVNT_T: case QVNT_ISWAP: case QVNT_SQRT_SWAP: case QVNT_SQRT_ISWAP:
            r.d.dagger = d.dagger ? 0u : 1u;                      // s.rs:39-44, ...
            break;
        case QVNT_RX: case QVNT_RXX: case QVNT_RY: case QVNT_RYY: case QVNT_RZ: case QVNT_RZZ:
            r.d.phase_re = -d.phase_re;                           // phase: -self.phase (rx.rs:42-47): the
            r.d.phase_im = -d.phase_im;                           // reference's dagger of a rotation is -R(t)
            break;
        case QVNT_U1:                                             // math/matrix.rs:32-35
            for (int rr = 0; rr < 2; ++rr)
                for (int cc = 0; cc < 2; ++cc) {
                    r.d.matrix[2 * (2 * rr + cc)] = d.matrix[2 * (2 * cc + rr)];
                    r.d.matrix[2 * (2 * rr + cc) + 1] = -d.matrix[2 * (2 * cc + rr) + 1];
                }
            break;
        case QVNT_U2:                                             // math/matrix.rs:76-96
            for (int rr = 0; rr < 4; ++rr)
                for (int cc = 0; cc < 4; ++cc) {
                    r.d.matrix[2 * (4 * rr + cc)] = d.matrix[2 * (4 * cc + rr)];
                    r.d.matrix[2 * (4 * rr + cc) + 1] = -d.matrix[2 * (4 * cc + rr) + 1];
                }
            break;
        default: break;
        }
        return r;
    }
    // Debug name "C{ctrl}_{gate}{mask}" (single/mod.rs:74-80) for the kinds whose name has no float
    std::string name() const {
        std::string g;
        const N a = d.a_mask;
        switch (d.kind) {
        case QVNT_ID: g = "Id"; break;
        case QVNT_X: g = "X" + std::to_string(a); break;
        case QVNT_Y: g = "Y" + std::to_string(a); break;
        case QVNT_Z: g = "Z" + std::to_string(a); break;
        case QVNT_S: g = "S" + std::to_string(a); break;
        case QVNT_T: g = "T" + std::to_string(a); break;
        case QVNT_H1: g = "H" + std::to_string(a); break;
        case QVNT_H2: g = "H" + std::to_string(a | d.b_mask); break;
        case QVNT_SWAP: g = "SWAP" + std::to_string(a); break;
        case QVNT_ISWAP: g = "iSWAP" + std::to_string(a); break;
        case QVNT_SQRT_SWAP: g = "sqrt(SWAP" + std::to_string(a) + ")"; break;
        case QVNT_SQRT_ISWAP: g = "sqrt(iSWAP" + std::to_string(a) + ")"; break;
        case QVNT_RX: g = "RX" + std::to_string(a); break;       // (+ "(angle)" in the reference)
        case QVNT_RY: g = "RY" + std::to_string(a); break;
        case QVNT_RZ: g = "RZ" + std::to_string(a); break;
        case QVNT_RXX: g = "RXX" + std::to_string(a); break;
        case QVNT_RYY: g = "RYY" + std::to_string(a); break;
        case QVNT_RZZ: g = "RZZ" + std::to_string(a); break;
        default: g = "U" + std::to_string(a | d.b_mask); break;
        }
        return (d.ctrl ? "C" + std::to_string(d.ctrl) + "_" : std::string()) + g;
    }
    bool operator==(const SingleOp &o) const {
        if (d.kind != o.d.kind || d.dagger != o.d.dagger || d.a_mask != o.d.a_mask || d.b_mask != o.d.b_mask ||
            d.ctrl != o.d.ctrl || d.phase_re != o.d.phase_re || d.phase_im != o.d.phase_im || act != o.act)
            return false;
        for (int i = 0; i < 32; ++i)
            if (d.matrix[i] != o.d.matrix[i]) return false;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// MultiOp = VecDeque<SingleOp>, applied front first (operator/multi/mod.rs:63,96-114)
// ---------------------------------------------------------------------------------------------
class MultiOp {
public:
    std::deque<SingleOp> ops;

    MultiOp() = default;
    MultiOp(const SingleOp &s) {                                   // From<SingleOp> drops "Id" (:135-146)
        if (s.d.kind != QVNT_ID) ops.push_back(s);
    }
    size_t len() const { return ops.size(); }
    N act_on() const {                                             // :116-118
        N m = 0;
        for (const auto &s : ops) m |= s.act_on();
        return m;
    }
    MultiOp dgr() const {                                          // :120-123: reverse + dagger each
        MultiOp r;
        for (auto it = ops.rbegin(); it != ops.rend(); ++it) r.ops.push_back(it->dgr());
        return r;
    }
    std::optional<MultiOp> c(N c_mask) const {                     // :125-132
        if (act_on() & c_mask) return std::nullopt;
        MultiOp r;
        for (const auto &s : ops) r.ops.push_back(*s.c(c_mask));
        return r;
    }
    MultiOp &operator*=(const MultiOp &rhs) {                      // Mul = concatenation (:148-191)
        ops.insert(ops.end(), rhs.ops.begin(), rhs.ops.end());
        return *this;
    }
    friend MultiOp operator*(MultiOp lhs, const MultiOp &rhs) { return lhs *= rhs; }
    std::string debug() const {                                    // "[H3, H12, C8_H3, C2_SWAP9]"
        std::string s = "[";
        for (size_t i = 0; i < ops.size(); ++i) s += (i ? ", " : "") + ops[i].name();
        return s + "]";
    }
    std::vector<qvnt_op_t> lower() const {
        std::vector<qvnt_op_t> v;
        v.reserve(ops.size());
        for (const auto &s : ops) v.push_back(s.d);
        return v;
    }
    bool operator==(const MultiOp &o) const { return ops == o.ops; }
};

// ---------------------------------------------------------------------------------------------
// op::* (operator/mod.rs:113-522) and the Option-returning single::* constructors
// ---------------------------------------------------------------------------------------------
namespace single {
inline std::optional<SingleOp> checked(SingleOp s, int bits) {     // single/mod.rs:4-11 (is_valid)
    if (count_ones(s.d.a_mask) != bits) return std::nullopt;
    return s;
}
inline C half(R phase) { return C(std::cos(phase / 2.0), std::sin(phase / 2.0)); }
inline SingleOp x(N a) { return SingleOp::make(QVNT_X, a); }
inline SingleOp y(N a) { return SingleOp::make(QVNT_Y, a); }
inline SingleOp z(N a) { return SingleOp::make(QVNT_Z, a); }
inline SingleOp s(N a) { return SingleOp::make(QVNT_S, a); }
inline SingleOp t(N a) { return SingleOp::make(QVNT_T, a); }
inline SingleOp h1(N a) { return SingleOp::make(QVNT_H1, a); }
inline SingleOp h2(N a, N b) { return SingleOp::make(QVNT_H2, a, b); }
inline std::optional<SingleOp> rx(N a, R p) { return checked(SingleOp::make(QVNT_RX, a, 0, half(p)), 1); }
inline std::optional<SingleOp> ry(N a, R p) { return checked(SingleOp::make(QVNT_RY, a, 0, half(p)), 1); }
inline std::optional<SingleOp> rz(N a, R p) { return checked(SingleOp::make(QVNT_RZ, a, 0, half(p)), 1); }
inline std::optional<SingleOp> rxx(N ab, R p) {                     // rxx.rs:10-14 uses phase * 0.5
    return checked(SingleOp::make(QVNT_RXX, ab, 0, C(std::cos(p * 0.5), std::sin(p * 0.5))), 2);
}
inline std::optional<SingleOp> ryy(N ab, R p) { return checked(SingleOp::make(QVNT_RYY, ab, 0, half(p)), 2); }
inline std::optional<SingleOp> rzz(N ab, R p) { return checked(SingleOp::make(QVNT_RZZ, ab, 0, half(p)), 2); }
inline std::optional<SingleOp> swap(N ab) { return checked(SingleOp::make(QVNT_SWAP, ab), 2); }
inline std::optional<SingleOp> i_swap(N ab) { return checked(SingleOp::make(QVNT_ISWAP, ab), 2); }
inline std::optional<SingleOp> sqrt_swap(N ab) { return checked(SingleOp::make(QVNT_SQRT_SWAP, ab), 2); }
inline std::optional<SingleOp> sqrt_i_swap(N ab) { return checked(SingleOp::make(QVNT_SQRT_ISWAP, ab), 2); }
}  // namespace single

namespace op {
inline MultiOp expect(const std::optional<SingleOp> &s, const char *msg) {   // `.expect(..)` panics
    if (!s) throw std::invalid_argument(msg);
    return MultiOp(*s);
}
inline MultiOp id() { return MultiOp(); }
inline MultiOp x(N a) { return MultiOp(single::x(a)); }
inline MultiOp y(N a) { return MultiOp(single::y(a)); }
inline MultiOp z(N a) { return MultiOp(single::z(a)); }
inline MultiOp s(N a) { return MultiOp(single::s(a)); }
inline MultiOp t(N a) { return MultiOp(single::t(a)); }
inline MultiOp rx(R p, N a) { return expect(single::rx(a, p), "Mask should contain 1 bit!"); }
inline MultiOp ry(R p, N a) { return expect(single::ry(a, p), "Mask should contain 1 bit!"); }
inline MultiOp rz(R p, N a) { return expect(single::rz(a, p), "Mask should contain 1 bit!"); }
inline MultiOp rxx(R p, N ab) { return expect(single::rxx(ab, p), "Mask should contain 2 bit!"); }
inline MultiOp ryy(R p, N ab) { return expect(single::ryy(ab, p), "Mask should contain 2 bit!"); }
inline MultiOp rzz(R p, N ab) { return expect(single::rzz(ab, p), "Mask should contain 2 bit!"); }
inline MultiOp swap(N ab) { return expect(single::swap(ab), "Mask should contain 2 bit!"); }
inline MultiOp i_swap(N ab) { return expect(single::i_swap(ab), "Mask should contain 2 bit!"); }
inline MultiOp sqrt_swap(N ab) { return expect(single::sqrt_swap(ab), "Mask should contain 2 bit!"); }
inline MultiOp sqrt_i_swap(N ab) { return expect(single::sqrt_i_swap(ab), "Mask should contain 2 bit!"); }

// multi/h.rs:14-45: set bits paired low->high into H2(hi, lo), trailing H1
inline MultiOp h(N a_mask) {
    const int count = count_ones(a_mask);
    if (count == 0) return MultiOp();
    if (count == 1) return MultiOp(single::h1(a_mask));
    MultiOp res;
    N first = 0;
    bool is_first = true;
    for (N bit = 1; bit && bit <= a_mask; bit <<= 1) {
        if (!(bit & a_mask)) continue;
        if (is_first) {
            first = bit;
            is_first = false;
        } else {
            res.ops.push_back(single::h2(bit, first));
            is_first = true;
        }
    }
    if (!is_first) res.ops.push_back(single::h1(first));
    return res;
}
inline MultiOp u1(R lam, N a) { return rz(lam, a); }                                   // mod.rs:472
inline MultiOp u2(R phi, R lam, N a) { return rz(lam, a) * ry(M_PI / 2.0, a) * rz(phi, a); }   // :480
inline MultiOp u3(R the, R phi, R lam, N a) { return rz(lam, a) * ry(the, a) * rz(phi, a); }   // :499
// multi/qft.rs:4-33: H1(v[i]) then RZ(v[i+j], pi * 0.5^j).c(v[i])
inline MultiOp qft(N a_mask) {
    std::vector<N> v;
    for (int i = 0; i < 64; ++i)
        if ((a_mask >> i) & 1) v.push_back(N(1) << i);
    const size_t count = v.size();
    if (count == 0) return MultiOp();
    if (count == 1) return h(a_mask);
    MultiOp res;
    for (size_t i = 0; i + 1 < count; ++i) {
        res *= h(v[i]);
        for (size_t j = 1; j < count - i; ++j)
            res.ops.push_back(*single::rz(v[i + j], M_PI * std::pow(0.5, (int)j))->c(v[i]));
    }
    res *= h(v[count - 1]);
    return res;
}
inline MultiOp qft_swapped(N a_mask) {                                                  // :35-52
    std::vector<N> v;
    for (int i = 0; i < 64; ++i)
        if ((a_mask >> i) & 1) v.push_back(N(1) << i);
    MultiOp res = qft(a_mask);
    for (size_t i = 0; i < v.size() / 2; ++i) res.ops.push_back(*single::swap(v[i] | v[v.size() - i - 1]));
    return res;
}
}  // namespace op

// ---------------------------------------------------------------------------------------------
// CReg (register/class.rs) and VReg (register/virtl.rs): host integers / mask sugar
// ---------------------------------------------------------------------------------------------
class CReg {
    N value_, q_num_, q_mask_;
public:
    explicit CReg(N q_num) : value_(0), q_num_(q_num), q_mask_(q_num >= 64 ? ~N(0) : (N(1) << q_num) - 1) {}
    static CReg with_state(N q_num, N state) {
        CReg r(q_num);
        r.value_ = state & r.q_mask_;
        return r;
    }
    N get() const { return value_; }
    N get_by_mask(N mask) const { return value_ & mask; }
    void set(N v, N mask) { value_ = (value_ & ~mask) | (v & mask & q_mask_); }
    void xor_(N v) { value_ ^= v & q_mask_; }
    N num() const { return q_num_; }
};

class VReg {
    std::vector<N> bits_;
    N mask_;
public:
    explicit VReg(N mask) : mask_(mask) {
        for (int i = 0; i < 64; ++i)
            if ((mask >> i) & 1) bits_.push_back(N(1) << i);
    }
    N operator[](size_t i) const { return bits_.at(i); }          // v[3]
    N all() const { return mask_; }                                // v[..]
    N of(std::initializer_list<size_t> idx) const {                // v[[0, 7]]
        N m = 0;
        for (size_t i : idx) m |= bits_.at(i);
        return m;
    }
    size_t len() const { return bits_.size(); }
};

// ---------------------------------------------------------------------------------------------
// QReg (register/quant.rs): the state vector, resident in B200 HBM
// ---------------------------------------------------------------------------------------------
class QReg {
    qvnt_reg_t *h_ = nullptr;
    N q_num_ = 0, q_mask_ = 0;
    explicit QReg(qvnt_reg_t *h, N q) : h_(h), q_num_(q), q_mask_(q >= 64 ? ~N(0) : (N(1) << q) - 1) {}
public:
    static QReg new_(N q_num) { return with_state(q_num, 0); }     // quant.rs:113 (`new` is a C++ keyword)
    static QReg with_state(N q_num, N state) {                     // quant.rs:129
        qvnt_reg_t *h = nullptr;
        check(qvnt_reg_create((uint32_t)q_num, state, &h));
        return QReg(h, q_num);
    }
    QReg(const QReg &o) : q_num_(o.q_num_), q_mask_(o.q_mask_) { check(qvnt_reg_clone(o.h_, &h_)); }   // Clone
    QReg(QReg &&o) noexcept : h_(o.h_), q_num_(o.q_num_), q_mask_(o.q_mask_) { o.h_ = nullptr; }
    QReg &operator=(QReg o) {
        std::swap(h_, o.h_);
        q_num_ = o.q_num_;
        q_mask_ = o.q_mask_;
        return *this;
    }
    ~QReg() { if (h_) qvnt_reg_destroy(h_); }                      // Drop

    N num() const { return q_num_; }
    // quant.rs:186-200 with GPUs for threads: the register continues on n GPUs (1, 2, 4 or 8 of this
    // box, sharded by its top qubits, state kept); nullopt for 0 or more than the box has
    std::optional<QReg> num_threads(size_t n) && {
        int ndev = 0;
        qvnt_device_count(&ndev);
        if (n == 0 || (n & (n - 1)) || n > 8 || n > (size_t)ndev) return std::nullopt;
        qvnt_reg_t *h = nullptr;
        check(qvnt_reg_set_gpus(h_, (uint32_t)n, &h));
        QReg out(std::move(*this));
        qvnt_reg_destroy(out.h_);
        out.h_ = h;
        return out;
    }
    void apply(const MultiOp &op) {                                // quant.rs:376
        const auto v = op.lower();
        if (!v.empty()) check(qvnt_reg_apply(h_, v.data(), v.size()));
    }
    void apply(const SingleOp &op) { check(qvnt_reg_apply(h_, &op.d, 1)); }
    CReg measure_mask(N mask, std::optional<double> u01 = std::nullopt) {   // quant.rs:490
        uint64_t out = 0;
        if (u01) check(qvnt_reg_measure_mask(h_, mask, *u01, &out, nullptr));
        else check(qvnt_reg_measure_mask_rng(h_, mask, &out));
        return CReg::with_state(q_num_, out);
    }
    CReg measure() { return measure_mask(q_mask_); }               // :505
    void collapse_mask(N idy, N mask) { check(qvnt_reg_collapse(h_, idy, mask)); }
    QReg &normalize() { check(qvnt_reg_normalize(h_)); return *this; }
    void reset(N state) { check(qvnt_reg_reset(h_, state)); }
    void reset_by_mask(N mask) { check(qvnt_reg_reset_by_mask(h_, mask)); }
    R get_absolute() { double v = 0; check(qvnt_reg_norm_sqr(h_, &v)); return v; }
    std::vector<R> get_probabilities() {                           // :434
        std::vector<R> p(size_t(1) << q_num_);
        check(qvnt_reg_probabilities(h_, 0, p.size(), p.data()));
        return p;
    }
    std::vector<std::pair<R, R>> get_polar() {                     // :417
        std::vector<std::pair<R, R>> p(size_t(1) << q_num_);
        check(qvnt_reg_polar(h_, 0, p.size(), reinterpret_cast<double *>(p.data())));
        return p;
    }
    std::vector<C> amplitudes() {                                  // psi (Debug fmt, tests)
        std::vector<C> a(size_t(1) << q_num_);
        check(qvnt_reg_read(h_, 0, a.size(), reinterpret_cast<double *>(a.data())));
        return a;
    }
    void write_amplitudes(const std::vector<C> &a, N off = 0) {
        check(qvnt_reg_write(h_, off, a.size(), reinterpret_cast<const double *>(a.data())));
    }
    VReg get_vreg() const { return VReg(q_mask_); }                // :232
    std::optional<VReg> get_vreg_by(N mask) const {                // :236
        if (mask & ~q_mask_) return std::nullopt;
        return VReg(mask);
    }
    friend QReg operator*(QReg &a, QReg &b) {                      // tensor_prod :330-371, Mul :625-636
        qvnt_reg_t *h = nullptr;
        check(qvnt_reg_tensor_prod(a.h_, b.h_, &h));
        return QReg(h, a.q_num_ + b.q_num_);
    }
    void sync() { check(qvnt_reg_sync(h_)); }
    qvnt_reg_t *handle() { return h_; }
};

// `Applicable::matrix(size)` (operator/applicable.rs:17-46): apply to each basis vector, transpose
template <class Op>
inline std::vector<std::vector<C>> matrix(const Op &op, N size) {
    const size_t dim = size_t(1) << size;
    std::vector<std::vector<C>> m(dim, std::vector<C>(dim));
    for (size_t idx = 0; idx < dim; ++idx) {
        QReg r = QReg::with_state(size, idx);
        r.apply(op);
        const auto col = r.amplitudes();
        for (size_t i = 0; i < dim; ++i) m[i][idx] = col[i];
    }
    return m;
}

}  // namespace qvnt
