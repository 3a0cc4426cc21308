// gates.cuh -- the 20 atomic gate formulas as device functions.
//
// Every function computes ONE output amplitude in the reference's gather form
// (out[i] = f(in[i], in[i ^ m], i)), in exactly the reference's floating-point
// operation order; the library is compiled with -fmad=false so no a*b+c is
// contracted (Rust never contracts), which makes device results bit-identical
// to the CPU oracle.  In-place safety comes from the callers: a thread loads
// the complete XOR-orbit of its outputs into registers before storing any.
//
// Reference formulas: /root/reference/src/operator/atomic/<kind>.rs::atomic_op,
// rotate(): src/math/mod.rs:41-50, complex multiply: num_complex 0.4.2
//   (a*b).re = a.re*b.re - a.im*b.im ; (a*b).im = a.re*b.im + a.im*b.re
#pragma once
#include "common.cuh"

namespace qv {

// All four products are formed before the two sums: the values are those of
// num_complex's Mul, and an in-place caller (v = c_mul(ph, v)) can then retire both
// inputs before the first output is written -- no register copies (see tile.cu).
__device__ __forceinline__ amp c_mul(amp a, amp b) {
    const double t0 = a.x * b.x, t1 = a.y * b.y, t2 = a.x * b.y, t3 = a.y * b.x;
    return make_double2(t0 - t1, t2 + t3);
}
__device__ __forceinline__ amp c_add(amp a, amp b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ amp c_neg(amp a) { return make_double2(-a.x, -a.y); }

// multiply by i^(q mod 4) with sign flips / swaps only (math/mod.rs:41-50)
__device__ __forceinline__ amp rotate_i(amp z, uint64_t q) {
    if (q & 2) z = c_neg(z);
    if (q & 1) z = make_double2(-z.y, z.x);
    return z;
}

__device__ __forceinline__ unsigned par(uint64_t v) { return __popcll(v) & 1u; }

// Index-independent gate parameters.  Every formula below takes the part of the
// amplitude index it depends on as a small selector (`cnt` = popcount(i & a) for the
// diagonal class, `sel` = popcount(i & a) & 1 for the pair class), so the direct sweeps
// (which derive it from the 64-bit index) and the fused tile pass (where it is a
// compile-time property of the register slot) share ONE copy of the arithmetic.
struct GP {
    double c, s;        // (ph_re, ph_im) = (cos t/2, sin t/2) as stored by the rotation ops
    uint32_t dagger;
    uint32_t ybase;     // y.rs:11: !(count_ones(mask) + 1) as u32
    const amp *mat;     // u1: 4, u2: 16 entries, row-major
};

__device__ __forceinline__ GP gp_of(const DevOp &op, const amp *mat) {
    GP g;
    g.c = op.ph_re;
    g.s = op.ph_im;
    g.dagger = op.dagger;
    g.ybase = ~(uint32_t)(__popcll(op.a) + 1);
    g.mat = mat;
    return g;
}

// ---- diagonal class: out[i] = f(i) * in[i] --------------------------------
// z.rs:15-21, s.rs:19-25, t.rs:24-35, rz.rs:19-25, rzz.rs:19-25
// cnt = popcount(i & a); only cnt mod 8 matters (the reference negates it in u64 for
// the daggers, and 2^64 = 0 mod 8).
template <int KIND>
__device__ __forceinline__ amp diag_out_s(const GP &g, amp p, uint32_t cnt) {
    if (KIND == QVNT_Z) {
        return (cnt & 1u) ? c_neg(p) : p;
    } else if (KIND == QVNT_S) {
        uint32_t count = cnt;
        if (g.dagger) count = 0u - count;
        return rotate_i(p, count);
    } else if (KIND == QVNT_T) {
        uint32_t count = cnt;
        if (g.dagger) count = 0u - count;
        amp r = rotate_i(p, count >> 1);
        if (count & 1u) r = c_mul(make_double2(QV_FRAC_1_SQRT_2, QV_FRAC_1_SQRT_2), r);
        return r;
    } else {  // QVNT_RZ (1-bit mask: cnt in {0,1}) / QVNT_RZZ (parity)
        amp ph = make_double2(g.c, g.s);
        if ((cnt & 1u) == 0) ph.y = -ph.y;
        return c_mul(ph, p);
    }
}

template <int KIND>
__device__ __forceinline__ amp diag_out(const DevOp &op, amp p, uint64_t i) {
    const GP g = gp_of(op, nullptr);
    return diag_out_s<KIND>(g, p, (uint32_t)__popcll(i & op.a));
}

// ---- pair class: out[i] = f(in[i], in[i ^ m], i) ---------------------------
// p0 = in[i], p1 = in[i ^ m], sel = popcount(i & a) & 1 (for 1-bit masks: "bit set").
// For the swap family the caller only invokes this on odd-parity indices (the
// even-parity branch is the identity).
//
// Every formula is split into its first level of operations (pair_mid: the products of a
// rotation, the sums of a Hadamard) and the second level (pair_fin).  The values are
// exactly the reference expressions'; the split only fixes the instruction ORDER so that an
// in-place update of a register-resident pair reads both members completely before it
// writes either (pair_update_s), which lets the register allocator reuse the input
// registers for the outputs instead of copying.
struct PairMid {
    double t[8];
};

template <int KIND>
__device__ __forceinline__ PairMid pair_mid(const GP &g, amp p0, amp p1, unsigned sel) {
    PairMid m;
    if (KIND == QVNT_X || KIND == QVNT_SWAP) {              // x.rs:15-17, swap.rs:16-22 (odd parity)
        m.t[0] = p1.x;
        m.t[1] = p1.y;
    } else if (KIND == QVNT_Y) {                            // y.rs:10-23
        uint32_t i_pow = g.ybase;
        if (sel == 0) i_pow ^= 2u;
        const amp r = rotate_i(p1, i_pow);
        m.t[0] = r.x;
        m.t[1] = r.y;
    } else if (KIND == QVNT_RX || KIND == QVNT_RXX) {       // rx.rs:18-24, rxx.rs:19-25
        m.t[0] = p0.x * g.c;
        m.t[1] = p1.y * g.s;
        m.t[2] = p0.y * g.c;
        m.t[3] = p1.x * g.s;
    } else if (KIND == QVNT_RY) {                           // ry.rs:19-29
        double s = g.s;
        if (sel == 0) s = -s;
        m.t[0] = p0.x * g.c;
        m.t[1] = p1.x * s;
        m.t[2] = p0.y * g.c;
        m.t[3] = p1.y * s;
    } else if (KIND == QVNT_RYY) {                          // ryy.rs:19-29
        double s = g.s;
        if (sel == 0) s = -s;
        m.t[0] = p0.x * g.c;
        m.t[1] = p1.y * s;
        m.t[2] = p0.y * g.c;
        m.t[3] = p1.x * s;
    } else if (KIND == QVNT_H1) {                           // h1.rs:16-22
        if (sel) p0 = c_neg(p0);
        m.t[0] = p0.x + p1.x;
        m.t[1] = p0.y + p1.y;
        m.t[2] = g.c;               // 1/sqrt(2) (planner.cu sets it; the halves of a split h2 carry 1 and 0.5)
    } else if (KIND == QVNT_U1) {                           // u1.rs:17-25
        // sel == 0: M[0]*p0 + M[1]*p1 ; sel == 1: M[2]*p1 + M[3]*p0   (p0 = self, p1 = partner)
        const amp A = sel == 0 ? g.mat[0] : g.mat[2], B = sel == 0 ? g.mat[1] : g.mat[3];
        const amp x = sel == 0 ? p0 : p1, y = sel == 0 ? p1 : p0;
        m.t[0] = A.x * x.x;
        m.t[1] = A.y * x.y;
        m.t[2] = A.x * x.y;
        m.t[3] = A.y * x.x;
        m.t[4] = B.x * y.x;
        m.t[5] = B.y * y.y;
        m.t[6] = B.x * y.y;
        m.t[7] = B.y * y.x;
    } else if (KIND == QVNT_ISWAP) {                        // i_swap.rs:20-37
        m.t[0] = g.dagger ? p1.y : -p1.y;
        m.t[1] = g.dagger ? -p1.x : p1.x;
    } else if (KIND == QVNT_SQRT_SWAP) {                    // sqrt_swap.rs:20-37
        if (g.dagger) {
            m.t[0] = p0.x + p0.y + p1.x - p1.y;
            m.t[1] = p0.y - p0.x + p1.y + p1.x;
        } else {
            m.t[0] = p0.x - p0.y + p1.x + p1.y;
            m.t[1] = p0.y + p0.x + p1.y - p1.x;
        }
    } else {                                                // QVNT_SQRT_ISWAP sqrt_i_swap.rs:20-37
        if (g.dagger) {
            m.t[0] = p0.x + p1.y;
            m.t[1] = p0.y - p1.x;
        } else {
            m.t[0] = p0.x - p1.y;
            m.t[1] = p0.y + p1.x;
        }
    }
    return m;
}

template <int KIND>
__device__ __forceinline__ amp pair_fin(const PairMid &m) {
    if (KIND == QVNT_RX || KIND == QVNT_RXX || KIND == QVNT_RYY) {
        return make_double2(m.t[0] + m.t[1], m.t[2] - m.t[3]);
    } else if (KIND == QVNT_RY) {
        return make_double2(m.t[0] + m.t[1], m.t[2] + m.t[3]);
    } else if (KIND == QVNT_H1) {
        return make_double2(m.t[0] * m.t[2], m.t[1] * m.t[2]);
    } else if (KIND == QVNT_U1) {
        return make_double2((m.t[0] - m.t[1]) + (m.t[4] - m.t[5]), (m.t[2] + m.t[3]) + (m.t[6] + m.t[7]));
    } else if (KIND == QVNT_SQRT_SWAP) {
        return make_double2(0.5 * m.t[0], 0.5 * m.t[1]);
    } else if (KIND == QVNT_SQRT_ISWAP) {
        return make_double2(QV_FRAC_1_SQRT_2 * m.t[0], QV_FRAC_1_SQRT_2 * m.t[1]);
    } else {                                                // x, y, swap, i_swap: no arithmetic left
        return make_double2(m.t[0], m.t[1]);
    }
}

template <int KIND>
__device__ __forceinline__ amp pair_out_s(const GP &g, amp p0, amp p1, unsigned sel) {
    return pair_fin<KIND>(pair_mid<KIND>(g, p0, p1, sel));
}

// A register copy the optimiser cannot turn into a renaming.  Pure permutations (x, swap,
// and the unsigned halves of y / i_swap) would otherwise become "v0 := old v1, v1 := old v0"
// at a control-flow merge, which makes old and new values of the same slot live at once and
// costs a copy of ALL register-resident amplitudes on every path through the interpreter.
__device__ __forceinline__ double opaque_copy(double x) {
    double r;
    asm volatile("mov.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// In-place update of one XOR pair held in registers; sel0 / sel1 are the selectors of the
// two members (for 1-bit masks 0 and 1; for 2-bit "both flipped" pairs they are equal).
template <int KIND>
__device__ __forceinline__ void pair_update_s(const GP &g, amp &v0, amp &v1, unsigned sel0, unsigned sel1) {
    const PairMid m0 = pair_mid<KIND>(g, v0, v1, sel0);
    const PairMid m1 = pair_mid<KIND>(g, v1, v0, sel1);
    if (KIND == QVNT_X || KIND == QVNT_SWAP || KIND == QVNT_Y || KIND == QVNT_ISWAP) {
        // no arithmetic after pair_mid: park member 0's result, then write both members
        const double t0 = opaque_copy(m0.t[0]), t1 = opaque_copy(m0.t[1]);
        v1.x = opaque_copy(m1.t[0]);
        v1.y = opaque_copy(m1.t[1]);
        v0.x = opaque_copy(t0);
        v0.y = opaque_copy(t1);
    } else {
        v0 = pair_fin<KIND>(m0);
        v1 = pair_fin<KIND>(m1);
    }
}

template <int KIND>
__device__ __forceinline__ void pair_update(const DevOp &op, const amp *__restrict__ mat, amp &v0, amp &v1,
                                            uint64_t i0, uint64_t i1) {
    const GP g = gp_of(op, mat);
    pair_update_s<KIND>(g, v0, v1, par(i0 & op.a), par(i1 & op.a));
}

// ---- quad class: h2 / u2 ---------------------------------------------------
// q[k] = in[base | (k&1 ? a : 0) | (k&2 ? b : 0)], base has both bits clear.
// quad_mid computes everything but the last operation of output `self` (same encoding),
// quad_fin the last one; the in-place update runs all four quad_mid before any quad_fin.
struct QuadMid {
    amp u, w;
};

template <int KIND>
__device__ __forceinline__ QuadMid quad_mid(const amp *__restrict__ mat, const amp (&q)[4], int self) {
    QuadMid m;
    if (KIND == QVNT_H2) {                                  // h2.rs:22-38
        amp p0 = q[self], p1 = q[self ^ 1], p2 = q[self ^ 2], p3 = q[self ^ 3];
        if (self & 1) { p0 = c_neg(p0); p2 = c_neg(p2); }
        if (self & 2) { p0 = c_neg(p0); p1 = c_neg(p1); }
        m.u = c_add(c_add(c_add(p0, p1), p2), p3);
        m.w = make_double2(0.0, 0.0);
    } else {                                                // u2.rs:22-49, row = 2*b_bit + a_bit
        // `self` already is 2*b_bit + a_bit; columns are (base, |a, |b, |a|b) = q[0..3]
        const amp *row = mat + 4 * self;
        amp r = c_mul(row[0], q[0]);
        r = c_add(r, c_mul(row[1], q[1]));
        r = c_add(r, c_mul(row[2], q[2]));
        m.u = r;
        m.w = c_mul(row[3], q[3]);
    }
    return m;
}

template <int KIND>
__device__ __forceinline__ amp quad_fin(const QuadMid &m) {
    if (KIND == QVNT_H2) return make_double2(m.u.x * 0.5, m.u.y * 0.5);
    return c_add(m.u, m.w);
}

template <int KIND>
__device__ __forceinline__ amp quad_out(const amp *__restrict__ mat, const amp (&q)[4], int self) {
    return quad_fin<KIND>(quad_mid<KIND>(mat, q, self));
}

template <int KIND>
__device__ __forceinline__ void quad_update(const amp *__restrict__ mat, amp (&q)[4]) {
    const QuadMid m0 = quad_mid<KIND>(mat, q, 0);
    const QuadMid m1 = quad_mid<KIND>(mat, q, 1);
    const QuadMid m2 = quad_mid<KIND>(mat, q, 2);
    const QuadMid m3 = quad_mid<KIND>(mat, q, 3);
    q[0] = quad_fin<KIND>(m0);
    q[1] = quad_fin<KIND>(m1);
    q[2] = quad_fin<KIND>(m2);
    q[3] = quad_fin<KIND>(m3);
}

}  // namespace qv
