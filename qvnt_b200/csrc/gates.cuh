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

__device__ __forceinline__ amp c_mul(amp a, amp b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
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

// ---- diagonal class: out[i] = f(i) * in[i] --------------------------------
// z.rs:15-21, s.rs:19-25, t.rs:24-35, rz.rs:19-25, rzz.rs:19-25
template <int KIND>
__device__ __forceinline__ amp diag_out(const DevOp &op, amp p, uint64_t i) {
    if (KIND == QVNT_Z) {
        return par(i & op.a) ? c_neg(p) : p;
    } else if (KIND == QVNT_S) {
        uint64_t count = (uint64_t)__popcll(i & op.a);
        if (op.dagger) count = ~count + 1ull;
        return rotate_i(p, count);
    } else if (KIND == QVNT_T) {
        uint64_t count = (uint64_t)__popcll(i & op.a);
        if (op.dagger) count = ~count + 1ull;
        amp r = rotate_i(p, count >> 1);
        if (count & 1ull) r = c_mul(make_double2(QV_FRAC_1_SQRT_2, QV_FRAC_1_SQRT_2), r);
        return r;
    } else if (KIND == QVNT_RZ) {
        amp ph = make_double2(op.ph_re, op.ph_im);
        if ((i & op.a) == 0) ph.y = -ph.y;
        return c_mul(ph, p);
    } else {  // QVNT_RZZ
        amp ph = make_double2(op.ph_re, op.ph_im);
        if (par(i & op.a) == 0) ph.y = -ph.y;
        return c_mul(ph, p);
    }
}

__device__ __forceinline__ amp diag_out_dyn(const DevOp &op, amp p, uint64_t i) {
    switch (op.kind) {
    case QVNT_Z: return diag_out<QVNT_Z>(op, p, i);
    case QVNT_S: return diag_out<QVNT_S>(op, p, i);
    case QVNT_T: return diag_out<QVNT_T>(op, p, i);
    case QVNT_RZ: return diag_out<QVNT_RZ>(op, p, i);
    default: return diag_out<QVNT_RZZ>(op, p, i);
    }
}

// ---- pair class: out[i] = f(in[i], in[i ^ m], i) ---------------------------
// p0 = in[i], p1 = in[i ^ m].  For the swap family the caller only invokes this
// on odd-parity indices (the even-parity branch is the identity).
// `mat` points at the op's 2x2 matrix (u1 only).
template <int KIND>
__device__ __forceinline__ amp pair_out(const DevOp &op, const amp *__restrict__ mat, amp p0, amp p1,
                                        uint64_t i) {
    if (KIND == QVNT_X) {                                   // x.rs:15-17
        return p1;
    } else if (KIND == QVNT_Y) {                            // y.rs:10-23
        uint64_t i_pow = (uint64_t)(uint32_t)~(uint32_t)(__popcll(op.a) + 1);
        if (par(i & op.a) == 0) i_pow ^= 2ull;
        return rotate_i(p1, i_pow);
    } else if (KIND == QVNT_RX || KIND == QVNT_RXX) {       // rx.rs:18-24, rxx.rs:19-25
        return make_double2(p0.x * op.ph_re + p1.y * op.ph_im, p0.y * op.ph_re - p1.x * op.ph_im);
    } else if (KIND == QVNT_RY) {                           // ry.rs:19-29
        double s = op.ph_im;
        if ((i & op.a) == 0) s = -s;
        return make_double2(p0.x * op.ph_re + p1.x * s, p0.y * op.ph_re + p1.y * s);
    } else if (KIND == QVNT_RYY) {                          // ryy.rs:19-29
        double s = op.ph_im;
        if (par(i & op.a) == 0) s = -s;
        return make_double2(p0.x * op.ph_re + p1.y * s, p0.y * op.ph_re - p1.x * s);
    } else if (KIND == QVNT_H1) {                           // h1.rs:16-22
        if (i & op.a) p0 = c_neg(p0);
        return make_double2((p0.x + p1.x) * QV_FRAC_1_SQRT_2, (p0.y + p1.y) * QV_FRAC_1_SQRT_2);
    } else if (KIND == QVNT_U1) {                           // u1.rs:17-25
        if ((i & op.a) == 0) return c_add(c_mul(mat[0], p0), c_mul(mat[1], p1));
        return c_add(c_mul(mat[2], p1), c_mul(mat[3], p0));
    } else if (KIND == QVNT_SWAP) {                         // swap.rs:16-22 (odd parity)
        return p1;
    } else if (KIND == QVNT_ISWAP) {                        // i_swap.rs:20-37
        return op.dagger ? make_double2(p1.y, -p1.x) : make_double2(-p1.y, p1.x);
    } else if (KIND == QVNT_SQRT_SWAP) {                    // sqrt_swap.rs:20-37
        if (op.dagger)
            return make_double2(0.5 * (p0.x + p0.y + p1.x - p1.y), 0.5 * (p0.y - p0.x + p1.y + p1.x));
        return make_double2(0.5 * (p0.x - p0.y + p1.x + p1.y), 0.5 * (p0.y + p0.x + p1.y - p1.x));
    } else {                                                // QVNT_SQRT_ISWAP sqrt_i_swap.rs:20-37
        if (op.dagger)
            return make_double2(QV_FRAC_1_SQRT_2 * (p0.x + p1.y), QV_FRAC_1_SQRT_2 * (p0.y - p1.x));
        return make_double2(QV_FRAC_1_SQRT_2 * (p0.x - p1.y), QV_FRAC_1_SQRT_2 * (p0.y + p1.x));
    }
}

// In-place update of one XOR pair {i0, i1 = i0 ^ m} held in registers.
template <int KIND>
__device__ __forceinline__ void pair_update(const DevOp &op, const amp *__restrict__ mat, amp &v0, amp &v1,
                                            uint64_t i0, uint64_t i1) {
    amp o0 = pair_out<KIND>(op, mat, v0, v1, i0);
    amp o1 = pair_out<KIND>(op, mat, v1, v0, i1);
    v0 = o0;
    v1 = o1;
}

__device__ __forceinline__ void pair_update_dyn(const DevOp &op, const amp *__restrict__ mat, amp &v0,
                                                amp &v1, uint64_t i0, uint64_t i1) {
    switch (op.kind) {
    case QVNT_X: pair_update<QVNT_X>(op, mat, v0, v1, i0, i1); break;
    case QVNT_Y: pair_update<QVNT_Y>(op, mat, v0, v1, i0, i1); break;
    case QVNT_RX: case QVNT_RXX: pair_update<QVNT_RX>(op, mat, v0, v1, i0, i1); break;
    case QVNT_RY: pair_update<QVNT_RY>(op, mat, v0, v1, i0, i1); break;
    case QVNT_RYY: pair_update<QVNT_RYY>(op, mat, v0, v1, i0, i1); break;
    case QVNT_H1: pair_update<QVNT_H1>(op, mat, v0, v1, i0, i1); break;
    case QVNT_U1: pair_update<QVNT_U1>(op, mat, v0, v1, i0, i1); break;
    case QVNT_SWAP: pair_update<QVNT_SWAP>(op, mat, v0, v1, i0, i1); break;
    case QVNT_ISWAP: pair_update<QVNT_ISWAP>(op, mat, v0, v1, i0, i1); break;
    case QVNT_SQRT_SWAP: pair_update<QVNT_SQRT_SWAP>(op, mat, v0, v1, i0, i1); break;
    default: pair_update<QVNT_SQRT_ISWAP>(op, mat, v0, v1, i0, i1); break;
    }
}

// ---- quad class: h2 / u2 ---------------------------------------------------
// q[k] = in[base | (k&1 ? a : 0) | (k&2 ? b : 0)], base has both bits clear.
// Returns the output at position `self` (same encoding).
template <int KIND>
__device__ __forceinline__ amp quad_out(const DevOp &op, const amp *__restrict__ mat, const amp q[4],
                                        int self) {
    if (KIND == QVNT_H2) {                                  // h2.rs:22-38
        amp p0 = q[self], p1 = q[self ^ 1], p2 = q[self ^ 2], p3 = q[self ^ 3];
        if (self & 1) { p0 = c_neg(p0); p2 = c_neg(p2); }
        if (self & 2) { p0 = c_neg(p0); p1 = c_neg(p1); }
        amp s = c_add(c_add(c_add(p0, p1), p2), p3);
        return make_double2(s.x * 0.5, s.y * 0.5);
    } else {                                                // u2.rs:22-49, row = 2*b_bit + a_bit
        // `self` already is 2*b_bit + a_bit; columns are (base, |a, |b, |a|b) = q[0..3]
        const amp *row = mat + 4 * self;
        amp r = c_mul(row[0], q[0]);
        r = c_add(r, c_mul(row[1], q[1]));
        r = c_add(r, c_mul(row[2], q[2]));
        r = c_add(r, c_mul(row[3], q[3]));
        return r;
    }
}

template <int KIND>
__device__ __forceinline__ void quad_update(const DevOp &op, const amp *__restrict__ mat, amp q[4]) {
    amp o0 = quad_out<KIND>(op, mat, q, 0);
    amp o1 = quad_out<KIND>(op, mat, q, 1);
    amp o2 = quad_out<KIND>(op, mat, q, 2);
    amp o3 = quad_out<KIND>(op, mat, q, 3);
    q[0] = o0; q[1] = o1; q[2] = o2; q[3] = o3;
}

}  // namespace qv
