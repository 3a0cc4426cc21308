//! src/operator/lower.rs (new, feature "b200"): host ops -> POD descriptors for the C ABI.
//! Every atomic `Op` already stores exactly the fields of `qvnt_op_t`, so lowering is a copy.
use qvnt_b200_sys::qvnt_op_t;

use crate::operator::{atomic::*, multi::MultiOp, single::SingleOp};

/// Added to `trait AtomicOp` (operator/atomic/dispatch.rs:28-80) and dispatched by
/// `enum_dispatch` like its other methods:  `fn lower(&self) -> qvnt_op_t;`
/// Per-atomic bodies (kind = variant index in `AtomicOpDispatch`, dispatch.rs:84-105):
pub(crate) mod per_atomic {
    use super::*;
    const Z: qvnt_op_t = qvnt_op_t::ZERO;
    pub fn id(_: &id::Op) -> qvnt_op_t { qvnt_op_t { kind: 0, ..Z } }
    pub fn x(o: &x::Op) -> qvnt_op_t { qvnt_op_t { kind: 1, a_mask: o.a_mask as u64, ..Z } }
    pub fn rx(o: &rx::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 2, a_mask: o.a_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn rxx(o: &rxx::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 3, a_mask: o.ab_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn y(o: &y::Op) -> qvnt_op_t { qvnt_op_t { kind: 4, a_mask: o.a_mask as u64, ..Z } }
    pub fn ry(o: &ry::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 5, a_mask: o.a_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn ryy(o: &ryy::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 6, a_mask: o.ab_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn z(o: &z::Op) -> qvnt_op_t { qvnt_op_t { kind: 7, a_mask: o.a_mask as u64, ..Z } }
    pub fn s(o: &s::Op) -> qvnt_op_t { qvnt_op_t { kind: 8, a_mask: o.a_mask as u64, dagger: o.dagger as u32, ..Z } }
    pub fn t(o: &t::Op) -> qvnt_op_t { qvnt_op_t { kind: 9, a_mask: o.a_mask as u64, dagger: o.dagger as u32, ..Z } }
    pub fn rz(o: &rz::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 10, a_mask: o.a_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn rzz(o: &rzz::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 11, a_mask: o.ab_mask as u64, phase_re: o.phase.re, phase_im: o.phase.im, ..Z }
    }
    pub fn u1(o: &u1::Op) -> qvnt_op_t {
        let mut d = qvnt_op_t { kind: 12, a_mask: o.a_mask as u64, ..Z };
        for (i, c) in o.matrix.iter().enumerate() { d.matrix[2 * i] = c.re; d.matrix[2 * i + 1] = c.im; }
        d
    }
    pub fn u2(o: &u2::Op) -> qvnt_op_t {
        let mut d = qvnt_op_t { kind: 13, a_mask: o.a_mask as u64, b_mask: o.b_mask as u64, ..Z };
        for (i, c) in o.matrix.iter().enumerate() { d.matrix[2 * i] = c.re; d.matrix[2 * i + 1] = c.im; }
        d
    }
    pub fn h1(o: &h1::Op) -> qvnt_op_t { qvnt_op_t { kind: 14, a_mask: o.a_mask as u64, ..Z } }
    pub fn h2(o: &h2::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 15, a_mask: o.a_mask as u64, b_mask: o.b_mask as u64, ..Z }
    }
    pub fn swap(o: &swap::Op) -> qvnt_op_t { qvnt_op_t { kind: 16, a_mask: o.ab_mask as u64, ..Z } }
    pub fn i_swap(o: &i_swap::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 17, a_mask: o.ab_mask as u64, dagger: o.dagger as u32, ..Z }
    }
    pub fn sqrt_swap(o: &sqrt_swap::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 18, a_mask: o.ab_mask as u64, dagger: o.dagger as u32, ..Z }
    }
    pub fn sqrt_i_swap(o: &sqrt_i_swap::Op) -> qvnt_op_t {
        qvnt_op_t { kind: 19, a_mask: o.ab_mask as u64, dagger: o.dagger as u32, ..Z }
    }
}

/// Crate-private companion of `Applicable`: what `QReg::apply` sends across the boundary.
pub(crate) trait Lower {
    fn lower(&self, out: &mut Vec<qvnt_op_t>);
}

impl Lower for SingleOp {
    fn lower(&self, out: &mut Vec<qvnt_op_t>) {
        let mut d = self.func.lower();      // AtomicOpDispatch -> per_atomic::* via enum_dispatch
        d.ctrl = self.ctrl as u64;          // single/mod.rs:43-47
        out.push(d);
    }
}

impl Lower for MultiOp {
    fn lower(&self, out: &mut Vec<qvnt_op_t>) {
        out.reserve(self.len());
        for op in self.iter() {             // queue order, front first (multi/mod.rs:96-114)
            op.lower(out);
        }
    }
}
