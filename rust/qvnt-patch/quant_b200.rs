//! src/register/quant.rs under `cfg(feature = "b200")`: the state vector lives in B200 HBM behind
//! an opaque handle; every public method keeps the reference's name and signature
//! (register/quant.rs:113-636) and forwards to one entry point of include/qvnt_b200.h.
use std::{ffi::CStr, fmt, ops::Mul, ptr::null_mut};

use qvnt_b200_sys as sys;
use rand::Rng;

use crate::{
    math::types::*,
    operator::{applicable::Applicable, lower::Lower},
    register::{CReg, VReg},
};

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::qvnt_last_error()) }.to_string_lossy().into_owned()
}
fn check(rc: i32) {
    // the reference's `apply` is infallible and panics on out-of-range masks (index OOB);
    // nothing ever unwinds across the C boundary itself
    if rc != sys::QVNT_OK {
        panic!("qvnt-b200: {}", last_error());
    }
}

pub struct Reg {
    dev: *mut sys::qvnt_reg_t,
    q_num: N,
    q_mask: N,
}
unsafe impl Send for Reg {}     // one host thread at a time, like `&mut QReg`

impl Reg {
    pub fn new(q_num: N) -> Self { Self::with_state(q_num, 0) }                       // :113
    pub fn with_state(q_num: N, state: N) -> Self {                                   // :129
        let mut dev = null_mut();
        check(unsafe { sys::qvnt_reg_create(q_num as u32, state as u64, &mut dev) });
        Self { dev, q_num, q_mask: (1usize << q_num).wrapping_sub(1) }
    }
    /// :186-200 with GPUs for threads: the register continues on `num_threads` GPUs (1, 2, 4 or 8 of
    /// this box, sharded by its top qubits, state kept); `None` for 0 or more than the box has
    pub fn num_threads(mut self, num_threads: usize) -> Option<Self> {
        let mut ndev = 0i32;
        unsafe { sys::qvnt_device_count(&mut ndev) };
        let n = num_threads;
        if n == 0 || n & (n - 1) != 0 || n > 8 || n > ndev as usize { return None; }
        let mut dev = null_mut();
        check(unsafe { sys::qvnt_reg_set_gpus(self.dev, n as u32, &mut dev) });
        unsafe { sys::qvnt_reg_destroy(self.dev) };
        self.dev = dev;          // (Drop destroys the new handle)
        Some(self)
    }
    pub fn num(&self) -> N { self.q_num }
    pub fn reset(&mut self, i_state: N) { check(unsafe { sys::qvnt_reg_reset(self.dev, i_state as u64) }) }
    pub fn reset_by_mask(&mut self, mask: N) { check(unsafe { sys::qvnt_reg_reset_by_mask(self.dev, mask as u64) }) }
    pub fn get_vreg(&self) -> VReg { VReg::new_with_mask(self.q_mask) }
    pub fn get_vreg_by(&self, mask: N) -> Option<VReg> {
        if mask & !self.q_mask != 0 { None } else { Some(VReg::new_with_mask(mask)) }
    }

    /// :376-395 -- the hot path: the whole op list crosses in one call so the scheduler can fuse
    pub fn apply<Op: Applicable + Lower>(&mut self, op: &Op) {
        let mut ops = Vec::new();
        op.lower(&mut ops);
        if !ops.is_empty() {
            check(unsafe { sys::qvnt_reg_apply(self.dev, ops.as_ptr(), ops.len()) });
        }
    }
    pub fn normalize(&mut self) -> &mut Self { check(unsafe { sys::qvnt_reg_normalize(self.dev) }); self }   // :397
    pub fn get_polar(&self) -> Vec<(R, R)> {                                          // :417
        let len = 1usize << self.q_num;
        let mut out = vec![(0.0, 0.0); len];
        check(unsafe { sys::qvnt_reg_polar(self.dev, 0, len as u64, out.as_mut_ptr() as *mut f64) });
        out
    }
    pub fn get_probabilities(&self) -> Vec<R> {                                       // :434
        let len = 1usize << self.q_num;
        let mut out = vec![0.0; len];
        check(unsafe { sys::qvnt_reg_probabilities(self.dev, 0, len as u64, out.as_mut_ptr()) });
        out
    }
    pub fn get_absolute(&self) -> R {                                                 // :458
        let mut v = 0.0;
        check(unsafe { sys::qvnt_reg_norm_sqr(self.dev, &mut v) });
        v
    }
    /// :490-501 -- the uniform variate comes from the same source the reference's
    /// `WeightedIndex` draws from (`thread_rng`), the cumulative search runs on the device;
    /// like the reference, the collapsed state is NOT renormalised
    pub fn measure_mask(&mut self, mask: N) -> CReg {
        let u: f64 = rand::thread_rng().gen();
        let mut out = 0u64;
        check(unsafe { sys::qvnt_reg_measure_mask(self.dev, mask as u64, u, &mut out, null_mut()) });
        CReg::with_state(self.q_num, out as N)
    }
    pub fn measure(&mut self) -> CReg { self.measure_mask(self.q_mask) }              // :505
    pub(crate) fn tensor_prod(self, other: Self) -> Self {                            // :330-371
        let mut dev = null_mut();
        check(unsafe { sys::qvnt_reg_tensor_prod(self.dev, other.dev, &mut dev) });
        Self { dev, q_num: self.q_num + other.q_num, q_mask: (1usize << (self.q_num + other.q_num)) - 1 }
    }
}

impl Clone for Reg {                                                                  // #[derive(Clone)] :102
    fn clone(&self) -> Self {
        let mut dev = null_mut();
        check(unsafe { sys::qvnt_reg_clone(self.dev, &mut dev) });
        Self { dev, q_num: self.q_num, q_mask: self.q_mask }
    }
}
impl Drop for Reg {
    fn drop(&mut self) { unsafe { sys::qvnt_reg_destroy(self.dev) }; }
}
impl Mul for Reg {                                                                    // :625-636
    type Output = Self;
    fn mul(self, other: Self) -> Self { self.tensor_prod(other) }
}
impl fmt::Debug for Reg {                                                             // :603-623
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        const MAX_LEN_TO_DISPLAY: usize = 8;
        let len = (1usize << self.q_num).min(MAX_LEN_TO_DISPLAY);
        let mut psi = vec![C::new(0.0, 0.0); len];
        check(unsafe { sys::qvnt_reg_read(self.dev, 0, len as u64, psi.as_mut_ptr() as *mut f64) });
        let mut map = f.debug_map();
        for (idx, z) in psi.iter().enumerate() { map.entry(&idx, z); }
        if (1usize << self.q_num) > MAX_LEN_TO_DISPLAY { map.finish_non_exhaustive() } else { map.finish() }
    }
}
