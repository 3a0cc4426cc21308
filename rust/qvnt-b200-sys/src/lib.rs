//! Raw FFI bindings to `libqvnt_b200.so` -- the checked-in equivalent of
//! `bindgen include/qvnt_b200.h`.  One `extern "C"` item per prototype of the header; the
//! reference-side safe wrapper lives in `rust/qvnt-patch/`.
//!
//! NOT COMPILED in this repository's image (no Rust toolchain there).  Kept in sync with the
//! header by `tests/test_abi.py::test_rust_bindings_cover_header`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const QVNT_OK: i32 = 0;
pub const QVNT_ERR_INVALID: i32 = 1;
pub const QVNT_ERR_BAD_MASK: i32 = 2;
pub const QVNT_ERR_OOM: i32 = 3;
pub const QVNT_ERR_CUDA: i32 = 4;
pub const QVNT_ERR_COMM: i32 = 5;
pub const QVNT_ERR_UNSUPPORTED: i32 = 6;

/// Gate kinds in the variant order of `AtomicOpDispatch` (operator/atomic/dispatch.rs:82-105).
#[repr(u32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum qvnt_kind {
    Id = 0, X, RX, RXX, Y, RY, RYY, Z, S, T, RZ, RZZ, U1, U2, H1, H2, Swap, ISwap, SqrtSwap, SqrtISwap,
}

/// One `SingleOp {act, ctrl, func}` (operator/single/mod.rs:43-47) lowered to POD (304 bytes).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct qvnt_op_t {
    pub kind: u32,
    pub dagger: u32,
    pub a_mask: u64,
    pub b_mask: u64,
    pub ctrl: u64,
    pub phase_re: f64,
    pub phase_im: f64,
    pub matrix: [f64; 32],
}

impl qvnt_op_t {
    pub const ZERO: qvnt_op_t = qvnt_op_t {
        kind: 0, dagger: 0, a_mask: 0, b_mask: 0, ctrl: 0, phase_re: 0.0, phase_im: 0.0, matrix: [0.0; 32],
    };
}

pub const QVNT_STATS_CLASSES: usize = 5;
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct qvnt_stats_t {
    pub launches: [u64; QVNT_STATS_CLASSES],
    pub ms: [f64; QVNT_STATS_CLASSES],
    pub alg_bytes: [u64; QVNT_STATS_CLASSES],
    pub ops_applied: u64,
    pub passes: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub peer_bytes: u64,
}

/// Opaque device-resident register (`QReg`).
#[repr(C)]
pub struct qvnt_reg_t {
    _private: [u8; 0],
}

pub const QVNT_IPC_BLOB_BYTES: usize = 256;

extern "C" {
    pub fn qvnt_version() -> i32;
    pub fn qvnt_last_error() -> *const c_char;
    pub fn qvnt_device_count(out: *mut i32) -> i32;
    pub fn qvnt_reg_create(q_num: u32, state: u64, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_combine(a: *mut qvnt_reg_t, b: *mut qvnt_reg_t, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_combine_unitary(a: *mut qvnt_reg_t, b: *mut qvnt_reg_t, matrix8: *const f64, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_linear_composition(reg: *mut qvnt_reg_t, other: *mut qvnt_reg_t, c0_re: f64, c0_im: f64, c1_re: f64, c1_im: f64) -> i32;
    pub fn qvnt_reg_sample_all(reg: *mut qvnt_reg_t, count: u64, seed: u64, host_out: *mut u64) -> i32;
    pub fn qvnt_reg_create_multi(q_num: u32, state: u64, n_gpus: u32, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_set_gpus(reg: *mut qvnt_reg_t, n_gpus: u32, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_create_sharded(q_num: u32, state: u64, rank: u32, world: u32, device: i32, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_export_ipc(reg: *mut qvnt_reg_t, blob: *mut c_void) -> i32;
    pub fn qvnt_reg_attach_peers(reg: *mut qvnt_reg_t, blobs: *const c_void) -> i32;
    pub fn qvnt_reg_clone(reg: *mut qvnt_reg_t, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_destroy(reg: *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_q_num(reg: *const qvnt_reg_t, out: *mut u32) -> i32;
    pub fn qvnt_reg_apply(reg: *mut qvnt_reg_t, ops: *const qvnt_op_t, n_ops: usize) -> i32;
    pub fn qvnt_plan_describe(q_num: u32, rank: u32, world: u32, peers_attached: i32, fuse: i32, tile_bits: i32, chunk_bits: i32, ops: *const qvnt_op_t, n_ops: usize, out: *mut c_char, cap: usize, needed: *mut usize) -> i32;
    pub fn qvnt_reg_norm_sqr(reg: *mut qvnt_reg_t, out: *mut f64) -> i32;
    pub fn qvnt_reg_probabilities(reg: *mut qvnt_reg_t, off: u64, cnt: u64, host_out: *mut f64) -> i32;
    pub fn qvnt_reg_polar(reg: *mut qvnt_reg_t, off: u64, cnt: u64, host_r_theta: *mut f64) -> i32;
    pub fn qvnt_reg_measure_mask(reg: *mut qvnt_reg_t, mask: u64, u01: f64, outcome: *mut u64, sampled: *mut u64) -> i32;
    pub fn qvnt_reg_measure_mask_rng(reg: *mut qvnt_reg_t, mask: u64, outcome: *mut u64) -> i32;
    pub fn qvnt_reg_collapse(reg: *mut qvnt_reg_t, idy: u64, mask: u64) -> i32;
    pub fn qvnt_reg_normalize(reg: *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_reset(reg: *mut qvnt_reg_t, state: u64) -> i32;
    pub fn qvnt_reg_reset_by_mask(reg: *mut qvnt_reg_t, mask: u64) -> i32;
    pub fn qvnt_reg_read(reg: *mut qvnt_reg_t, off: u64, cnt: u64, host_re_im: *mut f64) -> i32;
    pub fn qvnt_reg_write(reg: *mut qvnt_reg_t, off: u64, cnt: u64, host_re_im: *const f64) -> i32;
    pub fn qvnt_reg_tensor_prod(a: *mut qvnt_reg_t, b: *mut qvnt_reg_t, out: *mut *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_sync(reg: *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_set_option(reg: *mut qvnt_reg_t, key: *const c_char, value: i64) -> i32;
    pub fn qvnt_reg_stats(reg: *mut qvnt_reg_t, out: *mut qvnt_stats_t) -> i32;
    pub fn qvnt_reg_stats_reset(reg: *mut qvnt_reg_t) -> i32;
    pub fn qvnt_reg_mark(reg: *mut qvnt_reg_t, slot: i32) -> i32;
    pub fn qvnt_reg_elapsed_ms(reg: *mut qvnt_reg_t, slot_from: i32, slot_to: i32, ms: *mut f64) -> i32;
}
