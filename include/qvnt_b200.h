/*
 * qvnt_b200.h -- C ABI of the B200-native state-vector engine for QVNT's
 * gate-application hot path (libqvnt_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's `QReg` / `Applicable` path (citations relative to
 * /root/reference/src).  Plain pointers and sizes only; no C++ or torch types.
 * A Rust host binds it through the `qvnt-b200-sys` crate (rust/), a C++ host
 * through include/qvnt.hpp, Python through ctypes (qvnt_b200/_ffi.py); see
 * INTEGRATION.md for the reference-side patch.
 *
 * Conventions
 *   - every function returns a qvnt_status (0 = ok) and never unwinds/exits;
 *     qvnt_last_error() returns a thread-local message for the last failure.
 *   - a handle is used by one host thread at a time (Send, not Sync), exactly
 *     like `&mut QReg`.
 *   - calls enqueue on the handle's CUDA stream in order; only the functions
 *     that return data to the host (read/probabilities/norm/measure/sync/stats)
 *     block.
 *   - amplitude = complex f64, interleaved (re, im) = 16 bytes; bit k of the
 *     amplitude index is qubit k (reference: math/mod.rs:22-31, atomic/x.rs:15-17).
 *   - there is NO CPU fallback: without a CUDA device every call that needs one
 *     fails with QVNT_ERR_CUDA.
 */
#ifndef QVNT_B200_H
#define QVNT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QVNT_B200_VERSION 100 /* 0.1.0 */

typedef enum qvnt_status {
    QVNT_OK = 0,
    QVNT_ERR_INVALID = 1,     /* null pointer, bad argument, invalid op descriptor   */
    QVNT_ERR_BAD_MASK = 2,    /* a mask addresses qubits outside the register        */
    QVNT_ERR_OOM = 3,         /* device (or pinned host) allocation failed           */
    QVNT_ERR_CUDA = 4,        /* CUDA runtime error / no device                      */
    QVNT_ERR_COMM = 5,        /* peer-memory (NVLink IPC) setup or barrier failure   */
    QVNT_ERR_UNSUPPORTED = 6  /* valid request this build cannot serve               */
} qvnt_status;

/* Gate kinds, in the variant order of `AtomicOpDispatch`
 * (operator/atomic/dispatch.rs:82-105). */
typedef enum qvnt_kind {
    QVNT_ID = 0, QVNT_X, QVNT_RX, QVNT_RXX, QVNT_Y, QVNT_RY, QVNT_RYY, QVNT_Z, QVNT_S, QVNT_T,
    QVNT_RZ, QVNT_RZZ, QVNT_U1, QVNT_U2, QVNT_H1, QVNT_H2, QVNT_SWAP, QVNT_ISWAP,
    QVNT_SQRT_SWAP, QVNT_SQRT_ISWAP, QVNT_KIND_COUNT
} qvnt_kind;

/* One `SingleOp` = {act, ctrl, func} (operator/single/mod.rs:43-47) lowered to
 * POD.  The host keeps building ops exactly as the reference does (op::*,
 * .c(), .dgr()); only this descriptor crosses the boundary.
 *   a_mask   a_mask / ab_mask of the atomic op
 *   b_mask   second mask of h2/u2 (atomic/h2.rs:5-9, u2.rs:5-9), else 0
 *   ctrl     control mask: gate applies where (~idx & ctrl) == 0 (dispatch.rs:35)
 *   phase    (cos t/2, sin t/2) as stored by the rotation ops (e.g. rx.rs:10-14);
 *            `.dgr()` of a rotation negates BOTH parts on the host (rx.rs:42-47)
 *   dagger   s/t/i_swap/sqrt_swap/sqrt_i_swap flag (s.rs:39-44, ...)
 *   matrix   u1: 4, u2: 16 complex numbers row-major, interleaved re,im
 *            (u1.rs:17-25, u2.rs:22-49); `.dgr()` conj-transposes on the host
 */
typedef struct qvnt_op_t {
    uint32_t kind;
    uint32_t dagger;
    uint64_t a_mask;
    uint64_t b_mask;
    uint64_t ctrl;
    double phase_re, phase_im;
    double matrix[32];
} qvnt_op_t;

typedef struct qvnt_reg qvnt_reg_t; /* opaque: the device-resident `QReg` */

/* Counters of the work a handle has enqueued since the last stats_reset.
 * kernel classes: 0 direct sweep, 1 fused tile pass, 2 reduction/measure,
 * 3 init/collapse/scale, 4 cross-GPU barrier. */
#define QVNT_STATS_CLASSES 5
typedef struct qvnt_stats_t {
    uint64_t launches[QVNT_STATS_CLASSES];  /* kernel launches per class             */
    double   ms[QVNT_STATS_CLASSES];        /* device time per class (profile mode)  */
    uint64_t alg_bytes[QVNT_STATS_CLASSES]; /* algorithmic HBM bytes per class: 16 B read + 16 B
                                               written per amplitude a launch can change
                                               (reductions: 16 B read per amplitude)         */
    uint64_t ops_applied;                   /* SingleOps executed                    */
    uint64_t passes;                        /* HBM sweeps those ops were packed into */
    uint64_t h2d_bytes, d2h_bytes;          /* bytes moved across PCIe by this handle */
    uint64_t peer_bytes;                    /* bytes read+written in peer HBM (NVLink) */
} qvnt_stats_t;

/* ---- library ------------------------------------------------------------ */
int qvnt_version(void);
const char *qvnt_last_error(void);
int qvnt_device_count(int *out);

/* ---- lifetime: QReg::new / with_state / Clone / Drop ---------------------- */
/* QReg::with_state(q_num, state) (register/quant.rs:129-150) on the current
 * CUDA device; QReg::new(q) == with_state(q, 0) (:113-125). */
int qvnt_reg_create(uint32_t q_num, uint64_t state, qvnt_reg_t **out);
/* One shard of a register split by its top log2(world) qubits, one process per
 * GPU (replaces `QReg::num_threads`, quant.rs:186-200, as the parallelism
 * selector).  After creation every rank exports its handle blob, the host
 * exchanges the blobs (any transport), and attaches its peers. */
int qvnt_reg_create_sharded(uint32_t q_num, uint64_t state, uint32_t rank, uint32_t world,
                            int device, qvnt_reg_t **out);
/* The same sharding driven from ONE host process through ONE handle -- the drop-in for
 * `QReg::num_threads(n)` (quant.rs:186-200, threads.rs:42-52): n_gpus in {1, 2, 4, 8} of this
 * box, shard k on device k.  Every entry point of this header accepts the handle; one host thread
 * enqueues every shard's work before it waits for any. */
int qvnt_reg_create_multi(uint32_t q_num, uint64_t state, uint32_t n_gpus, qvnt_reg_t **out);
/* `QReg::num_threads(n)` on an existing register: a NEW handle on n_gpus GPUs holding the same
 * state (the caller destroys the old one, as `num_threads(self)` consumes it). */
int qvnt_reg_set_gpus(qvnt_reg_t *reg, uint32_t n_gpus, qvnt_reg_t **out);
#define QVNT_IPC_BLOB_BYTES 256
int qvnt_reg_export_ipc(qvnt_reg_t *reg, void *blob /* QVNT_IPC_BLOB_BYTES */);
int qvnt_reg_attach_peers(qvnt_reg_t *reg, const void *blobs /* world * QVNT_IPC_BLOB_BYTES */);
int qvnt_reg_clone(qvnt_reg_t *reg, qvnt_reg_t **out);             /* #[derive(Clone)] quant.rs:102 */
int qvnt_reg_destroy(qvnt_reg_t *reg);
int qvnt_reg_q_num(const qvnt_reg_t *reg, uint32_t *out);          /* QReg::num quant.rs:152 */

/* ---- the hot path --------------------------------------------------------- */
/* QReg::apply(&impl Applicable) (quant.rs:376-395) for a whole MultiOp
 * (multi/mod.rs:96-114) or a single SingleOp (n_ops == 1): ops are applied in
 * array order, in place, in HBM.  Result equals the reference's out-of-place
 * sweep per SingleOp (dispatch.rs:32-67). */
int qvnt_reg_apply(qvnt_reg_t *reg, const qvnt_op_t *ops, size_t n_ops);

/* The schedule qvnt_reg_apply would run for this op list on rank `rank` of a q_num-qubit
 * register sharded over `world` GPUs, as text (one line per pass / stage / op).  Host-only:
 * needs no CUDA device.  tile_bits / chunk_bits 0 = defaults.  *needed = bytes incl. NUL.
 * peers_attached: 0 / 1, 3 = attached with option "remap" off; fuse: 0 / 1, 3 = with option
 * "lower_two_bit" on. */
int qvnt_plan_describe(uint32_t q_num, uint32_t rank, uint32_t world, int peers_attached, int fuse,
                       int tile_bits, int chunk_bits, const qvnt_op_t *ops, size_t n_ops, char *out,
                       size_t cap, size_t *needed);

/* ---- measurement / normalisation ------------------------------------------ */
int qvnt_reg_norm_sqr(qvnt_reg_t *reg, double *out);               /* get_absolute  quant.rs:458-466 */
/* get_probabilities (quant.rs:434-454) for indices [off, off+cnt) of the full
 * register; on a sharded register the range must lie in this rank's shard. */
int qvnt_reg_probabilities(qvnt_reg_t *reg, uint64_t off, uint64_t cnt, double *host_out);
/* get_polar (quant.rs:417-431): (r, theta) pairs for [off, off+cnt). */
int qvnt_reg_polar(qvnt_reg_t *reg, uint64_t off, uint64_t cnt, double *host_r_theta);
/* measure_mask (quant.rs:490-501) with the uniform variate injected:
 * samples idx = first i with cumsum_i(p) > u01 * total (rand 0.8.5
 * WeightedIndex), collapses (no renormalisation, quant.rs:468-486), returns
 * idx & mask in *outcome and idx in *sampled (may be NULL). */
int qvnt_reg_measure_mask(qvnt_reg_t *reg, uint64_t mask, double u01, uint64_t *outcome,
                          uint64_t *sampled);
/* Same, drawing u01 from an internal splitmix64 stream (thread_rng stand-in). */
int qvnt_reg_measure_mask_rng(qvnt_reg_t *reg, uint64_t mask, uint64_t *outcome);
int qvnt_reg_collapse(qvnt_reg_t *reg, uint64_t idy, uint64_t mask); /* collapse_mask quant.rs:468-486 */
int qvnt_reg_normalize(qvnt_reg_t *reg);                           /* normalize quant.rs:397-414 */
int qvnt_reg_reset(qvnt_reg_t *reg, uint64_t state);               /* reset quant.rs:202-205 */
int qvnt_reg_reset_by_mask(qvnt_reg_t *reg, uint64_t mask);        /* reset_by_mask quant.rs:207-229 */

/* ---- data movement (Debug fmt, tests, tensor product) --------------------- */
int qvnt_reg_read(qvnt_reg_t *reg, uint64_t off, uint64_t cnt, double *host_re_im);
int qvnt_reg_write(qvnt_reg_t *reg, uint64_t off, uint64_t cnt, const double *host_re_im);
/* tensor_prod / Mul (quant.rs:330-371,625-636): out = a (low qubits) x b. */
int qvnt_reg_tensor_prod(qvnt_reg_t *a, qvnt_reg_t *b, qvnt_reg_t **out);
/* combine / combine_with_unitary / linear_composition (quant.rs:245-328; crate-private and untested
 * in the reference): out = (a | b) as the lower / upper half of a register one qubit larger,
 * optionally mixed by a 2x2 matrix (row-major, interleaved re,im); reg = reg * c0 + other * c1. */
int qvnt_reg_combine(qvnt_reg_t *a, qvnt_reg_t *b, qvnt_reg_t **out);
int qvnt_reg_combine_unitary(qvnt_reg_t *a, qvnt_reg_t *b, const double *matrix8, qvnt_reg_t **out);
int qvnt_reg_linear_composition(qvnt_reg_t *reg, qvnt_reg_t *other, double c0_re, double c0_im, double c1_re,
                                double c1_im);
/* sample_all (quant.rs:513-594): histogram of `count` shots (Gaussian approximation, no collapse)
 * for all 2^q_num indices into host_out; `seed` keys the device's counter-based normal generator
 * (the reference draws from thread_rng: statistical parity). */
int qvnt_reg_sample_all(qvnt_reg_t *reg, uint64_t count, uint64_t seed, uint64_t *host_out);
int qvnt_reg_sync(qvnt_reg_t *reg);

/* ---- tuning / instrumentation --------------------------------------------- */
/* keys: "fuse" (0/1); "tile_bits" (0 = auto, 6..12: amplitudes per tile of the fused pass);
 * "chunk_bits" (0 = auto, 2..12: contiguous amplitudes per chunk of a tile; at least two tile bits
 * are always left for gathered qubits); "tma" (tile loads: 1 cp.async.bulk + mbarrier, 0 16-byte
 * cp.async, -1 = auto, default: bulk copies for passes that read a peer shard); "tile_ctas" (0 =
 * auto, 3..5 CTAs per SM for 2^11-amplitude tiles); "ptx_ops" (1, default: the fast interpreter's
 * op loop as one inline-PTX block; 0: the C++ loop); "remap" (1, default: a pass on a global qubit
 * leaves it local -- logical -> physical qubit map; 0: exchange and write back);
 * "peer_chunk_bits", "peer_tile_bits" (the same two sizes for passes that start on a global
 * qubit); "single_ctrl" (1, default: diagonal ops with one control in a register slot run
 * through their own arms; 0: the generic predicated ones); "butterfly" (1, default: an uncontrolled
 * h adds / subtracts and its 1/sqrt(2) is folded into another gate of the pass); "lower_two_bit"
 * (0, default; 1: swap / i_swap / rxx / ryy of an op list run as products of cx, s, z, h, rzz so
 * that their pass needs no full interpreter -- measured slower on configs[4]); "double_buffer" (0; 1: two tile buffers
 * per CTA, 2: only for passes that read a peer shard), "prefetch" (experiments, off); "profile" (0/1: time every launch with CUDA
 * events); "seed". */
int qvnt_reg_set_option(qvnt_reg_t *reg, const char *key, int64_t value);
int qvnt_reg_stats(qvnt_reg_t *reg, qvnt_stats_t *out);
int qvnt_reg_stats_reset(qvnt_reg_t *reg);
/* CUDA-event stopwatch on the handle's stream: mark slot (0..15), then read the
 * elapsed milliseconds between two marked slots (blocks until both passed). */
int qvnt_reg_mark(qvnt_reg_t *reg, int slot);
int qvnt_reg_elapsed_ms(qvnt_reg_t *reg, int slot_from, int slot_to, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* QVNT_B200_H */
