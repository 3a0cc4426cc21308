/*
 * qvnt_oracle.c -- CPU restatement of the QVNT gate-application hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path (libqvnt_b200.so) never links, loads
 * or calls it.
 *
 * What it restates (all citations relative to /root/reference/src):
 *   - the sweep + control predicate        operator/atomic/dispatch.rs:32-67
 *   - the 20 atomic gate formulas          operator/atomic/{x,y,z,s,t,rx,rxx,ry,ryy,
 *                                          rz,rzz,u1,u2,h1,h2,swap,i_swap,sqrt_swap,
 *                                          sqrt_i_swap,id}.rs::atomic_op
 *   - i^q rotation helper                  math/mod.rs:41-50
 *   - MultiOp ping-pong application        operator/multi/mod.rs:96-114
 *   - register: new/with_state/apply/normalize/get_probabilities/get_absolute/
 *     collapse_mask/measure_mask/reset/reset_by_mask/tensor_prod
 *                                          register/quant.rs:113-150,202-229,330-371,
 *                                          376-501
 *   - WeightedIndex sampling (rand 0.8.5, un-vendored dependency, Cargo.lock):
 *     cumulative sums, x = u * total, partition_point(w <= x); the uniform
 *     variate u is an INPUT here because the reference seeds from entropy
 *     (register/quant.rs:496-497) -- the sampled outcome itself is "parity
 *     unpinned" by the reference, the distribution is pinned.
 *
 * The reference (Rust) cannot be compiled in this image (no rustc/cargo, crates
 * un-vendored), so this file is a "port"; it is pinned against every golden
 * vector the reference's own tests hold for the path (tests/test_oracle_golden.py).
 *
 * Floating point: plain IEEE f64, evaluated in exactly the reference's
 * operation order, compiled with -ffp-contract=off (Rust never contracts a*b+c).
 * Complex multiply follows num_complex 0.4.2:
 *     (a*b).re = a.re*b.re - a.im*b.im ; (a*b).im = a.re*b.im + a.im*b.re
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } qc;

/* Same memory layout as qvnt_op_t in include/qvnt_b200.h (checked by tests). */
typedef struct {
    uint32_t kind;      /* AtomicOpDispatch order, operator/atomic/dispatch.rs:84-105 */
    uint32_t dagger;    /* s/t/i_swap/sqrt_swap/sqrt_i_swap dagger flag */
    uint64_t a_mask;    /* a_mask or ab_mask */
    uint64_t b_mask;    /* h2/u2 only */
    uint64_t ctrl;      /* SingleOp.ctrl, operator/single/mod.rs:43-47 */
    double ph_re, ph_im;/* rotation phase (cos t/2, sin t/2) as stored by Op::new */
    double matrix[32];  /* u1: 4 complex row-major; u2: 16 complex row-major */
} qo_op;

enum {
    K_ID = 0, K_X, K_RX, K_RXX, K_Y, K_RY, K_RYY, K_Z, K_S, K_T, K_RZ, K_RZZ,
    K_U1, K_U2, K_H1, K_H2, K_SWAP, K_ISWAP, K_SQRTSWAP, K_SQRTISWAP, K_COUNT
};

#define FRAC_1_SQRT_2 0.70710678118654752440084436210485

static inline qc c_mul(qc a, qc b) {               /* num_complex Mul */
    qc r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static inline qc c_add(qc a, qc b) { qc r = { a.re + b.re, a.im + b.im }; return r; }
static inline qc c_neg(qc a) { qc r = { -a.re, -a.im }; return r; }
static inline qc c_scale(qc a, double t) { qc r = { a.re * t, a.im * t }; return r; }

/* math/mod.rs:41-50 : multiply by i^(q mod 4) by sign flips / swaps only. */
static inline qc rotate_i(qc z, uint64_t q) {
    if (q & 2) z = c_neg(z);
    if (q & 1) { double im = -z.im; z.im = z.re; z.re = im; }
    return z;
}

static inline unsigned pc(uint64_t v) { return (unsigned)__builtin_popcountll(v); }

/* One output amplitude of one atomic gate: operator/atomic/<kind>.rs::atomic_op */
static inline qc atomic_op(const qo_op *op, const qc *psi, uint64_t idx) {
    const uint64_t a = op->a_mask, b = op->b_mask;
    switch (op->kind) {
    case K_ID:                                              /* id.rs:7-9 */
        return psi[idx];
    case K_X:                                               /* x.rs:15-17 */
        return psi[idx ^ a];
    case K_Y: {                                             /* y.rs:10-23 */
        /* i_pow = !(count_ones + 1) evaluated in u32, zero-extended (y.rs:11) */
        uint64_t i_pow = (uint64_t)(uint32_t)~(uint32_t)(pc(a) + 1u);
        if ((pc(idx & a) & 1u) == 0) i_pow ^= 2;
        return rotate_i(psi[idx ^ a], i_pow);
    }
    case K_Z:                                               /* z.rs:15-21 */
        return (pc(idx & a) & 1u) ? c_neg(psi[idx]) : psi[idx];
    case K_S: {                                             /* s.rs:19-25 */
        uint64_t count = pc(idx & a);
        if (op->dagger) count = (~count) + 1u;
        return rotate_i(psi[idx], count);
    }
    case K_T: {                                             /* t.rs:24-35 */
        uint64_t count = pc(idx & a);
        if (op->dagger) count = (~count) + 1u;
        qc p = rotate_i(psi[idx], count >> 1);
        if (count & 1) { qc e = { FRAC_1_SQRT_2, FRAC_1_SQRT_2 }; return c_mul(e, p); }
        return p;
    }
    case K_RX:                                              /* rx.rs:18-24 */
    case K_RXX: {                                           /* rxx.rs:19-25 */
        qc p0 = psi[idx], p1 = psi[idx ^ a];
        qc r = { p0.re * op->ph_re + p1.im * op->ph_im,
                 p0.im * op->ph_re - p1.re * op->ph_im };
        return r;
    }
    case K_RY: {                                            /* ry.rs:19-29 */
        double c = op->ph_re, s = op->ph_im;
        qc p0 = psi[idx], p1 = psi[idx ^ a];
        if ((idx & a) == 0) s = -s;
        qc r = { p0.re * c + p1.re * s, p0.im * c + p1.im * s };
        return r;
    }
    case K_RYY: {                                           /* ryy.rs:19-29 */
        double c = op->ph_re, s = op->ph_im;
        qc p0 = psi[idx], p1 = psi[idx ^ a];
        if ((pc(idx & a) & 1u) == 0) s = -s;
        qc r = { p0.re * c + p1.im * s, p0.im * c - p1.re * s };
        return r;
    }
    case K_RZ: {                                            /* rz.rs:19-25 */
        qc ph = { op->ph_re, op->ph_im };
        if ((idx & a) == 0) ph.im = -ph.im;
        return c_mul(ph, psi[idx]);
    }
    case K_RZZ: {                                           /* rzz.rs:19-25 */
        qc ph = { op->ph_re, op->ph_im };
        if ((pc(idx & a) & 1u) == 0) ph.im = -ph.im;
        return c_mul(ph, psi[idx]);
    }
    case K_U1: {                                            /* u1.rs:17-25 */
        const qc *m = (const qc *)op->matrix;
        int a_bit = (idx & a) != 0;
        uint64_t base = idx & ~a;
        return c_add(c_mul(m[2 * a_bit + 0], psi[base]), c_mul(m[2 * a_bit + 1], psi[base | a]));
    }
    case K_U2: {                                            /* u2.rs:22-49 */
        const qc *m = (const qc *)op->matrix;
        int a_bit = (idx & a) != 0, b_bit = (idx & b) != 0;
        uint64_t base = idx & ~a & ~b;
        const qc *row = m + 4 * (2 * b_bit + a_bit);
        qc r = c_mul(row[0], psi[base]);
        r = c_add(r, c_mul(row[1], psi[base | a]));
        r = c_add(r, c_mul(row[2], psi[base | b]));
        r = c_add(r, c_mul(row[3], psi[base | a | b]));
        return r;
    }
    case K_H1: {                                            /* h1.rs:16-22 */
        qc p0 = psi[idx], p1 = psi[idx ^ a];
        if (idx & a) p0 = c_neg(p0);
        return c_scale(c_add(p0, p1), FRAC_1_SQRT_2);
    }
    case K_H2: {                                            /* h2.rs:22-38 */
        qc p0 = psi[idx], p1 = psi[idx ^ a], p2 = psi[idx ^ b], p3 = psi[idx ^ (a | b)];
        if (idx & a) { p0 = c_neg(p0); p2 = c_neg(p2); }
        if (idx & b) { p0 = c_neg(p0); p1 = c_neg(p1); }
        return c_scale(c_add(c_add(c_add(p0, p1), p2), p3), 0.5);
    }
    case K_SWAP:                                            /* swap.rs:16-22 */
        return (pc(idx & a) & 1u) ? psi[idx ^ a] : psi[idx];
    case K_ISWAP:                                           /* i_swap.rs:20-37 */
        if (pc(idx & a) & 1u) {
            qc p = psi[idx ^ a];
            qc r;
            if (op->dagger) { r.re = p.im; r.im = -p.re; }
            else            { r.re = -p.im; r.im = p.re; }
            return r;
        }
        return psi[idx];
    case K_SQRTSWAP:                                        /* sqrt_swap.rs:20-37 */
        if (pc(idx & a) & 1u) {
            qc p0 = psi[idx], p1 = psi[idx ^ a];
            qc r;
            if (op->dagger) {
                r.re = 0.5 * (p0.re + p0.im + p1.re - p1.im);
                r.im = 0.5 * (p0.im - p0.re + p1.im + p1.re);
            } else {
                r.re = 0.5 * (p0.re - p0.im + p1.re + p1.im);
                r.im = 0.5 * (p0.im + p0.re + p1.im - p1.re);
            }
            return r;
        }
        return psi[idx];
    case K_SQRTISWAP:                                       /* sqrt_i_swap.rs:20-37 */
        if (pc(idx & a) & 1u) {
            qc p0 = psi[idx], p1 = psi[idx ^ a];
            qc r;
            if (op->dagger) {
                r.re = FRAC_1_SQRT_2 * (p0.re + p1.im);
                r.im = FRAC_1_SQRT_2 * (p0.im - p1.re);
            } else {
                r.re = FRAC_1_SQRT_2 * (p0.re - p1.im);
                r.im = FRAC_1_SQRT_2 * (p0.im + p1.re);
            }
            return r;
        }
        return psi[idx];
    default:
        return psi[idx];
    }
}

/* AtomicOp::for_each / for_each_par (dispatch.rs:32-67): one full out-of-place
 * sweep over ALL len outputs, control predicate (!idx & ctrl) == 0. */
static void sweep(const qo_op *op, const qc *in, qc *out, uint64_t len, int threads) {
    const uint64_t ctrl = op->ctrl;
    if (threads <= 1) {
        for (uint64_t i = 0; i < len; ++i)
            out[i] = ((~i & ctrl) == 0) ? atomic_op(op, in, i) : in[i];
    } else {
#pragma omp parallel for schedule(static) num_threads(threads)
        for (uint64_t i = 0; i < len; ++i)
            out[i] = ((~i & ctrl) == 0) ? atomic_op(op, in, i) : in[i];
    }
}

/* ------------------------------------------------------------------------- */
/* Register (register/quant.rs)                                              */
/* ------------------------------------------------------------------------- */
#define MIN_BUFFER_LEN 8u                                   /* quant.rs:15 */

typedef struct {
    qc *psi;
    uint64_t len;       /* max(2^q_num, 8) */
    uint32_t q_num;
    uint64_t q_mask;
    int threads;        /* 1 = threading::Single, n = Multi(n) */
} qo_reg;

static qc *alloc_psi(uint64_t len) {
    void *p = NULL;
    if (posix_memalign(&p, 64, (size_t)len * sizeof(qc)) != 0) return NULL;
    return (qc *)p;
}

/* QReg::with_state (quant.rs:129-150); QReg::new == with_state(n, 0). */
qo_reg *qo_reg_create(uint32_t q_num, uint64_t state, int threads) {
    qo_reg *r = (qo_reg *)calloc(1, sizeof(qo_reg));
    if (!r) return NULL;
    uint64_t q_size = 1ull << q_num;
    r->q_num = q_num;
    r->q_mask = q_size - 1;
    r->len = q_size > MIN_BUFFER_LEN ? q_size : MIN_BUFFER_LEN;
    r->threads = threads < 1 ? 1 : threads;
    r->psi = alloc_psi(r->len);
    if (!r->psi) { free(r); return NULL; }
    memset(r->psi, 0, (size_t)r->len * sizeof(qc));
    r->psi[state & r->q_mask].re = 1.0;
    return r;
}

void qo_reg_destroy(qo_reg *r) { if (r) { free(r->psi); free(r); } }

qo_reg *qo_reg_clone(const qo_reg *r) {
    qo_reg *c = (qo_reg *)malloc(sizeof(qo_reg));
    *c = *r;
    c->psi = alloc_psi(r->len);
    memcpy(c->psi, r->psi, (size_t)r->len * sizeof(qc));
    return c;
}

uint64_t qo_reg_len(const qo_reg *r) { return r->len; }
uint32_t qo_reg_qnum(const qo_reg *r) { return r->q_num; }
qc *qo_reg_data(qo_reg *r) { return r->psi; }
void qo_reg_set_threads(qo_reg *r, int t) { r->threads = t < 1 ? 1 : t; }

/* QReg::reset (quant.rs:202-205) */
void qo_reg_reset(qo_reg *r, uint64_t state) {
    memset(r->psi, 0, (size_t)r->len * sizeof(qc));
    r->psi[r->q_mask & state].re = 1.0;
}

/* QReg::apply (quant.rs:376-395) + MultiOp::apply (multi/mod.rs:96-114):
 * fresh output buffer, copy of the input (to_vec), one sweep per SingleOp with
 * ping-pong, result swapped into the register.  3 state-sized buffers live. */
int qo_reg_apply(qo_reg *r, const qo_op *ops, uint64_t n_ops) {
    qc *out = alloc_psi(r->len);            /* Vec::with_capacity + set_len */
    qc *in = alloc_psi(r->len);             /* psi_i.to_vec() */
    if (!out || !in) { free(out); free(in); return 1; }
    memcpy(in, r->psi, (size_t)r->len * sizeof(qc));
    for (uint64_t k = 0; k < n_ops; ++k) {
        sweep(&ops[k], in, out, r->len, r->threads);
        qc *t = in; in = out; out = t;      /* mem::swap(&mut psi_i, psi_o) */
    }
    /* final swap: result is in `in` */
    free(r->psi);
    r->psi = in;
    free(out);
    return 0;
}

/* QReg::get_absolute (quant.rs:458-466): sum |a|^2 over the whole buffer. */
double qo_reg_norm_sqr(const qo_reg *r) {
    double s = 0.0;
    if (r->threads <= 1) {
        for (uint64_t i = 0; i < r->len; ++i)
            s += r->psi[i].re * r->psi[i].re + r->psi[i].im * r->psi[i].im;
    } else {
#pragma omp parallel for reduction(+ : s) schedule(static) num_threads(r->threads)
        for (uint64_t i = 0; i < r->len; ++i)
            s += r->psi[i].re * r->psi[i].re + r->psi[i].im * r->psi[i].im;
    }
    return s;
}

/* QReg::get_probabilities (quant.rs:434-454): |a_i|^2 * (1/sum), i < 2^n. */
void qo_reg_probabilities(const qo_reg *r, double *out) {
    double inv = 1.0 / qo_reg_norm_sqr(r);
    uint64_t n = 1ull << r->q_num;
#pragma omp parallel for schedule(static) num_threads(r->threads) if (r->threads > 1)
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (r->psi[i].re * r->psi[i].re + r->psi[i].im * r->psi[i].im) * inv;
}

/* QReg::normalize (quant.rs:397-414) */
void qo_reg_normalize(qo_reg *r) {
    double norm = sqrt(qo_reg_norm_sqr(r));
    if (norm <= 1e-15) { qo_reg_reset(r, 0); return; }
    if (1.0 - norm <= 1e-9) return;
    norm = 1.0 / norm;
#pragma omp parallel for schedule(static) num_threads(r->threads) if (r->threads > 1)
    for (uint64_t i = 0; i < r->len; ++i) { r->psi[i].re *= norm; r->psi[i].im *= norm; }
}

/* QReg::reset_by_mask (quant.rs:207-229) */
void qo_reg_reset_by_mask(qo_reg *r, uint64_t mask) {
    if ((mask & r->q_mask) == r->q_mask) { qo_reg_reset(r, 0); return; }
#pragma omp parallel for schedule(static) num_threads(r->threads) if (r->threads > 1)
    for (uint64_t i = 0; i < r->len; ++i)
        if (i & mask) { r->psi[i].re = 0.0; r->psi[i].im = 0.0; }
    qo_reg_normalize(r);
}

/* QReg::collapse_mask (quant.rs:468-486): zero where (idx ^ idy) & mask != 0.
 * No renormalisation. */
void qo_reg_collapse(qo_reg *r, uint64_t idy, uint64_t mask) {
#pragma omp parallel for schedule(static) num_threads(r->threads) if (r->threads > 1)
    for (uint64_t i = 0; i < r->len; ++i)
        if ((i ^ idy) & mask) { r->psi[i].re = 0.0; r->psi[i].im = 0.0; }
}

/* WeightedIndex::new + sample with an injected uniform variate u in [0,1):
 * total = sequential sum of weights; x = u * total; result = number of
 * cumulative sums (over all but the last weight) that are <= x.
 * (rand 0.8.5 distributions/weighted_index.rs; call site quant.rs:496-497.) */
uint64_t qo_weighted_index(const double *w, uint64_t n, double u) {
    double total = w[0];
    for (uint64_t i = 1; i < n; ++i) total += w[i];
    double x = u * total;
    double cum = w[0];
    uint64_t i = 0;
    /* cumulative_weights[k] = w0+..+wk for k in 0..n-1 (last one excluded) */
    while (i + 1 < n && cum <= x) { ++i; cum += w[i]; }
    return i;
}

/* QReg::measure_mask (quant.rs:490-501) with injected u.  Returns the CReg
 * value (rand_idx & mask); *sampled gets the full sampled index. */
uint64_t qo_reg_measure_mask(qo_reg *r, uint64_t mask, double u, uint64_t *sampled) {
    mask &= r->q_mask;
    if (mask == 0) { if (sampled) *sampled = 0; return 0; }
    uint64_t n = 1ull << r->q_num;
    double *p = (double *)malloc((size_t)n * sizeof(double));
    qo_reg_probabilities(r, p);
    uint64_t idx = qo_weighted_index(p, n, u);
    free(p);
    qo_reg_collapse(r, idx, mask);
    if (sampled) *sampled = idx;
    return idx & mask;
}

/* QReg::tensor_prod (quant.rs:330-371): out[idx] = a[idx & mask_a] * b[idx >> n_a] */
qo_reg *qo_reg_tensor_prod(const qo_reg *a, const qo_reg *b) {
    int th = a->threads > b->threads ? a->threads : b->threads;
    qo_reg *r = qo_reg_create(a->q_num + b->q_num, 0, th);
    uint64_t q_size = 1ull << r->q_num;
    for (uint64_t i = 0; i < r->len; ++i) {
        if (i < q_size)
            r->psi[i] = c_mul(a->psi[i & a->q_mask], b->psi[(i >> a->q_num) & b->q_mask]);
        else { r->psi[i].re = 0.0; r->psi[i].im = 0.0; }
    }
    return r;
}

/* Helpers for the harness (not reference functions). */
void qo_reg_read(const qo_reg *r, uint64_t off, uint64_t cnt, double *re_im) {
    memcpy(re_im, r->psi + off, (size_t)cnt * sizeof(qc));
}
void qo_reg_write(qo_reg *r, uint64_t off, uint64_t cnt, const double *re_im) {
    memcpy(r->psi + off, re_im, (size_t)cnt * sizeof(qc));
}
/* One raw sweep on caller buffers: SingleOp::apply (single/mod.rs:83-93). */
void qo_sweep(const qo_op *op, const double *in, double *out, uint64_t len, int threads) {
    sweep(op, (const qc *)in, (qc *)out, len, threads);
}
/* Time-only variant used by the CPU baseline: out-of-place sweeps between two
 * caller-owned buffers without the per-apply allocation (reported separately). */
int qo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
uint64_t qo_sizeof_op(void) { return sizeof(qo_op); }
