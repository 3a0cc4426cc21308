// tile.cu -- the fused tile pass: one in-place HBM sweep carrying many SingleOps.
//
// Replaces a RUN of the reference's per-SingleOp sweeps (AtomicOp::for_each,
// src/operator/atomic/dispatch.rs:32-67, driven by MultiOp::apply,
// src/operator/multi/mod.rs:96-114): instead of streaming the whole state through
// DRAM once per gate, a CTA stages a 2^T-amplitude tile in shared memory, applies every
// gate of the pass whose partner bits are tile bits, and writes the tile back.  Per
// pass the HBM traffic is 16 B read + 16 B written per amplitude, whatever the number
// of gates carried.
//
// Layout of a tile: T tile bits = low L bits (contiguous 16*2^L-byte chunks) + T-L gathered high
// bits.  A gathered bit may be a rank bit of a sharded register: that chunk is then read from /
// written to the peer GPU's HBM through its NVLink-mapped pointer (Segs), so a global-qubit gate
// needs no separate exchange step -- the "swap" is the tile's own load and store.
//
// Tile loads are TMA bulk copies: one `cp.async.bulk.shared.global` per chunk, issued by the lanes
// of warp 0, completing on an mbarrier every thread of the CTA waits on (SASS: UBLKCP + SYNCS).
// No thread spends instructions or registers on the load, and the data lands while the CTA runs
// the last stage of the previous tile.  (Option "bulk" = 0 keeps a 16-byte cp.async path for A/B
// measurements.)
//
// Shared-memory layout: chunk c lives at byte offset c * (16 * 2^L + 16) -- chunks stay linear (a
// bulk copy needs that) and ONE 16-byte pad per chunk rotates the bank group of consecutive
// chunks.  The address of tile-local index j is 16 * (j + (j >> L)): additive in the index bits,
// so the 16 addresses of a thread are one base plus stage-uniform constants.  A quarter-warp
// (8 lanes, 128 bytes) is conflict free when the three low lane bits are tile-local bits of the
// classes {0, L}, {1, L+1}, {2, L+2}; the planner assigns them that way whenever the stage's
// register slots leave one bit of each class free.
//
// Stages.  A stage keeps TILE_R = 4 tile bits in "register slots": every thread holds the 16
// amplitudes that differ in those bits, runs all of the stage's ops on them and only then touches
// shared memory again.  The last stage of a pass stores its registers straight to HBM, so the tile
// buffer is free as soon as that stage has read it and the next tile's load is issued there.
// An op is an 80-byte MOp in shared memory whose `code` selects a fully unrolled arm through ONE
// jump table; everything that depends on the amplitude index was split by the planner into
//   - register-slot part: compile-time per arm, or a 16-bit `okmask` (controls in register slots),
//   - thread part: one AND/compare on the thread's group number, only for ops flagged MOP_COND,
//   - tile part: one flag byte per op per tile (controls / diagonal-mask bits outside the
//     tile), computed once per tile and only for passes that have such ops (hdr.need_flags).
// Tiles none of whose ops is active (multi-controlled gates) are skipped without being read.
//
// Two interpreters share the plumbing: the FULL one carries every kind with the reference's
// formulas (gates.cuh); the FAST one (passes made of x, y, rx, ry, h1, z/s/t on one bit, rz, rzz)
// runs coefficient-driven arms written as in-place inline PTX -- see "FAST stage interpreter".
//
// Arithmetic: this file is compiled with FMA contraction ON (csrc/Makefile) and the planner may
// exchange COMMUTING ops, so the fused path is parity-checked at 1e-10, not bit-exact; the direct
// sweeps (direct.cu, -fmad=false) are the bit-exact path.
//
// Hazards the barriers cover: (1) between stages (threads exchange amplitudes through the tile
// buffer); (2) before the next tile's load overwrites the buffer (after the last stage's loads);
// (3) between a stage's loads and stores when the stage holds a lazy x on a thread bit, because
// threads then store into EACH OTHER's slots (TStage::sync_after_load); (4) across GPUs,
// dist_barrier() kernels around every pass that touches a peer shard, plus a system fence after
// the peer stores.
#include <mutex>

#include "engine.h"
#include "fastops_ptx.inc"
#include "gates.cuh"

namespace qv {

constexpr int NV = TILE_NV;
constexpr int TR = TILE_R;

// tile-local index -> 16-byte slot of the padded-linear tile buffer
__device__ __forceinline__ uint32_t slot16(uint32_t j, uint32_t L) { return j + (j >> L); }

// ---- op bodies of the FULL interpreter: templated on kind and register slot(s) ------------------
#define QV_FOR_K _Pragma("unroll") for (int K = 0; K < NV; ++K)
// Controls held in register slots are warp-uniform (okmask comes from the op descriptor): keep
// them as real branches -- the empty volatile asm stops the compiler from if-converting the
// guarded update into "compute both, then 4 selects per amplitude".
#define QV_GUARD(ALL) do { if (!(ALL)) asm volatile(""); } while (0)

// one-bit pair ops (x/y/rx/ry/h1/u1) on register slot RB
// The 8 pair updates of an op run as two halves separated by an always-taken branch the
// compiler cannot fold (gridDim.y == 1): ptxas schedules inside basic blocks, so at most 4 pairs'
// products are in flight at once -- enough to fill the FP64 pipe, and ~20 registers fewer than
// the full interleave, which is what keeps the interpreter's loop state out of local memory.
#define QV_HALF_SPLIT (gridDim.y == 1u)

template <int KIND, int RB, bool ALL>
__device__ __forceinline__ void body_p1(amp (&v)[NV], const GP &g, const uint32_t ok) {
    constexpr int HB = (RB == TR - 1) ? (1 << (TR - 2)) : (1 << (TR - 1));   // a slot bit other than RB
    QV_FOR_K {
        if ((K & (1 << RB)) || (K & HB)) continue;
        if (ALL || ((ok >> K) & 1u)) {
            QV_GUARD(ALL);
            pair_update_s<KIND>(g, v[K], v[K | (1 << RB)], 0u, 1u);
        }
    }
    if (QV_HALF_SPLIT) {
        QV_FOR_K {
            if ((K & (1 << RB)) || !(K & HB)) continue;
            if (ALL || ((ok >> K) & 1u)) {
                QV_GUARD(ALL);
                pair_update_s<KIND>(g, v[K], v[K | (1 << RB)], 0u, 1u);
            }
        }
    }
}

// rxx/ryy on slot pair P: partner differs in both slots; the two members share their parity
template <int KIND, int P, bool ALL>
__device__ __forceinline__ void body_p2x(amp (&v)[NV], const GP &g, const uint32_t ok) {
    constexpr int A = 1 << (2 * P), B = 2 << (2 * P);
    QV_FOR_K {
        if (K & A) continue;
        if (ALL || ((ok >> K) & 1u)) {
            QV_GUARD(ALL);
            const unsigned sel = (K & B) ? 1u : 0u;      // parity of (i & ab): a clear, b = K's bit
            pair_update_s<KIND>(g, v[K], v[K ^ (A | B)], sel, sel);
        }
    }
}

// swap family on slot pair P: only the odd-parity pair {a set, b set} changes
template <int KIND, int P, bool ALL>
__device__ __forceinline__ void body_odd(amp (&v)[NV], const GP &g, const uint32_t ok) {
    constexpr int A = 1 << (2 * P), B = 2 << (2 * P);
    QV_FOR_K {
        if (K & (A | B)) continue;
        if (ALL || ((ok >> K) & 1u)) {
            QV_GUARD(ALL);
            pair_update_s<KIND>(g, v[K | A], v[K | B], 1u, 1u);
        }
    }
}

// h2/u2: a = slot SA, b = slot SB
template <int KIND, int SA, int SB, bool ALL>
__device__ __forceinline__ void body_quad(amp (&v)[NV], const GP &g, const uint32_t ok) {
    constexpr int A = 1 << SA, B = 1 << SB;
    QV_FOR_K {
        if (K & (A | B)) continue;
        if (ALL || ((ok >> K) & 1u)) {
            QV_GUARD(ALL);
            amp q[4] = {v[K], v[K | A], v[K | B], v[K | A | B]};
            quad_update<KIND>(g.mat, q);
            v[K] = q[0];
            v[K | A] = q[1];
            v[K | B] = q[2];
            v[K | A | B] = q[3];
        }
    }
}

// diagonal ops; cnt_t = popcount of the target mask over the thread + tile part of the index
template <int KIND, bool PER_SLOT, bool ALL>
__device__ __forceinline__ void body_diag(amp (&v)[NV], const GP &g, const uint32_t ok, const uint32_t cnt_t,
                                          const uint32_t a_reg) {
    if (!PER_SLOT) {
        if (KIND == QVNT_Z && (cnt_t & 1u) == 0) return;
        QV_FOR_K {
            if (K >= NV / 2) continue;
            if (ALL || ((ok >> K) & 1u)) {
                QV_GUARD(ALL);
                v[K] = diag_out_s<KIND>(g, v[K], cnt_t);
            }
        }
        if (QV_HALF_SPLIT) {
            QV_FOR_K {
                if (K < NV / 2) continue;
                if (ALL || ((ok >> K) & 1u)) {
                    QV_GUARD(ALL);
                    v[K] = diag_out_s<KIND>(g, v[K], cnt_t);
                }
            }
        }
    } else {
        uint32_t b[TR];
#pragma unroll
        for (int j = 0; j < TR; ++j) b[j] = (a_reg >> j) & 1u;
        QV_FOR_K {
            if (ALL || ((ok >> K) & 1u)) {
                QV_GUARD(ALL);
                uint32_t cnt = cnt_t;
#pragma unroll
                for (int j = 0; j < TR; ++j)
                    if (K & (1 << j)) cnt += b[j];
                v[K] = diag_out_s<KIND>(g, v[K], cnt);
            }
        }
    }
}

#define QV_P1(KIND, base)                                                   \
    case base + 0: body_p1<KIND, 0, ALL>(v, g, ok); break;                  \
    case base + 1: body_p1<KIND, 1, ALL>(v, g, ok); break;                  \
    case base + 2: body_p1<KIND, 2, ALL>(v, g, ok); break;                  \
    case base + 3: body_p1<KIND, 3, ALL>(v, g, ok); break;
#define QV_P2X(KIND, base)                                                  \
    case base + 0: body_p2x<KIND, 0, ALL>(v, g, ok); break;                 \
    case base + 1: body_p2x<KIND, 1, ALL>(v, g, ok); break;
#define QV_ODD(KIND, base)                                                  \
    case base + 0: body_odd<KIND, 0, ALL>(v, g, ok); break;                 \
    case base + 1: body_odd<KIND, 1, ALL>(v, g, ok); break;
#define QV_DIAG(KIND, k5)                                                                              \
    case MC_DU + k5: body_diag<KIND, false, ALL>(v, g, ok, __popc(grp & m.a_thr) + (fl & 7u), 0); break; \
    case MC_DG + k5: body_diag<KIND, true, ALL>(v, g, ok, __popc(grp & m.a_thr) + (fl & 7u), m.a_reg & 0xFFFFu); break;

// One decoded op of the FULL interpreter: the first 32 bytes of the MOp (two 16-byte loads).
struct MDec {
    uint32_t w0;        // code | dagger << 8 | okmask << 16
    uint32_t ctrl_thr;
    uint32_t a_thr;
    uint32_t a_reg;     // low 16 bits (high 16: the op's index)
    double ph_re, ph_im;
};

// ALL: no control sits in a register slot (okmask = 0xFFFF) -- the common case runs without
// any per-slot predicate.
template <bool ALL>
__device__ __forceinline__ void apply_mop(const MDec &m, const uint32_t fl, const uint32_t grp,
                                          const amp *__restrict__ mats, amp (&v)[NV]) {
    GP g;
    g.c = m.ph_re;
    g.s = m.ph_im;
    g.dagger = (m.w0 >> 8) & 1u;
    g.ybase = ~2u;                      // y on ONE bit (the planner splits multi-bit x/y masks)
    g.mat = mats + m.a_thr;             // u1/u2 only
    const uint32_t ok = m.w0 >> 16;
    uint32_t code = m.w0 & 0xFFu;
    asm volatile("" : "+r"(code));      // keep the dispatch value 32-bit: ptxas then builds a jump table (BRX)
    switch (code) {
    QV_DIAG(QVNT_Z, 0)
    QV_DIAG(QVNT_S, 1)
    QV_DIAG(QVNT_T, 2)
    QV_DIAG(QVNT_RZ, 3)
    QV_DIAG(QVNT_RZZ, 4)
    QV_P1(QVNT_X, MC_P1 + 0)
    QV_P1(QVNT_Y, MC_P1 + 4)
    QV_P1(QVNT_RX, MC_P1 + 8)
    QV_P1(QVNT_RY, MC_P1 + 12)
    QV_P1(QVNT_H1, MC_P1 + 16)
    default: break;
    }
    switch (code) {
    QV_P1(QVNT_U1, MC_P1 + 20)
    QV_P2X(QVNT_RXX, MC_P2X + 0)
    QV_P2X(QVNT_RYY, MC_P2X + 2)
    QV_ODD(QVNT_SWAP, MC_ODD + 0)
    QV_ODD(QVNT_ISWAP, MC_ODD + 2)
    QV_ODD(QVNT_SQRT_SWAP, MC_ODD + 4)
    QV_ODD(QVNT_SQRT_ISWAP, MC_ODD + 6)
    case MC_H2 + 0: body_quad<QVNT_H2, 0, 1, ALL>(v, g, ok); break;
    case MC_H2 + 1: body_quad<QVNT_H2, 2, 3, ALL>(v, g, ok); break;
    case MC_U2 + 0: body_quad<QVNT_U2, 0, 1, ALL>(v, g, ok); break;
    case MC_U2 + 1: body_quad<QVNT_U2, 1, 0, ALL>(v, g, ok); break;
    case MC_U2 + 2: body_quad<QVNT_U2, 2, 3, ALL>(v, g, ok); break;
    case MC_U2 + 3: body_quad<QVNT_U2, 3, 2, ALL>(v, g, ok); break;
    default: break;
    }
}

// ---- stage plumbing shared by both interpreters ------------------------------------------------
// A stage gives the thread of group `grp` the 16 amplitudes whose tile-local indices are
// jl | (subset of the 4 register-slot bits).  All shared-memory addresses are 32-bit
// shared-window byte offsets into the padded-linear tile buffer: address(K) = mine_o + sum of c[j]
// over K's slot bits.
struct StageCtx {
    uint32_t r_lpos;    // 4 bytes: register slot j -> tile-local bit position
    uint32_t c[TR];     // byte offsets of the 4 register-slot bits
    uint32_t jl;        // tile-local index of slot pattern 0
    uint32_t mine_o;    // its byte offset
    unsigned long long goff;   // last stage, local tiles: byte offset of jl inside the tile's span of the shard
};

// Per-(stage, thread) constants are computed ONCE per kernel (every tile of the pass runs the same
// stages) and kept in shared memory: word = slot16(jl) << 16 | jl.
__device__ __forceinline__ uint32_t stage_jl(const TStage &st, const uint32_t n_t, const uint32_t grp) {
    uint32_t jl = 0;
    for (uint32_t k = 0; k < n_t; ++k) jl |= ((grp >> k) & 1u) << st.t_lpos[k];
    return jl;
}

__device__ __forceinline__ StageCtx stage_ctx(const TStage *st, const uint32_t *jlw, const uint4 *ctab) {
    StageCtx x;
    x.r_lpos = *reinterpret_cast<const uint32_t *>(st->r_lpos);
    const uint4 c = *ctab;
    x.c[0] = c.x;
    x.c[1] = c.y;
    x.c[2] = c.z;
    x.c[3] = c.w;
    const uint32_t w = *jlw;
    x.jl = w & 0xFFFFu;
    x.mine_o = (w >> 16) * 16u;
    x.goff = 0;
    return x;
}

__device__ __forceinline__ void stage_load(const uint32_t tile_s, const StageCtx &x, amp (&v)[NV]) {
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t a = tile_s + x.mine_o;
        if (K & 1) a += x.c[0];
        if (K & 2) a += x.c[1];
        if (K & 4) a += x.c[2];
        if (K & 8) a += x.c[3];
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v[K].x), "=d"(v[K].y) : "r"(a) : "memory");
    }
}

// After lazy inversions (FC_LI) register K of the thread holds the amplitude of slot pattern
// K ^ ib (ib = the inverted slots).  Folded into the store addressing, which is additive in the
// slot bits: base += step[j] and step[j] = -step[j] for every inverted slot j.  `inv` carries one
// byte per slot (0 or MOP_ALT_BYTES).
template <typename T>
__device__ __forceinline__ void fold_inv(const uint32_t inv, T &base, T (&step)[TR]) {
#pragma unroll
    for (int j = 0; j < TR; ++j)
        if ((inv >> (8 * j)) & 0xFFu) {
            base += step[j];
            step[j] = (T)0 - step[j];
        }
}

__device__ __forceinline__ void stage_store_smem(const uint32_t tile_s, const StageCtx &x, const uint32_t inv,
                                                 const amp (&v)[NV]) {
    uint32_t base = tile_s + x.mine_o, c[TR] = {x.c[0], x.c[1], x.c[2], x.c[3]};
    if (inv) fold_inv(inv, base, c);
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t a = base;
        if (K & 1) a += c[0];
        if (K & 2) a += c[1];
        if (K & 4) a += c[2];
        if (K & 8) a += c[3];
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(v[K].x), "d"(v[K].y) : "memory");
    }
}

// Last stage of a pass: the 16 amplitudes go straight from registers to HBM (or to the peer's
// HBM): address = chunk base pointer (tile-invariant table) + this tile's byte offset + offset
// inside the chunk.  The planner makes the low lane bits the lowest tile-local bits that are not
// register slots, so a warp's store instruction covers up to 512 contiguous bytes.
__device__ __forceinline__ void stage_store_global(const uint32_t ptr0_s, const uint32_t L,
                                                   const unsigned long long toff, const StageCtx &x,
                                                   const uint32_t inv, const amp (&v)[NV]) {
    uint32_t offc[TR], chc[TR];       // per register slot: its bit inside the chunk / in the chunk index
#pragma unroll
    for (int j = 0; j < TR; ++j) {
        const uint32_t lp = (x.r_lpos >> (8 * j)) & 0xFFu;
        offc[j] = lp < L ? (16u << lp) : 0u;
        chc[j] = lp < L ? 0u : (8u << (lp - L));
    }
    uint32_t off0 = (x.jl & ((1u << L) - 1u)) * 16u;
    uint32_t ch0 = ptr0_s + (x.jl >> L) * 8u;
    if (inv) {
        fold_inv(inv, off0, offc);
        fold_inv(inv, ch0, chc);
    }
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t off = off0, ch = ch0;
        if (K & 1) { off += offc[0]; ch += chc[0]; }
        if (K & 2) { off += offc[1]; ch += chc[1]; }
        if (K & 4) { off += offc[2]; ch += chc[2]; }
        if (K & 8) { off += offc[3]; ch += chc[3]; }
        unsigned long long base;
        asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(base) : "r"(ch) : "memory");
        *reinterpret_cast<amp *>(base + toff + off) = v[K];
    }
}

// Same, for tiles that lie entirely in this GPU's shard: the address is affine in the index bits,
// address(K) = base + goff(thread) + sum of the byte offsets of K's register-slot bits -- no table.
__device__ __forceinline__ void stage_store_global_local(const unsigned long long base, const uint8_t *s_gpos,
                                                         const StageCtx &x, const uint32_t inv,
                                                         const amp (&v)[NV]) {
    unsigned long long g[TR];
#pragma unroll
    for (int j = 0; j < TR; ++j) g[j] = 16ull << s_gpos[(x.r_lpos >> (8 * j)) & 0xFFu];
    unsigned long long a0 = base + x.goff;
    if (inv) fold_inv(inv, a0, g);
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        unsigned long long a = a0;
        if (K & 1) a += g[0];
        if (K & 2) a += g[1];
        if (K & 4) a += g[2];
        if (K & 8) a += g[3];
        *reinterpret_cast<amp *>(a) = v[K];
    }
}

// The op loop of the FULL interpreter: ops [ob, oe) of the pass on the register-resident amplitudes.
__device__ __forceinline__ void stage_ops_full(const uint32_t ops_s, const uint32_t flags_s, const uint32_t ob,
                                               const uint32_t oe, const uint32_t grp,
                                               const amp *__restrict__ mats, amp (&v)[NV]) {
    // The descriptor of op o+1 is fetched before op o's arithmetic is issued, so its
    // shared-memory latency hides under the FP64 burst of op o.
    MDec m_nx;
    uint32_t fl_nx;
    auto fetch = [&](uint32_t o) {
        const uint32_t op_a = ops_s + MOP_BYTES * o, fl_a = flags_s + o;
        asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(fl_nx) : "r"(fl_a) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(m_nx.w0), "=r"(m_nx.ctrl_thr), "=r"(m_nx.a_thr), "=r"(m_nx.a_reg)
                     : "r"(op_a)
                     : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n"
                     : "=d"(m_nx.ph_re), "=d"(m_nx.ph_im)
                     : "r"(op_a + 16u)
                     : "memory");
    };
    if (ob < oe) fetch(ob);
    for (uint32_t o = ob; o < oe; ++o) {
        const MDec m = m_nx;
        const uint32_t fl = fl_nx;
        if (!(fl & 0x80u) || (~grp & m.ctrl_thr)) {
            if (o + 1 < oe) fetch(o + 1);
            continue;
        }
        if ((m.w0 >> 16) == 0xFFFFu) apply_mop<true>(m, fl, grp, mats, v);
        else apply_mop<false>(m, fl, grp, mats, v);
        if (o + 1 < oe) fetch(o + 1);
    }
}

// ================================================================================================
// FAST stage interpreter.  Passes made only of the common kinds are lowered by the planner to the
// coefficient-driven forms of FCode (engine.h); every arm below updates the 16 register-resident
// amplitudes IN PLACE through inline PTX whose operands are tied ("+d"), so each amplitude keeps
// one home register through the whole op loop: no register shuffling at the merge points of the
// dispatch, no selects.  Per op the instruction stream is the FP64 arithmetic (4 per amplitude:
// 2 mul + 2 fma, the same count as the reference formulas with contraction) plus one 16-byte
// header load, one test of the header's flag bits, the jump table and two 16-byte coefficient
// loads inside the arm.  x / cx never move data unless the control itself sits in a register slot:
// they flip a bit of the thread's store index (FC_LX) or mark a register slot as inverted (FC_LI).
// ================================================================================================
__device__ __forceinline__ void f_pair_real(amp &p0, amp &p1, double a, double b, double c, double d) {
    asm volatile("{\n\t.reg .f64 t, u, w, z;\n\t"
                 "mul.rn.f64 t, %4, %0;\n\t"
                 "mul.rn.f64 u, %6, %0;\n\t"
                 "mul.rn.f64 w, %4, %1;\n\t"
                 "mul.rn.f64 z, %6, %1;\n\t"
                 "fma.rn.f64 %0, %5, %2, t;\n\t"
                 "fma.rn.f64 %1, %5, %3, w;\n\t"
                 "fma.rn.f64 %2, %7, %2, u;\n\t"
                 "fma.rn.f64 %3, %7, %3, z;\n\t}"
                 : "+d"(p0.x), "+d"(p0.y), "+d"(p1.x), "+d"(p1.y)
                 : "d"(a), "d"(b), "d"(c), "d"(d));
}
// new0 = a*p0 - i*b*p1 ; new1 = -i*c*p0 + d*p1
__device__ __forceinline__ void f_pair_cross(amp &p0, amp &p1, double a, double b, double c, double d) {
    asm volatile("{\n\t.reg .f64 t, u, w, z, nb, nc;\n\t"
                 "neg.f64 nb, %5;\n\t"
                 "neg.f64 nc, %6;\n\t"
                 "mul.rn.f64 t, %4, %0;\n\t"      // a * p0.x
                 "mul.rn.f64 w, %4, %1;\n\t"      // a * p0.y
                 "mul.rn.f64 u, %6, %1;\n\t"      // c * p0.y
                 "mul.rn.f64 z, nc, %0;\n\t"      // -c * p0.x
                 "fma.rn.f64 %0, %5, %3, t;\n\t"  // p0.x = b * p1.y + a * p0.x
                 "fma.rn.f64 %1, nb, %2, w;\n\t"  // p0.y = -b * p1.x + a * p0.y
                 "fma.rn.f64 %2, %7, %2, u;\n\t"  // p1.x = d * p1.x + c * p0.y
                 "fma.rn.f64 %3, %7, %3, z;\n\t}" // p1.y = d * p1.y - c * p0.x
                 : "+d"(p0.x), "+d"(p0.y), "+d"(p1.x), "+d"(p1.y)
                 : "d"(a), "d"(b), "d"(c), "d"(d));
}
__device__ __forceinline__ void f_swap(amp &p0, amp &p1) {
    asm volatile("{\n\t.reg .f64 t, u;\n\t"
                 "mov.f64 t, %0;\n\t"
                 "mov.f64 u, %1;\n\t"
                 "mov.f64 %0, %2;\n\t"
                 "mov.f64 %1, %3;\n\t"
                 "mov.f64 %2, t;\n\t"
                 "mov.f64 %3, u;\n\t}"
                 : "+d"(p0.x), "+d"(p0.y), "+d"(p1.x), "+d"(p1.y));
}
// p *= (fr, fi)
__device__ __forceinline__ void f_cmul(amp &p, double fr, double fi) {
    asm volatile("{\n\t.reg .f64 t, u, nfi;\n\t"
                 "neg.f64 nfi, %3;\n\t"
                 "mul.rn.f64 t, nfi, %1;\n\t"     // -fi * y
                 "mul.rn.f64 u, %3, %0;\n\t"      //  fi * x
                 "fma.rn.f64 %0, %2, %0, t;\n\t"  // x = fr * x - fi * y
                 "fma.rn.f64 %1, %2, %1, u;\n\t}" // y = fr * y + fi * x
                 : "+d"(p.x), "+d"(p.y)
                 : "d"(fr), "d"(fi));
}

// controls held in register slots while slots are inverted: register K holds slot pattern K ^ ib,
// so its predicate is okmask bit K ^ ib
__device__ __forceinline__ uint32_t perm_ok(uint32_t ok, const uint32_t inv) {
    if (inv & 0x000000FFu) ok = ((ok & 0x5555u) << 1) | ((ok >> 1) & 0x5555u);
    if (inv & 0x0000FF00u) ok = ((ok & 0x3333u) << 2) | ((ok >> 2) & 0x3333u);
    if (inv & 0x00FF0000u) ok = ((ok & 0x0F0Fu) << 4) | ((ok >> 4) & 0x0F0Fu);
    if (inv & 0xFF000000u) ok = ((ok & 0x00FFu) << 8) | ((ok >> 8) & 0x00FFu);
    return ok;
}

// the op's coefficient block: 32 bytes at +16, or the `alt` block behind it while slot RB is inverted
struct C4 { double c0, c1, c2, c3; };
__device__ __forceinline__ C4 coef4(const unsigned char *op, const uint32_t blk) {
    const double2 *c = reinterpret_cast<const double2 *>(op + 16u + blk);
    const double2 a = c[0], b = c[1];
    return {a.x, a.y, b.x, b.y};
}

template <int RB, bool ALL>
__device__ __forceinline__ void farm_pr(amp (&v)[NV], const unsigned char *op, const uint32_t inv, const uint32_t ok) {
    const C4 m = coef4(op, (inv >> (8 * RB)) & 0xFFu);
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_pair_real(v[K], v[K | (1 << RB)], m.c0, m.c1, m.c2, m.c3);
    }
}
template <int RB, bool ALL>
__device__ __forceinline__ void farm_px(amp (&v)[NV], const unsigned char *op, const uint32_t inv, const uint32_t ok) {
    const C4 m = coef4(op, (inv >> (8 * RB)) & 0xFFu);
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_pair_cross(v[K], v[K | (1 << RB)], m.c0, m.c1, m.c2, m.c3);
    }
}
template <int RB>
__device__ __forceinline__ void farm_sw(amp (&v)[NV], const uint32_t ok) {
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if ((ok >> K) & 1u) f_swap(v[K], v[K | (1 << RB)]);
    }
}
// diagonal, no target bit in a register slot: one factor for the whole thread
template <bool ALL>
__device__ __forceinline__ void farm_du(amp (&v)[NV], const unsigned char *op, const uint32_t w0, const uint32_t ok,
                                        const uint32_t par) {
    if (!par && (w0 & ((uint32_t)MOP_SKIP0 << 8))) return;
    const double2 f = *reinterpret_cast<const double2 *>(op + 16u + 16u * par);
    QV_FOR_K {
        if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], f.x, f.y);
    }
}
// diagonal, exactly one target bit in register slot RB (+ parity `par` of the target bits elsewhere):
// the slot's inversion and the outer parity both exchange the roles of f0 and f1 -- which is what
// the `alt` block holds
template <int RB, bool ALL>
__device__ __forceinline__ void farm_ds(amp (&v)[NV], const unsigned char *op, const uint32_t w0, const uint32_t inv,
                                        const uint32_t ok, const uint32_t par) {
    const uint32_t blk = ((inv >> (8 * RB)) & 0xFFu) ^ (par * MOP_ALT_BYTES);
    const C4 m = coef4(op, blk);
    const bool skip0 = (w0 & ((uint32_t)MOP_SKIP0 << 8)) != 0;      // the ORIGINAL f0 is 1
    if (!skip0 || blk != 0u) {
        QV_FOR_K {
            if (K & (1 << RB)) continue;
            if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], m.c0, m.c1);
        }
    }
    if (!skip0 || blk == 0u) {
        QV_FOR_K {
            if (!(K & (1 << RB))) continue;
            if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], m.c2, m.c3);
        }
    }
}
// diagonal, any set of target bits in register slots (rare: rzz with both bits in slots)
__device__ __forceinline__ void farm_dg(amp (&v)[NV], const unsigned char *op, const uint32_t a_reg, const uint32_t inv,
                                        const uint32_t ok, uint32_t par) {
    const C4 m = coef4(op, 0u);
    // register K holds slot pattern K ^ ib: the inverted target slots add their parity
    const uint32_t ib = ((inv >> 5) & 1u) | ((inv >> 12) & 2u) | ((inv >> 19) & 4u) | ((inv >> 26) & 8u);
    par = (par + __popc(ib & a_reg)) & 1u;
    QV_FOR_K {
        if ((ok >> K) & 1u) {
            if ((__popc((uint32_t)K & a_reg) + par) & 1u) f_cmul(v[K], m.c2, m.c3);
            else f_cmul(v[K], m.c0, m.c1);
        }
    }
}

#define QV_F4(ARM, ALLV, base, ...)                                    \
    case base + 0: ARM<0, ALLV>(__VA_ARGS__); break;                   \
    case base + 1: ARM<1, ALLV>(__VA_ARGS__); break;                   \
    case base + 2: ARM<2, ALLV>(__VA_ARGS__); break;                   \
    case base + 3: ARM<3, ALLV>(__VA_ARGS__); break;

// The op loop of the FAST interpreter.
__device__ __forceinline__ void stage_ops_fast(const unsigned char *ops, const volatile uint8_t *flags, const uint32_t ob,
                                               const uint32_t oe, const uint32_t grp, const uint32_t L, StageCtx &x,
                                               uint32_t &inv, amp (&v)[NV], const double2 *dtab, const uint32_t dstride,
                                               const volatile double2 *dout_v) {
    const double2 *dout = const_cast<const double2 *>(dout_v);
    uint32_t vgrp = grp;
    const unsigned char *p = ops + MOP_BYTES * ob;
    const unsigned char *const pe = ops + MOP_BYTES * oe;
    while (p != pe) {
        const unsigned char *const op = p;
        const uint4 h = *reinterpret_cast<const uint4 *>(op);     // w0 | ctrl_thr | a_thr | a_reg + idx << 16
        p += MOP_BYTES;
        uint32_t code = fc_generic(h.x & 0xFFu) % (uint32_t)FC_TOTAL;   // (the byte also carries the control class)
        if (h.x & ((uint32_t)MOP_COND << 8)) {
            bool skip = (~vgrp & h.y) != 0u;
            if (h.x & ((uint32_t)MOP_CONDB << 8)) skip = skip || !(flags[h.w >> 16] & 0x80u);
            if (skip) {
                if (code == (uint32_t)FC_DM || code == (uint32_t)(FC_DM + FC_MASKED)) p += MOP_BYTES * (h.w & 0xFFFFu);
                continue;
            }
        }
        // parity of the diagonal target bits on thread bits and outside the tile
        auto dpar = [&](const uint32_t w0, const uint32_t a_thr, const uint32_t idx) -> uint32_t {
            uint32_t c = (uint32_t)__popc(vgrp & a_thr);
            if (w0 & ((uint32_t)MOP_PARB << 8)) c += flags[idx];
            return c & 1u;
        };
        uint32_t ok = h.x >> 16;
        if (code >= (uint32_t)FC_MASKED && inv != 0u) ok = perm_ok(ok, inv);
        asm volatile("" : "+r"(code));      // keep the dispatch value 32-bit: one jump table (BRX)
        switch (code) {
        QV_F4(farm_pr, true, FC_PR, v, op, inv, ok)
        QV_F4(farm_px, true, FC_PX, v, op, inv, ok)
        QV_F4(farm_ds, true, FC_DS, v, op, h.x, inv, ok, dpar(h.x, h.z, h.w >> 16))
        case FC_DU: farm_du<true>(v, op, h.x, ok, dpar(h.x, h.z, h.w >> 16)); break;
        case FC_DG:
        case FC_DG + FC_MASKED: farm_dg(v, op, h.w & 0xFu, inv, ok, dpar(h.x, h.z, h.w >> 16)); break;
        case FC_LX: {
            // x between threads: the thread will store its amplitudes where the partner's were
            const uint32_t lp = h.w & 0xFFu, bit = 1u << lp, d = 16u * slot16(bit, L);
            vgrp ^= h.z;
            x.mine_o += (x.jl & bit) ? 0u - d : d;
            x.jl ^= bit;
            x.goff ^= 16ull << ((h.w >> 8) & 0xFFu);
            break;
        }
        case FC_LI: inv ^= MOP_ALT_BYTES << (8u * (h.w & 3u)); break;       // (a_reg = the slot)
        case FC_DM:
        case FC_DM + FC_MASKED: {
            // merged diagonal run: acc = product of the constituents' factors for this thread
            const uint32_t cnt = h.w & 0xFFFFu;
            double ar = 1.0, ai = 0.0;
            if (h.x & ((uint32_t)MOP_STATIC << 8)) {        // tabulated (engine.h MOP_STATIC)
                const double2 a = dtab[h.z * dstride];
                ar = a.x;
                ai = a.y;
                if (h.x & ((uint32_t)MOP_PARB << 8)) {
                    const double2 f = dout[h.z];
                    ar = a.x * f.x - a.y * f.y;
                    ai = a.x * f.y + a.y * f.x;
                }
                p += MOP_BYTES * cnt;
            } else
            for (uint32_t k = 0; k < cnt; ++k, p += MOP_BYTES) {
                const uint4 h2 = *reinterpret_cast<const uint4 *>(p);
                const uint32_t par = dpar(h2.x, h2.z, h2.w >> 16);
                if (!par && (h2.x & ((uint32_t)MOP_SKIP0 << 8))) continue;
                const double2 f = *reinterpret_cast<const double2 *>(p + 16u + 16u * par);
                const double t = ar * f.x - ai * f.y;
                ai = ar * f.y + ai * f.x;
                ar = t;
            }
            QV_FOR_K {
                if ((ok >> K) & 1u) f_cmul(v[K], ar, ai);
            }
            break;
        }
        QV_F4(farm_pr, false, FC_MASKED + FC_PR, v, op, inv, ok)
        QV_F4(farm_px, false, FC_MASKED + FC_PX, v, op, inv, ok)
        QV_F4(farm_ds, false, FC_MASKED + FC_DS, v, op, h.x, inv, ok, dpar(h.x, h.z, h.w >> 16))
        case FC_MASKED + FC_DU: farm_du<false>(v, op, h.x, ok, dpar(h.x, h.z, h.w >> 16)); break;
        case FC_SW + 0: farm_sw<0>(v, ok); break;
        case FC_SW + 1: farm_sw<1>(v, ok); break;
        case FC_SW + 2: farm_sw<2>(v, ok); break;
        case FC_SW + 3: farm_sw<3>(v, ok); break;
        default: break;
        }
    }
}

// The same op loop as ONE inline-PTX block (gen_fastops.py -> fastops_ptx.inc): header fetch,
// control test, a single brx.idx jump table and every arm.  The C++ loop above compiles to a
// compare tree plus divergence bookkeeping (~40 instructions and several dependent branches per op);
// this one spends ~20.  Option "ptx_ops" (default 1) selects it; the C++ loop stays as the readable
// specification and the A/B baseline.
template <bool SC>
__device__ __forceinline__ void stage_ops_fast_ptx(const uint32_t ops_s, const uint32_t flags_s, const uint32_t ob,
                                                   const uint32_t oe, const uint32_t grp, StageCtx &x, uint32_t &inv,
                                                   amp (&v)[NV], const uint32_t dtab_s, const uint32_t dstride_bytes,
                                                   const uint32_t dout_s) {
    uint32_t vgrp = grp;
    const uint32_t pb = ops_s + MOP_BYTES * ob;
    (void)oe;
    // SC: the flavour with the single-control arms (gen_fastops.py explains why there are two)
#define QV_FASTOPS_OPERANDS                                                                                          \
                 : "+d"(v[0].x), "+d"(v[0].y), "+d"(v[1].x), "+d"(v[1].y), "+d"(v[2].x), "+d"(v[2].y), "+d"(v[3].x), \
                   "+d"(v[3].y), "+d"(v[4].x), "+d"(v[4].y), "+d"(v[5].x), "+d"(v[5].y), "+d"(v[6].x), "+d"(v[6].y), \
                   "+d"(v[7].x), "+d"(v[7].y), "+d"(v[8].x), "+d"(v[8].y), "+d"(v[9].x), "+d"(v[9].y), "+d"(v[10].x),\
                   "+d"(v[10].y), "+d"(v[11].x), "+d"(v[11].y), "+d"(v[12].x), "+d"(v[12].y), "+d"(v[13].x),         \
                   "+d"(v[13].y), "+d"(v[14].x), "+d"(v[14].y), "+d"(v[15].x), "+d"(v[15].y), "+r"(vgrp), "+r"(inv), \
                   "+r"(x.jl), "+r"(x.mine_o), "+l"(x.goff)                                                          \
                 : "r"(pb), "r"(0u), "r"(flags_s), "r"(dtab_s), "r"(dout_s), "r"(dstride_bytes)                      \
                 : "memory"
    if (SC) asm volatile(QV_FASTOPS_PTX_SC QV_FASTOPS_OPERANDS);
    else asm volatile(QV_FASTOPS_PTX QV_FASTOPS_OPERANDS);
#undef QV_FASTOPS_OPERANDS
}

// ---- TMA bulk copy + mbarrier (tile loads) -------------------------------------------------------
__device__ __forceinline__ void mbar_init(const uint32_t mbar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t mbar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t mbar, const uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const unsigned long long src, const uint32_t bytes,
                                         const uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const unsigned long long src, const uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}

// Kernel configurations: <threads per CTA, min CTAs per SM>.
//   T == 12: 256 threads, 2 CTAs/SM, one 68 KiB buffer.
//   T <= 11: 128 threads, 4 CTAs/SM, one 34 KiB buffer each (default; option "tile_ctas" 3).
// One buffer per CTA: the last stage stores its registers straight to HBM, so the buffer is free
// as soon as that stage has read it -- the next tile's load is issued right there and overlaps
// the last stage's arithmetic and stores.
constexpr uint32_t META_SLOTS = 3;
__host__ __device__ inline bool need_flags_smem(const TPassHdr &h, bool full) { return full || h.need_flags != 0u; }      // per-tile op flags rotate through 3 slots: preparing tile i+1 must not
                                        // race with the threads still running the ops of tile i-1

// ---- timeline probe (build variant: make VARIANT=_trace EXTRA=-DQV_TRACE; tools/trace_pass.py) ----
// Thread 0 of the first QV_TRACE_CTAS CTAs stamps clock64 at the phase boundaries of its first
// QV_TRACE_TILES tiles; the last local pass and the last remap pass of a run are kept per device.
#ifdef QV_TRACE
constexpr int QV_TRACE_CTAS = 16, QV_TRACE_TILES = 48, QV_TRACE_PTS = 8;
__device__ unsigned long long g_trace[2][QV_TRACE_CTAS][QV_TRACE_TILES][QV_TRACE_PTS];
__device__ unsigned int g_trace_info[2][8];
#define QV_STAMP(k)                                                                                        \
    do {                                                                                                   \
        if (tid == 0 && (hdr.prefetch & 0x80000000u) && blockIdx.x < QV_TRACE_CTAS && tr_i < QV_TRACE_TILES) \
            g_trace[hdr.remap ? 1 : 0][blockIdx.x][tr_i][k] = (unsigned long long)clock64();               \
    } while (0)
#else
#define QV_STAMP(k) do { } while (0)
#endif

template <int THREADS, int MINB, bool FULL, bool BULK, bool PTXOPS, bool DB, bool SC = false>
__global__ void __launch_bounds__(THREADS, MINB)
k_tile_pass(const __grid_constant__ Segs segs, const __grid_constant__ TPassHdr hdr,
            const TStage *__restrict__ g_stages, const MOp *__restrict__ g_ops,
            const MBase *__restrict__ g_bases, const amp *__restrict__ mats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t T = hdr.T, L = hdr.L;
    const uint32_t n_ops = hdr.n_ops, n_stages = hdr.n_stages;
    const uint32_t tile_len = 1u << T;
    const uint32_t lmask = (1u << L) - 1u;
    const uint32_t n_chunks = 1u << (T - L);
    const uint32_t chunk_bytes = 16u << L, chunk_stride = chunk_bytes + 16u;
    const uint32_t buf_bytes = n_chunks * chunk_stride;
    const uint32_t flags_stride = (n_ops + 15u) & ~15u;

    // shared memory carve-up (16-byte aligned sections first)
    unsigned char *tile_b = smem_raw;                                              // (DB ? 2 : 1) * n_chunks * (16 * 2^L + 16)
    const uint32_t bufs_bytes = DB ? 2u * buf_bytes : buf_bytes;
    unsigned long long *s_mbar = reinterpret_cast<unsigned long long *>(smem_raw + bufs_bytes);   // 16
    MOp *s_ops = reinterpret_cast<MOp *>(smem_raw + bufs_bytes + 16u);             // 80 * (n_ops + n_stages)
    TStage *s_stages = reinterpret_cast<TStage *>(s_ops + n_ops + n_stages);       // 32 * n_stages
    MBase *s_bases = reinterpret_cast<MBase *>(s_stages + n_stages);               // 16 * n_ops
    const uint32_t n_bases = (FULL || hdr.need_flags) ? n_ops : 0u;                 // (per-tile flags only)
    unsigned long long *s_ptr0 = reinterpret_cast<unsigned long long *>(s_bases + n_bases);   // 8 * n_chunks
    uint4 *s_ctab = reinterpret_cast<uint4 *>(s_ptr0 + ((n_chunks + 1u) & ~1u));   // 16 * n_stages
    uint32_t *s_jltab = reinterpret_cast<uint32_t *>(s_ctab + n_stages);           // 4 * nthr * n_stages
    uint8_t *s_gpos = reinterpret_cast<uint8_t *>(s_jltab + nthr * n_stages);      // 16
    uint8_t *s_flags_all = s_gpos + 16;                                            // 3 * flags_stride (need_flags)
    // tabulated diagonal runs (MOP_STATIC): [run][thread] factor of the thread-bit members (once per
    // kernel) and [flag slot][run] factor of the members outside the tile (once per tile)
    const uint32_t n_static = hdr.n_static;
    const uint32_t flags_bytes = need_flags_smem(hdr, FULL) ? META_SLOTS * flags_stride : 16u;
    double2 *s_dtab = reinterpret_cast<double2 *>(s_flags_all + flags_bytes);      // 16 * nthr * n_static
    double2 *s_dout = s_dtab + (size_t)nthr * n_static;                            // 16 * META_SLOTS * n_static
    uint32_t *s_runhdr = reinterpret_cast<uint32_t *>(s_dout + META_SLOTS * n_static);   // 8 * n_static: position and op index of each run's header
    // PTX op loop: the code bytes as planned; the copies inside s_ops are rewritten per tile (patch_codes)
    uint8_t *s_code0 = reinterpret_cast<uint8_t *>(s_runhdr + 2u * n_static);      // n_ops + n_stages
    constexpr bool PATCH = PTXOPS && !FULL;

    const uint32_t shard_shift = segs.shift;
    const uint64_t shard_mask = (1ull << shard_shift) - 1ull;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_mbar);
    // the pass program: same for every tile this CTA processes
    {
        // Stage s's descriptors sit at s_ops[op index + s]: one SENTINEL descriptor (code MOP_END) follows
        // every stage, so the op loop needs neither an end pointer nor a compare -- it jumps on the code.
        const uint4 *src = reinterpret_cast<const uint4 *>(g_ops + hdr.op_begin);
        uint4 *dst = reinterpret_cast<uint4 *>(s_ops);
        for (uint32_t i = tid; i < (MOP_BYTES / 16u) * n_ops; i += nthr) {
            const uint32_t o = i / (MOP_BYTES / 16u);
            uint32_t sh = 0;
            while (sh < n_stages && o >= hdr.stage_end[sh]) ++sh;
            dst[i + sh * (MOP_BYTES / 16u)] = src[i];
        }
        for (uint32_t st = tid; st < n_stages; st += nthr)
            dst[(hdr.stage_end[st] + st) * (MOP_BYTES / 16u)] = make_uint4(MOP_END, 0u, 0u, 0u);
        src = reinterpret_cast<const uint4 *>(g_stages + hdr.stage_begin);
        dst = reinterpret_cast<uint4 *>(s_stages);
        for (uint32_t i = tid; i < 2 * n_stages; i += nthr) dst[i] = src[i];
        if (FULL || hdr.need_flags) {
            src = reinterpret_cast<const uint4 *>(g_bases + hdr.op_begin);
            dst = reinterpret_cast<uint4 *>(s_bases);
            for (uint32_t i = tid; i < n_ops; i += nthr) dst[i] = src[i];
        }
        // Chunk base pointers for the tile at offset 0: chunk c = the 2^L amplitudes whose gathered
        // tile bits spell c.  Rank bits among them select the shard (own HBM or a peer's, mapped
        // over NVLink); every tile of the pass adds the same byte offset to all of them.
        for (uint32_t c = tid; c < n_chunks; c += nthr) {
            uint64_t gidx = hdr.base_or;
            for (uint32_t j = 0; j < T - L; ++j)
                if ((c >> j) & 1u) gidx |= 1ull << hdr.gpos[L + j];
            s_ptr0[c] = (unsigned long long)(uintptr_t)(segs.seg[gidx >> shard_shift] + (gidx & shard_mask));
        }
        if (BULK && tid == 0) mbar_init(mbar, 1u);
    }
    const uint32_t n_t = T - TR;                     // thread bits per stage
    if (tid < 16u) s_gpos[tid] = hdr.gpos_store[tid];      // where the last stage stores each tile-local bit
    __syncthreads();                                 // the stage descriptors are in shared memory
    // PTX op loop: ops with controls OUTSIDE the tile (class 2: code >= 2 * FC_TOTAL) are decided once
    // per tile for all threads, so the loop does not test them: before a tile's first stage the code
    // byte in shared memory becomes the class 0 / 1 code (controls satisfied) or a skip code.
    int mine_patch = 0;
    if (PATCH) {
        for (uint32_t o = tid; o < n_ops + n_stages; o += nthr) {
            const uint32_t c0 = s_ops[o].code;
            s_code0[o] = (uint8_t)c0;
            mine_patch |= (int)(c0 >= 2u * (uint32_t)FC_TOTAL && c0 < 3u * (uint32_t)FC_TOTAL);
        }
    }
    // per-(stage, thread) and per-stage constants, once per kernel
    unsigned long long my_goff = 0;                  // last stage: byte offset of this thread's jl inside the tile's span
    for (uint32_t st = 0; st < n_stages; ++st) {
        const uint32_t jl = stage_jl(s_stages[st], n_t, tid);
        s_jltab[st * nthr + tid] = (slot16(jl, L) << 16) | jl;
        if (tid < (uint32_t)TR)
            reinterpret_cast<uint32_t *>(s_ctab + st)[tid] = 16u * slot16(1u << s_stages[st].r_lpos[tid], L);
        if (st + 1 == n_stages) {
            unsigned long long go = 0;
            for (uint32_t l = 0; l < T; ++l)
                if ((jl >> l) & 1u) go += 16ull << hdr.gpos_store[l];
            my_goff = go;
        }
    }
    if (!FULL && n_static) {
        // factor of every tabulated run's thread-bit members for THIS thread (its group number is its index)
        for (uint32_t o = 0; o < n_ops + n_stages; ++o) {        // (shared-memory positions: sentinels included)
            const MOp &hd = s_ops[o];
            if (hd.code == (uint8_t)MOP_END) continue;
            const uint32_t base_code = fc_generic(hd.code) % (uint32_t)FC_TOTAL;
            if ((base_code != (uint32_t)FC_DM && base_code != (uint32_t)(FC_DM + FC_MASKED)) || !(hd.dagger & MOP_STATIC)) continue;
            double ar = 1.0, ai = 0.0;
            for (uint32_t k = 1; k <= hd.a_reg; ++k) {
                const MOp &m = s_ops[o + k];
                if (!m.a_thr) continue;
                const uint32_t par = (uint32_t)__popc(tid & m.a_thr) & 1u;
                if (!par && (m.dagger & MOP_SKIP0)) continue;
                const double fr = par ? m.c2 : m.ph_re, fi = par ? m.c3 : m.ph_im;
                const double t = ar * fr - ai * fi;
                ai = ar * fi + ai * fr;
                ar = t;
            }
            s_dtab[hd.a_thr * nthr + tid] = make_double2(ar, ai);
            if (tid == 0) {
                s_runhdr[2u * hd.a_thr] = o;                   // position in shared memory
                s_runhdr[2u * hd.a_thr + 1u] = hd.idx;         // op index (s_bases, flags)
            }
            o += hd.a_reg;
        }
    }
    const uint32_t tile_s0 = (uint32_t)__cvta_generic_to_shared(tile_b);
    uint32_t tile_s = tile_s0;                       // the buffer of the tile being computed (DB: alternates)
    const unsigned long long shard_base = (unsigned long long)(uintptr_t)segs.seg[segs.rank];
    const uint32_t ops_s = (uint32_t)__cvta_generic_to_shared(s_ops);
    const uint32_t flags_all_s = (uint32_t)__cvta_generic_to_shared(s_flags_all);
    const uint32_t ptr0_s = (uint32_t)__cvta_generic_to_shared(s_ptr0);
    const bool active = tid < (1u << n_t);           // this thread owns a group of 16 amplitudes
    const bool need_flags = FULL || hdr.need_flags != 0u;
    // Remap pass: the last stage writes both halves of the tile into THIS shard (the peer half at the
    // pinned bit's other value), i.e. into the place the peer's CTA of the same tile loads ITS peer
    // half from.  Handshake per tile: after its load has landed a CTA stores the pass's epoch into
    // the peer's ack word of that tile; before its final stores it waits for its own ack word.  Both
    // GPUs walk the tiles in the same order with the same grid, and a load never waits, so the wait
    // always ends.  Both sides of the handshake are RELAXED system-scope accesses: the ack is sent after
    // the tile has landed in shared memory (the reads of the peer's copy are over), and the stores it
    // guards are issued behind a barrier that follows the polling load -- nothing else needs ordering.
    // (A release store here is a MEMBAR.SYS per tile on thread 0 with the whole CTA waiting at the
    // stage's barrier: the in-kernel timeline showed ~20k of a remap tile's 44k cycles in it.)
    const bool remap = hdr.remap != 0u;
    const unsigned long long store_keep = remap ? ~(16ull << hdr.remap_b) : ~0ull;
    unsigned int *const ack_mine = segs.ack[segs.rank];
    unsigned int *const ack_peer = remap ? segs.ack[segs.rank ^ (1u << (hdr.remap_g - segs.shift))] : nullptr;
    const bool has_patch = __syncthreads_or(mine_patch) != 0;
    auto patch_codes = [&](const uint32_t slot) {
        const uint8_t *flags = s_flags_all + slot * flags_stride;
        for (uint32_t o = tid; o < n_ops + n_stages; o += nthr) {
            const uint32_t c0 = s_code0[o];
            if (c0 < 2u * (uint32_t)FC_TOTAL || c0 >= 3u * (uint32_t)FC_TOTAL) continue;
            MOp &m = s_ops[o];
            const uint32_t arm = c0 - 2u * (uint32_t)FC_TOTAL;
            uint32_t c;
            if (flags[m.idx] & 0x80u) c = arm + (m.ctrl_thr ? (uint32_t)FC_TOTAL : 0u);
            else c = (arm == (uint32_t)FC_DM || arm == (uint32_t)(FC_DM + FC_MASKED)) ? MOP_NOP_RUN : MOP_NOP;
            m.code = (uint8_t)c;
        }
    };

    // Tile counter -> local base index (tile bits and ownership bits clear): the counter's bits
    // are spread over the positions hdr.fixed_mask leaves free.  The CTA's first tile is expanded
    // run by run; every further one is a masked addition (carries ripple through the fixed
    // positions because they are set to 1 for the addition).
    auto expand = [&](uint64_t b) -> uint64_t {
        for (uint32_t k = 0; k < hdr.n_runs; ++k) {
            const uint32_t p = hdr.run_pos[k], len = hdr.run_len[k];
            b = ((b >> p) << (p + len)) | (b & ((1ull << p) - 1ull));
        }
        return b;
    };
    const uint64_t fixed = hdr.fixed_mask;
    const uint64_t step = expand(gridDim.x);
    // Metadata of the first tile at or after (t, base), stepping by the grid, that some op of this
    // pass can change: per-op flags (bit 7: controls outside the tile satisfied; bits 0-2:
    // popcount of the diagonal target mask over the bits outside the tile, mod 8) and the tile's
    // byte offset inside the shard.  Tiles no op touches (multi-controlled gates) are skipped
    // without being read.  Passes whose ops do not look outside the tile skip all of this.
    auto prepare = [&](uint64_t &t, uint64_t &base, uint32_t slot, unsigned long long &toff) {
        uint8_t *flags = s_flags_all + slot * flags_stride;
        for (; t < hdr.n_tiles; t += gridDim.x, base = ((base | fixed) + step) & ~fixed) {
            const uint64_t lb = base | hdr.fx_val;
            toff = lb * 16ull;
            if (!need_flags) return;
            const uint64_t gb = lb | hdr.base_or;
            int any = 0;
            for (uint32_t o = tid; o < n_ops; o += nthr) {
                const MBase b = s_bases[o];
                const uint32_t okb = ((~gb & b.ctrl_base) == 0) ? 0x80u : 0u;
                flags[o] = (uint8_t)(okb | ((uint32_t)__popcll(gb & b.a_base) & 7u));
                any |= (int)okb;
            }
            if (!FULL && n_static) {
                // this tile's factor of every tabulated run's members OUTSIDE the tile: one warp per run,
                // one member per lane, product by shuffles (a serial walk by one thread cost a third of a
                // QFT pass: every other thread waits for it at the barrier below)
                const uint32_t lane = tid & 31u, nw = nthr >> 5;
                for (uint32_t r = tid >> 5; r < n_static; r += nw) {
                    const uint32_t o = s_runhdr[2u * r], ob = s_runhdr[2u * r + 1u];
                    const MOp &hd = s_ops[o];
                    if (!(hd.dagger & MOP_PARB)) continue;
                    double ar = 1.0, ai = 0.0;
                    for (uint32_t k = 1 + lane; k <= hd.a_reg; k += 32u) {
                        const uint64_t ab = s_bases[ob + k].a_base;
                        if (!ab) continue;
                        const MOp &m = s_ops[o + k];
                        const uint32_t par = (uint32_t)__popcll(gb & ab) & 1u;
                        if (!par && (m.dagger & MOP_SKIP0)) continue;
                        const double fr = par ? m.c2 : m.ph_re, fi = par ? m.c3 : m.ph_im;
                        const double t = ar * fr - ai * fi;
                        ai = ar * fi + ai * fr;
                        ar = t;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double br = __shfl_down_sync(0xffffffffu, ar, off), bi = __shfl_down_sync(0xffffffffu, ai, off);
                        const double t = ar * br - ai * bi;
                        ai = ar * bi + ai * br;
                        ar = t;
                    }
                    if (lane == 0) s_dout[slot * n_static + r] = make_double2(ar, ai);
                }
            }
            if (remap) any = 1;                          // every tile moves
            if (__syncthreads_or(any)) return;
        }
    };
    // BULK: one cp.async.bulk per chunk, all completing on the mbarrier; the copies are issued by lane
    // 0 of every warp (a warp's lanes would be serialised anyway: UBLKCP takes uniform operands), so
    // no warp is held up longer than n_chunks / warps issues.
    // !BULK: one 16-byte cp.async per amplitude; element j = tid + i * nthr, so with nthr a multiple
    // of the chunk length every iteration advances the chunk by nthr >> L and keeps the offset.
    const bool regular = (nthr & lmask) == 0u && (tile_len % nthr) == 0u;
    auto issue_load = [&](const unsigned long long toff, const uint32_t tile_s) {
        if (BULK) {
            if ((tid & 31u) == 0u) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the buffer was last touched by ld/st.shared
                    mbar_expect_tx(mbar, 16u << T);
                }
                const uint32_t nw = nthr >> 5;
                for (uint32_t c = tid >> 5; c < n_chunks; c += nw)
                    bulk_g2s(tile_s + c * chunk_stride, s_ptr0[c] + toff, chunk_bytes, mbar);
            }
        } else if (regular) {
            const unsigned long long mine = toff + (unsigned long long)(tid & lmask) * 16ull;
            const uint32_t cstep = (nthr >> L) * 8u, dstep = (nthr >> L) * chunk_stride;
            uint32_t ch = ptr0_s + (tid >> L) * 8u, d = tile_s + 16u * slot16(tid, L);
#pragma unroll 4
            for (uint32_t jb = 0; jb < tile_len; jb += nthr, ch += cstep, d += dstep) {
                unsigned long long p;
                asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(p) : "r"(ch) : "memory");
                p += mine;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(p) : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        } else {
            for (uint32_t j = tid; j < tile_len; j += nthr) {
                const unsigned long long p = s_ptr0[j >> L] + toff + (unsigned long long)(j & lmask) * 16ull;
                const uint32_t d = tile_s + 16u * slot16(j, L);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(p) : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
    };

    // The CTA has ONE tile buffer, so the shared-memory load of tile i+1 cannot start before the last
    // stage of tile i; to keep the HBM latency off that path the lines of tile i+1 are prefetched
    // into L2 a whole tile earlier (hdr.prefetch), and the late load finds them there.
    auto prefetch_tile = [&](const unsigned long long toff) {
        // one 128-byte line per thread and step: two instructions per 256 bytes, no registers held
        const uint32_t lines_per_chunk = chunk_bytes >> 7;
        if (lines_per_chunk == 0) return;
        const uint32_t n_lines = n_chunks * lines_per_chunk, lsh = L - 3u;      // chunk = line >> (L - 3)
        for (uint32_t ln = tid; ln < n_lines; ln += nthr) {
            const unsigned long long p = s_ptr0[ln >> lsh] + toff + (unsigned long long)(ln & (lines_per_chunk - 1u)) * 128ull;
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
        }
    };

    uint32_t mslot = 0, phase = 0;
    unsigned long long toff_cur = 0, toff_next = 0;
    uint64_t t_cur = blockIdx.x, base_cur = expand(blockIdx.x);
    prepare(t_cur, base_cur, mslot, toff_cur);
    if (t_cur < hdr.n_tiles) issue_load(toff_cur, tile_s);
#ifdef QV_TRACE
    uint32_t tr_i = 0;
    if (tid == 0 && blockIdx.x == 0 && (hdr.prefetch & 0x80000000u)) {
        unsigned int *inf = g_trace_info[hdr.remap ? 1 : 0];
        inf[0] = n_stages; inf[1] = n_ops; inf[2] = T; inf[3] = L; inf[4] = nthr; inf[5] = gridDim.x;
        inf[6] = (unsigned int)hdr.n_tiles; inf[7] = (BULK ? 1u : 0u) | (DB ? 2u : 0u) | (hdr.touches_peer ? 4u : 0u);
    }
#endif
    while (t_cur < hdr.n_tiles) {
        QV_STAMP(0);
        const uint32_t mnext = mslot + 1u == META_SLOTS ? 0u : mslot + 1u;
        uint64_t t_next = t_cur + gridDim.x, base_next = ((base_cur | fixed) + step) & ~fixed;
        if (BULK) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            if (need_flags || DB) __syncthreads();     // this tile's flag bytes; DB: the other buffer is free
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
        }
        QV_STAMP(1);
        // (the barrier above also says: nobody is still running the previous tile's ops)
        if (has_patch) patch_codes(mslot);
        prepare(t_next, base_next, mnext, toff_next);  // need_flags: ends in a barrier when there is a next tile
        const bool has_next = t_next < hdr.n_tiles;
        if (has_patch && !has_next) __syncthreads();
        if (DB && has_next) {
            // Two buffers: everybody is done with the previous tile (the barrier above), so its
            // buffer takes the next tile NOW -- a whole tile of compute ahead of its use.
            issue_load(toff_next, tile_s ^ tile_s0 ^ (tile_s0 + buf_bytes));
        }
        QV_STAMP(2);
        const uint8_t *flags = s_flags_all + mslot * flags_stride;
        const uint32_t flags_s = flags_all_s + mslot * flags_stride;
        if (remap && tid == 0)       // "I have read your copy of tile t_cur"
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(ack_peer + t_cur), "r"(hdr.epoch) : "memory");
        if ((hdr.prefetch & 1u) && has_next) prefetch_tile(toff_next);

        for (uint32_t s = 0; s < n_stages; ++s) {
            const bool last = s + 1 == n_stages;
            const uint32_t ob = s ? hdr.stage_end[s - 1] : 0u, oe = hdr.stage_end[s];
            amp v[NV];
            StageCtx x;
            if (active) {
                x = stage_ctx(s_stages + s, s_jltab + (s * nthr + tid), s_ctab + s);
                if (last) x.goff = my_goff;
                stage_load(tile_s, x, v);
            }
            if (last && DB) {
                if (remap) {
                    if (tid == 0) {
                        unsigned int got;
                        unsigned long long spins = 0;
                        do {
                            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(got) : "l"(ack_mine + t_cur) : "memory");
                            if (got != hdr.epoch && ++spins > (1ull << 24)) {
                                __nanosleep(1000);
                                if (spins > (1ull << 24) + 20000000ull) __trap();
                            }
                        } while (got != hdr.epoch);
                    }
                    __syncthreads();
                }
            } else if (last) {
                if (remap && tid == 0) {               // the peer has read what the stores below overwrite
                    // (bounded: a peer kernel that never runs -- shards sharing a GPU, a dead peer -- must
                    // surface as an error, not as a hung GPU)
                    unsigned int got;
                    unsigned long long spins = 0;
                    do {
                        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(got) : "l"(ack_mine + t_cur) : "memory");
                        if (got != hdr.epoch && ++spins > (1ull << 24)) {
                            __nanosleep(1000);
                            if (spins > (1ull << 24) + 20000000ull) __trap();      // ~20 s
                        }
                    } while (got != hdr.epoch);
                }
                __syncthreads();                       // every thread holds its amplitudes: the buffer is free
                if (has_next) issue_load(toff_next, tile_s);
            } else if (s_stages[s].sync_after_load) {
                // a lazy x makes threads store into each other's slots of the tile buffer: nobody
                // may store before everybody has loaded
                __syncthreads();
            }
            if (last) QV_STAMP(5);
            if (active) {
                uint32_t inv = 0;
                if (FULL) stage_ops_full(ops_s + MOP_BYTES * s, flags_s, ob, oe, tid, mats, v);
                else if (PTXOPS)
                    stage_ops_fast_ptx<SC>(ops_s + MOP_BYTES * s, flags_s, ob, oe, tid, x, inv, v,
                                       (uint32_t)__cvta_generic_to_shared(s_dtab + tid), 16u * nthr,
                                       (uint32_t)__cvta_generic_to_shared(s_dout + mslot * n_static));
                else
                    stage_ops_fast(reinterpret_cast<const unsigned char *>(s_ops + s), flags, ob, oe, tid, L, x, inv, v,
                                   s_dtab + tid, nthr, s_dout + mslot * n_static);
                if (last) QV_STAMP(6);
                if (!last) stage_store_smem(tile_s, x, inv, v);
                else if (hdr.touches_peer && !remap) stage_store_global(ptr0_s, L, toff_cur, x, inv, v);
                else stage_store_global_local(shard_base + (toff_cur & store_keep), s_gpos, x, inv, v);
            }
            if (!last) __syncthreads();
            if (!last && s < 2u) QV_STAMP(3 + s);
        }
        QV_STAMP(7);
#ifdef QV_TRACE
        ++tr_i;
#endif
        mslot = mnext;
        if (DB) tile_s = tile_s ^ tile_s0 ^ (tile_s0 + buf_bytes);
        t_cur = t_next;
        base_cur = base_next;
        toff_cur = toff_next;
    }
    // Stores into a peer's HBM must be performed at SYSTEM scope before the barrier kernel that
    // follows signals the peers: the barrier's own fence is executed by other threads and does not
    // cover them.
    if (hdr.touches_peer && !remap) __threadfence_system();
}

constexpr size_t TILE_SMEM_MAX = 227u * 1024u;

static size_t tile_smem_bytes(const TPassHdr &h, int threads, bool db = false) {
    const bool flags = h.full || h.need_flags;
    return (db ? 2 : 1) * (((size_t)16 << h.T) + ((size_t)16 << (h.T - h.L))) + 16 + (size_t)MOP_BYTES * (h.n_ops + h.n_stages) +
           (flags ? sizeof(MBase) * h.n_ops : 0) + (size_t)32 * h.n_stages + (((size_t)8 << (h.T - h.L)) + 8) +
           (size_t)16 * h.n_stages + (size_t)4 * threads * h.n_stages + 16 +
           (flags ? (size_t)META_SLOTS * ((h.n_ops + 15u) & ~15u) : 16) +
           (size_t)16 * h.n_static * ((size_t)threads + META_SLOTS) + (size_t)8 * h.n_static +
           (((size_t)h.n_ops + h.n_stages + 15u) & ~(size_t)15u);
}

typedef void (*tile_kernel_t)(const Segs, const TPassHdr, const TStage *, const MOp *, const MBase *, const amp *);

template <int THREADS, int MINB>
static tile_kernel_t pick_kernel(bool full, bool bulk, bool ptx) {
    if (full) return bulk ? k_tile_pass<THREADS, MINB, true, true, false, false> : k_tile_pass<THREADS, MINB, true, false, false, false>;
    if (ptx) return bulk ? k_tile_pass<THREADS, MINB, false, true, true, false> : k_tile_pass<THREADS, MINB, false, false, true, false>;
    return bulk ? k_tile_pass<THREADS, MINB, false, true, false, false> : k_tile_pass<THREADS, MINB, false, false, false, false>;
}

static tile_kernel_t select_kernel(int threads, int ctas, bool full, bool bulk, bool ptx, bool db, bool sc = false) {
    // passes with single-control ops: the flavour of the PTX op loop that has those arms (3 CTAs per SM only)
    if (sc && !full && ptx) {
        if (threads == 256)
            return bulk ? k_tile_pass<256, 2, false, true, true, false, true> : k_tile_pass<256, 2, false, false, true, false, true>;
        if (db) return bulk ? k_tile_pass<128, 3, false, true, true, true, true> : k_tile_pass<128, 3, false, false, true, true, true>;
        return bulk ? k_tile_pass<128, 3, false, true, true, false, true> : k_tile_pass<128, 3, false, false, true, false, true>;
    }
    // two tile buffers: the fast path with the PTX op loop, 3 CTAs of 128 threads per SM
    if (db && threads == 128 && !full && ptx)
        return bulk ? k_tile_pass<128, 3, false, true, true, true> : k_tile_pass<128, 3, false, false, true, true>;
    if (threads == 256) return pick_kernel<256, 2>(full, bulk, ptx);
    if (ctas == 4) return pick_kernel<128, 4>(full, bulk, ptx);
    if (ctas == 5 && !full && !bulk && ptx) return k_tile_pass<128, 5, false, false, true, false>;
    return pick_kernel<128, 3>(full, bulk, ptx);
}

// cudaFuncSetAttribute is per device and costs a driver call per kernel: once per process and device.
int tile_kernel_setup() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return 0;
    bool ok = true;
    for (int full = 0; full < 2; ++full)
        for (int bulk = 0; bulk < 2; ++bulk)
            for (int ptx = 0; ptx < 2; ++ptx)
                for (int cfg = 0; cfg < 5; ++cfg)
                    for (int sc = 0; sc < 2; ++sc) {
                        const tile_kernel_t k = select_kernel(cfg == 0 ? 256 : 128, cfg == 1 ? 3 : cfg == 4 ? 5 : 4, full, bulk,
                                                              ptx, cfg == 3, sc != 0);
                        ok = ok && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)TILE_SMEM_MAX) == cudaSuccess;
                    }
    done[dev] = ok;
    return ok ? 0 : -1;
}

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const MOp *d_ops, const MBase *d_bases, const amp *mat_table, int sm_count,
                     const TileKnobs &knobs) {
    if (hdr.T < TILE_MIN_BITS || hdr.T > TILE_MAX_BITS || hdr.L > hdr.T || hdr.T - hdr.L > TILE_MAX_HIGH ||
        hdr.n_tiles == 0 || (hdr.n_ops == 0 && !hdr.remap) || hdr.n_ops > (uint32_t)TILE_MAX_OPS || hdr.n_stages == 0 ||
        hdr.n_stages > (uint32_t)TILE_MAX_STAGES)
        return -1;
    int threads;
    if (hdr.T >= 12) {
        threads = 256;
    } else {
        threads = 1 << (hdr.T - TILE_R);
        if (threads < 32) threads = 32;
        if (threads > 128) threads = 128;
    }
    // Two tile buffers (the next tile is loaded a whole tile of compute ahead): measured no gain, for
    // local and for peer passes alike -- the exposed wait is not the load's latency -- so only on request.
    const bool bulk = knobs.bulk > 0 || (knobs.bulk < 0 && hdr.touches_peer);
    const bool want_db = knobs.double_buffer == 1 || (knobs.double_buffer == 2 && hdr.touches_peer);
    bool db = want_db && threads == 128 && hdr.T == 11 && !hdr.full && knobs.ptx_ops &&
              2 * (tile_smem_bytes(hdr, threads, true) + 1024) <= TILE_SMEM_MAX;
    const tile_kernel_t kern = select_kernel(threads == 256 ? 256 : 128, knobs.ctas_per_sm, hdr.full != 0, bulk,
                                             knobs.ptx_ops != 0, db, hdr.uses_sc != 0);
    const size_t smem = tile_smem_bytes(hdr, threads, db);
    if (smem > TILE_SMEM_MAX) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    TPassHdr h2 = hdr;
    h2.prefetch = knobs.prefetch ? 1u : 0u;
#ifdef QV_TRACE
    {   // the probe keeps the pass with the most ops of each kind (local / remap)
        static uint32_t best[2] = {0u, 0u};
        const int kind = hdr.remap ? 1 : 0;
        if (hdr.n_ops >= best[kind]) {
            best[kind] = hdr.n_ops;
            h2.prefetch |= 0x80000000u;
        }
    }
#endif
    kern<<<(unsigned)grid, threads, smem, st>>>(segs, h2, d_stages, d_ops, d_bases, mat_table);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace qv

#ifdef QV_TRACE
// debug builds only: copies device `dev`'s probe (kind 0: last local pass, 1: last remap pass) to the host
extern "C" int qvnt_debug_trace(int dev, int kind, unsigned long long *stamps, unsigned int *info) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(dev);
    cudaDeviceSynchronize();
    const size_t n = sizeof(unsigned long long) * qv::QV_TRACE_CTAS * qv::QV_TRACE_TILES * qv::QV_TRACE_PTS;
    cudaError_t e = cudaMemcpyFromSymbol(stamps, qv::g_trace, n, n * (size_t)kind);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(info, qv::g_trace_info, 32, 32 * (size_t)kind);
    cudaSetDevice(cur);
    return e == cudaSuccess ? 0 : -1;
}
#endif
