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
// Layout of a tile: T tile bits = low L bits (contiguous 16*2^L-byte chunks, loaded with
// coalesced 16-byte cp.async) + T-L gathered high bits.  A gathered bit may be a rank bit
// of a sharded register: that chunk is then read from / written to the peer GPU's HBM
// through its NVLink-mapped pointer (Segs), so a global-qubit gate needs no separate
// exchange step -- the "swap" is the tile's own load and store.
//
// Shared memory is XOR-swizzled at 16-byte granularity (slot ^= fold of the upper index
// bits into the low 3) so that, whichever bits a stage keeps in registers, the 8 lanes
// of a quarter-warp hit 8 different 16-byte bank groups (the planner picks the lane bits).
// The swizzle is GF(2)-linear, so a thread's 16 slot addresses are one base XOR 16
// stage-uniform constants.
//
// Stages.  A stage keeps TILE_R = 4 tile bits in "register slots": every thread holds the 16
// amplitudes that differ in those bits, runs all of the stage's ops on them and only then touches
// shared memory again.  The last stage of a pass stores its registers straight to HBM, so the tile
// buffer is free as soon as that stage has read it and the next tile's cp.async loads are issued
// there.  An op is a 48-byte MOp in shared memory whose `code` selects a fully unrolled arm;
// everything that depends on the amplitude index was split by the planner into
//   - register-slot part: compile-time per arm, or a 16-bit `okmask` (controls in register slots),
//   - thread part: one AND/compare on the thread's group number per op,
//   - tile part: one flag byte per op per tile (controls / diagonal-mask bits outside the
//     tile), computed once per tile.
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
// buffer); (2) before the next tile's loads overwrite the buffer (after the last stage's loads);
// (3) between a stage's loads and stores when the stage holds a lazy x, because threads then store
// into EACH OTHER's slots (TStage::sync_after_load); (4) across GPUs, dist_barrier() kernels
// around every pass that touches a peer shard, plus a system fence after the peer stores.
#include "engine.h"
#include "gates.cuh"

namespace qv {

constexpr int NV = TILE_NV;
constexpr int TR = TILE_R;

__device__ __forceinline__ uint32_t swz(uint32_t j) {
    return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9)) & 7u);
}

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

// ---- op bodies: templated on kind and register slot(s), fully unrolled over the 16 slots -----
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

// ALL: no control sits in a register slot (okmask = 0xFFFF) -- the common case runs without
// any per-slot predicate.
// One decoded op: the 32-byte MOp as it sits in shared memory (two 16-byte loads).
struct MDec {
    uint32_t w0;        // code | dagger << 8 | okmask << 16
    uint32_t ctrl_thr;
    uint32_t a_thr;
    uint32_t a_reg;     // low 16 bits
    double ph_re, ph_im;
};

// FULL: the interpreter carries the arms of every kind.  Passes made only of the common kinds
// (diagonal class, x, y, rx, ry, h1) run the lean instance instead: the rare arms (u1/u2 keep a
// whole matrix live, the two-bit ops have the most temporaries) set the register pressure of the
// whole op loop, and without them its loop state stays in registers.
template <bool ALL, bool FULL>
__device__ __forceinline__ void apply_mop(const MDec &m, const uint32_t fl, const uint32_t grp,
                                          const amp *__restrict__ mats, amp (&v)[NV]) {
    GP g;
    g.c = m.ph_re;
    g.s = m.ph_im;
    g.dagger = (m.w0 >> 8) & 0xFFu;
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
    if (!FULL) return;
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
// jl ^ (subset of the 4 register-slot bits).  All shared-memory addresses are 32-bit
// shared-window offsets, XOR-swizzled (swz is GF(2)-linear: address(K) = mine_o ^ c[..]).
struct StageCtx {
    uint32_t r_lpos;    // 4 bytes: register slot j -> tile-local bit position
    uint32_t c[TR];     // swizzled byte offsets of the 4 register-slot bits
    uint32_t jl;        // tile-local index of slot pattern 0
    uint32_t mine_o;    // its swizzled byte offset
    unsigned long long goff;   // last stage, local tiles: byte offset of jl inside the tile's span of the shard
};

// Per-(stage, thread) constants are computed ONCE per kernel (every tile of the pass runs the same
// stages) and kept in shared memory: word = (16 * swz(jl)) << 16 | jl.
__device__ __forceinline__ uint32_t stage_jl(const uint32_t stage_s, const uint32_t n_t, const uint32_t grp) {
    uint32_t sw[8];      // the TStage: op_begin, op_end, r_lpos[4], t_lpos[16], pad
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(sw[0]), "=r"(sw[1]), "=r"(sw[2]), "=r"(sw[3]) : "r"(stage_s) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(sw[4]), "=r"(sw[5]), "=r"(sw[6]), "=r"(sw[7]) : "r"(stage_s + 16u) : "memory");
    uint32_t jl = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) {
        if (k < n_t) jl |= ((grp >> k) & 1u) << ((sw[3 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu);
    }
    return jl;
}

__device__ __forceinline__ StageCtx stage_ctx(const uint32_t stage_s, const uint32_t jltab_s, const uint32_t ctab_s) {
    StageCtx x;
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(x.r_lpos) : "r"(stage_s + 8u) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(x.c[0]), "=r"(x.c[1]), "=r"(x.c[2]), "=r"(x.c[3]) : "r"(ctab_s) : "memory");
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(w) : "r"(jltab_s) : "memory");
    x.jl = w & 0xFFFFu;
    x.mine_o = w >> 16;
    x.goff = 0;
    return x;
}

__device__ __forceinline__ void stage_load(const uint32_t tile_s, const StageCtx &x, amp (&v)[NV]) {
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t a = x.mine_o;
        if (K & 1) a ^= x.c[0];
        if (K & 2) a ^= x.c[1];
        if (K & 4) a ^= x.c[2];
        if (K & 8) a ^= x.c[3];
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v[K].x), "=d"(v[K].y) : "r"(tile_s + a) : "memory");
    }
}

__device__ __forceinline__ void stage_store_smem(const uint32_t tile_s, const StageCtx &x, const amp (&v)[NV]) {
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t a = x.mine_o;
        if (K & 1) a ^= x.c[0];
        if (K & 2) a ^= x.c[1];
        if (K & 4) a ^= x.c[2];
        if (K & 8) a ^= x.c[3];
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(tile_s + a), "d"(v[K].x), "d"(v[K].y) : "memory");
    }
}

// Last stage of a pass: the 16 amplitudes go straight from registers to HBM (or to the peer's
// HBM): address = chunk base pointer (tile-invariant table) + this tile's byte offset + offset
// inside the chunk.  The planner makes the low lane bits the lowest tile-local bits that are not
// register slots, so a warp's store instruction covers up to 512 contiguous bytes.
__device__ __forceinline__ void stage_store_global(const uint32_t ptr0_s, const uint32_t L,
                                                   const unsigned long long toff, const StageCtx &x,
                                                   const amp (&v)[NV]) {
    uint32_t offc[TR], chc[TR];       // per register slot: its bit inside the chunk / in the chunk index
#pragma unroll
    for (int j = 0; j < TR; ++j) {
        const uint32_t lp = (x.r_lpos >> (8 * j)) & 0xFFu;
        offc[j] = lp < L ? (16u << lp) : 0u;
        chc[j] = lp < L ? 0u : (8u << (lp - L));
    }
    const unsigned long long off0 = toff + (x.jl & ((1u << L) - 1u)) * 16u;
    const uint32_t ch0 = ptr0_s + (x.jl >> L) * 8u;
#pragma unroll
    for (int K = 0; K < NV; ++K) {
        uint32_t off = 0, ch = ch0;
        if (K & 1) { off |= offc[0]; ch += chc[0]; }      // (ch is an address: add, the table is not
        if (K & 2) { off |= offc[1]; ch += chc[1]; }      //  aligned to its own size)
        if (K & 4) { off |= offc[2]; ch += chc[2]; }
        if (K & 8) { off |= offc[3]; ch += chc[3]; }
        unsigned long long base;
        asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(base) : "r"(ch) : "memory");
        *reinterpret_cast<amp *>(base + off0 + off) = v[K];
    }
}

// Same, for tiles that lie entirely in this GPU's shard: the address is affine in the index bits,
// address(K) = base + goff(thread) + sum of the byte offsets of K's register-slot bits -- no table.
__device__ __forceinline__ void stage_store_global_local(const unsigned long long base, const uint32_t gpos_s,
                                                         const StageCtx &x, const amp (&v)[NV]) {
    unsigned long long g[TR];
#pragma unroll
    for (int j = 0; j < TR; ++j) {
        uint32_t gp;
        asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(gp) : "r"(gpos_s + ((x.r_lpos >> (8 * j)) & 0xFFu)) : "memory");
        g[j] = 16ull << gp;
    }
    const unsigned long long a0 = base + x.goff;
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
        if ((m.w0 >> 16) == 0xFFFFu) apply_mop<true, true>(m, fl, grp, mats, v);
        else apply_mop<false, true>(m, fl, grp, mats, v);
        if (o + 1 < oe) fetch(o + 1);
    }
}

// ================================================================================================
// FAST stage interpreter.  Passes made only of the common kinds are lowered by the planner to the
// coefficient-driven forms of FCode (engine.h); every arm below updates the 16 register-resident
// amplitudes IN PLACE through inline PTX whose operands are tied ("+d"), so each amplitude keeps
// one home register through the whole op loop: no register shuffling at the merge points of the
// dispatch, no selects -- per op the instruction stream is the FP64 arithmetic plus ~15
// instructions of fetch / test / dispatch.  4 FP64 instructions per amplitude and op
// (2 mul + 2 fma), the same count as the reference formulas with contraction.
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
// new0 = (p0 + p1) * s ; new1 = (p0 - p1) * s
__device__ __forceinline__ void f_pair_addsub(amp &p0, amp &p1, double sc) {
    asm volatile("{\n\t.reg .f64 t, u, w, z;\n\t"
                 "add.rn.f64 t, %0, %2;\n\t"
                 "sub.rn.f64 u, %0, %2;\n\t"
                 "add.rn.f64 w, %1, %3;\n\t"
                 "sub.rn.f64 z, %1, %3;\n\t"
                 "mul.rn.f64 %0, t, %4;\n\t"
                 "mul.rn.f64 %2, u, %4;\n\t"
                 "mul.rn.f64 %1, w, %4;\n\t"
                 "mul.rn.f64 %3, z, %4;\n\t}"
                 : "+d"(p0.x), "+d"(p0.y), "+d"(p1.x), "+d"(p1.y)
                 : "d"(sc));
}
// new0 = a*p0 - i*b*p1 ; new1 = -i*c*p0 + d*p1   (nb = -b, nc = -c)
__device__ __forceinline__ void f_pair_cross(amp &p0, amp &p1, double a, double b, double nb, double c, double nc,
                                             double d) {
    asm volatile("{\n\t.reg .f64 t, u, w, z;\n\t"
                 "mul.rn.f64 t, %4, %0;\n\t"      // a * p0.x
                 "mul.rn.f64 w, %4, %1;\n\t"      // a * p0.y
                 "mul.rn.f64 u, %7, %1;\n\t"      // c * p0.y
                 "mul.rn.f64 z, %8, %0;\n\t"      // -c * p0.x
                 "fma.rn.f64 %0, %5, %3, t;\n\t"  // p0.x = b * p1.y + a * p0.x
                 "fma.rn.f64 %1, %6, %2, w;\n\t"  // p0.y = -b * p1.x + a * p0.y
                 "fma.rn.f64 %2, %9, %2, u;\n\t"  // p1.x = d * p1.x + c * p0.y
                 "fma.rn.f64 %3, %9, %3, z;\n\t}" // p1.y = d * p1.y - c * p0.x
                 : "+d"(p0.x), "+d"(p0.y), "+d"(p1.x), "+d"(p1.y)
                 : "d"(a), "d"(b), "d"(nb), "d"(c), "d"(nc), "d"(d));
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
// p *= (fr, fi)   (nfi = -fi)
__device__ __forceinline__ void f_cmul(amp &p, double fr, double fi, double nfi) {
    asm volatile("{\n\t.reg .f64 t, u;\n\t"
                 "mul.rn.f64 t, %4, %1;\n\t"      // -fi * y
                 "mul.rn.f64 u, %3, %0;\n\t"      //  fi * x
                 "fma.rn.f64 %0, %2, %0, t;\n\t"  // x = fr * x - fi * y
                 "fma.rn.f64 %1, %2, %1, u;\n\t}" // y = fr * y + fi * x
                 : "+d"(p.x), "+d"(p.y)
                 : "d"(fr), "d"(fi), "d"(nfi));
}

struct FDec {
    uint32_t w0;        // code | flags << 8 | okmask << 16
    uint32_t ctrl_thr;
    uint32_t a_thr;
    uint32_t a_reg;     // low 16 bits
    double c0, c1, c2, c3;
};

template <int RB, bool ALL>
__device__ __forceinline__ void farm_pr(amp (&v)[NV], const FDec &m, const uint32_t ok) {
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_pair_real(v[K], v[K | (1 << RB)], m.c0, m.c1, m.c2, m.c3);
    }
}
template <int RB, bool ALL>
__device__ __forceinline__ void farm_pa(amp (&v)[NV], const FDec &m, const uint32_t ok) {
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_pair_addsub(v[K], v[K | (1 << RB)], m.c0);
    }
}
template <int RB, bool ALL>
__device__ __forceinline__ void farm_px(amp (&v)[NV], const FDec &m, const uint32_t ok) {
    const double nb = -m.c1, nc = -m.c2;
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_pair_cross(v[K], v[K | (1 << RB)], m.c0, m.c1, nb, m.c2, nc, m.c3);
    }
}
template <int RB, bool ALL>
__device__ __forceinline__ void farm_sw(amp (&v)[NV], const uint32_t ok) {
    QV_FOR_K {
        if (K & (1 << RB)) continue;
        if (ALL || ((ok >> K) & 1u)) f_swap(v[K], v[K | (1 << RB)]);
    }
}
// diagonal, no target bit in a register slot: one factor for the whole thread
template <bool ALL>
__device__ __forceinline__ void farm_du(amp (&v)[NV], const FDec &m, const uint32_t ok, const uint32_t par) {
    if (!par && (m.w0 & ((uint32_t)MOP_SKIP0 << 8))) return;
    const double fr = par ? m.c2 : m.c0, fi = par ? m.c3 : m.c1, nfi = -fi;
    QV_FOR_K {
        if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], fr, fi, nfi);
    }
}
// diagonal, exactly one target bit in register slot RB (+ parity `par` of the target bits elsewhere)
template <int RB, bool ALL>
__device__ __forceinline__ void farm_ds(amp (&v)[NV], const FDec &m, const uint32_t ok, const uint32_t par) {
    double f0r = m.c0, f0i = m.c1, f1r = m.c2, f1i = m.c3;
    bool skip0 = (m.w0 & ((uint32_t)MOP_SKIP0 << 8)) != 0;
    if (par) {                    // rzz with its second bit outside the slots: the roles swap
        f0r = m.c2; f0i = m.c3; f1r = m.c0; f1i = m.c1;
        skip0 = false;
    }
    const double n0 = -f0i, n1 = -f1i;
    if (!skip0) {
        QV_FOR_K {
            if (K & (1 << RB)) continue;
            if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], f0r, f0i, n0);
        }
    }
    QV_FOR_K {
        if (!(K & (1 << RB))) continue;
        if (ALL || ((ok >> K) & 1u)) f_cmul(v[K], f1r, f1i, n1);
    }
}
// diagonal, any set of target bits in register slots (rare: rzz with both bits in slots)
__device__ __forceinline__ void farm_dg(amp (&v)[NV], const FDec &m, const uint32_t ok, const uint32_t par) {
    const uint32_t a_reg = m.a_reg & 0xFu;
    const double n0 = -m.c1, n1 = -m.c3;
    QV_FOR_K {
        if ((ok >> K) & 1u) {
            if ((__popc((uint32_t)K & a_reg) + par) & 1u) f_cmul(v[K], m.c2, m.c3, n1);
            else f_cmul(v[K], m.c0, m.c1, n0);
        }
    }
}

#define QV_F4(ARM, ALLV, base, ...)                                    \
    case base + 0: ARM<0, ALLV>(__VA_ARGS__); break;                   \
    case base + 1: ARM<1, ALLV>(__VA_ARGS__); break;                   \
    case base + 2: ARM<2, ALLV>(__VA_ARGS__); break;                   \
    case base + 3: ARM<3, ALLV>(__VA_ARGS__); break;

// One jump table over (form, slot, "no control sits in a register slot"): the planner adds
// FC_ALL to the code when okmask == 0xFFFF, so the common case runs without per-slot predicates.
__device__ __forceinline__ void apply_fop(const FDec &m, const uint32_t fl, const uint32_t grp, amp (&v)[NV]) {
    const uint32_t ok = m.w0 >> 16;
    uint32_t code = m.w0 & 0xFFu;
    asm volatile("" : "+r"(code));
    switch (code) {
#ifndef QV_EXP_SMALL
    QV_F4(farm_pr, false, FC_PR, v, m, ok)
    QV_F4(farm_px, false, FC_PX, v, m, ok)
#endif
    QV_F4(farm_sw, false, FC_SW, v, ok)
#ifndef QV_EXP_SMALL
    case FC_DU: farm_du<false>(v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u); break;
    QV_F4(farm_ds, false, FC_DS, v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u)
    case FC_DG: farm_dg(v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u); break;
    QV_F4(farm_pa, false, FC_PA, v, m, ok)
#endif
    QV_F4(farm_pr, true, FC_ALL + FC_PR, v, m, ok)
    QV_F4(farm_px, true, FC_ALL + FC_PX, v, m, ok)
    QV_F4(farm_sw, true, FC_ALL + FC_SW, v, ok)
    case FC_ALL + FC_DU: farm_du<true>(v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u); break;
    QV_F4(farm_ds, true, FC_ALL + FC_DS, v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u)
    case FC_ALL + FC_DG: farm_dg(v, m, ok, (__popc(grp & m.a_thr) + fl) & 1u); break;
    QV_F4(farm_pa, true, FC_ALL + FC_PA, v, m, ok)
    default: break;
    }
}

// The op loop of the FAST interpreter.
// FC_LX ("lazy x"): an x whose target and controls all sit on thread / outer bits is a permutation
// BETWEEN threads: it costs three XORs -- the thread flips the bit in the index it will store its
// amplitudes to (x.jl / x.mine_o) and in the virtual group number `vgrp` the later ops of the
// stage test their thread-bit controls and parities against.
__device__ __forceinline__ void stage_ops_fast(const uint32_t ops_s, const uint32_t flags_s, const uint32_t ob,
                                               const uint32_t oe, const uint32_t grp, StageCtx &x, amp (&v)[NV]) {
    uint32_t vgrp = grp;
    for (uint32_t o = ob; o < oe; ++o) {
        const uint32_t op_a = ops_s + MOP_BYTES * o;
        uint32_t fl;
        FDec m;
        asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(fl) : "r"(flags_s + o) : "memory");
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(m.w0), "=r"(m.ctrl_thr), "=r"(m.a_thr), "=r"(m.a_reg)
                     : "r"(op_a)
                     : "memory");
        if (!(fl & 0x80u) || (~vgrp & m.ctrl_thr)) {
            if ((m.w0 & 0xFFu) == (uint32_t)FC_DM) o += m.a_reg & 0xFFFFu;     // the whole run shares the controls
            continue;
        }
        if ((m.w0 & 0xFFu) == (uint32_t)FC_DM) {
            // merged diagonal run: acc = product of the constituents' factors for this thread
            const uint32_t cnt = m.a_reg & 0xFFFFu;
            double ar = 1.0, ai = 0.0;
            for (uint32_t k = 1; k <= cnt; ++k) {
                const uint32_t a2 = op_a + MOP_BYTES * k;
                uint32_t fl2, w2, athr2;
                asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(fl2) : "r"(flags_s + o + k) : "memory");
                asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(w2) : "r"(a2) : "memory");
                asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(athr2) : "r"(a2 + 8u) : "memory");
                const uint32_t par = (__popc(vgrp & athr2) + fl2) & 1u;
                if (!par && (w2 & ((uint32_t)MOP_SKIP0 << 8))) continue;
                double fr, fi;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(fr), "=d"(fi) : "r"(a2 + 16u + 16u * par) : "memory");
                const double t = ar * fr - ai * fi;
                ai = ar * fi + ai * fr;
                ar = t;
            }
            const uint32_t ok = m.w0 >> 16;
            const double nai = -ai;
            QV_FOR_K {
                if ((ok >> K) & 1u) f_cmul(v[K], ar, ai, nai);
            }
            o += cnt;
            continue;
        }
        if ((m.w0 & 0xFFu) == (uint32_t)FC_LX) {
            vgrp ^= m.a_thr;
            x.jl ^= 1u << (m.a_reg & 0xFFu);
            x.mine_o ^= 16u * swz(1u << (m.a_reg & 0xFFu));
            x.goff ^= 16ull << (m.a_reg >> 8);           // (a_reg high byte: the bit's position in the shard)
            continue;
        }
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(m.c0), "=d"(m.c1) : "r"(op_a + 16u) : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(m.c2), "=d"(m.c3) : "r"(op_a + 32u) : "memory");
        apply_fop(m, fl, vgrp, v);
    }
}

// Kernel configurations: <threads per CTA, min CTAs per SM, tile buffers>.
//   T == 12: 256 threads, 2 CTAs/SM, one 64 KiB buffer.  The last stage stores its registers
//            straight to HBM, so the buffer is free as soon as that stage has read it: the next
//            tile's cp.async loads are issued right there and overlap the last stage's arithmetic
//            and stores.
//   T <= 11: 128 threads, 4 CTAs/SM, one 32 KiB buffer each (default): the kernel is bound by
//            instruction issue and latency, not by HBM, so the extra resident warps pay more
//            than a second buffer does (option "tile_nbuf" = 2: 3 CTAs/SM with two buffers).
constexpr uint32_t META_SLOTS = 3;      // per-tile op flags rotate through 3 slots: preparing tile i+1 must not
                                        // race with the threads still running the ops of tile i-1

template <int THREADS, int MINB, int NB, bool FULL>
__global__ void __launch_bounds__(THREADS, MINB)
k_tile_pass(const __grid_constant__ Segs segs, const __grid_constant__ TPassHdr hdr,
            const TStage *__restrict__ g_stages, const MOp *__restrict__ g_ops,
            const MBase *__restrict__ g_bases, const amp *__restrict__ mats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t T = hdr.T, L = hdr.L;
    const uint32_t n_ops = hdr.n_ops, n_stages = hdr.n_stages;
    const uint32_t tile_len = 1u << T;
    const uint32_t tile_bytes = 16u << T;
    const uint32_t lmask = (1u << L) - 1u;
    const uint32_t n_chunks = 1u << (T - L);
    const uint32_t flags_stride = (n_ops + 15u) & ~15u;

    // shared memory carve-up (16-byte aligned sections first)
    unsigned char *tiles_b = smem_raw;                                            // NB * (16 << T)
    MOp *s_ops = reinterpret_cast<MOp *>(smem_raw + (size_t)NB * tile_bytes);      // 48 * n_ops
    TStage *s_stages = reinterpret_cast<TStage *>(s_ops + n_ops);                  // 32 * n_stages
    MBase *s_bases = reinterpret_cast<MBase *>(s_stages + n_stages);               // 16 * n_ops
    unsigned long long *s_ptr0 = reinterpret_cast<unsigned long long *>(s_bases + n_ops);   // 8 * n_chunks
    unsigned long long *s_goff = s_ptr0 + ((n_chunks + 1u) & ~1u);                 // 8 * nthr (last stage)
    uint32_t *s_ctab = reinterpret_cast<uint32_t *>(s_goff + nthr);                // 16 * n_stages
    uint32_t *s_jltab = s_ctab + 4u * n_stages;                                    // 4 * nthr * n_stages
    uint8_t *s_gpos = reinterpret_cast<uint8_t *>(s_jltab + nthr * n_stages);      // 16
    uint8_t *s_flags_all = s_gpos + 16;                                            // 3 * flags_stride

    const uint32_t shard_shift = segs.shift;
    const uint64_t shard_mask = (1ull << shard_shift) - 1ull;
    // the pass program: same for every tile this CTA processes
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_ops + hdr.op_begin);
        uint4 *dst = reinterpret_cast<uint4 *>(s_ops);
        for (uint32_t i = tid; i < (MOP_BYTES / 16u) * n_ops; i += nthr) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g_stages + hdr.stage_begin);
        dst = reinterpret_cast<uint4 *>(s_stages);
        for (uint32_t i = tid; i < 2 * n_stages; i += nthr) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g_bases + hdr.op_begin);
        dst = reinterpret_cast<uint4 *>(s_bases);
        for (uint32_t i = tid; i < n_ops; i += nthr) dst[i] = src[i];
        // Chunk base pointers for the tile at offset 0: chunk c = the 2^L amplitudes whose gathered
        // tile bits spell c.  Rank bits among them select the shard (own HBM or a peer's, mapped
        // over NVLink); every tile of the pass adds the same byte offset to all of them.
        for (uint32_t c = tid; c < n_chunks; c += nthr) {
            uint64_t gidx = hdr.base_or;
            for (uint32_t j = 0; j < T - L; ++j)
                if ((c >> j) & 1u) gidx |= 1ull << hdr.gpos[L + j];
            s_ptr0[c] = (unsigned long long)(uintptr_t)(segs.seg[gidx >> shard_shift] + (gidx & shard_mask));
        }
    }
    const uint32_t n_t = T - TR;                     // thread bits per stage
    if (tid < 16u) s_gpos[tid] = hdr.gpos[tid];
    __syncthreads();                                 // the stage descriptors are in shared memory
    // per-(stage, thread) and per-stage constants, once per kernel
    for (uint32_t st = 0; st < n_stages; ++st) {
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(s_stages + st);
        const uint32_t jl = stage_jl(stage_s, n_t, tid);
        s_jltab[st * nthr + tid] = ((16u * swz(jl)) << 16) | jl;
        if (tid < (uint32_t)TR) s_ctab[4u * st + tid] = 16u * swz(1u << s_stages[st].r_lpos[tid]);
        if (st + 1 == n_stages) {
            unsigned long long go = 0;
            for (uint32_t l = 0; l < T; ++l)
                if ((jl >> l) & 1u) go += 16ull << hdr.gpos[l];
            s_goff[tid] = go;
        }
    }
    const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles_b);
    const uint32_t jltab_s = (uint32_t)__cvta_generic_to_shared(s_jltab);
    const uint32_t ctab_s = (uint32_t)__cvta_generic_to_shared(s_ctab);
    const uint32_t gpos_s = (uint32_t)__cvta_generic_to_shared(s_gpos);
    const unsigned long long shard_base = (unsigned long long)(uintptr_t)segs.seg[segs.rank];
    const uint32_t ops_s = (uint32_t)__cvta_generic_to_shared(s_ops);
    const uint32_t stages_s = (uint32_t)__cvta_generic_to_shared(s_stages);
    const uint32_t flags_all_s = (uint32_t)__cvta_generic_to_shared(s_flags_all);
    const uint32_t ptr0_s = (uint32_t)__cvta_generic_to_shared(s_ptr0);
    const bool active = tid < (1u << n_t);           // this thread owns a group of 16 amplitudes
    __syncthreads();

    // Metadata of the first tile >= t (stepping by the grid) that some op of this pass can change:
    // per-op flags (bit 7: controls outside the tile satisfied; bits 0-2: popcount of the diagonal
    // target mask over the bits outside the tile, mod 8) and the tile's byte offset inside the
    // shard.  Tiles no op touches (multi-controlled gates) are skipped without being read.
    // Returns n_tiles if there is none.
    auto prepare = [&](uint64_t t, uint32_t slot, unsigned long long &toff) -> uint64_t {
        uint8_t *flags = s_flags_all + slot * flags_stride;
        for (; t < hdr.n_tiles; t += gridDim.x) {
            uint64_t base = t;               // tile counter -> local base index (tile bits clear)
            for (uint32_t k = 0; k < hdr.n_runs; ++k) {
                const uint32_t p = hdr.run_pos[k], len = hdr.run_len[k];
                base = ((base >> p) << (p + len)) | (base & ((1ull << p) - 1ull));
            }
            base |= hdr.fx_val;
            toff = base * 16ull;
            base |= hdr.base_or;
            int any = 0;
            for (uint32_t o = tid; o < n_ops; o += nthr) {
                const MBase b = s_bases[o];
                const uint32_t okb = ((~base & b.ctrl_base) == 0) ? 0x80u : 0u;
                flags[o] = (uint8_t)(okb | ((uint32_t)__popcll(base & b.a_base) & 7u));
                any |= (int)okb;
            }
            if (__syncthreads_or(any)) return t;
        }
        return hdr.n_tiles;
    };
    // 2^(T-L) chunks of 2^L contiguous amplitudes, 16-byte cp.async, swizzled; one commit group.
    // Element j = tid + i * nthr: swz is GF(2)-linear and the thread count a power of two, so the
    // shared-memory slot is swz(tid) ^ swz(i * nthr); with nthr a multiple of the chunk length the
    // offset inside the chunk is the thread's own and the chunk index advances by nthr >> L.
    const uint32_t my_slot = 16u * swz(tid);
    const bool regular = (nthr & lmask) == 0u && (tile_len % nthr) == 0u;
    auto issue_load = [&](const unsigned long long toff, uint32_t bslot) {
        const uint32_t tb_s = tiles_s + bslot * tile_bytes;
        if (hdr.sysload) {
            // peer tiles: system-scope loads (never served from a requester-side cache)
            for (uint32_t j = tid; j < tile_len; j += nthr) {
                const unsigned long long p = s_ptr0[j >> L] + toff + (unsigned long long)(j & lmask) * 16ull;
                double a, b;
                asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];\n" : "=d"(a), "=d"(b) : "l"(p) : "memory");
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(tb_s + 16u * swz(j)), "d"(a), "d"(b) : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            return;
        }
        if (regular) {
            const unsigned long long mine = toff + (unsigned long long)(tid & lmask) * 16ull;
            const uint32_t cstep = (nthr >> L) * 8u;
            uint32_t ch = ptr0_s + (tid >> L) * 8u;
            for (uint32_t jb = 0; jb < tile_len; jb += nthr, ch += cstep) {
                unsigned long long p;
                asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(p) : "r"(ch) : "memory");
                p += mine;
                const uint32_t d = tb_s + (my_slot ^ (16u * swz(jb)));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(p) : "memory");
            }
        } else {
            for (uint32_t j = tid; j < tile_len; j += nthr) {
                const unsigned long long p = s_ptr0[j >> L] + toff + (unsigned long long)(j & lmask) * 16ull;
                const uint32_t d = tb_s + 16u * swz(j);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(p) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    uint32_t mslot = 0, bslot = 0;
    unsigned long long toff_cur = 0, toff_next = 0;
    uint64_t t_cur = prepare(blockIdx.x, mslot, toff_cur);
    if (t_cur < hdr.n_tiles) issue_load(toff_cur, bslot);
    while (t_cur < hdr.n_tiles) {
        const uint32_t mnext = mslot + 1u == META_SLOTS ? 0u : mslot + 1u;
        const uint64_t t_next = prepare(t_cur + gridDim.x, mnext, toff_next);
        const bool has_next = t_next < hdr.n_tiles;
        if (NB == 2 && has_next) {
            issue_load(toff_next, bslot ^ 1u);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const uint32_t tile_s = tiles_s + bslot * tile_bytes;
        const uint32_t flags_s = flags_all_s + mslot * flags_stride;

        for (uint32_t s = 0; s < n_stages; ++s) {
            const bool last = s + 1 == n_stages;
            const uint32_t ob = s ? hdr.stage_end[s - 1] : 0u, oe = hdr.stage_end[s];
            amp v[NV];
            StageCtx x;
            if (active) {
                x = stage_ctx(stages_s + 32u * s, jltab_s + 4u * (s * nthr + tid), ctab_s + 16u * s);
                if (last) x.goff = s_goff[tid];
                stage_load(tile_s, x, v);
            }
            if (last && NB == 1) {
                __syncthreads();                       // every thread holds its amplitudes: the buffer is free
                if (has_next) issue_load(toff_next, bslot);
            } else if (!last && s_stages[s].sync_after_load) {
                // a lazy x makes threads store into each other's slots of the tile buffer: nobody
                // may store before everybody has loaded
                __syncthreads();
            }
            if (active) {
                if (FULL) stage_ops_full(ops_s, flags_s, ob, oe, tid, mats, v);
                else stage_ops_fast(ops_s, flags_s, ob, oe, tid, x, v);
                if (!last) stage_store_smem(tile_s, x, v);
                else if (hdr.touches_peer) stage_store_global(ptr0_s, L, toff_cur, x, v);
                else stage_store_global_local(shard_base + toff_cur, gpos_s, x, v);
            }
            if (!last) __syncthreads();
        }
        mslot = mnext;
        if (NB == 2) bslot ^= 1u;
        t_cur = t_next;
        toff_cur = toff_next;
    }
    // Stores into a peer's HBM must be performed at SYSTEM scope before the barrier kernel that
    // follows signals the peers: the barrier's own fence is executed by other threads and does not
    // cover them.
    if (hdr.touches_peer) __threadfence_system();
}

constexpr size_t TILE_SMEM_MAX = 227u * 1024u;

static size_t tile_smem_bytes(const TPassHdr &h, int nb, int threads = 256) {
    return (size_t)nb * ((size_t)16 << h.T) + (size_t)(MOP_BYTES + sizeof(MBase)) * h.n_ops + (size_t)32 * h.n_stages +
           ((size_t)8 << (h.T - h.L)) + 8 + (size_t)META_SLOTS * ((h.n_ops + 15u) & ~15u) +
           (size_t)8 * threads + (size_t)16 * h.n_stages + (size_t)4 * threads * h.n_stages + 16;
}

typedef void (*tile_kernel_t)(const Segs, const TPassHdr, const TStage *, const MOp *, const MBase *, const amp *);

template <int THREADS, int MINB, int NB>
static tile_kernel_t pick_kernel(bool full) {
    return full ? k_tile_pass<THREADS, MINB, NB, true> : k_tile_pass<THREADS, MINB, NB, false>;
}

int tile_kernel_setup() {
    bool ok = true;
    for (int full = 0; full < 2; ++full) {
        const tile_kernel_t ks[6] = {pick_kernel<128, 3, 2>(full), pick_kernel<128, 3, 1>(full),
                                     pick_kernel<256, 2, 1>(full), pick_kernel<256, 1, 2>(full),
                                     pick_kernel<128, 4, 1>(full), pick_kernel<128, 5, 1>(full)};
        for (tile_kernel_t k : ks)
            ok = ok && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TILE_SMEM_MAX) == cudaSuccess;
    }
    return ok ? 0 : -1;
}

int g_tile_stagger = 0;      // experiment knob (option "tile_stagger"): start offset between the CTAs of an SM, cycles (no measurable effect)
int g_tile_sysload = 0;    // debug/safety knob (option "tile_sysload"): system-scope loads for peer tiles
int g_tile_nbuf = 0;     // tuning knob (option "tile_nbuf"): 0 = auto, 1 / 2 = force the buffer count

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const MOp *d_ops, const MBase *d_bases, const amp *mat_table, int sm_count) {
    if (hdr.T < TILE_MIN_BITS || hdr.T > TILE_MAX_BITS || hdr.L > hdr.T || hdr.T - hdr.L > TILE_MAX_HIGH ||
        hdr.n_tiles == 0 || hdr.n_ops == 0 || hdr.n_ops > (uint32_t)TILE_MAX_OPS ||
        hdr.n_stages > (uint32_t)TILE_MAX_STAGES)
        return -1;
    tile_kernel_t kern;
    int threads, nb;
    if (hdr.T >= 12) {
        threads = 256;
        nb = g_tile_nbuf == 2 ? 2 : 1;
        if (nb == 2 && tile_smem_bytes(hdr, 2) > TILE_SMEM_MAX) nb = 1;
        kern = nb == 2 ? pick_kernel<256, 1, 2>(hdr.full != 0) : pick_kernel<256, 2, 1>(hdr.full != 0);
    } else {
        threads = 1 << (hdr.T - TILE_R);
        if (threads < 32) threads = 32;
        if (threads > 128) threads = 128;
        nb = g_tile_nbuf == 2 ? 2 : 1;
        if (nb == 2 && tile_smem_bytes(hdr, 2) > TILE_SMEM_MAX) nb = 1;
        kern = nb == 2 ? pick_kernel<128, 3, 2>(hdr.full != 0)
                       : g_tile_nbuf == 3 ? pick_kernel<128, 3, 1>(hdr.full != 0)
                       : g_tile_nbuf == 5 ? pick_kernel<128, 5, 1>(hdr.full != 0) : pick_kernel<128, 4, 1>(hdr.full != 0);
    }
    const size_t smem = tile_smem_bytes(hdr, nb, threads);
    if (smem > TILE_SMEM_MAX) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    TPassHdr h2 = hdr;
    h2.waves = (uint32_t)per_sm;
    h2.stagger_cycles = grid > (uint64_t)sm_count ? (uint32_t)g_tile_stagger : 0u;
    h2.sysload = (g_tile_sysload && hdr.touches_peer) ? 1u : 0u;
    kern<<<(unsigned)grid, threads, smem, st>>>(segs, h2, d_stages, d_ops, d_bases, mat_table);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace qv
