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
// The stage interpreter.  A stage keeps TILE_R = 4 tile bits in "register slots": every
// thread holds the 16 amplitudes that differ in those bits.  An op is a 32-byte MOp in
// shared memory whose `code` selects a fully unrolled body (kind x slot); everything
// that depends on the amplitude index was split by the planner into
//   - register-slot part: compile-time per slot (signs of ry/h1/y/ryy, rows of u1/u2) or a
//     16-bit `okmask` (controls sitting in register slots),
//   - thread part: one AND/compare on the thread's group number per op,
//   - tile part: one flag byte per op per tile (controls / diagonal-mask bits outside the
//     tile), computed once per tile.
// Tiles none of whose ops is active (multi-controlled gates) are skipped without being read.
//
// Arithmetic: every gate uses the reference's formula (gates.cuh) with FMA contraction
// off, on the same operands as the reference's gather form; only the ORDER of commuting
// gates may differ from the op list (planner.cu).
#include "engine.h"
#include "gates.cuh"

namespace qv {

#ifndef QV_OP_PREFETCH
#define QV_OP_PREFETCH 0
#endif
constexpr int NV = TILE_NV;
constexpr int TR = TILE_R;

__device__ __forceinline__ uint32_t swz(uint32_t j) {
    return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9)) & 7u);
}

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
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

// One stage on one tile: every thread takes the 16 amplitudes of its group(s) into registers,
// runs the stage's ops on them and writes them back.  All addresses are 32-bit shared-window
// offsets; the op range [ob, oe) is relative to the pass's first op.
template <bool FULL>
__device__ __forceinline__ void run_stage(const uint32_t tile_s, const uint32_t stage_s, const uint32_t ops_s,
                                          const uint32_t flags_s, const uint32_t ob, const uint32_t oe,
                                          const uint32_t n_t, const amp *__restrict__ mats) {
    uint32_t sw[8];      // the TStage: op_begin, op_end, r_lpos[4], t_lpos[16], pad
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(sw[0]), "=r"(sw[1]), "=r"(sw[2]), "=r"(sw[3]) : "r"(stage_s) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(sw[4]), "=r"(sw[5]), "=r"(sw[6]), "=r"(sw[7]) : "r"(stage_s + 16u) : "memory");
    uint32_t c[TR];
#pragma unroll
    for (int j = 0; j < TR; ++j) c[j] = 16u * swz(1u << ((sw[2] >> (8 * j)) & 0xFFu));
    const uint32_t groups = 1u << n_t;
    for (uint32_t grp = threadIdx.x; grp < groups; grp += blockDim.x) {
        uint32_t jl = 0;
#pragma unroll
        for (uint32_t k = 0; k < 16; ++k) {
            if (k < n_t) jl |= ((grp >> k) & 1u) << ((sw[3 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu);
        }
        const uint32_t mine_o = 16u * swz(jl);      // byte offset of slot pattern 0 in the tile
        // (the XOR applies to the offset inside the tile: the tile itself is only 16-byte aligned)
        amp v[NV];
#pragma unroll
        for (int K = 0; K < NV; ++K) {
            uint32_t a = mine_o;
            if (K & 1) a ^= c[0];
            if (K & 2) a ^= c[1];
            if (K & 4) a ^= c[2];
            if (K & 8) a ^= c[3];
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n"
                         : "=d"(v[K].x), "=d"(v[K].y)
                         : "r"(tile_s + a)
                         : "memory");
        }
        // ob / oe come from the kernel parameters (uniform registers): the loop control costs no
        // vector registers, which the arms below need for their 16 amplitudes + products.
        // The descriptor of op o+1 is fetched before op o's arithmetic is issued, so its
        // shared-memory latency hides under the FP64 burst of op o.
        MDec m_nx;
        uint32_t fl_nx;
        auto fetch = [&](uint32_t o) {
            const uint32_t op_a = ops_s + 32u * o, fl_a = flags_s + o;
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
#if QV_OP_PREFETCH
            if (o + 1 < oe) fetch(o + 1);
#endif
            if (!(fl & 0x80u) || (~grp & m.ctrl_thr)) {
#if !QV_OP_PREFETCH
                if (o + 1 < oe) fetch(o + 1);
#endif
                continue;
            }
            if ((m.w0 >> 16) == 0xFFFFu) apply_mop<true, FULL>(m, fl, grp, mats, v);
            else apply_mop<false, FULL>(m, fl, grp, mats, v);
#if !QV_OP_PREFETCH
            if (o + 1 < oe) fetch(o + 1);
#endif
        }
        // recompute the 16 slot addresses instead of keeping them live across the op loop
        uint32_t mine_o2 = mine_o;
        asm volatile("" : "+r"(mine_o2));
#pragma unroll
        for (int K = 0; K < NV; ++K) {
            uint32_t a = mine_o2;
            if (K & 1) a ^= c[0];
            if (K & 2) a ^= c[1];
            if (K & 4) a ^= c[2];
            if (K & 8) a ^= c[3];
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(tile_s + a), "d"(v[K].x), "d"(v[K].y)
                         : "memory");
        }
    }
}

// Kernel configurations: <threads per CTA, min CTAs per SM, tile buffers>.
//   T <= 11: 128 threads, 3 CTAs/SM, TWO tile buffers -- the next tile's cp.async loads are in
//            flight while the current tile is computed, stores are fire-and-forget, so the HBM
//            phases overlap the FP64 phases inside each CTA.
//   T == 12: 256 threads, 2 CTAs/SM, one 64 KiB buffer (two do not fit twice per SM).
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
    MOp *s_ops = reinterpret_cast<MOp *>(smem_raw + (size_t)NB * tile_bytes);      // 32 * n_ops
    TStage *s_stages = reinterpret_cast<TStage *>(s_ops + n_ops);                  // 32 * n_stages
    amp **s_cptr_all = reinterpret_cast<amp **>(s_stages + n_stages);              // NB * 8 * n_chunks
    uint8_t *s_flags_all = reinterpret_cast<uint8_t *>(s_cptr_all + NB * n_chunks);  // NB * flags_stride

    // the pass program: same for every tile this CTA processes
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_ops + hdr.op_begin);
        uint4 *dst = reinterpret_cast<uint4 *>(s_ops);
        for (uint32_t i = tid; i < 2 * n_ops; i += nthr) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g_stages + hdr.stage_begin);
        dst = reinterpret_cast<uint4 *>(s_stages);
        for (uint32_t i = tid; i < 2 * n_stages; i += nthr) dst[i] = src[i];
    }
    const uint32_t shard_shift = segs.shift;
    const uint64_t shard_mask = (1ull << shard_shift) - 1ull;
    const uint32_t n_t = T - TR;                     // thread bits per stage
    const MBase *bases = g_bases + hdr.op_begin;
    const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles_b);
    const uint32_t ops_s = (uint32_t)__cvta_generic_to_shared(s_ops);
    const uint32_t stages_s = (uint32_t)__cvta_generic_to_shared(s_stages);
    const uint32_t flags_all_s = (uint32_t)__cvta_generic_to_shared(s_flags_all);
    __syncthreads();

    // Metadata of the first tile >= t (stepping by the grid) that some op of this pass can change:
    // per-op flags (bit 7: controls outside the tile satisfied; bits 0-2: popcount of the diagonal
    // target mask over the bits outside the tile, mod 8) and the chunk pointers.  Tiles no op
    // touches (multi-controlled gates) are skipped without being read.  Returns n_tiles if none.
    auto prepare = [&](uint64_t t, uint32_t slot) -> uint64_t {
        uint8_t *flags = s_flags_all + slot * flags_stride;
        amp **cptr = s_cptr_all + slot * n_chunks;
        for (; t < hdr.n_tiles; t += gridDim.x) {
            uint64_t base = t;               // tile counter -> global base index (tile bits clear)
            for (uint32_t k = 0; k < hdr.n_runs; ++k) {
                const uint32_t p = hdr.run_pos[k], len = hdr.run_len[k];
                base = ((base >> p) << (p + len)) | (base & ((1ull << p) - 1ull));
            }
            base |= hdr.fx_val | hdr.base_or;
            int any = 0;
            for (uint32_t o = tid; o < n_ops; o += nthr) {
                const MBase b = bases[o];
                const uint32_t okb = ((~base & b.ctrl_base) == 0) ? 0x80u : 0u;
                flags[o] = (uint8_t)(okb | ((uint32_t)__popcll(base & b.a_base) & 7u));
                any |= (int)okb;
            }
            for (uint32_t c = tid; c < n_chunks; c += nthr) {
                uint64_t gidx = base;
                for (uint32_t j = 0; j < T - L; ++j)
                    if ((c >> j) & 1u) gidx |= 1ull << hdr.gpos[L + j];
                cptr[c] = segs.seg[gidx >> shard_shift] + (gidx & shard_mask);
            }
            if (__syncthreads_or(any)) return t;
        }
        return hdr.n_tiles;
    };
    // 2^(T-L) chunks of 2^L contiguous amplitudes, 16-byte cp.async, swizzled; one commit group
    auto issue_load = [&](uint32_t slot) {
        amp *const *cptr = s_cptr_all + slot * n_chunks;
        unsigned char *tb = tiles_b + slot * tile_bytes;
        for (uint32_t j = tid; j < tile_len; j += nthr) cp_async_16(tb + 16u * swz(j), cptr[j >> L] + (j & lmask));
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    uint32_t slot = 0;
    uint64_t t_cur = prepare(blockIdx.x, slot);
    if (t_cur < hdr.n_tiles) issue_load(slot);
    while (t_cur < hdr.n_tiles) {
        uint64_t t_next = hdr.n_tiles;
        if (NB == 2) {
            t_next = prepare(t_cur + gridDim.x, slot ^ 1u);
            if (t_next < hdr.n_tiles) {
                issue_load(slot ^ 1u);
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const uint32_t tile_s = tiles_s + slot * tile_bytes;
        const uint32_t flags_s = flags_all_s + slot * flags_stride;

        // ---- stages: 16 amplitudes per thread in registers ------------------------------------
        for (uint32_t s = 0; s < n_stages; ++s) {
            run_stage<FULL>(tile_s, stages_s + 32u * s, ops_s, flags_s, s ? hdr.stage_end[s - 1] : 0u, hdr.stage_end[s],
                      n_t, mats);
            __syncthreads();
        }

        // ---- store back in place (fire and forget) --------------------------------------------
        {
            amp *const *cptr = s_cptr_all + slot * n_chunks;
            const unsigned char *tb = tiles_b + slot * tile_bytes;
            for (uint32_t j = tid; j < tile_len; j += nthr)
                *(cptr[j >> L] + (j & lmask)) = *reinterpret_cast<const amp *>(tb + 16u * swz(j));
        }
        __syncthreads();                 // this slot's tile buffer and metadata are free again
        if (NB == 1) {
            t_next = prepare(t_cur + gridDim.x, slot);
            if (t_next < hdr.n_tiles) issue_load(slot);
        } else {
            slot ^= 1u;
        }
        t_cur = t_next;
    }
}

constexpr size_t TILE_SMEM_MAX = 227u * 1024u;

static size_t tile_smem_bytes(const TPassHdr &h, int nb) {
    return (size_t)nb * ((size_t)16 << h.T) + (size_t)32 * h.n_ops + (size_t)32 * h.n_stages +
           (size_t)nb * ((size_t)8 << (h.T - h.L)) + (size_t)nb * ((h.n_ops + 15u) & ~15u);
}

typedef void (*tile_kernel_t)(const Segs, const TPassHdr, const TStage *, const MOp *, const MBase *, const amp *);

template <int THREADS, int MINB, int NB>
static tile_kernel_t pick_kernel(bool full) {
    return full ? k_tile_pass<THREADS, MINB, NB, true> : k_tile_pass<THREADS, MINB, NB, false>;
}

int tile_kernel_setup() {
    bool ok = true;
    for (int full = 0; full < 2; ++full) {
        const tile_kernel_t ks[3] = {pick_kernel<128, 3, 2>(full), pick_kernel<128, 3, 1>(full),
                                     pick_kernel<256, 2, 1>(full)};
        for (tile_kernel_t k : ks)
            ok = ok && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TILE_SMEM_MAX) == cudaSuccess;
    }
    return ok ? 0 : -1;
}

int g_tile_nbuf = 0;     // tuning knob (option "tile_nbuf"): 0 = auto, 1 / 2 = force for T <= 11

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const MOp *d_ops, const MBase *d_bases, const amp *mat_table, int sm_count) {
    if (hdr.T < TILE_MIN_BITS || hdr.T > TILE_MAX_BITS || hdr.L > hdr.T || hdr.T - hdr.L > TILE_MAX_HIGH ||
        hdr.n_tiles == 0 || hdr.n_ops == 0 || hdr.n_ops > (uint32_t)TILE_MAX_OPS ||
        hdr.n_stages > (uint32_t)TILE_MAX_STAGES)
        return -1;
    tile_kernel_t kern;
    int threads, nb;
    if (hdr.T >= 12) {
        kern = pick_kernel<256, 2, 1>(hdr.full != 0);
        threads = 256;
        nb = 1;
    } else {
        threads = 1 << (hdr.T - TILE_R);
        if (threads < 32) threads = 32;
        if (threads > 128) threads = 128;
        nb = 2;
        kern = pick_kernel<128, 3, 2>(hdr.full != 0);
        if (g_tile_nbuf == 1 || tile_smem_bytes(hdr, 2) > TILE_SMEM_MAX) {   // (very long pass programs)
            nb = 1;
            kern = pick_kernel<128, 3, 1>(hdr.full != 0);
        }
    }
    const size_t smem = tile_smem_bytes(hdr, nb);
    if (smem > TILE_SMEM_MAX) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    kern<<<(unsigned)grid, threads, smem, st>>>(segs, hdr, d_stages, d_ops, d_bases, mat_table);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace qv
