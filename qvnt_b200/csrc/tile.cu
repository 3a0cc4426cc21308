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
// bits into the low 3) so that, whichever 3 bits a stage keeps in registers, the 8 lanes
// of a quarter-warp hit 8 different 16-byte bank groups (the planner picks the lane bits).
//
// Arithmetic: every gate uses the reference's formula (gates.cuh) with FMA contraction
// off, on the same operands as the reference's gather form; only the ORDER of commuting
// gates may differ from the op list (planner.cu).
#include "engine.h"
#include "gates.cuh"

namespace qv {

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

// global index of register slot K of a thread whose register bits are all clear in g
template <int K>
__device__ __forceinline__ uint64_t gi(uint64_t g, const uint64_t (&rg)[TILE_R]) {
    uint64_t i = g;
    if (K & 1) i |= rg[0];
    if (K & 2) i |= rg[1];
    if (K & 4) i |= rg[2];
    return i;
}

template <int K, int END, typename F>
struct StaticFor {
    static __device__ __forceinline__ void run(F &f) {
        f.template operator()<K>();
        StaticFor<K + 1, END, F>::run(f);
    }
};
template <int END, typename F>
struct StaticFor<END, END, F> {
    static __device__ __forceinline__ void run(F &) {}
};

constexpr int NV = 1 << TILE_R;

// ---- per-form bodies, templated on the gate kind and the register bits involved ---------
template <int KIND>
struct DiagBody {
    const DevOp &d;
    amp (&v)[NV];
    uint64_t g;
    const uint64_t (&rg)[TILE_R];
    template <int K>
    __device__ __forceinline__ void operator()() {
        const uint64_t i = gi<K>(g, rg);
        if ((~i & d.ctrl) == 0) v[K] = diag_out<KIND>(d, v[K], i);
    }
};

template <int KIND, int RB>
struct Pair1Body {
    const DevOp &d;
    const amp *m;
    amp (&v)[NV];
    uint64_t g;
    const uint64_t (&rg)[TILE_R];
    template <int K>
    __device__ __forceinline__ void operator()() {
        if (K & (1 << RB)) return;
        constexpr int K1 = K | (1 << RB);
        const uint64_t i0 = gi<K>(g, rg), i1 = gi<K1>(g, rg);
        if ((~i0 & d.ctrl) == 0) pair_update<KIND>(d, m, v[K], v[K1], i0, i1);
    }
};

template <int KIND, int RA, int RB>
struct Pair2XBody {   // partner differs in both register bits
    const DevOp &d;
    const amp *m;
    amp (&v)[NV];
    uint64_t g;
    const uint64_t (&rg)[TILE_R];
    template <int K>
    __device__ __forceinline__ void operator()() {
        if (K & (1 << RA)) return;
        constexpr int K1 = K ^ (1 << RA) ^ (1 << RB);
        const uint64_t i0 = gi<K>(g, rg), i1 = gi<K1>(g, rg);
        if ((~i0 & d.ctrl) == 0) pair_update<KIND>(d, m, v[K], v[K1], i0, i1);
    }
};

template <int KIND, int RA, int RB>
struct Odd2Body {     // swap family: only the odd-parity pair {a set, b set} changes
    const DevOp &d;
    const amp *m;
    amp (&v)[NV];
    uint64_t g;
    const uint64_t (&rg)[TILE_R];
    template <int K>
    __device__ __forceinline__ void operator()() {
        if (K & ((1 << RA) | (1 << RB))) return;
        constexpr int K0 = K | (1 << RA), K1 = K | (1 << RB);
        const uint64_t i0 = gi<K0>(g, rg), i1 = gi<K1>(g, rg);
        if ((~i0 & d.ctrl) == 0) pair_update<KIND>(d, m, v[K0], v[K1], i0, i1);
    }
};

template <int KIND, int RA, int RB>
struct QuadBody {     // h2 / u2: a = register bit RA, b = register bit RB
    const DevOp &d;
    const amp *m;
    amp (&v)[NV];
    uint64_t g;
    const uint64_t (&rg)[TILE_R];
    template <int K>
    __device__ __forceinline__ void operator()() {
        if (K & ((1 << RA) | (1 << RB))) return;
        const uint64_t i0 = gi<K>(g, rg);
        if ((~i0 & d.ctrl) != 0) return;
        amp q[4] = {v[K], v[K | (1 << RA)], v[K | (1 << RB)], v[K | (1 << RA) | (1 << RB)]};
        quad_update<KIND>(d, m, q);
        v[K] = q[0];
        v[K | (1 << RA)] = q[1];
        v[K | (1 << RB)] = q[2];
        v[K | (1 << RA) | (1 << RB)] = q[3];
    }
};

#define COMMA ,
#define QV_RUN(BODY)                                   \
    do {                                               \
        BODY body_{d, m, v, g, rg};                    \
        StaticFor<0, NV, BODY>::run(body_);            \
    } while (0)
#define QV_RUN_D(BODY)                                 \
    do {                                               \
        BODY body_{d, v, g, rg};                       \
        StaticFor<0, NV, BODY>::run(body_);            \
    } while (0)

template <int RB>
__device__ __forceinline__ void run_pair1(const DevOp &d, const amp *m, amp (&v)[NV], uint64_t g,
                                          const uint64_t (&rg)[TILE_R]) {
    switch (d.kind) {
    case QVNT_X: QV_RUN(Pair1Body<QVNT_X COMMA RB>); break;
    case QVNT_Y: QV_RUN(Pair1Body<QVNT_Y COMMA RB>); break;
    case QVNT_RX: QV_RUN(Pair1Body<QVNT_RX COMMA RB>); break;
    case QVNT_RY: QV_RUN(Pair1Body<QVNT_RY COMMA RB>); break;
    case QVNT_H1: QV_RUN(Pair1Body<QVNT_H1 COMMA RB>); break;
    default: QV_RUN(Pair1Body<QVNT_U1 COMMA RB>); break;
    }
}
template <int RA, int RB>
__device__ __forceinline__ void run_pair2x(const DevOp &d, const amp *m, amp (&v)[NV], uint64_t g,
                                           const uint64_t (&rg)[TILE_R]) {
    if (d.kind == QVNT_RXX) QV_RUN(Pair2XBody<QVNT_RXX COMMA RA COMMA RB>);
    else QV_RUN(Pair2XBody<QVNT_RYY COMMA RA COMMA RB>);
}
template <int RA, int RB>
__device__ __forceinline__ void run_odd2(const DevOp &d, const amp *m, amp (&v)[NV], uint64_t g,
                                         const uint64_t (&rg)[TILE_R]) {
    switch (d.kind) {
    case QVNT_SWAP: QV_RUN(Odd2Body<QVNT_SWAP COMMA RA COMMA RB>); break;
    case QVNT_ISWAP: QV_RUN(Odd2Body<QVNT_ISWAP COMMA RA COMMA RB>); break;
    case QVNT_SQRT_SWAP: QV_RUN(Odd2Body<QVNT_SQRT_SWAP COMMA RA COMMA RB>); break;
    default: QV_RUN(Odd2Body<QVNT_SQRT_ISWAP COMMA RA COMMA RB>); break;
    }
}
template <int RA, int RB>
__device__ __forceinline__ void run_quad(const DevOp &d, const amp *m, amp (&v)[NV], uint64_t g,
                                         const uint64_t (&rg)[TILE_R]) {
    if (d.kind == QVNT_H2) QV_RUN(QuadBody<QVNT_H2 COMMA RA COMMA RB>);
    else QV_RUN(QuadBody<QVNT_U2 COMMA RA COMMA RB>);
}

__device__ __forceinline__ void apply_op(const TOp *__restrict__ top, const amp *__restrict__ mats, amp (&v)[NV],
                                      uint64_t g, const uint64_t (&rg)[TILE_R]) {
    DevOp d = top->d;
    const uint32_t form = top->form;
    const int ra = top->ra, rb = top->rb;
    const amp *m = mats + d.mat;
    switch (form) {
    case TF_DIAG:
        switch (d.kind) {
        case QVNT_Z: QV_RUN_D(DiagBody<QVNT_Z>); break;
        case QVNT_S: QV_RUN_D(DiagBody<QVNT_S>); break;
        case QVNT_T: QV_RUN_D(DiagBody<QVNT_T>); break;
        case QVNT_RZ: QV_RUN_D(DiagBody<QVNT_RZ>); break;
        default: QV_RUN_D(DiagBody<QVNT_RZZ>); break;
        }
        break;
    case TF_PAIR1:
        if (ra == 0) run_pair1<0>(d, m, v, g, rg);
        else if (ra == 1) run_pair1<1>(d, m, v, g, rg);
        else run_pair1<2>(d, m, v, g, rg);
        break;
    case TF_PAIR2X: {
        const int lo = ra < rb ? ra : rb, hi = ra < rb ? rb : ra;
        if (lo == 0 && hi == 1) run_pair2x<0, 1>(d, m, v, g, rg);
        else if (lo == 0) run_pair2x<0, 2>(d, m, v, g, rg);
        else run_pair2x<1, 2>(d, m, v, g, rg);
        break;
    }
    case TF_ODD2: {
        const int lo = ra < rb ? ra : rb, hi = ra < rb ? rb : ra;
        if (lo == 0 && hi == 1) run_odd2<0, 1>(d, m, v, g, rg);
        else if (lo == 0) run_odd2<0, 2>(d, m, v, g, rg);
        else run_odd2<1, 2>(d, m, v, g, rg);
        break;
    }
    default:  // TF_QUAD
        switch (ra * 3 + rb) {
        case 1: run_quad<0, 1>(d, m, v, g, rg); break;
        case 2: run_quad<0, 2>(d, m, v, g, rg); break;
        case 3: run_quad<1, 0>(d, m, v, g, rg); break;
        case 5: run_quad<1, 2>(d, m, v, g, rg); break;
        case 6: run_quad<2, 0>(d, m, v, g, rg); break;
        default: run_quad<2, 1>(d, m, v, g, rg); break;
        }
        break;
    }
}

__global__ void __launch_bounds__(TILE_THREADS, 3)
k_tile_pass(const __grid_constant__ Segs segs, const __grid_constant__ TPassHdr hdr,
            const TStage *__restrict__ stages, const TOp *__restrict__ ops, const amp *__restrict__ mats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    amp *tile = reinterpret_cast<amp *>(smem_raw);
    __shared__ uint64_t s_hoff[1 << TILE_MAX_HIGH];   // gathered-bit pattern -> global index bits
    __shared__ amp *s_seg[MAX_WORLD];
    __shared__ uint8_t s_gpos[16];
    __shared__ uint8_t s_fxpos[64];

    const uint32_t tid = threadIdx.x;
    const uint32_t T = hdr.T, L = hdr.L;
    const uint32_t tile_len = 1u << T;
    const uint32_t lmask = (1u << L) - 1u;
    const uint32_t n_fx = hdr.fx.n;
    if (tid < 16) s_gpos[tid] = hdr.gpos[tid];
    if (tid < 64) s_fxpos[tid] = hdr.fx.pos[tid];
    if (tid < MAX_WORLD) s_seg[tid] = segs.seg[tid];
    __syncthreads();
    for (uint32_t h = tid; h < (1u << (T - L)); h += TILE_THREADS) {
        uint64_t gb = 0;
        for (uint32_t j = 0; j < T - L; ++j)
            if ((h >> j) & 1u) gb |= 1ull << s_gpos[L + j];
        s_hoff[h] = gb;
    }
    __syncthreads();
    const uint32_t shard_shift = segs.shift;
    const uint64_t shard_mask = (1ull << shard_shift) - 1ull;
    const uint32_t n_t = T - TILE_R;                 // thread bits per stage
    const uint32_t groups = 1u << n_t;

    for (uint64_t tile_i = blockIdx.x; tile_i < hdr.n_tiles; tile_i += gridDim.x) {
        // tile counter -> base index (tile bits clear)
        uint64_t base = tile_i;
        for (uint32_t k = 0; k < n_fx; ++k) {
            const uint32_t p = s_fxpos[k];
            base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
        }
        base |= hdr.fx.val | hdr.base_or;

        // ---- load: 2^(T-L) chunks of 2^L contiguous amplitudes, 16-byte cp.async, swizzled ----
        for (uint32_t j = tid; j < tile_len; j += TILE_THREADS) {
            const uint64_t gidx = base | s_hoff[j >> L] | (uint64_t)(j & lmask);
            const amp *src = s_seg[gidx >> shard_shift] + (gidx & shard_mask);
            cp_async_16(&tile[swz(j)], src);
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- stages: 2^TILE_R amplitudes per thread in registers ---------------------------
        for (uint32_t s = 0; s < hdr.n_stages; ++s) {
            const TStage *st = stages + hdr.stage_begin + s;
            uint32_t rl[TILE_R];
            uint64_t rg[TILE_R];
#pragma unroll
            for (int j = 0; j < TILE_R; ++j) {
                const uint32_t lp = st->r_lpos[j];
                rl[j] = 1u << lp;
                rg[j] = 1ull << s_gpos[lp];
            }
            const uint32_t ob = st->op_begin, oe = st->op_end;
            for (uint32_t grp = tid; grp < groups; grp += TILE_THREADS) {
                uint32_t jl = 0;
                uint64_t g = base;
                for (uint32_t k = 0; k < n_t; ++k) {
                    if ((grp >> k) & 1u) {
                        const uint32_t lp = st->t_lpos[k];
                        jl |= 1u << lp;
                        g |= 1ull << s_gpos[lp];
                    }
                }
                amp v[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const uint32_t j = jl | ((k & 1) ? rl[0] : 0u) | ((k & 2) ? rl[1] : 0u) | ((k & 4) ? rl[2] : 0u);
                    v[k] = tile[swz(j)];
                }
                for (uint32_t o = ob; o < oe; ++o) apply_op(ops + o, mats, v, g, rg);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const uint32_t j = jl | ((k & 1) ? rl[0] : 0u) | ((k & 2) ? rl[1] : 0u) | ((k & 4) ? rl[2] : 0u);
                    tile[swz(j)] = v[k];
                }
            }
            __syncthreads();
        }

        // ---- store back in place ----------------------------------------------------------
        for (uint32_t j = tid; j < tile_len; j += TILE_THREADS) {
            const uint64_t gidx = base | s_hoff[j >> L] | (uint64_t)(j & lmask);
            amp *dst = s_seg[gidx >> shard_shift] + (gidx & shard_mask);
            *dst = tile[swz(j)];
        }
        __syncthreads();
    }
}

int tile_kernel_setup() {
    static bool done = false;
    static cudaError_t err = cudaSuccess;
    // per-device attribute; cheap enough to set on every device we see
    err = cudaFuncSetAttribute(k_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(amp) << TILE_MAX_BITS));
    done = true;
    (void)done;
    return err == cudaSuccess ? 0 : -1;
}

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const TOp *d_ops, const amp *mat_table, int sm_count) {
    if (hdr.T < TILE_MIN_BITS || hdr.T > TILE_MAX_BITS || hdr.L > hdr.T || hdr.T - hdr.L > TILE_MAX_HIGH ||
        hdr.n_tiles == 0)
        return -1;
    const size_t smem = sizeof(amp) << hdr.T;
    int per_sm = (int)((200u * 1024u) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > hdr.n_tiles) grid = hdr.n_tiles;
    k_tile_pass<<<(unsigned)grid, TILE_THREADS, smem, st>>>(segs, hdr, d_stages, d_ops, mat_table);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace qv
