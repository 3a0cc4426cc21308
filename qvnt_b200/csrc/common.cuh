// common.cuh -- shared types for the sm_100a state-vector engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qvnt_b200.h"

namespace qv {

typedef double2 amp;  // complex f64, x = re, y = im (16 B, interleaved like num_complex::Complex<f64>)

constexpr int MAX_WORLD = 8;

// Where the shards of the (possibly multi-GPU) register live.  A global amplitude
// index i resolves to seg[i >> shift] + (i & lmask).  world == 1: shift = q_num.
struct Segs {
    amp *seg[MAX_WORLD];
    unsigned int *ack[MAX_WORLD];   // per-tile handshake words of every shard (remap passes, tile.cu)
    uint32_t shift;      // local index bits n_local
    uint32_t rank;       // this GPU's rank
    uint32_t world_bits; // log2(world)
    uint32_t _pad;
};

__device__ __forceinline__ amp *resolve(const Segs &s, uint64_t i) {
    return s.seg[i >> s.shift] + (i & ((1ull << s.shift) - 1ull));
}

// Device-side gate descriptor in GLOBAL index space (64 B header + matrix in a side table).
struct DevOp {
    uint32_t kind;
    uint32_t dagger;
    uint64_t a;       // a_mask / ab_mask
    uint64_t b;       // b_mask (h2/u2)
    uint64_t ctrl;
    double ph_re, ph_im;
    uint32_t mat;     // index (in amp units) into the matrix table (u1: 4, u2: 16 entries)
    uint32_t _pad;
};

// Gate classes (how many amplitudes one output couples).
enum OpClass { CLS_DIAG = 0, CLS_PAIR = 1, CLS_QUAD = 2, CLS_NONE = 3 };

__host__ __device__ inline int op_class(uint32_t kind) {
    switch (kind) {
    case QVNT_ID: return CLS_NONE;
    case QVNT_Z: case QVNT_S: case QVNT_T: case QVNT_RZ: case QVNT_RZZ: return CLS_DIAG;
    case QVNT_H2: case QVNT_U2: return CLS_QUAD;
    default: return CLS_PAIR;
    }
}
// Swap family: identity on the even-parity subspace of ab_mask.
__host__ __device__ inline bool op_odd_only(uint32_t kind) {
    return kind == QVNT_SWAP || kind == QVNT_ISWAP || kind == QVNT_SQRT_SWAP || kind == QVNT_SQRT_ISWAP;
}

#define QV_FRAC_1_SQRT_2 0.70710678118654752440084436210485

}  // namespace qv
