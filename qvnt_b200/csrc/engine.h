// engine.h -- internal interfaces between the kernel translation units and the C ABI.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

#include "common.cuh"

namespace qv {

// ---- direct sweep kernels (direct.cu) ---------------------------------------
// Enumerates work items t in [0, items) and scatters them over the index bits
// that are NOT fixed: positions pos[0..n) (ascending) are skipped and filled
// from `val`.  Used to walk only the amplitudes a gate can change
// (ctrl bits = 1, pivot bit = 0, ...).
struct Fixed {
    uint32_t n;
    uint32_t _pad;
    uint64_t val;
    uint8_t pos[64];
};

// One SingleOp as one in-place sweep over this GPU's shard.  `op` masks are
// local (global control bits already resolved by the caller); `idx_or` carries
// the rank bits so diagonal phases / sign choices see the global index.
// Returns the number of kernel launches enqueued (0 if nothing to do), <0 on error.
int launch_direct(cudaStream_t st, amp *psi, uint32_t n_local, const DevOp &op, const amp *mat_table,
                  uint64_t idx_or, uint64_t *touched_amps);

// ---- fused tile pass (tile.cu) ----------------------------------------------
constexpr int TILE_MAX_BITS = 13;
constexpr int TILE_MAX_HIGH = 13;

// Tile-local form of one op inside a pass.  Bits of the tile are numbered
// 0..T-1 (local index j); everything outside the tile is constant per tile and
// evaluated against the tile's base index.
struct TileOp {
    uint32_t kind;
    uint32_t dagger;
    uint32_t cls;        // OpClass
    uint32_t mix;        // pair: local XOR mask; quad: local a bit mask
    uint32_t mix_b;      // quad: local b bit mask
    uint32_t pivot;      // pair: highest local bit of mix (index); quad: unused
    uint32_t ctrl_in;    // control bits that are tile bits (local numbering)
    uint32_t a_in;       // a_mask bits inside the tile (local numbering)
    uint64_t ctrl_out;   // control bits outside the tile (global numbering)
    uint64_t a_out;      // a_mask bits outside the tile (global numbering)
    uint64_t a_glob;     // full a_mask (global numbering) -- popcount for y
    double ph_re, ph_im;
    uint32_t mat;        // matrix table offset
    uint32_t sync_after; // 1: __syncthreads() needed after this op
};

struct TilePass {
    uint32_t T;                       // tile bits
    uint32_t chunk_bits;              // contiguous low bits of the tile (one bulk copy each)
    uint32_t n_high;                  // T - chunk_bits gathered bits
    uint8_t high_pos[TILE_MAX_HIGH];  // their global positions, ascending
    // tile enumeration: counter c -> base index via Fixed (tile bits + ownership bits fixed)
    Fixed fx;
    uint64_t n_tiles;                 // tiles this rank processes
    uint32_t op_begin, op_end;        // range in the pass's TileOp array
    uint32_t touches_peer;            // some tile bit is a rank bit
};

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TilePass &pass, const TileOp *d_ops,
                     const amp *mat_table, int use_tma, int sm_count);
size_t tile_smem_bytes(uint32_t T);

// ---- measurement / utility kernels (measure.cu) -----------------------------
constexpr int REDUCE_BLOCKS_MAX = 4096;
// sum |a|^2 over psi[0..len): deterministic two-stage reduction; result in *d_out.
int launch_norm_sqr(cudaStream_t st, const amp *psi, uint64_t len, double *d_partials, double *d_out,
                    int sm_count);
int launch_probabilities(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, double inv,
                         double *d_out);
int launch_polar(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, double *d_out);
// zero where ((i | idx_or) ^ idy) & mask != 0
int launch_collapse(cudaStream_t st, amp *psi, uint64_t len, uint64_t idx_or, uint64_t idy, uint64_t mask);
// zero where (i | idx_or) & mask != 0
int launch_zero_mask(cudaStream_t st, amp *psi, uint64_t len, uint64_t idx_or, uint64_t mask);
int launch_scale(cudaStream_t st, amp *psi, uint64_t len, double f);
int launch_set_basis(cudaStream_t st, amp *psi, uint64_t len, uint64_t one_at /* >= len: none */);
int launch_tensor_prod(cudaStream_t st, const amp *a, uint32_t qa, const amp *b, uint32_t qb, amp *out,
                       uint64_t out_off, uint64_t out_len);
// measurement sampling, blocked-sequential cumulative order (see measure.cu)
constexpr uint64_t SAMPLE_BLOCK = 1ull << 12;
int launch_block_weights(cudaStream_t st, const amp *psi, uint64_t len, double inv, double *d_l1,
                         double *d_l2);
int launch_total(cudaStream_t st, const double *d_l2, uint64_t n2, double *d_out);
int launch_locate(cudaStream_t st, const amp *psi, uint64_t len, double inv, const double *d_l1,
                  uint64_t n1, const double *d_l2, uint64_t n2, double prefix, double x,
                  uint64_t *d_result /* [0]=index, [1]=found, [2]=bits of running sum */);

}  // namespace qv
