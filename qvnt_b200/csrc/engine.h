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
// One pass = one in-place sweep over the shard carrying MANY SingleOps.  A tile is the
// set of 2^T amplitudes obtained by varying T "tile bits" of the global index: the low L
// bits (one contiguous 16*2^L-byte chunk in HBM) plus T-L gathered high bits, which may be
// rank bits of a sharded register (the chunk then lives in a peer GPU's HBM, reached over
// NVLink through the mapped peer pointer).  The CTA stages the tile in shared memory
// (cp.async, XOR-swizzled), runs the pass's stages on it and writes it back.
// A stage gives every thread 2^TILE_R amplitudes in registers (TILE_R "register bits"
// of the tile) and applies all of the stage's ops to them before touching shared memory
// again; ops whose partner bits are not register bits wait for a later stage.
constexpr int TILE_R = 3;
constexpr int TILE_MAX_BITS = 12;
constexpr int TILE_MIN_BITS = 4;
constexpr int TILE_MAX_HIGH = 6;   // gathered (non-contiguous) tile bits: T - L <= 6
constexpr int TILE_THREADS = 256;

enum TForm : uint32_t {
    TF_DIAG = 0,    // z/s/t/rz/rzz: no partner
    TF_PAIR1 = 1,   // x/y/rx/ry/h1/u1 on one bit: partner = i ^ bit(ra)
    TF_PAIR2X = 2,  // rxx/ryy: partner = i ^ bit(ra) ^ bit(rb)
    TF_ODD2 = 3,    // swap family: odd-parity pair {bit(ra) set, bit(rb) set}
    TF_QUAD = 4     // h2/u2: a = bit(ra), b = bit(rb)
};

struct TOp {          // 64 bytes
    DevOp d;          // masks in GLOBAL numbering (signs / phases / controls test the global index)
    uint32_t form;
    uint8_t ra, rb;   // register-bit index (0..TILE_R-1) of the op's partner bit(s)
    uint8_t _p[2];
};

struct TStage {       // 32 bytes
    uint32_t op_begin, op_end;   // range in the pass's TOp array
    uint8_t r_lpos[4];           // register bit j -> tile-local bit position
    uint8_t t_lpos[16];          // thread bit k   -> tile-local bit position (T - TILE_R entries)
    uint32_t _pad;
};

struct TPassHdr {
    uint32_t T, L;               // tile bits, contiguous low bits
    uint32_t n_stages;
    uint32_t stage_begin;        // first stage of this pass in the uploaded stage array
    uint8_t gpos[16];            // tile-local bit -> global bit position (gpos[l] = l for l < L)
    Fixed fx;                    // tile counter -> LOCAL base index (tile-local bits 0, ownership bits fixed)
    uint64_t n_tiles;            // tiles this rank processes
    uint64_t base_or;            // this rank's bits for the global qubits that are NOT tile bits
    uint32_t touches_peer;       // some tile bit is a rank bit
    uint32_t _pad;
};

int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const TOp *d_ops, const amp *mat_table, int sm_count);
int tile_kernel_setup();

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
