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
// A stage gives every thread 2^TILE_R amplitudes in registers (TILE_R "register slots",
// each holding one tile bit) and applies all of the stage's ops to them before touching
// shared memory again; ops whose partner bits are not in register slots wait for a later
// stage.  Everything index-dependent an op needs (controls, diagonal-phase parities) is
// split on the host into a register-slot part (compile-time per slot), a thread part
// (one 32-bit test per op per thread) and a tile part (one flag per op per tile), so the
// inner loop is FP64 arithmetic plus a handful of integer instructions.
constexpr int TILE_R = 4;
constexpr int TILE_NV = 1 << TILE_R;
constexpr int TILE_MAX_BITS = 12;
constexpr int TILE_MIN_BITS = 4;
// Defaults measured on configs[1] (tools/exp.sh): 2^11-amplitude tiles (32 KiB) made of 2^4-amplitude
// chunks (256 B) + 7 gathered bits, 4 CTAs of 128 threads per SM -> 8.0k gates/s; 2^12 tiles with
// 2 CTAs of 256 threads reach 7.6k (fewer passes, but half the warps to hide latency with).
constexpr int TILE_DEFAULT_BITS = 11;
constexpr int TILE_DEFAULT_CHUNK = 4;
constexpr int TILE_MAX_HIGH = 8;   // gathered (non-contiguous) tile bits: T - L <= 8
constexpr int TILE_THREADS = 256;  // = 2^(TILE_MAX_BITS - TILE_R)
constexpr int TILE_MAX_OPS = 2048; // SingleOps per pass (their descriptors live in shared memory)
constexpr int TILE_MAX_STAGES = 64;

enum TForm : uint32_t {
    TF_DIAG = 0,    // z/s/t/rz/rzz: no partner
    TF_PAIR1 = 1,   // x/y/rx/ry/h1/u1 on one bit: partner differs in slot ra
    TF_PAIR2X = 2,  // rxx/ryy: partner differs in slots ra and rb
    TF_ODD2 = 3,    // swap family: odd-parity pair {slot ra set, slot rb set}
    TF_QUAD = 4,    // h2/u2: a = slot ra, b = slot rb
    TF_LAZYX = 5,   // x between threads: no register slot
    TF_LAZYI = 7,   // x on a register slot, no control in a slot: the slot is marked inverted
    TF_GROUP = 6    // header of a merged diagonal run (not an op of its own)
};

// Dispatch codes of the stage interpreter (tile.cu).  Two-bit ops sit on the slot pairs
// (0,1) or (2,3) only; the planner assigns slots accordingly.
enum MCode : uint8_t {
    MC_DU = 0,    // + {Z,S,T,RZ,RZZ}: diagonal, no target bit in a register slot (one factor per thread)
    MC_DG = 5,    // + {Z,S,T,RZ,RZZ}: diagonal, per-slot factor
    MC_P1 = 10,   // + 4*{X,Y,RX,RY,H1,U1} + slot
    MC_P2X = 34,  // + 2*{RXX,RYY} + pair
    MC_ODD = 38,  // + 2*{SWAP,ISWAP,SQRT_SWAP,SQRT_ISWAP} + pair
    MC_H2 = 46,   // + pair
    MC_U2 = 48,   // + 2*pair + (a sits in the odd slot)
    MC_COUNT = 52
};

// Dispatch codes of the FAST stage interpreter (tile.cu): passes made only of the common kinds
// (x, y, rx, ry, h1 on one bit; z/s/t on one bit; rz; rzz) are lowered by the planner to a few
// coefficient-driven forms, so one arm serves several kinds:
//   FC_PR  real pair      new0 = c0*p0 + c1*p1 ; new1 = c2*p0 + c3*p1            (ry, h1)
//   FC_PX  crossed pair   new0 = c0*p0 - i*c1*p1 ; new1 = -i*c2*p0 + c3*p1        (rx, y)
//   FC_DU / FC_DS / FC_DG diagonal: p *= (parity of i & target ? f1 : f0), f0 = (c0,c1), f1 = (c2,c3);
//          DU: no target bit in a register slot; DS: exactly one (slot j); DG: any a_reg
//   FC_SW  x whose target AND a control sit in register slots: the selected pairs trade places
//   FC_LX / FC_LI  "lazy" x: a permutation that costs no data movement, see below
// The codes are small and dense so the interpreter's switch is one jump table.  INVARIANT: every code
// whose arm works on ONE register slot j has (code & 3) == j, in every variant (masked, control class,
// single-control, butterfly): the op loop's prologue takes the slot's inversion byte -- hence the
// coefficient block -- from the two low bits of the code byte.  So FC_MASKED and FC_TOTAL are
// multiples of 4.
enum FCode : uint8_t {
    FC_PR = 0,    // + slot
    FC_PX = 4,    // + slot
    FC_DS = 8,    // + slot
    FC_DU = 12,
    FC_DG = 13,
    FC_LX = 14,   // lazy x, target on a thread bit: target and controls on thread / outer bits (a_thr =
                  //         target's thread bit, a_reg = its tile-local position | position in the shard << 8);
                  //         the thread flips the bit in the index it will STORE its amplitudes to
    FC_LI = 15,   // lazy x, target in register slot j (a_reg = j), controls on thread /
                  //         outer bits: the thread marks slot j as inverted -- its register K now holds the
                  //         amplitude of slot pattern K ^ (1 << j).  Later ops on that slot read their
                  //         coefficients from the descriptor's `alt` block (roles of the pair exchanged),
                  //         and the stage's stores go to the exchanged addresses.  3 instructions.
    FC_DM = 16,   // header of a run of a_reg diagonal ops that share their controls and have no target
                  //         bit in a register slot: their factors are multiplied into ONE complex number
                  //         per thread, applied to the 16 amplitudes once
    FC_COUNT = 17,
    FC_MASKED = 20, // added to FC_PR / FC_PX / FC_DS / FC_DU / FC_DG / FC_DM codes when a control sits in a
                    // register slot (okmask != 0xFFFF): per-slot-pattern predicates
    FC_SW = 40,   // + slot (always masked)
    FC_TOTAL = 44,
    // The code byte of a fast MOp also carries the op's CONTROL CLASS: code = arm + FC_TOTAL * cls,
    // cls 0: unconditional, 1: controls on thread bits, 2: ... and outside the tile (MOP_COND /
    // MOP_CONDB say the same).  The PTX op loop jumps on the whole byte (class 1 lands in a stub that
    // tests the controls first; class 2 is rewritten per tile, tile.cu patch_codes), so
    // unconditional ops never pay for the test.
    //
    // SINGLE-CONTROL arms (bytes 3 * FC_TOTAL .. FC_SPECIAL_END - 1, class 0 only): diagonal forms whose
    // ONLY control sits in register slot c (qft's controlled phases).  The generic masked arms test a
    // predicate per slot pattern -- ptxas turns that into all the arithmetic plus a select per result,
    // 3x the instructions of the unmasked arm; these touch exactly the 8 patterns with bit c set.
    FC_DS1 = 132, // + 4 * c + j: FC_MASKED + FC_DS + j with okmask == "slot bit c set"  (code & 3 == j: the prologue
                  //   of the op loop takes the target slot's inversion byte from the two low bits of the code)
    FC_DU1 = 148, // + c: FC_MASKED + FC_DU
    FC_DM1 = 152, // + c: FC_MASKED + FC_DM
    // BUTTERFLY h (class 0, no control anywhere): FC_PR + j with coefficients (1, 1, 1, -1) -- the arm adds and
    // subtracts (32 FP64 instructions instead of 64) and ignores the coefficient block.  The gate's 1/sqrt(2) is a
    // factor on EVERY amplitude, so the planner multiplies the factors of a pass's butterflies into the
    // coefficients of one unconditional pair op of the same pass (planner.cu, flush).
    FC_HB = 156,  // + j
    FC_SPECIAL_END = 160
};
// the generic (masked) code a single-control byte stands for; any other byte unchanged
__host__ __device__ inline uint32_t fc_generic(uint32_t byte) {
    if (byte < (uint32_t)FC_DS1 || byte >= (uint32_t)FC_SPECIAL_END) return byte;
    if (byte < (uint32_t)FC_DU1) return (uint32_t)(FC_MASKED + FC_DS) + ((byte - (uint32_t)FC_DS1) & 3u);
    if (byte < (uint32_t)FC_DM1) return (uint32_t)(FC_MASKED + FC_DU);
    if (byte < (uint32_t)FC_HB) return (uint32_t)(FC_MASKED + FC_DM);
    return (uint32_t)FC_PR + (byte - (uint32_t)FC_HB);
}
constexpr uint8_t MOP_SKIP0 = 0x02;   // flags bit 1 (diagonal forms): f0 == 1, even parity untouched
constexpr uint8_t MOP_COND = 0x04;    // flags bit 2: the op has controls on thread bits or outside the tile
                                      //   (ctrl_thr != 0 or ctrl_base != 0): test before dispatch
constexpr uint8_t MOP_PARB = 0x08;    // flags bit 3 (diagonal forms): target bits outside the tile (a_base != 0):
                                      //   the per-tile flag byte carries their parity
constexpr uint8_t MOP_CONDB = 0x10;   // flags bit 4: ... and some of them outside the tile (ctrl_base != 0): read the flag byte
constexpr uint8_t MOP_ATHR = 0x20;    // flags bit 5 (diagonal forms): target bits on thread bits (a_thr != 0)
constexpr uint8_t MOP_STATIC = 0x40;  // flags bit 6 (FC_DM header): every member's target is either on thread bits or outside the
                                      //   tile and no lazy x precedes the run in its stage: the per-thread factor of the thread-bit
                                      //   members is tabulated once per kernel (slot = header's a_thr), the factor of the
                                      //   outside members once per tile (MOP_PARB on the header: there are such members)
constexpr int TILE_MAX_STATIC = 12;   // tabulated runs per pass (2 KiB of shared memory each at 128 threads)
constexpr uint32_t MOP_END = 255;     // code of the sentinel descriptor the kernel puts behind every stage in shared memory
constexpr uint32_t MOP_NOP = 254;     // PTX op loop: code the kernel writes over a class-2 op whose controls outside the tile are
                                      //   not satisfied in the CURRENT tile (a satisfied one gets its class 0 / 1 code)
constexpr uint32_t MOP_NOP_RUN = 253; // ... the same for a merged run's header: its members are skipped with it
constexpr uint32_t MOP_ALT_BYTES = 32; // byte distance from the coefficient block to the `alt` block

struct __align__(16) MOp {   // 80 bytes, staged in shared memory
    uint8_t code;        // MCode (full interpreter) or FCode (fast interpreter)
    uint8_t dagger;      // bit 0: dagger (full); fast: MOP_* flags
    uint16_t okmask;     // bit K: register slot pattern K satisfies the controls held in register slots
    uint32_t ctrl_thr;   // controls on thread bits (bit k = thread bit k of the stage)
    uint32_t a_thr;      // diagonal class: target-mask bits on thread bits; u1/u2: matrix table index
    uint16_t a_reg;      // diagonal class: target-mask bits on register slots (LX / LI / DM: see FCode)
    uint16_t idx;        // the op's index inside its pass (per-tile flag byte)
    double ph_re, ph_im; // full: (cos t/2, sin t/2); fast: c0, c1
    double c2, c3;       // fast only
    double alt[4];       // fast only: the coefficients to use while the op's register slot is inverted
                         // (FC_LI): pair forms (c3, c2, c1, c0); FC_DS (c2, c3, c0, c1)
};
static_assert(sizeof(MOp) == 80, "MOp layout");
constexpr uint32_t MOP_BYTES = 80;

struct MBase {           // per-op masks over the index bits that are NOT tile bits (global numbering)
    uint64_t ctrl_base;
    uint64_t a_base;     // diagonal class only
};

struct TStage {       // 32 bytes
    uint32_t op_begin, op_end;   // range in the uploaded MOp array
    uint8_t r_lpos[4];           // register slot j -> tile-local bit position
    uint8_t t_lpos[16];          // thread bit k    -> tile-local bit position (T - TILE_R entries)
    uint32_t sync_after_load;    // the stage holds a lazy x: threads store to OTHER threads' shared-memory
                                 // slots, so every thread must have loaded before any thread stores
};
static_assert(sizeof(TStage) == 32, "TStage layout");

struct TPassHdr {
    uint32_t T, L;               // tile bits, contiguous low bits
    uint32_t n_stages;
    uint32_t stage_begin;        // first stage of this pass in the uploaded stage array
    uint32_t op_begin, n_ops;    // this pass's ops in the uploaded MOp / MBase arrays
    uint8_t gpos[16];            // tile-local bit -> global bit position (gpos[l] = l for l < L)
    // tile counter -> LOCAL base index (tile-local bits 0, ownership bits fixed): runs of
    // fixed positions, ascending; base = expand(counter) | fx_val
    uint32_t n_runs;
    uint8_t run_pos[16], run_len[16];
    uint64_t fx_val;
    uint64_t n_tiles;            // tiles this rank processes
    uint64_t base_or;            // this rank's bits for the global qubits that are NOT tile bits
    uint32_t touches_peer;       // some tile bit is a rank bit
    // Remap pass: one tile bit is the rank bit remap_g and the pinned local bit is remap_b; the pass
    // stores BOTH values of g into this rank's shard, at bit position b -- afterwards index bits g
    // and b have traded places (the planner's logical -> physical qubit map records it).  Loads
    // read the peer half through NVLink as in any peer pass; stores are all local, into places the
    // PEER's CTA of the same tile reads, hence the per-tile handshake (tile.cu).
    uint8_t remap, remap_g, remap_b, _pad3;
    uint32_t epoch;              // remap handshake value of this pass (filled at launch)
    uint8_t gpos_store[16];      // tile-local bit -> bit position the last stage stores it to (== gpos unless remap)
    uint32_t full;               // some op needs the full interpreter (u1/u2, two-bit pair ops, multi-bit masks)
    uint32_t need_flags;         // some op of the pass depends on index bits outside the tile (per-tile flag bytes)
    uint32_t n_static;           // tabulated diagonal runs (MOP_STATIC headers) of the pass
    uint32_t prefetch;           // filled by launch_tile_pass (TileKnobs): L2 prefetch of the next tile's chunks
    uint32_t uses_sc;            // some op of the pass has a single-control code (FC_DS1 ...): the kernel flavour with those arms
    uint64_t fixed_mask;         // local index bits the tile counter does NOT enumerate (tile bits + ownership bits)
    uint16_t stage_end[TILE_MAX_STAGES];   // ops of stage s: [stage_end[s-1], stage_end[s]) relative to op_begin
    Fixed fx;                    // the same enumeration bit by bit (host side: describe / tests)
};

// Per-handle tuning knobs of the tile pass (qvnt_reg_set_option): no process-global state.
struct TileKnobs {
    int ctas_per_sm = 0;   // 0 = auto (3 x 128 threads at 168 registers for 2^11 tiles: no spills; 2 x 256 for 2^12); 4: forced
    int bulk = -1;         // tile loads: 1 = cp.async.bulk (TMA) + mbarrier, 0 = 16-byte cp.async, -1 = auto: bulk copies
                           //    for passes that read a peer shard (fewer, larger NVLink requests: measured +7 % at
                           //    2 GPUs), cp.async for local passes (a 256-byte bulk copy costs its issuing lane ~17
                           //    instructions: measured -8 % on one GPU)
    int prefetch = 0;      // 1: cp.async.bulk.prefetch.L2 of tile i+1 while tile i computes (measured: no gain, off)
    int double_buffer = 0; // 1: two tile buffers per CTA (2^11 tiles); 2: only for passes that read a peer shard
    int single_ctrl = 1;   // planner: single-control arms (FC_DS1 / FC_DU1 / FC_DM1) instead of the generic masked ones
    int lower_two_bit = 0; // planner: swap / i_swap / rxx / ryy of an op list as products of fast kinds (no full-interpreter
                           //   pass); off: measured slower on configs[4] (planner.cu lower_ops)
    int butterfly = 1;     // planner: uncontrolled h as add / subtract (FC_HB), its scale folded into another op of the pass
    int ptx_ops = 1;       // 1: the fast interpreter's op loop as one inline-PTX block (fastops_ptx.inc); 0: C++ loop
};
int launch_tile_pass(cudaStream_t st, const Segs &segs, const TPassHdr &hdr, const TStage *d_stages,
                     const MOp *d_ops, const MBase *d_bases, const amp *mat_table, int sm_count,
                     const TileKnobs &knobs);
int tile_kernel_setup();   // once per process and device (qvnt_reg_create)

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
int launch_combine_unitary(cudaStream_t st, const amp *a, const amp *b, uint32_t q, const double *m8, amp *out);
int launch_linear_composition(cudaStream_t st, amp *self, const amp *other, uint64_t len, amp c0, amp c1);
// sample_all: sum over i of sqrt(p_i) g_i -> *d_out; then the counts of [off, off + cnt) and their sum
int launch_sample_noise_sum(cudaStream_t st, const amp *psi, uint64_t len, uint64_t idx_or, double inv, uint64_t seed,
                            double *d_partials, double *d_out, int sm_count);
int launch_sample_counts(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, uint64_t idx_or, double inv,
                         uint64_t seed, double c, double n_sum, unsigned long long *d_out, unsigned long long *d_total);
// measurement sampling, blocked-sequential cumulative order (see measure.cu)
constexpr uint64_t SAMPLE_BLOCK = 1ull << 12;
int launch_block_weights(cudaStream_t st, const amp *psi, uint64_t len, double inv, double *d_l1,
                         double *d_l2);
int launch_total(cudaStream_t st, const double *d_l2, uint64_t n2, double *d_out);
int launch_locate(cudaStream_t st, const amp *psi, uint64_t len, double inv, const double *d_l1,
                  uint64_t n1, const double *d_l2, uint64_t n2, double prefix, double x,
                  uint64_t *d_result /* [0]=index, [1]=found, [2]=bits of running sum */);

}  // namespace qv
