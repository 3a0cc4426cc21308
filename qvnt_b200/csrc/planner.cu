// planner.cu -- host-side scheduler of qvnt_reg_apply: validates the op list, packs
// SingleOps into fused tile passes and enqueues the kernels.  Replaces the reference's
// "one full out-of-place sweep per SingleOp" driver: QReg::apply
// (src/register/quant.rs:376-395) -> MultiOp::apply (src/operator/multi/mod.rs:96-114)
// -> SingleOp::apply (src/operator/single/mod.rs:83-93).
//
// Scheduling.  Every op has MIX bits (qubits whose basis state it changes: the XOR
// partners of its gather) and DIAGONAL bits (controls and the targets of z/s/t/rz/rzz:
// the op is block-diagonal in them).  Two ops commute when every shared qubit is a
// diagonal bit of both.  A pass is built greedily in list order: an op joins the pass if
// it commutes with every earlier op that was left behind and its mix bits fit the tile
// (low L bits are always tile bits, T-L more can be gathered -- rank bits of a sharded
// register included, which turns the pass into the NVLink exchange).  Diagonal ops and
// controls never constrain the tile.  Inside a pass the same greedy rule splits the ops
// into stages by the TILE_R bits each thread keeps in registers.  Only the order of commuting
// ops ever changes, so the result equals the op-by-op order up to f64 rounding
// (|diff| ~ 1e-16, parity bar 1e-10); with option "fuse" = 0 every op is its own sweep
// and the result is bit-identical to the reference arithmetic.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "reg.h"

namespace qv {

static inline int pc64(uint64_t v) { return __builtin_popcountll(v); }
static inline int ctz64(uint64_t v) { return __builtin_ctzll(v); }

struct POp {
    DevOp d;        // masks in PHYSICAL index-bit numbering (== qubit numbering until a remap pass moves a qubit)
    uint64_t la, lb, lctrl;   // the same masks in LOGICAL (qubit) numbering, as the caller gave them
    int cls;
    uint64_t mix;   // bits that must lie inside a tile / in registers
    uint64_t dg;    // bits the op is diagonal in (controls, diagonal targets)
    uint32_t src;   // index of the SingleOp in the caller's array
};

struct PlanCfg {
    uint32_t q_num, n_local, rank, world;
    bool peers, fuse;
    bool remap = true;       // global-qubit passes leave the qubit local (logical -> physical map), see build_plan
    int tile_bits, chunk_bits;
    uint8_t perm[64];        // logical qubit -> physical index bit at the start of the op list
    uint64_t ack_cap = ~0ull; // tiles the per-tile handshake array of a remap pass can hold
    int peer_chunk_bits = 0;  // chunk size of passes that START on a global qubit (0 = same as chunk_bits): larger
                              // contiguous runs make better NVLink requests
    int peer_tile_bits = 0;   // tile size of those passes (0 = same as tile_bits): more gathered bits carry more
                              // gates per NVLink exchange
    int force_pinned = -1;    // restore_layout: the local bit a remap pass must trade the global bit with
    bool single_ctrl = true;  // single-control arms for diagonal forms with one control in a register slot
    bool butterfly = true;    // uncontrolled h as a butterfly (FC_HB), its scale folded into another op of the pass
    bool lower_two_bit = false; // swap / i_swap / rxx / ryy in op lists as products of kinds the fast interpreter has
                                // (option; off: measured slower on the 36-qubit QASM circuit, see lower_ops)
    uint64_t q_mask() const { return q_num >= 64 ? ~0ull : ((1ull << q_num) - 1ull); }
};

static int validate(const PlanCfg &r_, const qvnt_op_t &o, size_t k) {
    struct { uint64_t q_mask; uint32_t q_num; } rr{r_.q_mask(), r_.q_num}, *r = &rr;
    if (o.kind >= QVNT_KIND_COUNT) {
        set_error("op %zu: unknown kind %u", k, o.kind);
        return QVNT_ERR_INVALID;
    }
    const uint64_t all = o.a_mask | o.b_mask | o.ctrl;
    if (all & ~r->q_mask) {
        set_error("op %zu (kind %u): mask 0x%llx addresses qubits outside the %u-qubit register", k, o.kind,
                  (unsigned long long)all, r->q_num);
        return QVNT_ERR_BAD_MASK;
    }
    int need_a = -1, need_b = 0;
    switch (o.kind) {
    case QVNT_RX: case QVNT_RY: case QVNT_RZ: case QVNT_U1: case QVNT_H1: need_a = 1; break;
    case QVNT_RXX: case QVNT_RYY: case QVNT_RZZ: case QVNT_SWAP: case QVNT_ISWAP: case QVNT_SQRT_SWAP:
    case QVNT_SQRT_ISWAP: need_a = 2; break;
    case QVNT_H2: case QVNT_U2: need_a = 1; need_b = 1; break;
    default: break;
    }
    if ((need_a >= 0 && pc64(o.a_mask) != need_a) || (need_b && (pc64(o.b_mask) != 1 || o.a_mask == o.b_mask))) {
        set_error("op %zu (kind %u): invalid target mask (is_valid() of the reference op fails)", k, o.kind);
        return QVNT_ERR_INVALID;
    }
    const uint64_t act = o.a_mask | (need_b ? o.b_mask : 0);
    if (o.ctrl & act) {
        set_error("op %zu (kind %u): control mask overlaps the target mask", k, o.kind);
        return QVNT_ERR_INVALID;
    }
    return QVNT_OK;
}

// ---- lowering: qvnt_op_t -> POp ------------------------------------------------------------
static int lower_ops(const PlanCfg &r, const qvnt_op_t *ops, size_t n_ops, std::vector<POp> &pl,
                     std::vector<amp> &mats) {
    const bool tile_possible = r.n_local >= (uint32_t)TILE_MIN_BITS;
    // a lone SingleOp keeps the reference's exact arithmetic (one direct sweep); op LISTS are
    // refactored for the fused pass (exact factorisations, FMA arithmetic: parity bar 1e-10)
    const bool split_fused = r.fuse && tile_possible && n_ops > 1;
    pl.reserve(n_ops + 16);
    for (size_t k = 0; k < n_ops; ++k) {
        int rc = validate(r, ops[k], k);
        if (rc) return rc;
        const qvnt_op_t &o = ops[k];
        POp p;
        memset(&p, 0, sizeof(p));
        p.src = (uint32_t)k;
        p.cls = op_class(o.kind);
        if (p.cls == CLS_NONE) continue;                     // Id
        if (p.cls == CLS_PAIR && o.a_mask == 0) continue;    // x(0) / y(0): identity (i^0)
        if (p.cls == CLS_DIAG && o.a_mask == 0 && o.kind != QVNT_RZ && o.kind != QVNT_RZZ) continue;
        p.d.kind = o.kind;
        p.d.dagger = o.dagger ? 1u : 0u;
        p.d.a = o.a_mask;
        p.d.b = (o.kind == QVNT_H2 || o.kind == QVNT_U2) ? o.b_mask : 0;
        p.d.ctrl = o.ctrl;
        p.d.ph_re = o.phase_re;
        p.d.ph_im = o.phase_im;
        if (o.kind == QVNT_H1) p.d.ph_re = QV_FRAC_1_SQRT_2;     // the butterfly's scale (h1.rs:21)
        if (o.kind == QVNT_U1 || o.kind == QVNT_U2) {
            const int cnt = o.kind == QVNT_U1 ? 4 : 16;
            p.d.mat = (uint32_t)mats.size();
            for (int i = 0; i < cnt; ++i) mats.push_back(make_double2(o.matrix[2 * i], o.matrix[2 * i + 1]));
        }
        const bool split_masks = (r.fuse && tile_possible) || r.world > 1;
        if (split_masks && (o.kind == QVNT_X || o.kind == QVNT_Y) && pc64(o.a_mask) > 1) {
            // x(m) = prod_b x(b), y(m) = prod_b y(b): permutations and i-power sign flips only,
            // so the factorisation is exact (y: i^(2*ones-k), atomic/y.rs:10-23).
            uint64_t m = o.a_mask;
            while (m) {
                POp q = p;
                q.d.a = m & (~m + 1);
                q.mix = q.d.a;
                q.dg = q.d.ctrl;
                pl.push_back(q);
                m &= m - 1;
            }
            continue;
        }
        if (split_fused && (o.kind == QVNT_Z || o.kind == QVNT_S || o.kind == QVNT_T) &&
            pc64(o.a_mask) > 1) {
            // z/s/t(m) = prod_b z/s/t(b): the phase of an amplitude is w^popcount(i & m)
            // (z.rs:15-21, s.rs:19-25, t.rs:24-35), a product of one factor per set bit.
            uint64_t m = o.a_mask;
            while (m) {
                POp q = p;
                q.d.a = m & (~m + 1);
                q.mix = 0;
                q.dg = q.d.ctrl | q.d.a;
                pl.push_back(q);
                m &= m - 1;
            }
            continue;
        }
        if (split_fused && o.kind == QVNT_H2) {
            // h2(a, b) = h1(a) h1(b) (h2.rs:22-38 is the product of the two butterflies)
            for (int half = 0; half < 2; ++half) {
                POp q = p;
                q.d.kind = QVNT_H1;
                q.cls = CLS_PAIR;
                q.d.a = half ? o.b_mask : o.a_mask;
                q.d.b = 0;
                q.d.ph_re = half ? 0.5 : 1.0;      // h2.rs:37 scales the sum of four once, by 0.5 (exact powers of
                                                   // two: the golden vectors of quant.rs:653-673 stay bit-exact)
                q.d.ph_im = 0.0;
                q.mix = q.d.a;
                q.dg = q.d.ctrl;
                pl.push_back(q);
            }
            continue;
        }
        if (split_fused && r.lower_two_bit && pc64(o.a_mask) == 2 &&
            (o.kind == QVNT_SWAP || o.kind == QVNT_ISWAP || o.kind == QVNT_RXX || o.kind == QVNT_RYY)) {
            // Two-bit kinds the fast stage interpreter has no form for, as products of kinds it has (a pass
            // with one full-interpreter op runs the full interpreter for ALL its ops):
            //   swap(a, b)    = cx(a->b) cx(b->a) cx(a->b)                      (swap.rs:16-22; exact)
            //   i_swap(a, b)  = swap(a, b) * diag(1, i, i, 1),  diag = s(a) s(b) cz(a, b)   (i_swap.rs:20-37; exact)
            //   rxx(a, b)     = h(a) h(b) rzz(a, b) h(a) h(b)                   (X(x)X = (H(x)H)(Z(x)Z)(H(x)H))
            //   ryy(a, b)     = s(a) s(b) rxx(a, b) s^-1(a) s^-1(b)             (Y = S X S^-1)
            // Controls go on the middle factor only where the outer ones cancel without it.
            // OFF by default (option "lower_two_bit"): configs[4] at 36 qubits on 8 GPUs ran 1.98 s with it
            // against 1.66 s without (profiles/r02w_, r02u_sharded_36q_qasm_8gpu.json) -- three controlled
            // swaps of register slots cost more than the full interpreter's one permutation, and a cx whose
            // target is a global qubit asks for a remap pass of its own.
            const uint64_t ba = o.a_mask & (~o.a_mask + 1), bb = o.a_mask & ~ba;
            auto emit = [&](uint32_t kind, uint64_t a, uint64_t ctrl, uint32_t dagger, double re, double im) {
                POp q;
                memset(&q, 0, sizeof(q));
                q.src = (uint32_t)k;
                q.cls = op_class(kind);
                q.d.kind = kind;
                q.d.dagger = dagger;
                q.d.a = a;
                q.d.ctrl = ctrl;
                q.d.ph_re = re;
                q.d.ph_im = im;
                q.mix = q.cls == CLS_PAIR ? a : 0;
                q.dg = ctrl | (q.cls == CLS_DIAG ? a : 0);
                pl.push_back(q);
            };
            const uint32_t dg = o.dagger ? 1u : 0u;
            if (o.kind == QVNT_SWAP || o.kind == QVNT_ISWAP) {
                if (o.kind == QVNT_ISWAP) {
                    emit(QVNT_S, ba, o.ctrl, dg, 0.0, 0.0);
                    emit(QVNT_S, bb, o.ctrl, dg, 0.0, 0.0);
                    emit(QVNT_Z, bb, o.ctrl | ba, 0u, 0.0, 0.0);
                }
                emit(QVNT_X, bb, o.ctrl | ba, 0u, 0.0, 0.0);
                emit(QVNT_X, ba, o.ctrl | bb, 0u, 0.0, 0.0);
                emit(QVNT_X, bb, o.ctrl | ba, 0u, 0.0, 0.0);
            } else {
                if (o.kind == QVNT_RYY) {
                    emit(QVNT_S, ba, 0, 1u, 0.0, 0.0);
                    emit(QVNT_S, bb, 0, 1u, 0.0, 0.0);
                }
                emit(QVNT_H1, ba, 0, 0u, QV_FRAC_1_SQRT_2, 0.0);
                emit(QVNT_H1, bb, 0, 0u, QV_FRAC_1_SQRT_2, 0.0);
                emit(QVNT_RZZ, o.a_mask, o.ctrl, dg, o.phase_re, o.phase_im);
                emit(QVNT_H1, ba, 0, 0u, QV_FRAC_1_SQRT_2, 0.0);
                emit(QVNT_H1, bb, 0, 0u, QV_FRAC_1_SQRT_2, 0.0);
                if (o.kind == QVNT_RYY) {
                    emit(QVNT_S, ba, 0, 0u, 0.0, 0.0);
                    emit(QVNT_S, bb, 0, 0u, 0.0, 0.0);
                }
            }
            continue;
        }
        p.mix = p.cls == CLS_PAIR ? p.d.a : (p.cls == CLS_QUAD ? (p.d.a | p.d.b) : 0);
        p.dg = p.d.ctrl | (p.cls == CLS_DIAG ? p.d.a : 0);
        pl.push_back(p);
    }
    // logical (qubit) masks -> physical index bits under the register's current qubit map
    auto xl = [&](uint64_t m) {
        uint64_t o = 0;
        for (; m; m &= m - 1) o |= 1ull << r.perm[ctz64(m)];
        return o;
    };
    for (POp &p : pl) {
        p.la = p.d.a;
        p.lb = p.d.b;
        p.lctrl = p.d.ctrl;
        p.d.a = xl(p.d.a);
        p.d.b = xl(p.d.b);
        p.d.ctrl = xl(p.d.ctrl);
        p.mix = xl(p.mix);
        p.dg = xl(p.dg);
    }
    return QVNT_OK;
}

static inline uint64_t swap_bits(uint64_t m, int i, int j) {
    const uint64_t d = ((m >> i) ^ (m >> j)) & 1ull;
    return m ^ ((d << i) | (d << j));
}

// ---- one direct sweep -----------------------------------------------------------------------
static int run_direct(qvnt_reg *r, const POp &p) {
    const uint64_t lmask = r->local_len - 1;
    const uint64_t gmask = r->q_mask & ~lmask;
    const uint64_t rbits = (uint64_t)r->rank << r->n_local;
    r->stats.passes += 1;
    if ((p.d.ctrl & gmask) & ~rbits) return QVNT_OK;     // a control on a rank bit this GPU does not satisfy
    DevOp d = p.d;
    d.ctrl &= lmask;
    LaunchScope ls(r, 0);
    uint64_t touched = 0;
    int n = launch_direct(r->stream, r->psi, r->n_local, d, r->d_mat, rbits, &touched);
    ls.done(n);
    r->stats.alg_bytes[0] += touched * 32;
    r->stats.h2d_bytes += sizeof(DevOp) + sizeof(Fixed);      // kernel parameters
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "direct sweep launch");
        set_error("internal: direct sweep rejected op kind %u", d.kind);
        return QVNT_ERR_INVALID;
    }
    return QVNT_OK;
}

// ---- pass / stage construction --------------------------------------------------------------
struct PassPlan {
    TPassHdr hdr;
    std::vector<int> ops;      // indices into the POp list, in execution order
    bool direct = false;       // single op, local: run as a direct sweep
};

// Greedy selection in list order under a bit budget.  `always` bits are free; at most
// `budget` further bits may be added to `set`.  `allowed` = bits that may be added at all.
// Returns the selected indices (subsequence of `cand`), and leaves the rest in `rest`.
static void greedy_select(const std::vector<POp> &pl, const std::vector<int> &cand, uint64_t always,
                          uint64_t allowed, int budget, uint64_t all_bits, size_t window,
                          std::vector<int> &sel, std::vector<int> &rest, uint64_t &set_out) {
    uint64_t set = always, bw = 0, br = 0;
    sel.clear();
    rest.clear();
    size_t scanned = 0;
    size_t i = 0;
    for (; i < cand.size(); ++i) {
        const POp &p = pl[cand[i]];
        bool conflict = (p.mix & (bw | br)) || (p.dg & bw);
        if (!conflict) {
            const uint64_t need = p.mix & ~set;
            if ((need & ~allowed) || pc64(need) > budget) conflict = true;
            else {
                set |= need;
                budget -= pc64(need);
                sel.push_back(cand[i]);
            }
        }
        if (conflict) {
            bw |= p.mix;
            br |= p.dg;
            rest.push_back(cand[i]);
            if ((bw & all_bits) == all_bits) { ++i; break; }     // everything later is blocked
        }
        if (++scanned >= window && !sel.empty() && !rest.empty()) { ++i; break; }
    }
    for (; i < cand.size(); ++i) rest.push_back(cand[i]);
    set_out = set;
}

// Kinds the fast stage interpreter carries (tile.cu, FCode)
static bool fast_kind(const POp &p) {
    switch (p.d.kind) {
    case QVNT_ID: return true;      // placeholder of an empty remap pass (restore_layout): occupies a slot, emits nothing
    case QVNT_X: case QVNT_Y: case QVNT_RX: case QVNT_RY: case QVNT_H1:
    case QVNT_Z: case QVNT_S: case QVNT_T: case QVNT_RZ:
        return pc64(p.d.a) == 1;
    case QVNT_RZZ:
        return true;
    default:
        return false;
    }
}

struct MInfo { int src; int form, ra, rb; };   // host-side description of every MOp (describe / tests)

struct Plan {
    uint8_t perm[64];          // logical qubit -> physical index bit after the last pass
    std::vector<PassPlan> passes;
    std::vector<TStage> stages;
    std::vector<MOp> mops;
    std::vector<MBase> bases;
    std::vector<MInfo> minfo;
};

// ---- stage selection: which ops run while WHICH tile bits sit in the register slots ----------
// Greedy in list order like greedy_select, but with slot placement: one-bit ops take any
// slot, two-bit ops (rxx/ryy, swap family, h2/u2) must sit on the slot pair (0,1) or (2,3)
// because the interpreter only carries bodies for those (tile.cu).
struct Slots {
    uint64_t banned = 0; // bits a lazy x of this stage tests or flips: they must stay thread bits
    int bit[TILE_R];     // global bit held by slot j, -1 = free
    int find(int q) const {
        for (int j = 0; j < TILE_R; ++j)
            if (bit[j] == q) return j;
        return -1;
    }
};

static bool place_op(const POp &p, Slots &sl, int &ra, int &rb) {
    ra = rb = 0;
    if (p.mix == 0) return true;
    if (pc64(p.mix) == 1) {
        const int q = ctz64(p.mix);
        int j = sl.find(q);
        if (j < 0 && ((sl.banned >> q) & 1)) return false;
        if (j < 0) {
            // prefer a free slot whose pair partner is already taken (keeps whole pairs free)
            for (int k = TILE_R - 1; k >= 0 && j < 0; --k)
                if (sl.bit[k] < 0 && sl.bit[k ^ 1] >= 0) j = k;
            for (int k = TILE_R - 1; k >= 0 && j < 0; --k)
                if (sl.bit[k] < 0) j = k;
            if (j < 0) return false;
            sl.bit[j] = q;
        }
        ra = j;
        return true;
    }
    // two mix bits: x = a (or the lower bit of ab), y = b (or the upper bit)
    int x, y;
    if (p.cls == CLS_QUAD) {
        x = ctz64(p.d.a);
        y = ctz64(p.d.b);
    } else {
        x = ctz64(p.mix);
        y = ctz64(p.mix & (p.mix - 1));
    }
    int jx = sl.find(x), jy = sl.find(y);
    if ((jx < 0 && ((sl.banned >> x) & 1)) || (jy < 0 && ((sl.banned >> y) & 1))) return false;
    if (jx >= 0 && jy >= 0) {
        if ((jx ^ jy) != 1) return false;
    } else if (jx >= 0) {
        if (sl.bit[jx ^ 1] >= 0) return false;
        jy = jx ^ 1;
        sl.bit[jy] = y;
    } else if (jy >= 0) {
        if (sl.bit[jy ^ 1] >= 0) return false;
        jx = jy ^ 1;
        sl.bit[jx] = x;
    } else {
        int pr = -1;
        for (int k = 0; k + 1 < TILE_R; k += 2)
            if (sl.bit[k] < 0 && sl.bit[k + 1] < 0) pr = k;
        if (pr < 0) return false;
        jx = pr;
        jy = pr + 1;
        sl.bit[jx] = x;
        sl.bit[jy] = y;
    }
    ra = jx;
    rb = jy;
    return true;
}

struct StageSel { int idx, ra, rb; };

// lazy_x (fast passes): an x whose target and controls are not register-slot bits runs as a
// permutation between threads (FC_LX, tile.cu) -- three integer instructions instead of 48-96
// register moves.  Such an x is therefore DEFERRED while one of its bits sits in a slot and other
// work remains to form a later stage; once placed, its bits are banned from becoming slots.
static void stage_select(const std::vector<POp> &pl, const std::vector<int> &cand, uint64_t tile_set,
                         size_t max_ops, std::vector<StageSel> &sel, std::vector<int> &rest, Slots &sl,
                         bool lazy_x) {
    for (int j = 0; j < TILE_R; ++j) sl.bit[j] = -1;
    sl.banned = 0;
    uint64_t bw = 0, br = 0;
    sel.clear();
    rest.clear();
    size_t i = 0;
    for (; i < cand.size(); ++i) {
        const POp &p = pl[cand[i]];
        bool conflict = (p.mix & (bw | br)) || (p.dg & bw) || sel.size() >= max_ops;
        if (!conflict && lazy_x && p.d.kind == QVNT_X && pc64(p.mix) == 1) {
            uint64_t slot_bits = 0;
            for (int j = 0; j < TILE_R; ++j)
                if (sl.bit[j] >= 0) slot_bits |= 1ull << sl.bit[j];
            const uint64_t bits = p.mix | p.d.ctrl;
            // (the TILE_R slots are always filled: enough unbanned tile bits must remain)
            const bool room = pc64(tile_set) - pc64((sl.banned | bits) & tile_set) >= TILE_R;
            if (!(bits & slot_bits) && room) {
                sl.banned |= bits & tile_set;
                sel.push_back({cand[i], -1, -1});
                continue;
            }
        }
        if (!conflict) {
            Slots trial = sl;
            int ra, rb;
            if (place_op(p, trial, ra, rb)) {
                sl = trial;
                sel.push_back({cand[i], ra, rb});
            } else {
                conflict = true;
            }
        }
        if (conflict) {
            bw |= p.mix;
            br |= p.dg;
            rest.push_back(cand[i]);
            if ((bw & tile_set) == tile_set) { ++i; break; }     // everything later is blocked
        }
    }
    for (; i < cand.size(); ++i) rest.push_back(cand[i]);
}

// Pure host function (no CUDA): op list -> passes -> stages.
static int build_plan(const PlanCfg &c, std::vector<POp> &pl, Plan &plan) {
    const uint32_t n_local = c.n_local;
    const uint64_t lmask = (1ull << n_local) - 1ull;
    const uint64_t qmask = c.q_mask();
    const uint64_t gmask = qmask & ~lmask;
    const bool peers = c.world > 1 && c.peers;
    uint32_t wbits = 0;
    while ((1u << wbits) < c.world) ++wbits;

    uint32_t T = c.tile_bits ? (uint32_t)c.tile_bits : (uint32_t)TILE_DEFAULT_BITS;
    if (T > (uint32_t)TILE_MAX_BITS) T = TILE_MAX_BITS;
    if (T > n_local) T = n_local;      // (a tile with k rank bits pins k local bits outside the tile)
    uint32_t L = c.chunk_bits ? (uint32_t)c.chunk_bits : (uint32_t)TILE_DEFAULT_CHUNK;
    if (L > n_local) L = n_local;
    if (L > T) L = T;
    if (T - L > (uint32_t)TILE_MAX_HIGH) L = T - TILE_MAX_HIGH;
    if (L + 2 > T) L = T >= 2 ? T - 2 : 0;      // keep room for the two gathered bits a two-qubit op may need
    if (c.force_pinned >= 0 && (uint32_t)c.force_pinned < L) L = (uint32_t)c.force_pinned;   // the pinned bit is no tile bit
    if (T - L > (uint32_t)TILE_MAX_HIGH) T = L + TILE_MAX_HIGH;
    const bool tile_ok = T >= (uint32_t)TILE_MIN_BITS && n_local >= (uint32_t)TILE_MIN_BITS;
    const uint64_t low = (1ull << L) - 1ull;
    const uint64_t rank_bits = (uint64_t)c.rank << n_local;

    memcpy(plan.perm, c.perm, sizeof(plan.perm));
    std::vector<PassPlan> &passes = plan.passes;
    std::vector<int> cand(pl.size()), sel, rest;
    for (size_t i = 0; i < pl.size(); ++i) cand[i] = (int)i;
    while (!cand.empty()) {
        const POp &first = pl[cand[0]];
        PassPlan pp;
        memset(&pp.hdr, 0, sizeof(pp.hdr));
        const bool first_global = (first.mix & gmask) != 0;
        if (first_global && !peers) {
            set_error("gate on a sharded (global) qubit needs qvnt_reg_attach_peers first");
            return QVNT_ERR_COMM;
        }
        if (first_global && !tile_ok) {
            set_error("shard too small (%u local qubits) for a global-qubit gate", n_local);
            return QVNT_ERR_UNSUPPORTED;
        }
        if ((!c.fuse || !tile_ok) && !first_global) {
            pp.direct = true;
            pp.ops.push_back(cand[0]);
            cand.erase(cand.begin());
            passes.push_back(pp);
            continue;
        }
        uint64_t set = 0;
        const uint64_t allowed = peers ? qmask : lmask;
        // passes that start on a global qubit may use larger chunks (option "peer_chunk_bits")
        uint32_t Lp = L, Tp = T;
        if (first_global && c.force_pinned < 0) {
            if (c.peer_tile_bits > (int)T) Tp = (uint32_t)c.peer_tile_bits;
            if (Tp > (uint32_t)TILE_MAX_BITS) Tp = TILE_MAX_BITS;
            if (Tp > n_local) Tp = n_local;
            if (c.peer_chunk_bits > (int)L) Lp = (uint32_t)c.peer_chunk_bits;
            if (Lp + 2 > Tp) Lp = Tp - 2;
            if (Tp - Lp > (uint32_t)TILE_MAX_HIGH) Lp = Tp - TILE_MAX_HIGH;
        }
        const uint64_t low_p = (1ull << Lp) - 1ull;
        if (!c.fuse) {
            sel.assign(1, cand[0]);
            rest.assign(cand.begin() + 1, cand.end());
            set = low_p | first.mix;
        } else {
            greedy_select(pl, cand, low_p, allowed, (int)(Tp - Lp), qmask, 1u << 16, sel, rest, set);
        }
        if (sel.empty() && !first_global) {
            // no tile geometry holds the op's mix bits (tiny shards): one direct sweep
            pp.direct = true;
            pp.ops.push_back(cand[0]);
            cand.erase(cand.begin());
            passes.push_back(pp);
            continue;
        }
        if (sel.empty()) {
            set_error("internal: planner could not place op kind %u (mix 0x%llx) in a tile", first.d.kind,
                      (unsigned long long)first.mix);
            return QVNT_ERR_UNSUPPORTED;
        }
        if (sel.size() == 1 && !(pl[sel[0]].mix & gmask)) {
            pp.direct = true;
            pp.ops = sel;
            cand.swap(rest);
            passes.push_back(pp);
            continue;
        }
        // fill the tile up to T bits with the lowest unused local bits (keeps it contiguous)
        for (uint32_t b = Lp; pc64(set) < (int)Tp && b < n_local; ++b)
            if ((int)b != c.force_pinned) set |= 1ull << b;
        // tile-local numbering: ascending global position
        TPassHdr &h = pp.hdr;
        h.T = (uint32_t)pc64(set);
        h.L = Lp;
        uint32_t lp = 0;
        for (uint64_t m = set; m; m &= m - 1) h.gpos[lp++] = (uint8_t)ctz64(m);
        const uint64_t tile_g = set & gmask;
        const int kg = pc64(tile_g);
        h.touches_peer = kg ? 1u : 0u;
        // ownership: kg local non-tile bits are pinned to this rank's tile-global bits.
        // Remap pass (kg == 1, option "remap"): the pass does not write the peer half back -- every
        // rank keeps BOTH values of the global bit g for its value of the pinned local bit b, i.e.
        // index bits g and b trade places and the qubit that was global is local from here on
        // (tile.cu: stores go to the local address with bit b := value of g).  The pinned bit is then
        // chosen as the local bit whose qubit is needed latest (as a mix bit) by the remaining ops.
        uint64_t own_mask = 0, own_val = 0;
        const bool remap = c.remap && kg == 1 && peers &&
                           (1ull << (n_local - (uint32_t)pc64(set & lmask) - 1u)) <= c.ack_cap;
        if (remap) {
            const int gb = ctz64(tile_g);
            int best = c.force_pinned;
            size_t best_d = 0;
            for (int b = (int)n_local - 1; b >= 0 && c.force_pinned < 0; --b) {
                if ((set >> b) & 1) continue;
                size_t d = rest.size() + 1;                  // never needed again
                for (size_t k = 0; k < rest.size(); ++k)
                    if ((pl[rest[k]].mix >> b) & 1) {
                        d = k;
                        break;
                    }
                if (best < 0 || d > best_d) {
                    best = b;
                    best_d = d;
                }
            }
            if (best < 0) {
                set_error("shard too small for a tile with a global bit");
                return QVNT_ERR_UNSUPPORTED;
            }
            own_mask = 1ull << best;
            if ((rank_bits >> gb) & 1ull) own_val = own_mask;
            h.remap = 1;
            h.remap_g = (uint8_t)gb;
            h.remap_b = (uint8_t)best;
        } else {
            uint64_t tg = tile_g;
            int b = (int)n_local - 1;
            while (tg) {
                while (b >= 0 && ((set >> b) & 1)) --b;
                if (b < 0) {
                    set_error("shard too small for a tile with %d global bits", kg);
                    return QVNT_ERR_UNSUPPORTED;
                }
                const int gb = ctz64(tg);
                own_mask |= 1ull << b;
                if ((rank_bits >> gb) & 1ull) own_val |= 1ull << b;
                --b;
                tg &= tg - 1;
            }
        }
        const uint64_t fixed = (set & lmask) | own_mask;
        h.fx.n = 0;
        h.fx._pad = 0;
        h.fx.val = own_val;
        for (uint64_t m = fixed; m; m &= m - 1) h.fx.pos[h.fx.n++] = (uint8_t)ctz64(m);
        h.n_tiles = 1ull << (n_local - h.fx.n);
        h.base_or = rank_bits & ~tile_g;
        pp.ops = sel;
        cand.swap(rest);
        passes.push_back(pp);
        if (h.remap) {
            // the ops still to be scheduled see the new layout
            const int gb = h.remap_g, lb = h.remap_b;
            for (int idx : cand) {
                POp &q = pl[idx];
                q.d.a = swap_bits(q.d.a, gb, lb);
                q.d.b = swap_bits(q.d.b, gb, lb);
                q.d.ctrl = swap_bits(q.d.ctrl, gb, lb);
                q.mix = swap_bits(q.mix, gb, lb);
                q.dg = swap_bits(q.dg, gb, lb);
            }
            for (uint32_t q = 0; q < c.q_num; ++q) {
                if (plan.perm[q] == gb) plan.perm[q] = (uint8_t)lb;
                else if (plan.perm[q] == lb) plan.perm[q] = (uint8_t)gb;
            }
        }
    }

    // ---- stages of every tile pass ----
    std::vector<PassPlan> out_passes;
    for (PassPlan &pp : passes) {
        if (pp.direct) {
            out_passes.push_back(pp);
            continue;
        }
        TPassHdr &h = pp.hdr;
        uint64_t set = 0;
        int lpos_of[64];
        for (uint32_t l = 0; l < h.T; ++l) {
            set |= 1ull << h.gpos[l];
            lpos_of[h.gpos[l]] = (int)l;
        }
        // runs of fixed positions for the kernel's tile-counter expansion
        h.n_runs = 0;
        for (uint32_t k = 0; k < h.fx.n;) {
            uint32_t e = k + 1;
            while (e < h.fx.n && h.fx.pos[e] == h.fx.pos[e - 1] + 1) ++e;
            if (h.n_runs >= 16) {
                set_error("internal: tile geometry needs more than 16 runs");
                return QVNT_ERR_UNSUPPORTED;
            }
            h.run_pos[h.n_runs] = h.fx.pos[k];
            h.run_len[h.n_runs] = (uint8_t)(e - k);
            ++h.n_runs;
            k = e;
        }
        h.fx_val = h.fx.val;
        for (uint32_t l = 0; l < 16; ++l)
            h.gpos_store[l] = (h.remap && l < h.T && h.gpos[l] == h.remap_g) ? h.remap_b : h.gpos[l];
        h.fixed_mask = 0;
        for (uint32_t k = 0; k < h.fx.n; ++k) h.fixed_mask |= 1ull << h.fx.pos[k];

        bool pass_fast = true;
        for (int idx : pp.ops) pass_fast = pass_fast && fast_kind(pl[idx]);
        PassPlan cur = pp;           // same geometry; ops / stages filled below (split if too long)
        cur.hdr.full = pass_fast ? 0u : 1u;
        cur.ops.clear();
        cur.hdr.stage_begin = (uint32_t)plan.stages.size();
        cur.hdr.op_begin = (uint32_t)plan.mops.size();
        int cur_static = 0;
        double pass_scale = 1.0;      // product of the factors of the pass's butterfly h (FC_HB), folded at flush
        const bool use_butterfly = c.butterfly;
        auto flush = [&]() {
            cur.hdr.n_static = (uint32_t)cur_static;
            cur_static = 0;
            cur.hdr.n_stages = (uint32_t)plan.stages.size() - cur.hdr.stage_begin;
            cur.hdr.n_ops = (uint32_t)plan.mops.size() - cur.hdr.op_begin;
            for (uint32_t k = 0; k < cur.hdr.n_stages; ++k)
                cur.hdr.stage_end[k] = (uint16_t)(plan.stages[cur.hdr.stage_begin + k].op_end - cur.hdr.op_begin);
            cur.hdr.need_flags = 0;
            cur.hdr.uses_sc = 0;
            if (pass_scale != 1.0 && !cur.hdr.full) {
                // the butterflies' common factor: into the first unconditional, unmasked pair op of the pass (it runs
                // on every amplitude of every tile); a pass without one turns its last butterfly back into a pair op
                MOp *carrier = nullptr, *last_hb = nullptr;
                for (uint32_t k = 0; k < cur.hdr.n_ops; ++k) {
                    MOp &mk = plan.mops[cur.hdr.op_begin + k];
                    if (mk.code >= (uint8_t)FC_HB && mk.code < (uint8_t)(FC_HB + TILE_R)) last_hb = &mk;
                    else if (!carrier && mk.code < (uint8_t)(FC_PX + TILE_R) && mk.okmask == 0xFFFFu &&
                             !(mk.dagger & (MOP_COND | MOP_CONDB)))
                        carrier = &mk;
                }
                if (!carrier && last_hb) {
                    last_hb->code = (uint8_t)(FC_PR + (last_hb->code - FC_HB));
                    carrier = last_hb;
                }
                if (carrier) {
                    carrier->ph_re *= pass_scale; carrier->ph_im *= pass_scale;
                    carrier->c2 *= pass_scale; carrier->c3 *= pass_scale;
                    for (int q = 0; q < 4; ++q) carrier->alt[q] *= pass_scale;
                }
            }
            pass_scale = 1.0;
            uint32_t member_until = 0;
            for (uint32_t k = 0; k < cur.hdr.n_ops; ++k) {
                MOp &mk = plan.mops[cur.hdr.op_begin + k];
                mk.idx = (uint16_t)k;
                if (!cur.hdr.full && mk.code < (uint8_t)FC_TOTAL) {      // (butterflies already carry their final code)
                    const int cls = (mk.dagger & MOP_CONDB) ? 2 : (mk.dagger & MOP_COND) ? 1 : 0;
                    // one control, in register slot c: the single-control arm (engine.h FC_DS1 ...); the members
                    // of a run are data of their header, their codes stay
                    int c1 = -1;
                    for (int c = 0; c < TILE_R; ++c) {
                        uint32_t want = 0;
                        for (int K = 0; K < TILE_NV; ++K)
                            if (K & (1 << c)) want |= 1u << K;
                        if (mk.okmask == want) c1 = c;
                    }
                    const uint8_t g = mk.code;
                    if (cls == 0 && c1 >= 0 && k >= member_until && c.single_ctrl) {
                        if (g >= (uint8_t)(FC_MASKED + FC_DS) && g < (uint8_t)(FC_MASKED + FC_DS + TILE_R))
                            mk.code = (uint8_t)(FC_DS1 + 4 * c1 + (g - (FC_MASKED + FC_DS)));
                        else if (g == (uint8_t)(FC_MASKED + FC_DU)) mk.code = (uint8_t)(FC_DU1 + c1);
                        else if (g == (uint8_t)(FC_MASKED + FC_DM)) mk.code = (uint8_t)(FC_DM1 + c1);
                    }
                    if (g == (uint8_t)FC_DM || g == (uint8_t)(FC_MASKED + FC_DM)) member_until = k + 1u + mk.a_reg;
                    if (mk.code < (uint8_t)FC_TOTAL)                  // control class into the code byte (engine.h)
                        mk.code = (uint8_t)(mk.code + FC_TOTAL * cls);
                    else if (mk.code < (uint8_t)FC_HB) cur.hdr.uses_sc = 1;
                }
                const MBase &mb = plan.bases[cur.hdr.op_begin + k];
                if (mb.ctrl_base | mb.a_base) cur.hdr.need_flags = 1;
            }
            if (cur.hdr.n_ops || (cur.hdr.remap && cur.hdr.n_stages)) out_passes.push_back(cur);
            cur.ops.clear();
            cur.hdr.stage_begin = (uint32_t)plan.stages.size();
            cur.hdr.op_begin = (uint32_t)plan.mops.size();
        };
        std::vector<int> c2 = pp.ops, r2;
        std::vector<StageSel> s2;
        while (!c2.empty()) {
            Slots sl;
            // Program size per pass, bounded by the CTA's shared memory next to the tile (tile.cu,
            // tile_smem_bytes: 67 B per op, 48 B + 4 B x threads per stage): 2^12-amplitude tiles
            // (64 KiB, 256 threads) leave room for half of what 2^11 tiles do.
            const size_t max_ops = h.T >= 12 ? (size_t)TILE_MAX_OPS / 2 : (size_t)TILE_MAX_OPS;
            const size_t max_stages = h.T >= 12 ? (size_t)TILE_MAX_STAGES / 2 : (size_t)TILE_MAX_STAGES;
            const size_t used = plan.mops.size() - cur.hdr.op_begin;
            const size_t room = used < max_ops ? (max_ops - used) * 3 / 4 : 0;      // (+ run headers)
            stage_select(pl, c2, set, room, s2, r2, sl, pass_fast);
            if (s2.empty() || plan.stages.size() - cur.hdr.stage_begin >= max_stages) {
                if (plan.mops.size() == cur.hdr.op_begin) {
                    set_error("internal: stage construction stalled");
                    return QVNT_ERR_UNSUPPORTED;
                }
                flush();               // program full: continue in a new pass over the same tiles
                continue;
            }
            // register slots: the stage's mix bits + filler (highest free tile bits)
            uint64_t rset = 0;
            for (int j = 0; j < TILE_R; ++j)
                if (sl.bit[j] >= 0) rset |= 1ull << sl.bit[j];
            for (int j = 0; j < TILE_R; ++j) {
                if (sl.bit[j] >= 0) continue;
                for (int l = (int)h.T - 1; l >= 0; --l)
                    if (!(((rset | sl.banned) >> h.gpos[l]) & 1)) {
                        sl.bit[j] = h.gpos[l];
                        rset |= 1ull << h.gpos[l];
                        break;
                    }
            }
            TStage st;
            memset(&st, 0, sizeof(st));
            for (int j = 0; j < TILE_R; ++j) st.r_lpos[j] = (uint8_t)lpos_of[sl.bit[j]];
            // thread bits: lane bit j (j = 0..2) takes tile-local bit j or L + j -- in the padded-linear
            // tile buffer (tile.cu) those move an address by 2^j bank groups, so a quarter-warp's eight
            // 16-byte accesses fall into eight different groups; the rest ascend
            std::vector<int> order;
            {
                std::vector<int> nr;
                for (uint32_t l = 0; l < h.T; ++l)
                    if (!((rset >> h.gpos[l]) & 1)) nr.push_back((int)l);
                std::vector<char> used(nr.size(), 0);
                for (int j = 0; j < 3; ++j) {
                    int pick = -1;
                    for (size_t i = 0; i < nr.size() && pick < 0; ++i)
                        if (!used[i] && nr[i] == j) pick = (int)i;
                    for (size_t i = 0; i < nr.size() && pick < 0; ++i)
                        if (!used[i] && nr[i] == (int)h.L + j) pick = (int)i;
                    for (size_t i = 0; i < nr.size() && pick < 0; ++i)
                        if (!used[i]) pick = (int)i;
                    if (pick >= 0) {
                        used[pick] = 1;
                        order.push_back(nr[pick]);
                    }
                }
                for (size_t i = 0; i < nr.size(); ++i)
                    if (!used[i]) order.push_back(nr[i]);
                for (size_t i = 0; i < order.size(); ++i) st.t_lpos[i] = (uint8_t)order[i];
            }
            // index-bit classes of this stage, in global numbering
            uint64_t reg_of_bit[TILE_R];
            for (int j = 0; j < TILE_R; ++j) reg_of_bit[j] = 1ull << sl.bit[j];
            auto split = [&](uint64_t m, uint32_t &reg, uint32_t &thr, uint64_t &base) {
                reg = thr = 0;
                for (int j = 0; j < TILE_R; ++j)
                    if (m & reg_of_bit[j]) reg |= 1u << j;
                for (size_t k = 0; k < order.size(); ++k)
                    if ((m >> h.gpos[order[k]]) & 1ull) thr |= 1u << k;
                base = m & ~set;
            };
            st.op_begin = (uint32_t)plan.mops.size();
            for (const StageSel &ss : s2) {
                const POp &p = pl[ss.idx];
                if (p.d.kind == QVNT_ID) continue;
                MOp m;
                MBase b;
                MInfo mi;
                memset(&m, 0, sizeof(m));
                memset(&b, 0, sizeof(b));
                uint32_t creg, cthr;
                split(p.d.ctrl, creg, cthr, b.ctrl_base);
                m.ctrl_thr = cthr;
                m.okmask = 0;
                for (int K = 0; K < TILE_NV; ++K)
                    if ((~(uint32_t)K & creg) == 0) m.okmask |= (uint16_t)(1u << K);
                m.dagger = (uint8_t)p.d.dagger;
                m.ph_re = p.d.ph_re;
                m.ph_im = p.d.ph_im;
                mi.src = ss.idx;
                mi.ra = ss.ra;
                mi.rb = ss.rb;
                if (pass_fast) {
                    const double c = p.d.ph_re, sn = p.d.ph_im, hs = QV_FRAC_1_SQRT_2;
                    const bool dg = p.d.dagger != 0;
                    double f0r = 1.0, f0i = 0.0, f1r = 1.0, f1i = 0.0;
                    mi.form = p.cls == CLS_DIAG ? TF_DIAG : TF_PAIR1;
                    m.dagger = 0;       // fast forms: MOP_* flags only
                    auto pair4 = [&](uint8_t code, double c0, double c1, double c2, double c3) {
                        m.code = (uint8_t)(code + ss.ra);
                        m.ph_re = c0; m.ph_im = c1; m.c2 = c2; m.c3 = c3;
                        // the same 2x2 with the roles of the pair's members exchanged (slot inverted, FC_LI)
                        m.alt[0] = c3; m.alt[1] = c2; m.alt[2] = c1; m.alt[3] = c0;
                    };
                    switch (p.d.kind) {
                    // h1.rs:16-22: (p0 + p1) * s, (p0 - p1) * s
                    case QVNT_H1:
                        if (use_butterfly && !creg && !cthr && !b.ctrl_base) {
                            // add / subtract only; the factor (1/sqrt(2), or the 1 / 0.5 of a split h2) goes into
                            // the pass's scale and from there into one pair op's coefficients (flush)
                            pair4(FC_PR, 1.0, 1.0, 1.0, -1.0);
                            m.code = (uint8_t)(FC_HB + ss.ra);
                            pass_scale *= p.d.ph_re;
                        } else {
                            pair4(FC_PR, p.d.ph_re, p.d.ph_re, p.d.ph_re, -p.d.ph_re);
                        }
                        break;
                    case QVNT_RY: pair4(FC_PR, c, -sn, sn, c); break;
                    case QVNT_RX: pair4(FC_PX, c, sn, sn, c); break;
                    case QVNT_Y: pair4(FC_PX, 0.0, 1.0, -1.0, 0.0); break;
                    case QVNT_X:
                        if (ss.ra < 0) {
                            uint32_t treg, tthr;
                            uint64_t tbase;
                            split(p.d.a, treg, tthr, tbase);
                            m.code = (uint8_t)FC_LX;
                            st.sync_after_load = 1;
                            m.a_thr = tthr;
                            {
                                const int pos = ctz64(p.d.a);
                                m.a_reg = (uint16_t)(lpos_of[pos] | (((h.remap && pos == h.remap_g) ? h.remap_b : pos) << 8));
                            }
                            {
                                // payload in the coefficient words (both blocks): byte distance of the bit in
                                // the padded-linear tile buffer | the bit << 32, and its byte distance in the shard
                                const uint32_t lp = (uint32_t)lpos_of[ctz64(p.d.a)], bit = 1u << lp;
                                const uint64_t w = (uint64_t)(16u * (bit + (bit >> h.L))) | ((uint64_t)bit << 32);
                                const int pos = ctz64(p.d.a);       // (a remap pass stores the rank bit at the pinned bit)
                                const uint64_t g = 16ull << ((h.remap && pos == h.remap_g) ? h.remap_b : pos);
                                const uint64_t at = (uint64_t)tthr;              // the target's thread bit
                                memcpy(&m.ph_re, &w, 8);
                                memcpy(&m.ph_im, &g, 8);
                                memcpy(&m.c2, &at, 8);
                                m.alt[0] = m.ph_re;
                                m.alt[1] = m.ph_im;
                                m.alt[2] = m.c2;
                            }
                            mi.form = TF_LAZYX;
                            mi.ra = mi.rb = 0;
                        } else if (creg == 0) {
                            // no control in a register slot: the thread marks the slot as inverted
                            m.code = (uint8_t)FC_LI;
                            m.a_reg = (uint16_t)ss.ra;
                            {
                                const uint64_t w = (uint64_t)MOP_ALT_BYTES << (8 * ss.ra);   // the slot's inversion byte
                                memcpy(&m.ph_re, &w, 8);
                                m.alt[0] = m.ph_re;
                            }
                            mi.form = TF_LAZYI;
                        } else {
                            m.code = (uint8_t)(FC_SW + ss.ra);
                        }
                        break;
                    case QVNT_Z: f1r = -1.0; break;
                    case QVNT_S: f1r = 0.0; f1i = dg ? -1.0 : 1.0; break;
                    case QVNT_T: f1r = hs; f1i = dg ? -hs : hs; break;
                    default: f0r = c; f0i = -sn; f1r = c; f1i = sn; break;      // rz, rzz
                    }
                    if (p.cls == CLS_DIAG) {
                        uint32_t areg, athr;
                        split(p.d.a, areg, athr, b.a_base);
                        m.a_reg = (uint16_t)areg;
                        m.a_thr = athr;
                        m.code = areg == 0 ? (uint8_t)FC_DU
                                 : pc64(areg) == 1 ? (uint8_t)(FC_DS + ctz64(areg)) : (uint8_t)FC_DG;
                        m.ph_re = f0r; m.ph_im = f0i; m.c2 = f1r; m.c3 = f1i;
                        m.alt[0] = f1r; m.alt[1] = f1i; m.alt[2] = f0r; m.alt[3] = f0i;
                        if (m.code == (uint8_t)FC_DG) {      // per-slot-pattern factors: no role exchange, always the masked arm
                            m.alt[0] = f0r; m.alt[1] = f0i; m.alt[2] = f1r; m.alt[3] = f1i;
                            m.code = (uint8_t)(FC_DG + FC_MASKED);
                        }
                        if (f0r == 1.0 && f0i == 0.0) m.dagger |= MOP_SKIP0;
                        if (b.a_base) m.dagger |= MOP_PARB;
                        if (athr) m.dagger |= MOP_ATHR;
                        mi.ra = mi.rb = 0;
                    }
                    if (cthr || b.ctrl_base) m.dagger |= MOP_COND;
                    if (b.ctrl_base) m.dagger |= MOP_CONDB;
                    if (m.okmask != 0xFFFFu && m.code < (uint8_t)FC_MASKED) m.code = (uint8_t)(m.code + FC_MASKED);
                } else if (p.cls == CLS_DIAG) {
                    uint32_t areg, athr;
                    split(p.d.a, areg, athr, b.a_base);
                    m.a_reg = (uint16_t)areg;
                    m.a_thr = athr;
                    int k5 = p.d.kind == QVNT_Z ? 0 : p.d.kind == QVNT_S ? 1 : p.d.kind == QVNT_T ? 2
                             : p.d.kind == QVNT_RZ ? 3 : 4;
                    m.code = (uint8_t)((areg ? MC_DG : MC_DU) + k5);
                    mi.form = TF_DIAG;
                    mi.ra = mi.rb = 0;
                } else if (p.cls == CLS_QUAD) {
                    mi.form = TF_QUAD;
                    const int pair = ss.ra >> 1;
                    if (p.d.kind == QVNT_H2) m.code = (uint8_t)(MC_H2 + pair);
                    else {
                        m.code = (uint8_t)(MC_U2 + 2 * pair + (ss.ra & 1));
                        m.a_thr = p.d.mat;
                    }
                } else if (pc64(p.d.a) == 1) {
                    mi.form = TF_PAIR1;
                    int k6;
                    switch (p.d.kind) {
                    case QVNT_X: k6 = 0; break;
                    case QVNT_Y: k6 = 1; break;
                    case QVNT_RX: k6 = 2; break;
                    case QVNT_RY: k6 = 3; break;
                    case QVNT_H1: k6 = 4; break;
                    case QVNT_U1: k6 = 5; break;
                    default:
                        set_error("internal: kind %u is not a one-bit pair op", p.d.kind);
                        return QVNT_ERR_INVALID;
                    }
                    m.code = (uint8_t)(MC_P1 + 4 * k6 + ss.ra);
                    if (p.d.kind == QVNT_U1) m.a_thr = p.d.mat;
                } else {
                    const int pair = ss.ra >> 1;
                    if (op_odd_only(p.d.kind)) {
                        mi.form = TF_ODD2;
                        const int k4 = p.d.kind == QVNT_SWAP ? 0 : p.d.kind == QVNT_ISWAP ? 1
                                       : p.d.kind == QVNT_SQRT_SWAP ? 2 : 3;
                        m.code = (uint8_t)(MC_ODD + 2 * k4 + pair);
                    } else if (p.d.kind == QVNT_RXX || p.d.kind == QVNT_RYY) {
                        mi.form = TF_PAIR2X;
                        m.code = (uint8_t)(MC_P2X + 2 * (p.d.kind == QVNT_RYY) + pair);
                    } else {
                        set_error("internal: kind %u with a %d-bit mask reached the tile planner", p.d.kind,
                                  pc64(p.d.a));
                        return QVNT_ERR_INVALID;
                    }
                }
                plan.mops.push_back(m);
                plan.bases.push_back(b);
                plan.minfo.push_back(mi);
                cur.ops.push_back(ss.idx);
            }
            // ---- merge runs of diagonal ops (fast passes) ----
            // Consecutive diagonal ops with no target bit in a register slot and identical controls
            // multiply every amplitude of a thread by thread-wide factors: a header op (FC_DM) makes
            // the kernel fold the run into ONE complex factor per thread (4 FP64 instructions per
            // constituent instead of 64) -- qft's chains of controlled rz behind every h.
            if (pass_fast) {
                const size_t b0 = st.op_begin;
                std::vector<MOp> mo(plan.mops.begin() + b0, plan.mops.end());
                std::vector<MBase> ba(plan.bases.begin() + b0, plan.bases.end());
                std::vector<MInfo> mf(plan.minfo.begin() + b0, plan.minfo.end());
                plan.mops.resize(b0);
                plan.bases.resize(b0);
                plan.minfo.resize(b0);
                bool stage_has_lx = false;
                for (const MOp &q : mo) stage_has_lx = stage_has_lx || q.code == (uint8_t)FC_LX;
                auto is_du = [&](size_t k) {
                    return mo[k].code == (uint8_t)FC_DU || mo[k].code == (uint8_t)(FC_MASKED + FC_DU);
                };
                for (size_t k = 0; k < mo.size();) {
                    size_t e = k;
                    if (is_du(k)) {
                        e = k + 1;
                        while (e < mo.size() && is_du(e) && mo[e].okmask == mo[k].okmask &&
                               mo[e].ctrl_thr == mo[k].ctrl_thr && ba[e].ctrl_base == ba[k].ctrl_base && e - k < 4096)
                            ++e;
                    }
                    if (e - k >= 3) {
                        MOp hd;
                        MBase hb;
                        memset(&hd, 0, sizeof(hd));
                        memset(&hb, 0, sizeof(hb));
                        hd.code = (uint8_t)(mo[k].okmask == 0xFFFFu ? FC_DM : FC_DM + FC_MASKED);
                        hd.dagger = (uint8_t)(mo[k].dagger & (MOP_COND | MOP_CONDB));
                        // tabulated run (engine.h MOP_STATIC): targets on thread bits XOR outside the tile, and
                        // the thread's group number still is its thread index (no lazy x in the stage)
                        bool is_static = !stage_has_lx && cur_static < TILE_MAX_STATIC;
                        bool outer = false;
                        for (size_t q = k; q < e && is_static; ++q) {
                            is_static = (mo[q].a_thr != 0) != (ba[q].a_base != 0);
                            outer = outer || ba[q].a_base != 0;
                        }
                        if (is_static) {
                            hd.dagger |= MOP_STATIC;
                            if (outer) hd.dagger |= MOP_PARB;
                            hd.a_thr = (uint32_t)cur_static++;
                        }
                        hd.okmask = mo[k].okmask;
                        hd.ctrl_thr = mo[k].ctrl_thr;
                        hd.a_reg = (uint16_t)(e - k);
                        hb.ctrl_base = ba[k].ctrl_base;
                        plan.mops.push_back(hd);
                        plan.bases.push_back(hb);
                        plan.minfo.push_back({mf[k].src, TF_GROUP, 0, 0});
                    } else if (e == k) {
                        e = k + 1;
                    }
                    for (size_t q = k; q < e; ++q) {
                        plan.mops.push_back(mo[q]);
                        plan.bases.push_back(ba[q]);
                        plan.minfo.push_back(mf[q]);
                    }
                    k = e;
                }
            }
            st.op_end = (uint32_t)plan.mops.size();
            plan.stages.push_back(st);
            c2.swap(r2);
        }
        flush();
    }
    passes.swap(out_passes);
    return QVNT_OK;
}

struct Uploaded {
    TStage *d_stages = nullptr;
    MOp *d_mops = nullptr;
    MBase *d_bases = nullptr;
};

// the pass programs of a plan -> device (one H2D copy from pinned staging)
static int upload_plan(qvnt_reg *r, const Plan &plan, Uploaded &u) {
    if (plan.stages.empty()) return QVNT_OK;
    const size_t sb = plan.stages.size() * sizeof(TStage), ob = plan.mops.size() * sizeof(MOp),
                 bb = plan.bases.size() * sizeof(MBase);
    const size_t sb_al = (sb + 255) & ~(size_t)255, ob_al = (ob + 255) & ~(size_t)255;
    const size_t total = sb_al + ob_al + bb;
    int rc = ensure_stage(r, total);
    if (rc) return rc;
    if ((rc = ensure_dev(&r->d_ops, &r->d_ops_cap, total))) return rc;
    memcpy(r->h_stage, plan.stages.data(), sb);
    if (ob) memcpy((char *)r->h_stage + sb_al, plan.mops.data(), ob);
    if (bb) memcpy((char *)r->h_stage + sb_al + ob_al, plan.bases.data(), bb);
    QV_CUDA(cudaMemcpyAsync(r->d_ops, r->h_stage, total, cudaMemcpyHostToDevice, r->stream));
    QV_CUDA(cudaEventRecord(r->stage_free, r->stream));
    r->stage_busy = true;
    r->stats.h2d_bytes += total;
    u.d_stages = (TStage *)r->d_ops;
    u.d_mops = (MOp *)((char *)r->d_ops + sb_al);
    u.d_bases = (MBase *)((char *)r->d_ops + sb_al + ob_al);
    return QVNT_OK;
}

// one pass of a plan, with the cross-GPU barriers it needs (need_barrier: peers may still be
// writing into / reading from this shard)
static int enqueue_pass(qvnt_reg *r, const std::vector<POp> &pl, const PassPlan &pp, const Uploaded &u,
                        bool &need_barrier) {
    const uint32_t n_local = r->n_local;
    if (pp.direct) {
        if (need_barrier) {
            int rc = dist_barrier(r);
            if (rc) return rc;
            need_barrier = false;
        }
        return run_direct(r, pl[pp.ops[0]]);
    }
    const TPassHdr &h = pp.hdr;
    if (h.touches_peer || need_barrier) {
        int rc = dist_barrier(r);
        if (rc) return rc;
    }
    // (a remap pass writes only this shard, and only places the peer has acknowledged reading)
    need_barrier = h.touches_peer != 0 && !h.remap;
    TPassHdr hl = h;
    if (h.remap) hl.epoch = ++r->remap_epoch;
    LaunchScope ls(r, 1);
    int n = launch_tile_pass(r->stream, r->segs, hl, u.d_stages, u.d_mops, u.d_bases, r->d_mat, r->sm_count, r->knobs);
    ls.done(n);
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "tile pass launch");
        set_error("internal: tile pass rejected (T=%u L=%u)", h.T, h.L);
        return QVNT_ERR_INVALID;
    }
    r->stats.passes += 1;
    r->stats.h2d_bytes += sizeof(TPassHdr) + sizeof(Segs);    // kernel parameters
    const uint64_t amps = h.n_tiles << h.T;
    r->stats.alg_bytes[1] += amps * 32;
    if (h.touches_peer) {
        int kg = 0;
        for (uint32_t l = 0; l < h.T; ++l) kg += h.gpos[l] >= n_local;
        r->stats.peer_bytes += h.remap ? amps * 8 : ((amps * 32) >> kg) * ((1ull << kg) - 1ull);
    }
    return QVNT_OK;
}

static int run_plan(qvnt_reg *r, const std::vector<POp> &pl, Plan &plan) {
    Uploaded u;
    int rc = upload_plan(r, plan, u);
    if (rc) return rc;
    bool need_barrier = false;
    for (const PassPlan &pp : plan.passes)
        if ((rc = enqueue_pass(r, pl, pp, u, need_barrier))) return rc;
    if (need_barrier) return dist_barrier(r);
    return QVNT_OK;
}

static PlanCfg cfg_of(const qvnt_reg *r) {
    PlanCfg c;
    c.q_num = r->q_num;
    c.n_local = r->n_local;
    c.rank = r->rank;
    c.world = r->world;
    c.peers = r->peers_attached;
    c.fuse = r->opt_fuse != 0;
    c.tile_bits = r->opt_tile_bits;
    c.chunk_bits = r->opt_chunk_bits;
    c.remap = r->opt_remap != 0 && r->remap_possible;
    c.peer_chunk_bits = r->opt_peer_chunk_bits;
    c.peer_tile_bits = r->opt_peer_tile_bits;
    c.single_ctrl = r->knobs.single_ctrl != 0;
    c.butterfly = r->knobs.butterfly != 0;
    c.lower_two_bit = r->knobs.lower_two_bit != 0;
    c.ack_cap = r->ack_cap;
    memcpy(c.perm, r->perm, sizeof(c.perm));
    return c;
}

int run_ops(qvnt_reg *r, const qvnt_op_t *ops, size_t n_ops) {
    std::vector<POp> pl;
    std::vector<amp> mats;
    const PlanCfg c = cfg_of(r);
    int rc = lower_ops(c, ops, n_ops, pl, mats);
    if (rc) return rc;
    r->stats.ops_applied += n_ops;
    if (pl.empty()) return QVNT_OK;
    Plan plan;
    if ((rc = build_plan(c, pl, plan))) return rc;

    // matrices of u1/u2 ops -> device table (pageable source: cudaMemcpyAsync stages it
    // before returning, so `mats` may go out of scope)
    if (!mats.empty()) {
        const size_t bytes = mats.size() * sizeof(amp);
        if ((rc = ensure_dev((void **)&r->d_mat, &r->d_mat_cap, bytes))) return rc;
        QV_CUDA(cudaMemcpyAsync(r->d_mat, mats.data(), bytes, cudaMemcpyHostToDevice, r->stream));
        r->stats.h2d_bytes += bytes;
    }
    if ((rc = run_plan(r, pl, plan))) return rc;
    memcpy(r->perm, plan.perm, sizeof(r->perm));
    return QVNT_OK;
}

// One remap pass that carries no gate: index bits g (global) and b (local) trade places.
static int empty_remap_pass(qvnt_reg *r, const PlanCfg &base, uint32_t g, uint32_t b) {
    PlanCfg c = base;
    c.remap = true;
    c.fuse = true;
    c.force_pinned = (int)b;
    std::vector<POp> pl(1);
    memset(&pl[0], 0, sizeof(POp));
    pl[0].d.kind = QVNT_ID;
    pl[0].cls = CLS_PAIR;
    pl[0].mix = 1ull << g;
    pl[0].d.a = pl[0].la = 1ull << g;
    Plan plan;
    int rc = build_plan(c, pl, plan);
    if (rc) return rc;
    if (plan.passes.size() != 1 || !plan.passes[0].hdr.remap) {
        set_error("internal: could not build the remap pass that restores qubit order (bits %u, %u)", g, b);
        return QVNT_ERR_UNSUPPORTED;
    }
    return run_plan(r, pl, plan);
}

// Undo the qubit remapping of earlier qvnt_reg_apply calls: every API that addresses amplitudes by
// index (read / write / probabilities / measurement order / collapse) expects qubit q at index bit q.
// Global bits first -- one EMPTY remap pass each (data movement only: the pinned bit is the local bit
// that holds the qubit belonging to the global position); what is left is a permutation of local
// bits, undone by swap gates on the index bits (pure permutations: exact).
int restore_layout(qvnt_reg *g) {
    // (a group handle restores all its shards in step; they share one qubit map)
    qvnt_reg *r = g->shards.empty() ? g : g->shards[0];
    bool ident = true;
    for (uint32_t q = 0; q < r->q_num; ++q) ident = ident && r->perm[q] == q;
    if (ident) return QVNT_OK;
    PlanCfg c = cfg_of(r);
    for (uint32_t q = 0; q < 64; ++q) c.perm[q] = (uint8_t)q;       // the passes below are built in PHYSICAL numbering
    uint8_t at[64];                                                  // at[p] = qubit at index bit p
    for (uint32_t q = 0; q < r->q_num; ++q) at[r->perm[q]] = (uint8_t)q;
    std::vector<POp> pl;
    Plan plan;
    memcpy(plan.perm, c.perm, sizeof(plan.perm));
    auto swap_pos = [&](uint32_t a, uint32_t b) {
        const uint8_t qa = at[a], qb = at[b];
        at[a] = qb;
        at[b] = qa;
        r->perm[qa] = (uint8_t)b;
        r->perm[qb] = (uint8_t)a;
        for (qvnt_reg *s : g->shards) memcpy(s->perm, r->perm, sizeof(r->perm));
    };
    auto empty_pass = [&](uint32_t gb, uint32_t lb) -> int {
        if (g->shards.empty()) return empty_remap_pass(r, c, gb, lb);
        for (qvnt_reg *s : g->shards) {
            QV_CUDA(cudaSetDevice(s->device));
            PlanCfg cs = cfg_of(s);
            memcpy(cs.perm, c.perm, sizeof(cs.perm));
            int rc = empty_remap_pass(s, cs, gb, lb);
            if (rc) return rc;
        }
        return QVNT_OK;
    };
    // 1. global positions
    for (uint32_t gp = r->n_local; gp < r->q_num; ++gp) {
        while (at[gp] != gp) {
            uint32_t p = r->perm[gp];                // where the qubit that belongs at gp sits now
            if (p >= r->n_local) {
                // it sits at ANOTHER global position: bring it to a local bit first
                uint32_t b = r->n_local - 1;
                int rc = empty_pass(p, b);
                if (rc) return rc;
                swap_pos(p, b);
                continue;
            }
            int rc = empty_pass(gp, p);
            if (rc) return rc;
            swap_pos(gp, p);
        }
    }
    // 2. local positions: cycle decomposition into swaps of index bits
    std::vector<qvnt_op_t> sw;
    for (uint32_t p = 0; p < r->n_local; ++p) {
        while (at[p] != p) {
            const uint32_t q = r->perm[p];           // qubit p sits at index bit q
            qvnt_op_t o;
            memset(&o, 0, sizeof(o));
            o.kind = QVNT_SWAP;
            o.a_mask = (1ull << p) | (1ull << q);
            sw.push_back(o);
            swap_pos(p, q);
        }
    }
    // (perm is the identity now, so run_ops plans these swaps in plain index-bit numbering)
    if (!sw.empty()) {
        const int keep = r->opt_remap;
        const uint64_t ops_before = r->stats.ops_applied;
        r->opt_remap = 0;
        for (qvnt_reg *s : g->shards) s->opt_remap = 0;
        int rc = g->shards.empty() ? run_ops(r, sw.data(), sw.size()) : run_ops_group(g, sw.data(), sw.size());
        r->stats.ops_applied = ops_before;
        r->opt_remap = keep;
        for (qvnt_reg *s : g->shards) s->opt_remap = keep;
        if (rc) return rc;
    }
    return QVNT_OK;
}

// ---- group handles: one host thread drives every shard -----------------------------------------
// Pass k of every shard is enqueued before pass k + 1 of any: a shard's barrier kernel spins until
// the other shards' barrier kernels run, so no shard may get a whole plan ahead of the others (a
// full launch queue would block the host with the peers' kernels not yet enqueued).
static int run_plans_group(qvnt_reg *g, std::vector<std::vector<POp>> &pls, std::vector<Plan> &plans) {
    const size_t P = g->shards.size();
    std::vector<Uploaded> ups(P);
    for (size_t k = 0; k < P; ++k) {
        if (plans[k].passes.size() != plans[0].passes.size()) {
            set_error("internal: shards planned different pass counts");
            return QVNT_ERR_INVALID;
        }
        QV_CUDA(cudaSetDevice(g->shards[k]->device));
        int rc = upload_plan(g->shards[k], plans[k], ups[k]);
        if (rc) return rc;
    }
    std::vector<char> need(P, 0);
    for (size_t i = 0; i < plans[0].passes.size(); ++i)
        for (size_t k = 0; k < P; ++k) {
            QV_CUDA(cudaSetDevice(g->shards[k]->device));
            bool nb = need[k] != 0;
            int rc = enqueue_pass(g->shards[k], pls[k], plans[k].passes[i], ups[k], nb);
            need[k] = nb;
            if (rc) return rc;
        }
    for (size_t k = 0; k < P; ++k)
        if (need[k]) {
            QV_CUDA(cudaSetDevice(g->shards[k]->device));
            int rc = dist_barrier(g->shards[k]);
            if (rc) return rc;
        }
    return QVNT_OK;
}

int run_ops_group(qvnt_reg *g, const qvnt_op_t *ops, size_t n_ops) {
    const size_t P = g->shards.size();
    std::vector<std::vector<POp>> pls(P);
    std::vector<Plan> plans(P);
    for (size_t k = 0; k < P; ++k) {
        qvnt_reg *r = g->shards[k];
        std::vector<amp> mats;
        const PlanCfg c = cfg_of(r);
        int rc = lower_ops(c, ops, n_ops, pls[k], mats);
        if (rc) return rc;
        r->stats.ops_applied += k == 0 ? n_ops : 0;
        if (pls[k].empty()) continue;
        if ((rc = build_plan(c, pls[k], plans[k]))) return rc;
        if (!mats.empty()) {
            QV_CUDA(cudaSetDevice(r->device));
            const size_t bytes = mats.size() * sizeof(amp);
            if ((rc = ensure_dev((void **)&r->d_mat, &r->d_mat_cap, bytes))) return rc;
            QV_CUDA(cudaMemcpyAsync(r->d_mat, mats.data(), bytes, cudaMemcpyHostToDevice, r->stream));
            QV_CUDA(cudaStreamSynchronize(r->stream));       // (mats is a local)
            r->stats.h2d_bytes += bytes;
        }
    }
    if (pls[0].empty()) return QVNT_OK;
    int rc = run_plans_group(g, pls, plans);
    if (rc) return rc;
    for (size_t k = 0; k < P; ++k) memcpy(g->shards[k]->perm, plans[k].perm, sizeof(plans[k].perm));
    return QVNT_OK;
}

// Text dump of the schedule (tests, DESIGN.md, debugging); needs no device.
int describe_plan(uint32_t q_num, uint32_t rank, uint32_t world, int peers, int fuse, int tile_bits,
                  int chunk_bits, const qvnt_op_t *ops, size_t n_ops, std::string &out) {
    PlanCfg c;
    c.q_num = q_num;
    c.world = world ? world : 1;
    c.rank = rank;
    uint32_t wb = 0;
    while ((1u << wb) < c.world) ++wb;
    if (q_num < wb || q_num > 40 || (c.world & (c.world - 1)) || c.world > (uint32_t)MAX_WORLD || rank >= c.world) {
        set_error("bad register shape");
        return QVNT_ERR_INVALID;
    }
    c.n_local = q_num - wb;
    c.peers = peers != 0;
    c.remap = !(peers & 2);          // (peers_attached == 3: peers attached, remap passes off)
    c.peer_chunk_bits = tile_bits ? 0 : 6;     // the library defaults (reg.h) unless the caller fixes the geometry
    c.peer_tile_bits = tile_bits ? 0 : 12;
    for (uint32_t q = 0; q < 64; ++q) c.perm[q] = (uint8_t)q;
    c.fuse = (fuse & 1) != 0;
    c.lower_two_bit = (fuse & 2) != 0;       // (fuse == 3: with option "lower_two_bit")
    c.tile_bits = tile_bits;
    c.chunk_bits = chunk_bits;
    std::vector<POp> pl;
    std::vector<amp> mats;
    int rc = lower_ops(c, ops, n_ops, pl, mats);
    if (rc) return rc;
    Plan plan;
    if ((rc = build_plan(c, pl, plan))) return rc;
    char buf[1024];
    auto op_line = [&](const POp &p, int form, int ra, int rb) {
        // scale: the butterfly factor of an h1 (1/sqrt2; the halves of a split h2 carry 1 and 0.5)
        // a / b / ctrl: LOGICAL (qubit) masks as the caller gave them; pa / pb / pctrl: the index bits they
        // occupied when the op was scheduled (differ after a remap pass)
        snprintf(buf, sizeof(buf),
                 "op src=%u kind=%u dagger=%u a=%llu b=%llu ctrl=%llu pa=%llu pb=%llu pctrl=%llu form=%d ra=%d rb=%d "
                 "scale=%a\n",
                 p.src, p.d.kind, p.d.dagger, (unsigned long long)p.la, (unsigned long long)p.lb,
                 (unsigned long long)p.lctrl, (unsigned long long)p.d.a, (unsigned long long)p.d.b,
                 (unsigned long long)p.d.ctrl, form, ra, rb, p.d.kind == QVNT_H1 ? p.d.ph_re : 0.0);
        out += buf;
    };
    for (const PassPlan &pp : plan.passes) {
        if (pp.direct) {
            out += "pass direct\n";
            op_line(pl[pp.ops[0]], -1, -1, -1);
            continue;
        }
        const TPassHdr &h = pp.hdr;
        snprintf(buf, sizeof(buf),
                 "pass tile T=%u L=%u n_tiles=%llu base_or=%llu peer=%u full=%u fx_val=%llu remap=%u rg=%u rb=%u gpos=",
                 h.T, h.L, (unsigned long long)h.n_tiles, (unsigned long long)h.base_or, h.touches_peer, h.full,
                 (unsigned long long)h.fx.val, h.remap, h.remap_g, h.remap_b);
        out += buf;
        for (uint32_t l = 0; l < h.T; ++l) out += std::to_string(h.gpos[l]) + (l + 1 < h.T ? "," : "");
        out += " fx_pos=";
        for (uint32_t k = 0; k < h.fx.n; ++k) out += std::to_string(h.fx.pos[k]) + (k + 1 < h.fx.n ? "," : "");
        out += "\n";
        for (uint32_t s = 0; s < h.n_stages; ++s) {
            const TStage &st = plan.stages[h.stage_begin + s];
            out += "stage r=";
            for (int j = 0; j < TILE_R; ++j) out += std::to_string(st.r_lpos[j]) + (j + 1 < TILE_R ? "," : "");
            out += " t=";
            for (uint32_t k = 0; k + TILE_R < h.T; ++k) out += std::to_string(st.t_lpos[k]) + (k + 1 + TILE_R < h.T ? "," : "");
            out += " sync=" + std::to_string(st.sync_after_load) + "\n";
            for (uint32_t o = st.op_begin; o < st.op_end; ++o) {
                op_line(pl[plan.minfo[o].src], plan.minfo[o].form, plan.minfo[o].ra, plan.minfo[o].rb);
                // the encoded micro-op exactly as the kernel reads it (tests/test_tile_emulator.py)
                const MOp &m = plan.mops[o];
                const MBase &b = plan.bases[o];
                snprintf(buf, sizeof(buf),
                         "mop code=%u flags=%u okmask=%u ctrl_thr=%u a_thr=%u a_reg=%u idx=%u c0=%a c1=%a c2=%a c3=%a "
                         "a0=%a a1=%a a2=%a a3=%a ctrl_base=%llu a_base=%llu\n",
                         m.code, m.dagger, m.okmask, m.ctrl_thr, m.a_thr, m.a_reg, m.idx, m.ph_re, m.ph_im, m.c2, m.c3,
                         m.alt[0], m.alt[1], m.alt[2], m.alt[3], (unsigned long long)b.ctrl_base,
                         (unsigned long long)b.a_base);
                out += buf;
            }
        }
    }
    return QVNT_OK;
}

}  // namespace qv
