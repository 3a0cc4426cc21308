// planner.cu -- host-side scheduler of qvnt_reg_apply: validates the op list,
// packs runs of adjacent SingleOps into fused tile passes and enqueues the
// kernels.  Replaces the reference's "one full out-of-place sweep per SingleOp"
// driver: QReg::apply (src/register/quant.rs:376-395) -> MultiOp::apply
// (src/operator/multi/mod.rs:96-114) -> SingleOp::apply (single/mod.rs:83-93).
//
// Scheduling rule (adjacency fusion, order preserving => results identical to
// the op-by-op order): walk the list front to back and keep extending the
// current pass while the union of the ops' MIX bits (the bits an op XORs when it
// gathers its partners) still fits the tile: low `chunk` bits are always in the
// tile, at most T - chunk further bits can be gathered.  Diagonal ops and
// control bits never constrain the tile -- outside the tile they are per-tile
// constants.  A pass of one op that needs no peer memory runs as a direct sweep.
#include <algorithm>
#include <cstring>

#include "reg.h"

namespace qv {

static inline int pc64(uint64_t v) { return __builtin_popcountll(v); }

struct POp {
    DevOp d;        // masks in GLOBAL numbering, ctrl already stripped of satisfied rank bits
    int cls;
    uint64_t mix;   // bits that must lie inside a tile (0 for diagonal ops)
};

static int validate(const qvnt_reg *r, const qvnt_op_t &o, size_t k) {
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

int run_ops(qvnt_reg *r, const qvnt_op_t *ops, size_t n_ops) {
    const uint64_t lmask = r->local_len - 1;
    const uint64_t gmask = r->q_mask & ~lmask;
    const uint64_t rbits = (uint64_t)r->rank << r->n_local;

    std::vector<POp> pl;
    pl.reserve(n_ops);
    std::vector<amp> mats;
    for (size_t k = 0; k < n_ops; ++k) {
        int rc = validate(r, ops[k], k);
        if (rc) return rc;
        const qvnt_op_t &o = ops[k];
        POp p;
        memset(&p, 0, sizeof(p));
        p.cls = op_class(o.kind);
        if (p.cls == CLS_NONE) continue;                     // Id
        if (p.cls == CLS_PAIR && o.a_mask == 0) continue;    // x(0) / y(0): identity
        if (p.cls == CLS_DIAG && o.a_mask == 0 && o.kind != QVNT_RZ && o.kind != QVNT_RZZ) continue;
        p.d.kind = o.kind;
        p.d.dagger = o.dagger ? 1u : 0u;
        p.d.a = o.a_mask;
        p.d.b = (o.kind == QVNT_H2 || o.kind == QVNT_U2) ? o.b_mask : 0;
        p.d.ctrl = o.ctrl;
        p.d.ph_re = o.phase_re;
        p.d.ph_im = o.phase_im;
        if (o.kind == QVNT_U1 || o.kind == QVNT_U2) {
            const int cnt = o.kind == QVNT_U1 ? 4 : 16;
            p.d.mat = (uint32_t)mats.size();
            for (int i = 0; i < cnt; ++i) mats.push_back(make_double2(o.matrix[2 * i], o.matrix[2 * i + 1]));
        }
        p.mix = p.cls == CLS_PAIR ? p.d.a : (p.cls == CLS_QUAD ? (p.d.a | p.d.b) : 0);
        pl.push_back(p);
    }
    r->stats.ops_applied += n_ops;
    if (pl.empty()) return QVNT_OK;

    // matrices of u1/u2 ops -> device table
    if (!mats.empty()) {
        const size_t bytes = mats.size() * sizeof(amp);
        int rc = ensure_stage(r, bytes);
        if (rc) return rc;
        if ((rc = ensure_dev((void **)&r->d_mat, &r->d_mat_cap, bytes))) return rc;
        memcpy(r->h_stage, mats.data(), bytes);
        QV_CUDA(cudaMemcpyAsync(r->d_mat, r->h_stage, bytes, cudaMemcpyHostToDevice, r->stream));
        QV_CUDA(cudaEventRecord(r->stage_free, r->stream));
        r->stage_busy = true;
        r->stats.h2d_bytes += bytes;
    }

    for (size_t k = 0; k < pl.size(); ++k) {
        POp &p = pl[k];
        if (p.mix & gmask) {
            set_error("gate on a global (sharded) qubit needs attached peers; not available in this call path");
            return QVNT_ERR_UNSUPPORTED;
        }
        // control bits on rank bits: this GPU either takes part or idles
        const uint64_t cg = p.d.ctrl & gmask;
        r->stats.passes += 1;
        if (cg & ~rbits) continue;
        DevOp d = p.d;
        d.ctrl &= lmask;
        LaunchScope ls(r, 0);
        uint64_t touched = 0;
        int n = launch_direct(r->stream, r->psi, r->n_local, d, r->d_mat, rbits, &touched);
        ls.done(n);
        r->stats.alg_bytes[0] += touched * 32;
        if (n < 0) {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "direct sweep launch");
            set_error("internal: direct sweep rejected op kind %u", d.kind);
            return QVNT_ERR_INVALID;
        }
    }
    return QVNT_OK;
}

}  // namespace qv
