// api.cu -- the C ABI (include/qvnt_b200.h): register lifetime, measurement,
// data movement, instrumentation.  The hot path (qvnt_reg_apply) is scheduled in
// planner.cu.  Reference functions replaced are cited in the header.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "reg.h"

namespace qv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return e == cudaErrorMemoryAllocation ? QVNT_ERR_OOM : QVNT_ERR_CUDA;
}

int ensure_dev(void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *p) return QVNT_OK;
    if (*p) QV_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = bytes < 4096 ? 4096 : bytes;
    QV_CUDA(cudaMalloc(p, want));
    *cap = want;
    return QVNT_OK;
}

// Pinned staging buffer the op descriptors are copied from.  Before the host
// overwrites it, wait until the previous H2D copy out of it has executed.
int ensure_stage(qvnt_reg *r, size_t bytes) {
    if (r->stage_busy) {
        QV_CUDA(cudaEventSynchronize(r->stage_free));
        r->stage_busy = false;
    }
    if (r->h_stage_cap >= bytes) return QVNT_OK;
    if (r->h_stage) QV_CUDA(cudaFreeHost(r->h_stage));
    r->h_stage = nullptr;
    r->h_stage_cap = 0;
    size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
    QV_CUDA(cudaMallocHost(&r->h_stage, want));
    r->h_stage_cap = want;
    return QVNT_OK;
}

LaunchScope::LaunchScope(qvnt_reg *reg, int c) : r(reg), cls(c) {
    if (!r->opt_profile) return;
    auto get = [&]() {
        cudaEvent_t e;
        if (!r->event_pool.empty()) {
            e = r->event_pool.back();
            r->event_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, r->stream);
}
void LaunchScope::done(int n_launches) {
    if (n_launches > 0) r->stats.launches[cls] += (uint64_t)n_launches;
    if (!r->opt_profile) return;
    cudaEventRecord(b, r->stream);
    r->timed.push_back({cls, a, b});
}

static int fold_timed(qvnt_reg *r) {
    if (r->timed.empty()) return QVNT_OK;
    QV_CUDA(cudaStreamSynchronize(r->stream));
    for (auto &t : r->timed) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.a, t.b);
        r->stats.ms[t.cls] += (double)ms;
        r->event_pool.push_back(t.a);
        r->event_pool.push_back(t.b);
    }
    r->timed.clear();
    return QVNT_OK;
}

static int use(const qvnt_reg *r) {
    if (!r) {
        set_error("null register handle");
        return QVNT_ERR_INVALID;
    }
    QV_CUDA(cudaSetDevice(r->shards.empty() ? r->device : r->shards[0]->device));
    return QVNT_OK;
}

static uint64_t rank_bits(const qvnt_reg *r) { return (uint64_t)r->rank << r->n_local; }

static int init_state(qvnt_reg *r, uint64_t state) {
    for (uint32_t q = 0; q < 64; ++q) r->perm[q] = (uint8_t)q;      // nothing to keep: the qubit map starts over
    state &= r->q_mask;
    const uint64_t owner = state >> r->n_local;
    const uint64_t local = state & (r->local_len - 1);
    LaunchScope ls(r, 3);
    int n = launch_set_basis(r->stream, r->psi, r->local_len, owner == r->rank ? local : ~0ull);
    ls.done(n);
    r->stats.alg_bytes[3] += r->local_len * 16;
    if (n < 0) return cuda_fail(cudaGetLastError(), "set_basis");
    return QVNT_OK;
}

static int create_common(uint32_t q_num, uint64_t state, uint32_t rank, uint32_t world, int device,
                         qvnt_reg_t **out) {
    if (!out) {
        set_error("null out pointer");
        return QVNT_ERR_INVALID;
    }
    *out = nullptr;
    if (world == 0 || world > (uint32_t)MAX_WORLD || (world & (world - 1)) || rank >= world) {
        set_error("world must be 1/2/4/8 and rank < world (got rank %u world %u)", rank, world);
        return QVNT_ERR_INVALID;
    }
    uint32_t wb = 0;
    while ((1u << wb) < world) ++wb;
    if (q_num > 40 || q_num < wb) {
        set_error("q_num %u out of range for world %u", q_num, world);
        return QVNT_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return QVNT_ERR_CUDA;
    }
    if (device < 0) QV_CUDA(cudaGetDevice(&device));
    if (device >= ndev) {
        set_error("device %d out of range (%d devices)", device, ndev);
        return QVNT_ERR_INVALID;
    }
    QV_CUDA(cudaSetDevice(device));
    if (tile_kernel_setup() != 0) return cuda_fail(cudaGetLastError(), "tile kernel attributes");
    qvnt_reg *r = new (std::nothrow) qvnt_reg();
    if (!r) return QVNT_ERR_OOM;
    r->device = device;
    cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, device);
    r->q_num = q_num;
    r->rank = rank;
    r->world = world;
    r->world_bits = wb;
    r->n_local = q_num - wb;
    r->q_mask = q_num >= 64 ? ~0ull : ((1ull << q_num) - 1ull);
    r->local_len = 1ull << r->n_local;
    int rc = QVNT_OK;
    auto fail = [&](int code) {
        qvnt_reg_destroy(r);
        return code;
    };
    if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaStreamCreate"));
    e = cudaMalloc((void **)&r->psi, r->local_len * sizeof(amp));
    if (e != cudaSuccess) {
        set_error("cannot allocate %llu bytes of HBM for a %u-qubit shard: %s",
                  (unsigned long long)(r->local_len * sizeof(amp)), r->n_local, cudaGetErrorString(e));
        cudaGetLastError();
        return fail(QVNT_ERR_OOM);
    }
    if (cudaMalloc((void **)&r->d_partials, REDUCE_BLOCKS_MAX * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&r->d_scalars, 16 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&r->d_result, 8 * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc((void **)&r->mailbox, 4096) != cudaSuccess ||
        cudaMallocHost((void **)&r->h_scalars, 32 * sizeof(double)) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->stage_free, cudaEventDisableTiming) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "scratch allocation"));
    QV_CUDA(cudaMemsetAsync(r->mailbox, 0, 4096, r->stream));
    for (uint32_t q = 0; q < 64; ++q) r->perm[q] = (uint8_t)q;
    if (world > 1) {
        // one handshake word per tile of a remap pass (tile.cu); 2^22 words cover 2^11-amplitude tiles
        // of a 2^33-amplitude shard, smaller shards get a word per 2^8 amplitudes
        uint64_t cap = r->local_len >> 8;
        if (cap < 1) cap = 1;
        if (cap > (1ull << 22)) cap = 1ull << 22;
        if (cudaMalloc((void **)&r->ack, cap * sizeof(unsigned int)) != cudaSuccess)
            return fail(cuda_fail(cudaGetLastError(), "handshake words"));
        QV_CUDA(cudaMemsetAsync(r->ack, 0, cap * sizeof(unsigned int), r->stream));
        r->ack_cap = cap;
    }
    {
        // Everything the hot path and measure_mask need is allocated up front: allocation calls
        // synchronise the whole device, which must not happen while a peer shard's barrier kernel
        // is waiting for this shard (and costs milliseconds on the hot path).
        const uint64_t n1 = (r->local_len + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
        const uint64_t n2b = (n1 + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
        if ((rc = ensure_dev((void **)&r->d_l1, &r->l1_cap, n1 * sizeof(double))) ||
            (rc = ensure_dev((void **)&r->d_l2, &r->l2_cap, n2b * sizeof(double))) ||
            (rc = ensure_stage(r, (size_t)2 << 20)) || (rc = ensure_dev(&r->d_ops, &r->d_ops_cap, (size_t)2 << 20)) ||
            (rc = ensure_dev((void **)&r->d_mat, &r->d_mat_cap, (size_t)1 << 16)))
            return fail(rc);
    }
    for (int i = 0; i < MAX_WORLD; ++i) r->segs.seg[i] = nullptr;
    for (int i = 0; i < MAX_WORLD; ++i) r->segs.ack[i] = nullptr;
    r->segs.seg[rank] = r->psi;
    r->segs.ack[rank] = r->ack;
    r->segs.shift = r->n_local;
    r->segs.rank = rank;
    r->segs.world_bits = wb;
    r->mail[rank] = r->mailbox;
    rc = init_state(r, state);
    if (rc != QVNT_OK) return fail(rc);
    *out = r;
    return QVNT_OK;
}

static int norm_sqr_local(qvnt_reg *r, double *out) {
    LaunchScope ls(r, 2);
    int n = launch_norm_sqr(r->stream, r->psi, r->local_len, r->d_partials, r->d_scalars, r->sm_count);
    ls.done(n);
    r->stats.alg_bytes[2] += r->local_len * 16;
    if (n < 0) return cuda_fail(cudaGetLastError(), "norm_sqr");
    QV_CUDA(cudaMemcpyAsync(r->h_scalars, r->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2h_bytes += sizeof(double);
    *out = r->h_scalars[0];
    return QVNT_OK;
}

// sum |a|^2 over the whole register (all shards, accumulated in rank order)
static int norm_sqr_global(qvnt_reg *r, double *out) {
    double local = 0.0;
    int rc = norm_sqr_local(r, &local);
    if (rc) return rc;
    if (r->world == 1) {
        *out = local;
        return QVNT_OK;
    }
    double all[MAX_WORLD];
    rc = dist_allgather_double(r, local, all);
    if (rc) return rc;
    double s = 0.0;
    for (uint32_t k = 0; k < r->world; ++k) s += all[k];
    *out = s;
    return QVNT_OK;
}

static int normalize_impl(qvnt_reg *r) {
    double n2 = 0.0;
    int rc = norm_sqr_global(r, &n2);
    if (rc) return rc;
    const double norm = sqrt(n2);
    if (norm <= 1e-15) return init_state(r, 0);      // quant.rs:399-402
    if (1.0 - norm <= 1e-9) return QVNT_OK;          // quant.rs:403-405
    LaunchScope ls(r, 3);
    int n = launch_scale(r->stream, r->psi, r->local_len, 1.0 / norm);
    ls.done(n);
    r->stats.alg_bytes[3] += r->local_len * 32;
    return n < 0 ? cuda_fail(cudaGetLastError(), "scale") : QVNT_OK;
}

static int collapse_impl(qvnt_reg *r, uint64_t idy, uint64_t mask) {
    LaunchScope ls(r, 3);
    int n = launch_collapse(r->stream, r->psi, r->local_len, rank_bits(r), idy, mask);
    ls.done(n);
    r->stats.alg_bytes[3] += r->local_len * 16;   // write-only: zeroes the mismatching amplitudes
    return n < 0 ? cuda_fail(cudaGetLastError(), "collapse") : QVNT_OK;
}

static int check_local_range(qvnt_reg *r, uint64_t off, uint64_t cnt, uint64_t *local_off) {
    const uint64_t lo = rank_bits(r);
    if (off < lo || cnt > r->local_len || off - lo > r->local_len - cnt) {
        set_error("range [%llu, +%llu) is outside rank %u's shard", (unsigned long long)off,
                  (unsigned long long)cnt, r->rank);
        return QVNT_ERR_INVALID;
    }
    *local_off = off - lo;
    return QVNT_OK;
}

// ---- group handles (qvnt_reg_create_multi): one process, one host thread, every shard ---------
// The collectives of the one-process-per-GPU path (mailbox all-gathers, host syncs per rank) are
// replaced by "enqueue on every shard, then sync every shard, combine on the host" -- same kernels,
// same arithmetic order (rank order), so a group and an SPMD register give identical results.
static bool is_group(const qvnt_reg *r) { return r && !r->shards.empty(); }

static int g_each(qvnt_reg *g, int (*fn)(qvnt_reg *)) {
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        int rc = fn(s);
        if (rc) return rc;
    }
    return QVNT_OK;
}
static int g_sync(qvnt_reg *g) {
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        QV_CUDA(cudaStreamSynchronize(s->stream));
    }
    return QVNT_OK;
}

static int g_norm_sqr(qvnt_reg *g, double *out) {
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        LaunchScope ls(s, 2);
        int n = launch_norm_sqr(s->stream, s->psi, s->local_len, s->d_partials, s->d_scalars, s->sm_count);
        ls.done(n);
        s->stats.alg_bytes[2] += s->local_len * 16;
        if (n < 0) return cuda_fail(cudaGetLastError(), "norm_sqr");
        QV_CUDA(cudaMemcpyAsync(s->h_scalars, s->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        s->stats.d2h_bytes += sizeof(double);
    }
    int rc = g_sync(g);
    if (rc) return rc;
    double sum = 0.0;
    for (qvnt_reg *s : g->shards) sum += s->h_scalars[0];      // rank order, like norm_sqr_global
    *out = sum;
    return QVNT_OK;
}

static int g_init_state(qvnt_reg *g, uint64_t state) {
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        int rc = init_state(s, state);
        if (rc) return rc;
    }
    return QVNT_OK;
}

static int g_normalize(qvnt_reg *g) {
    double n2 = 0.0;
    int rc = g_norm_sqr(g, &n2);
    if (rc) return rc;
    const double norm = sqrt(n2);
    if (norm <= 1e-15) return g_init_state(g, 0);
    if (1.0 - norm <= 1e-9) return QVNT_OK;
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        LaunchScope ls(s, 3);
        int n = launch_scale(s->stream, s->psi, s->local_len, 1.0 / norm);
        ls.done(n);
        s->stats.alg_bytes[3] += s->local_len * 32;
        if (n < 0) return cuda_fail(cudaGetLastError(), "scale");
    }
    return QVNT_OK;
}

static int g_collapse(qvnt_reg *g, uint64_t idy, uint64_t mask) {
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        int rc = collapse_impl(s, idy, mask);
        if (rc) return rc;
    }
    return QVNT_OK;
}

static int g_measure(qvnt_reg *g, uint64_t mask, double u01, uint64_t *outcome, uint64_t *sampled) {
    qvnt_reg *r0 = g->shards[0];
    mask &= r0->q_mask;
    if (mask == 0) {
        *outcome = 0;
        if (sampled) *sampled = 0;
        return QVNT_OK;
    }
    double n2 = 0.0;
    int rc = g_norm_sqr(g, &n2);
    if (rc) return rc;
    if (!(n2 > 0.0)) {
        set_error("measure_mask on a register whose amplitudes are all zero");
        return QVNT_ERR_INVALID;
    }
    const double inv = 1.0 / n2;
    const uint64_t n1 = (r0->local_len + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    const uint64_t n2b = (n1 + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    for (qvnt_reg *s : g->shards) {
        QV_CUDA(cudaSetDevice(s->device));
        LaunchScope ls(s, 2);
        int n = launch_block_weights(s->stream, s->psi, s->local_len, inv, s->d_l1, s->d_l2);
        int m = launch_total(s->stream, s->d_l2, n2b, s->d_scalars + 1);
        ls.done(n + m);
        s->stats.alg_bytes[2] += s->local_len * 16;
        if (n < 0 || m < 0) return cuda_fail(cudaGetLastError(), "block_weights");
        QV_CUDA(cudaMemcpyAsync(s->h_scalars, s->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        s->stats.d2h_bytes += sizeof(double);
    }
    if ((rc = g_sync(g))) return rc;
    double total = 0.0;
    std::vector<double> prefix(g->shards.size());
    for (size_t k = 0; k < g->shards.size(); ++k) {
        prefix[k] = total;
        total += g->shards[k]->h_scalars[0];
    }
    const double x = u01 * total;
    for (size_t k = 0; k < g->shards.size(); ++k) {
        qvnt_reg *s = g->shards[k];
        QV_CUDA(cudaSetDevice(s->device));
        LaunchScope ls(s, 2);
        int n = launch_locate(s->stream, s->psi, s->local_len, inv, s->d_l1, n1, s->d_l2, n2b, prefix[k], x, s->d_result);
        ls.done(n);
        if (n < 0) return cuda_fail(cudaGetLastError(), "locate");
        QV_CUDA(cudaMemcpyAsync(s->h_scalars + 8, s->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
        s->stats.d2h_bytes += 4 * sizeof(uint64_t);
    }
    if ((rc = g_sync(g))) return rc;
    uint64_t idx = r0->q_mask;
    bool found = false;
    for (size_t k = 0; k < g->shards.size() && !found; ++k) {
        const uint64_t *h = (const uint64_t *)(g->shards[k]->h_scalars + 8);
        if (h[1]) {
            idx = rank_bits(g->shards[k]) | h[0];
            found = true;
        }
    }
    for (size_t k = g->shards.size(); k-- > 0 && !found;) {
        const uint64_t *h = (const uint64_t *)(g->shards[k]->h_scalars + 8);
        if (h[3] != ~0ull) {
            idx = rank_bits(g->shards[k]) | h[3];
            found = true;
        }
    }
    if ((rc = g_collapse(g, idx, mask))) return rc;
    *outcome = idx & mask;
    if (sampled) *sampled = idx;
    return QVNT_OK;
}

// [off, off + cnt) of the whole register, shard by shard
template <typename F>
static int g_ranges(qvnt_reg *g, uint64_t off, uint64_t cnt, F fn) {
    qvnt_reg *r0 = g->shards[0];
    if (off > r0->q_mask + 1 || cnt > r0->q_mask + 1 - off) {
        set_error("range [%llu, +%llu) is outside the register", (unsigned long long)off, (unsigned long long)cnt);
        return QVNT_ERR_INVALID;
    }
    uint64_t done = 0;
    while (done < cnt) {
        const uint64_t i = off + done;
        qvnt_reg *s = g->shards[i >> r0->n_local];
        const uint64_t lo = i & (s->local_len - 1);
        const uint64_t c = (cnt - done) < (s->local_len - lo) ? (cnt - done) : (s->local_len - lo);
        QV_CUDA(cudaSetDevice(s->device));
        int rc = fn(s, lo, c, done);
        if (rc) return rc;
        done += c;
    }
    return QVNT_OK;
}

}  // namespace qv

using namespace qv;

extern "C" {

int qvnt_version(void) { return QVNT_B200_VERSION; }
const char *qvnt_last_error(void) { return qv::g_err; }

int qvnt_device_count(int *out) {
    if (!out) return QVNT_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out = n;
    return QVNT_OK;
}

int qvnt_reg_create(uint32_t q_num, uint64_t state, qvnt_reg_t **out) {
    return create_common(q_num, state, 0, 1, -1, out);
}

int qvnt_reg_create_sharded(uint32_t q_num, uint64_t state, uint32_t rank, uint32_t world, int device,
                            qvnt_reg_t **out) {
    return create_common(q_num, state, rank, world, device, out);
}

int qvnt_reg_create_multi(uint32_t q_num, uint64_t state, uint32_t n_gpus, qvnt_reg_t **out) {
    if (!out) {
        set_error("null out pointer");
        return QVNT_ERR_INVALID;
    }
    *out = nullptr;
    if (n_gpus <= 1) return create_common(q_num, state, 0, 1, -1, out);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available; this library has no CPU fallback");
        return QVNT_ERR_CUDA;
    }
    // QVNT_MULTI_SHARE_DEVICES (read here only, a test facility): shards may share devices (shard k on
    // device k mod visible devices), so the one-handle path can be exercised on a box with fewer GPUs
    const bool share = getenv("QVNT_MULTI_SHARE_DEVICES") != nullptr;
    if ((n_gpus & (n_gpus - 1)) || n_gpus > (uint32_t)MAX_WORLD || (n_gpus > (uint32_t)ndev && !share)) {
        set_error("n_gpus must be 1, 2, 4 or 8 and at most the %d visible devices (got %u)", ndev, n_gpus);
        return QVNT_ERR_INVALID;
    }
    qvnt_reg *g = new (std::nothrow) qvnt_reg();
    if (!g) return QVNT_ERR_OOM;
    g->q_num = q_num;
    g->q_mask = q_num >= 64 ? ~0ull : ((1ull << q_num) - 1ull);
    g->world = n_gpus;
    int rc = QVNT_OK;
    for (uint32_t k = 0; k < n_gpus && rc == QVNT_OK; ++k) {
        qvnt_reg *sh = nullptr;
        rc = create_common(q_num, state, k, n_gpus, (int)(k % (uint32_t)ndev), &sh);
        if (rc == QVNT_OK) g->shards.push_back(sh);
    }
    // the shards find each other through the same blobs the one-process-per-GPU path exchanges
    std::vector<unsigned char> blobs((size_t)n_gpus * QVNT_IPC_BLOB_BYTES);
    for (uint32_t k = 0; k < n_gpus && rc == QVNT_OK; ++k)
        rc = qvnt_reg_export_ipc(g->shards[k], blobs.data() + (size_t)k * QVNT_IPC_BLOB_BYTES);
    for (uint32_t k = 0; k < n_gpus && rc == QVNT_OK; ++k) rc = qvnt_reg_attach_peers(g->shards[k], blobs.data());
    if (rc != QVNT_OK) {
        qvnt_reg_destroy(g);
        return rc;
    }
    g->n_local = g->shards[0]->n_local;
    g->local_len = g->shards[0]->local_len;
    *out = g;
    return QVNT_OK;
}

int qvnt_reg_set_gpus(qvnt_reg_t *r, uint32_t n_gpus, qvnt_reg_t **out) {
    int rc = use(r);
    if (rc) return rc;
    if (!out) return QVNT_ERR_INVALID;
    if (r->shards.empty() && r->world != 1) {
        set_error("set_gpus: the register is one shard of a one-process-per-GPU register");
        return QVNT_ERR_UNSUPPORTED;
    }
    if ((rc = restore_layout(r))) return rc;
    if ((rc = qvnt_reg_create_multi(r->q_num, 0, n_gpus, out))) return rc;
    qvnt_reg *d = *out;
    // copy the state shard by shard (device to device, across GPUs where the sharding differs)
    const uint64_t len = 1ull << r->q_num;
    const uint64_t src_len = r->shards.empty() ? len : r->local_len, dst_len = d->shards.empty() ? len : d->local_len;
    const uint64_t step = src_len < dst_len ? src_len : dst_len;
    if ((rc = qvnt_reg_sync(r)) || (rc = qvnt_reg_sync(d))) return rc;
    for (uint64_t i = 0; i < len; i += step) {
        qvnt_reg *ss = r->shards.empty() ? r : r->shards[i / src_len];
        qvnt_reg *ds = d->shards.empty() ? d : d->shards[i / dst_len];
        QV_CUDA(cudaMemcpyPeer(ds->psi + (i % dst_len), ds->device, ss->psi + (i % src_len), ss->device,
                               step * sizeof(amp)));
    }
    d->opt_fuse = r->opt_fuse;
    d->rng_state = r->rng_state;
    return QVNT_OK;
}

int qvnt_reg_destroy(qvnt_reg_t *r) {
    if (!r) return QVNT_OK;
    if (!r->shards.empty()) {
        for (qvnt_reg *s : r->shards) qvnt_reg_destroy(s);
        r->shards.clear();
        delete r;
        return QVNT_OK;
    }
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    for (int i = 0; i < MAX_WORLD; ++i) {
        if (r->peer_ptr[i]) cudaIpcCloseMemHandle(r->peer_ptr[i]);
        if (r->peer_mail[i]) cudaIpcCloseMemHandle(r->peer_mail[i]);
        if (r->peer_ack[i]) cudaIpcCloseMemHandle(r->peer_ack[i]);
    }
    for (auto &t : r->timed) {
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    for (auto e : r->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 16; ++i)
        if (r->marks[i]) cudaEventDestroy(r->marks[i]);
    if (r->stage_free) cudaEventDestroy(r->stage_free);
    cudaFree(r->psi);
    cudaFree(r->d_partials);
    cudaFree(r->d_scalars);
    cudaFree(r->d_result);
    cudaFree(r->d_l1);
    cudaFree(r->d_l2);
    cudaFree(r->d_tmp);
    cudaFree(r->d_ops);
    cudaFree(r->d_mat);
    cudaFree(r->mailbox);
    cudaFree(r->ack);
    if (r->h_stage) cudaFreeHost(r->h_stage);
    if (r->h_scalars) cudaFreeHost(r->h_scalars);
    if (r->stream) cudaStreamDestroy(r->stream);
    cudaGetLastError();
    delete r;
    return QVNT_OK;
}

int qvnt_reg_clone(qvnt_reg_t *r, qvnt_reg_t **out) {
    int rc = use(r);
    if (rc) return rc;
    if (is_group(r)) {
        if ((rc = restore_layout(r))) return rc;
        if ((rc = qvnt_reg_create_multi(r->q_num, 0, r->world, out))) return rc;
        qvnt_reg *c = *out;
        if ((rc = g_sync(r)) || (rc = g_sync(c))) return rc;
        for (size_t k = 0; k < r->shards.size(); ++k) {
            qvnt_reg *a = r->shards[k], *b = c->shards[k];
            QV_CUDA(cudaSetDevice(a->device));
            QV_CUDA(cudaMemcpyAsync(b->psi, a->psi, a->local_len * sizeof(amp), cudaMemcpyDeviceToDevice, a->stream));
            b->opt_fuse = a->opt_fuse;
            b->opt_tile_bits = a->opt_tile_bits;
            b->opt_chunk_bits = a->opt_chunk_bits;
            b->opt_remap = a->opt_remap;
            b->opt_peer_chunk_bits = a->opt_peer_chunk_bits;
            b->opt_peer_tile_bits = a->opt_peer_tile_bits;
            b->knobs = a->knobs;
        }
        c->rng_state = r->rng_state;
        return g_sync(r);
    }
    if (r->world != 1) {
        set_error("clone of one shard of a one-process-per-GPU register: clone every rank's shard");
        return QVNT_ERR_UNSUPPORTED;
    }
    rc = create_common(r->q_num, 0, 0, 1, r->device, out);
    if (rc) return rc;
    qvnt_reg *c = *out;
    QV_CUDA(cudaStreamSynchronize(c->stream));
    QV_CUDA(cudaMemcpyAsync(c->psi, r->psi, r->local_len * sizeof(amp), cudaMemcpyDeviceToDevice, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    c->opt_fuse = r->opt_fuse;
    c->opt_tile_bits = r->opt_tile_bits;
    c->opt_chunk_bits = r->opt_chunk_bits;
    c->knobs = r->knobs;
    c->rng_state = r->rng_state;
    return QVNT_OK;
}

int qvnt_reg_q_num(const qvnt_reg_t *r, uint32_t *out) {
    if (!r || !out) return QVNT_ERR_INVALID;
    *out = r->q_num;
    return QVNT_OK;
}

int qvnt_reg_apply(qvnt_reg_t *r, const qvnt_op_t *ops, size_t n_ops) {
    int rc = use(r);
    if (rc) return rc;
    if (n_ops == 0) return QVNT_OK;
    if (!ops) {
        set_error("null op array");
        return QVNT_ERR_INVALID;
    }
    return is_group(r) ? run_ops_group(r, ops, n_ops) : run_ops(r, ops, n_ops);
}

int qvnt_plan_describe(uint32_t q_num, uint32_t rank, uint32_t world, int peers_attached, int fuse,
                       int tile_bits, int chunk_bits, const qvnt_op_t *ops, size_t n_ops, char *out,
                       size_t cap, size_t *needed) {
    if (!ops && n_ops) return QVNT_ERR_INVALID;
    std::string s;
    int rc = describe_plan(q_num, rank, world, peers_attached, fuse, tile_bits, chunk_bits, ops, n_ops, s);
    if (rc) return rc;
    if (needed) *needed = s.size() + 1;
    if (out && cap) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return QVNT_OK;
}

int qvnt_reg_norm_sqr(qvnt_reg_t *r, double *out) {
    int rc = use(r);
    if (rc) return rc;
    if (!out) return QVNT_ERR_INVALID;
    return is_group(r) ? g_norm_sqr(r, out) : norm_sqr_global(r, out);
}

static int stream_out(qvnt_reg *r, uint64_t loff, uint64_t cnt, double *host, int per, double inv) {
    // per = doubles per amplitude in the output (1: probabilities, 2: polar)
    const uint64_t chunk = 1ull << 21;
    int rc = ensure_dev((void **)&r->d_tmp, &r->tmp_cap,
                        (size_t)(cnt < chunk ? cnt : chunk) * per * sizeof(double));
    if (rc) return rc;
    for (uint64_t done = 0; done < cnt; done += chunk) {
        const uint64_t c = cnt - done < chunk ? cnt - done : chunk;
        LaunchScope ls(r, 2);
        int n = per == 1 ? launch_probabilities(r->stream, r->psi, loff + done, c, inv, r->d_tmp)
                         : launch_polar(r->stream, r->psi, loff + done, c, r->d_tmp);
        ls.done(n);
        if (n < 0) return cuda_fail(cudaGetLastError(), "probabilities");
        QV_CUDA(cudaMemcpyAsync(host + done * per, r->d_tmp, c * per * sizeof(double), cudaMemcpyDeviceToHost,
                                r->stream));
        QV_CUDA(cudaStreamSynchronize(r->stream));
        r->stats.d2h_bytes += c * per * sizeof(double);
    }
    return QVNT_OK;
}

int qvnt_reg_probabilities(qvnt_reg_t *r, uint64_t off, uint64_t cnt, double *host_out) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if (cnt == 0) return QVNT_OK;
    if (!host_out) return QVNT_ERR_INVALID;
    if (is_group(r)) {
        double n2 = 0.0;
        if ((rc = g_norm_sqr(r, &n2))) return rc;
        const double inv = 1.0 / n2;
        return g_ranges(r, off, cnt, [&](qvnt_reg *s, uint64_t lo, uint64_t c, uint64_t done) {
            return stream_out(s, lo, c, host_out + done, 1, inv);
        });
    }
    uint64_t loff = 0;
    if ((rc = check_local_range(r, off, cnt, &loff))) return rc;
    double n2 = 0.0;
    if ((rc = norm_sqr_global(r, &n2))) return rc;
    return stream_out(r, loff, cnt, host_out, 1, 1.0 / n2);
}

int qvnt_reg_polar(qvnt_reg_t *r, uint64_t off, uint64_t cnt, double *host_r_theta) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if (cnt == 0) return QVNT_OK;
    if (!host_r_theta) return QVNT_ERR_INVALID;
    if (is_group(r))
        return g_ranges(r, off, cnt, [&](qvnt_reg *s, uint64_t lo, uint64_t c, uint64_t done) {
            return stream_out(s, lo, c, host_r_theta + 2 * done, 2, 0.0);
        });
    uint64_t loff = 0;
    if ((rc = check_local_range(r, off, cnt, &loff))) return rc;
    return stream_out(r, loff, cnt, host_r_theta, 2, 0.0);
}

int qvnt_reg_collapse(qvnt_reg_t *r, uint64_t idy, uint64_t mask) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    return is_group(r) ? g_collapse(r, idy, mask) : collapse_impl(r, idy, mask);
}

int qvnt_reg_measure_mask(qvnt_reg_t *r, uint64_t mask, double u01, uint64_t *outcome, uint64_t *sampled) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if (!outcome) return QVNT_ERR_INVALID;
    if (!(u01 >= 0.0 && u01 < 1.0)) {
        set_error("u01 must be in [0, 1)");
        return QVNT_ERR_INVALID;
    }
    if (is_group(r)) return g_measure(r, mask, u01, outcome, sampled);
    mask &= r->q_mask;
    if (mask == 0) {                      // quant.rs:491-494
        *outcome = 0;
        if (sampled) *sampled = 0;
        return QVNT_OK;
    }
    double n2 = 0.0;
    if ((rc = norm_sqr_global(r, &n2))) return rc;
    if (!(n2 > 0.0)) {                    // WeightedIndex::new(..).unwrap() panics on an all-zero register
        set_error("measure_mask on a register whose amplitudes are all zero");
        return QVNT_ERR_INVALID;
    }
    const double inv = 1.0 / n2;          // get_probabilities: abs = 1 / sum
    const uint64_t n1 = (r->local_len + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    const uint64_t n2b = (n1 + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    if ((rc = ensure_dev((void **)&r->d_l1, &r->l1_cap, n1 * sizeof(double)))) return rc;
    if ((rc = ensure_dev((void **)&r->d_l2, &r->l2_cap, n2b * sizeof(double)))) return rc;
    {
        LaunchScope ls(r, 2);
        int n = launch_block_weights(r->stream, r->psi, r->local_len, inv, r->d_l1, r->d_l2);
        int m = launch_total(r->stream, r->d_l2, n2b, r->d_scalars + 1);
        ls.done(n + m);
        r->stats.alg_bytes[2] += r->local_len * 16;
        if (n < 0 || m < 0) return cuda_fail(cudaGetLastError(), "block_weights");
    }
    QV_CUDA(cudaMemcpyAsync(r->h_scalars, r->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2h_bytes += sizeof(double);
    const double my_total = r->h_scalars[0];
    double totals[MAX_WORLD] = {my_total};
    if (r->world > 1 && (rc = dist_allgather_double(r, my_total, totals))) return rc;
    double total = 0.0, prefix = 0.0;
    for (uint32_t k = 0; k < r->world; ++k) {
        if (k == r->rank) prefix = total;
        total += totals[k];
    }
    const double x = u01 * total;         // UniformFloat::sample: u * scale + low, low = 0
    {
        LaunchScope ls(r, 2);
        int n = launch_locate(r->stream, r->psi, r->local_len, inv, r->d_l1, n1, r->d_l2, n2b, prefix, x,
                              r->d_result);
        ls.done(n);
        if (n < 0) return cuda_fail(cudaGetLastError(), "locate");
    }
    uint64_t *hres = (uint64_t *)(r->h_scalars + 8);
    QV_CUDA(cudaMemcpyAsync(hres, r->d_result, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2h_bytes += 4 * sizeof(uint64_t);
    uint64_t idx;
    if (r->world == 1) {
        idx = hres[0];                     // not found (x rounds past the total): the last index of non-zero weight
    } else {
        // first rank (in rank order) whose shard contains the crossing; none => the last index of
        // non-zero weight over all shards (bit 63 tags "fallback" entries)
        const uint64_t FB = 1ull << 63;
        uint64_t mine = hres[1] ? (rank_bits(r) | hres[0])
                                : (hres[3] != ~0ull ? (FB | rank_bits(r) | hres[3]) : ~0ull);
        uint64_t all[MAX_WORLD];
        if ((rc = dist_allgather_u64(r, mine, all))) return rc;
        idx = r->q_mask;
        bool found = false;
        for (uint32_t k = 0; k < r->world && !found; ++k)
            if (all[k] != ~0ull && !(all[k] & FB)) {
                idx = all[k];
                found = true;
            }
        for (uint32_t k = r->world; k-- > 0 && !found;)
            if (all[k] != ~0ull) {
                idx = all[k] & ~FB;
                found = true;
            }
    }
    if ((rc = collapse_impl(r, idx, mask))) return rc;
    *outcome = idx & mask;
    if (sampled) *sampled = idx;
    return QVNT_OK;
}

int qvnt_reg_measure_mask_rng(qvnt_reg_t *r, uint64_t mask, uint64_t *outcome) {
    if (!r) return QVNT_ERR_INVALID;
    // splitmix64 -> 53-bit uniform in [0,1); all ranks of a sharded register share the seed
    uint64_t z = (r->rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    return qvnt_reg_measure_mask(r, mask, u, outcome, nullptr);
}

int qvnt_reg_normalize(qvnt_reg_t *r) {
    int rc = use(r);
    if (rc) return rc;
    return is_group(r) ? g_normalize(r) : normalize_impl(r);
}

int qvnt_reg_reset(qvnt_reg_t *r, uint64_t state) {
    int rc = use(r);
    if (rc) return rc;
    return is_group(r) ? g_init_state(r, state) : init_state(r, state);
}

int qvnt_reg_reset_by_mask(qvnt_reg_t *r, uint64_t mask) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if ((mask & r->q_mask) == r->q_mask) return is_group(r) ? g_init_state(r, 0) : init_state(r, 0);   // quant.rs:208-210
    if (is_group(r)) {
        for (qvnt_reg *s : r->shards) {
            QV_CUDA(cudaSetDevice(s->device));
            LaunchScope ls(s, 3);
            int n = launch_zero_mask(s->stream, s->psi, s->local_len, rank_bits(s), mask);
            ls.done(n);
            if (n < 0) return cuda_fail(cudaGetLastError(), "zero_mask");
        }
        return g_normalize(r);
    }
    {
        LaunchScope ls(r, 3);
        int n = launch_zero_mask(r->stream, r->psi, r->local_len, rank_bits(r), mask);
        ls.done(n);
        if (n < 0) return cuda_fail(cudaGetLastError(), "zero_mask");
    }
    return normalize_impl(r);
}

int qvnt_reg_read(qvnt_reg_t *r, uint64_t off, uint64_t cnt, double *host) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if (cnt == 0) return QVNT_OK;
    if (!host) return QVNT_ERR_INVALID;
    if (is_group(r))
        return g_ranges(r, off, cnt, [&](qvnt_reg *s, uint64_t lo, uint64_t c, uint64_t done) {
            QV_CUDA(cudaMemcpyAsync(host + 2 * done, s->psi + lo, c * sizeof(amp), cudaMemcpyDeviceToHost, s->stream));
            QV_CUDA(cudaStreamSynchronize(s->stream));
            s->stats.d2h_bytes += c * sizeof(amp);
            return (int)QVNT_OK;
        });
    uint64_t loff = 0;
    if ((rc = check_local_range(r, off, cnt, &loff))) return rc;
    QV_CUDA(cudaMemcpyAsync(host, r->psi + loff, cnt * sizeof(amp), cudaMemcpyDeviceToHost, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2h_bytes += cnt * sizeof(amp);
    return QVNT_OK;
}

int qvnt_reg_write(qvnt_reg_t *r, uint64_t off, uint64_t cnt, const double *host) {
    int rc = use(r);
    if (rc) return rc;
    if ((rc = restore_layout(r))) return rc;      // qubits back at their own index bits
    if (cnt == 0) return QVNT_OK;
    if (!host) return QVNT_ERR_INVALID;
    if (is_group(r))
        return g_ranges(r, off, cnt, [&](qvnt_reg *s, uint64_t lo, uint64_t c, uint64_t done) {
            QV_CUDA(cudaMemcpyAsync(s->psi + lo, host + 2 * done, c * sizeof(amp), cudaMemcpyHostToDevice, s->stream));
            QV_CUDA(cudaStreamSynchronize(s->stream));
            s->stats.h2d_bytes += c * sizeof(amp);
            return (int)QVNT_OK;
        });
    uint64_t loff = 0;
    if ((rc = check_local_range(r, off, cnt, &loff))) return rc;
    QV_CUDA(cudaMemcpyAsync(r->psi + loff, host, cnt * sizeof(amp), cudaMemcpyHostToDevice, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.h2d_bytes += cnt * sizeof(amp);
    return QVNT_OK;
}

int qvnt_reg_tensor_prod(qvnt_reg_t *a, qvnt_reg_t *b, qvnt_reg_t **out) {
    int rc = use(a);
    if (rc) return rc;
    if (!b || !out) return QVNT_ERR_INVALID;
    if (is_group(a) || is_group(b) || a->world != 1 || b->world != 1 || a->device != b->device) {
        set_error("tensor_prod needs two single-GPU registers on the same device");
        return QVNT_ERR_UNSUPPORTED;
    }
    rc = create_common(a->q_num + b->q_num, 0, 0, 1, a->device, out);
    if (rc) return rc;
    qvnt_reg *o = *out;
    QV_CUDA(cudaStreamSynchronize(a->stream));
    QV_CUDA(cudaStreamSynchronize(b->stream));
    LaunchScope ls(o, 3);
    int n = launch_tensor_prod(o->stream, a->psi, a->q_num, b->psi, b->q_num, o->psi, 0, o->local_len);
    ls.done(n);
    if (n < 0) return cuda_fail(cudaGetLastError(), "tensor_prod");
    QV_CUDA(cudaStreamSynchronize(o->stream));
    return QVNT_OK;
}

// combine (quant.rs:245-271): a register of q + 1 qubits whose lower half is `a` and upper half `b`
static int combine_common(qvnt_reg_t *a, qvnt_reg_t *b, const double *m8, qvnt_reg_t **out) {
    int rc = use(a);
    if (rc) return rc;
    if (!b || !out) return QVNT_ERR_INVALID;
    if (is_group(a) || is_group(b) || a->world != 1 || b->world != 1 || a->device != b->device) {
        set_error("combine needs two single-GPU registers on the same device");
        return QVNT_ERR_UNSUPPORTED;
    }
    if (a->q_num != b->q_num) {                      // the reference answers None
        set_error("combine: registers of %u and %u qubits", a->q_num, b->q_num);
        return QVNT_ERR_INVALID;
    }
    if (a->q_num < 3) {
        set_error("combine: registers below 3 qubits keep padding amplitudes the halves would interleave");
        return QVNT_ERR_UNSUPPORTED;
    }
    if ((rc = create_common(a->q_num + 1, 0, 0, 1, a->device, out))) return rc;
    qvnt_reg *o = *out;
    QV_CUDA(cudaStreamSynchronize(a->stream));
    QV_CUDA(cudaStreamSynchronize(b->stream));
    if (!m8) {
        QV_CUDA(cudaMemcpyAsync(o->psi, a->psi, a->local_len * sizeof(amp), cudaMemcpyDeviceToDevice, o->stream));
        QV_CUDA(cudaMemcpyAsync(o->psi + a->local_len, b->psi, b->local_len * sizeof(amp), cudaMemcpyDeviceToDevice,
                                o->stream));
    } else {
        LaunchScope ls(o, 3);
        int n = launch_combine_unitary(o->stream, a->psi, b->psi, a->q_num, m8, o->psi);
        ls.done(n);
        if (n < 0) return cuda_fail(cudaGetLastError(), "combine");
    }
    QV_CUDA(cudaStreamSynchronize(o->stream));
    return QVNT_OK;
}
int qvnt_reg_combine(qvnt_reg_t *a, qvnt_reg_t *b, qvnt_reg_t **out) { return combine_common(a, b, nullptr, out); }
int qvnt_reg_combine_unitary(qvnt_reg_t *a, qvnt_reg_t *b, const double *matrix8, qvnt_reg_t **out) {
    if (!matrix8) return QVNT_ERR_INVALID;
    return combine_common(a, b, matrix8, out);
}

int qvnt_reg_linear_composition(qvnt_reg_t *r, qvnt_reg_t *other, double c0_re, double c0_im, double c1_re,
                                double c1_im) {
    int rc = use(r);
    if (rc) return rc;
    if (!other) return QVNT_ERR_INVALID;
    if (is_group(r) || is_group(other) || r->world != 1 || other->world != 1 || r->device != other->device ||
        r->q_num != other->q_num) {
        set_error("linear_composition needs two single-GPU registers of the same size on the same device");
        return QVNT_ERR_UNSUPPORTED;
    }
    QV_CUDA(cudaStreamSynchronize(other->stream));
    LaunchScope ls(r, 3);
    int n = launch_linear_composition(r->stream, r->psi, other->psi, r->local_len, make_double2(c0_re, c0_im),
                                      make_double2(c1_re, c1_im));
    ls.done(n);
    r->stats.alg_bytes[3] += r->local_len * 48;
    return n < 0 ? cuda_fail(cudaGetLastError(), "linear_composition") : QVNT_OK;
}

// sample_all (quant.rs:513-594): the histogram of `count` shots in the reference's Gaussian
// approximation, computed on the device (two reads of the state, no 2^n-sized scratch beyond the
// output staging), copied out in chunks; the final +-delta correction of the reference runs on
// the host over the returned array.
int qvnt_reg_sample_all(qvnt_reg_t *r, uint64_t count, uint64_t seed, uint64_t *host_out) {
    int rc = use(r);
    if (rc) return rc;
    if (!host_out) return QVNT_ERR_INVALID;
    if ((rc = restore_layout(r))) return rc;
    if (!is_group(r) && r->world != 1) {
        set_error("sample_all: call it on a single-GPU register or on a one-handle multi-GPU register");
        return QVNT_ERR_UNSUPPORTED;
    }
    std::vector<qvnt_reg *> sh = is_group(r) ? r->shards : std::vector<qvnt_reg *>{r};
    double n2 = 0.0;
    if ((rc = is_group(r) ? g_norm_sqr(r, &n2) : norm_sqr_global(r, &n2))) return rc;
    if (!(n2 > 0.0)) {
        set_error("sample_all on a register whose amplitudes are all zero");
        return QVNT_ERR_INVALID;
    }
    const double inv = 1.0 / n2, c = (double)count;
    // pass 1: sum of the noise terms (rank order)
    for (qvnt_reg *s : sh) {
        QV_CUDA(cudaSetDevice(s->device));
        LaunchScope ls(s, 2);
        int n = launch_sample_noise_sum(s->stream, s->psi, s->local_len, rank_bits(s), inv, seed, s->d_partials,
                                        s->d_scalars + 2, s->sm_count);
        ls.done(n);
        s->stats.alg_bytes[2] += s->local_len * 16;
        if (n < 0) return cuda_fail(cudaGetLastError(), "sample_all");
        QV_CUDA(cudaMemcpyAsync(s->h_scalars + 2, s->d_scalars + 2, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    double n_sum = 0.0;
    for (qvnt_reg *s : sh) {
        QV_CUDA(cudaSetDevice(s->device));
        QV_CUDA(cudaStreamSynchronize(s->stream));
        n_sum += s->h_scalars[2];
    }
    // pass 2: the counts, chunk by chunk through the staging buffer
    const uint64_t chunk = 1ull << 21;
    unsigned long long total = 0;
    for (qvnt_reg *s : sh) {
        QV_CUDA(cudaSetDevice(s->device));
        if ((rc = ensure_dev((void **)&s->d_tmp, &s->tmp_cap, (size_t)chunk * sizeof(uint64_t) + 64))) return rc;
        unsigned long long *d_cnt = (unsigned long long *)s->d_tmp;
        unsigned long long *d_total = d_cnt + chunk;
        QV_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s->stream));
        for (uint64_t done = 0; done < s->local_len; done += chunk) {
            const uint64_t cnt = s->local_len - done < chunk ? s->local_len - done : chunk;
            LaunchScope ls(s, 2);
            int n = launch_sample_counts(s->stream, s->psi, done, cnt, rank_bits(s), inv, seed, c, n_sum, d_cnt, d_total);
            ls.done(n);
            if (n < 0) return cuda_fail(cudaGetLastError(), "sample_all");
            QV_CUDA(cudaMemcpyAsync(host_out + (rank_bits(s) | done), d_cnt, cnt * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, s->stream));
            QV_CUDA(cudaStreamSynchronize(s->stream));
            s->stats.d2h_bytes += cnt * sizeof(uint64_t);
        }
        s->stats.alg_bytes[2] += s->local_len * 16;
        unsigned long long t = 0;
        QV_CUDA(cudaMemcpy(&t, d_total, sizeof(t), cudaMemcpyDeviceToHost));
        total += t;
    }
    // quant.rs:568-591: spread the rounding surplus / deficit over the histogram
    const uint64_t len = 1ull << r->q_num, q_mask = r->q_mask;
    if (total < count) {
        const uint64_t delta = count - total;
        const uint64_t all = delta >> r->q_num, first = q_mask ? delta % q_mask : 0;
        for (uint64_t i = 0; i < len; ++i) host_out[i] += all + (i < first ? 1 : 0);
    } else if (total > count) {
        uint64_t delta = total - count;
        for (uint64_t idx = 0; delta; ++idx) {
            uint64_t &v = host_out[idx & q_mask];
            if (v == 0) continue;
            v -= 1;
            delta -= 1;
        }
    }
    return QVNT_OK;
}

int qvnt_reg_sync(qvnt_reg_t *r) {
    int rc = use(r);
    if (rc) return rc;
    if (is_group(r)) return g_sync(r);
    QV_CUDA(cudaStreamSynchronize(r->stream));
    return QVNT_OK;
}

int qvnt_reg_set_option(qvnt_reg_t *r, const char *key, int64_t value) {
    if (!r || !key) return QVNT_ERR_INVALID;
    if (is_group(r)) {
        if (!strcmp(key, "seed")) {
            r->rng_state = (uint64_t)value;
            return QVNT_OK;
        }
        if (!strcmp(key, "remap") && !value) {
            int rc = restore_layout(r);
            if (rc) return rc;
        }
        for (qvnt_reg *s : r->shards) {
            if (!strcmp(key, "remap")) {
                s->opt_remap = value != 0;
                continue;
            }
            if (!strcmp(key, "peer_chunk_bits")) {
                s->opt_peer_chunk_bits = (int)value;
                continue;
            }
            if (!strcmp(key, "peer_tile_bits")) {
                s->opt_peer_tile_bits = (int)value;
                continue;
            }
            int rc = qvnt_reg_set_option(s, key, value);
            if (rc) return rc;
        }
        return QVNT_OK;
    }
    if (!strcmp(key, "fuse")) r->opt_fuse = value != 0;
    else if (!strcmp(key, "tile_bits")) {
        if (value != 0 && (value < 6 || value > TILE_MAX_BITS)) {
            set_error("tile_bits must be 0 (auto) or 6..%d", TILE_MAX_BITS);
            return QVNT_ERR_INVALID;
        }
        r->opt_tile_bits = (int)value;
    } else if (!strcmp(key, "chunk_bits")) {
        if (value != 0 && (value < 2 || value > TILE_MAX_BITS)) {
            set_error("chunk_bits must be 0 (auto) or 2..%d", TILE_MAX_BITS);
            return QVNT_ERR_INVALID;
        }
        r->opt_chunk_bits = (int)value;
    } else if (!strcmp(key, "tma")) r->knobs.bulk = value < 0 ? -1 : value != 0;
    else if (!strcmp(key, "remap")) {
        int rc = use(r);
        if (rc) return rc;
        if (!value && (rc = restore_layout(r))) return rc;
        r->opt_remap = value != 0;
    } else if (!strcmp(key, "peer_chunk_bits")) {
        if (value != 0 && (value < 2 || value > TILE_MAX_BITS)) {
            set_error("peer_chunk_bits must be 0 (same as chunk_bits) or 2..%d", TILE_MAX_BITS);
            return QVNT_ERR_INVALID;
        }
        r->opt_peer_chunk_bits = (int)value;
    } else if (!strcmp(key, "peer_tile_bits")) {
        if (value != 0 && (value < 6 || value > TILE_MAX_BITS)) {
            set_error("peer_tile_bits must be 0 (same as tile_bits) or 6..%d", TILE_MAX_BITS);
            return QVNT_ERR_INVALID;
        }
        r->opt_peer_tile_bits = (int)value;
    } else if (!strcmp(key, "prefetch")) r->knobs.prefetch = value != 0;
    else if (!strcmp(key, "ptx_ops")) r->knobs.ptx_ops = value != 0;
    else if (!strcmp(key, "single_ctrl")) r->knobs.single_ctrl = value != 0;
    else if (!strcmp(key, "butterfly")) r->knobs.butterfly = value != 0;
    else if (!strcmp(key, "lower_two_bit")) r->knobs.lower_two_bit = value != 0;
    else if (!strcmp(key, "double_buffer")) r->knobs.double_buffer = value < 0 || value > 2 ? 0 : (int)value;
    else if (!strcmp(key, "tile_ctas")) {
        if (value != 0 && (value < 3 || value > 5)) {
            set_error("tile_ctas must be 0 (auto) or 3..5");
            return QVNT_ERR_INVALID;
        }
        r->knobs.ctas_per_sm = (int)value;
    } else if (!strcmp(key, "profile")) {
        int rc = use(r);
        if (rc) return rc;
        if (!value) fold_timed(r);
        r->opt_profile = value != 0;
    } else if (!strcmp(key, "seed")) r->rng_state = (uint64_t)value;
    else {
        set_error("unknown option '%s'", key);
        return QVNT_ERR_INVALID;
    }
    return QVNT_OK;
}

int qvnt_reg_stats(qvnt_reg_t *r, qvnt_stats_t *out) {
    int rc = use(r);
    if (rc) return rc;
    if (!out) return QVNT_ERR_INVALID;
    if (is_group(r)) {
        // launches / bytes: summed over the shards; device time per class: the slowest shard;
        // ops_applied / passes: per register (shard 0)
        memset(out, 0, sizeof(*out));
        for (size_t k = 0; k < r->shards.size(); ++k) {
            qvnt_stats_t st;
            if ((rc = qvnt_reg_stats(r->shards[k], &st))) return rc;
            for (int c = 0; c < QVNT_STATS_CLASSES; ++c) {
                out->launches[c] += st.launches[c];
                out->alg_bytes[c] += st.alg_bytes[c];
                if (st.ms[c] > out->ms[c]) out->ms[c] = st.ms[c];
            }
            out->h2d_bytes += st.h2d_bytes;
            out->d2h_bytes += st.d2h_bytes;
            out->peer_bytes += st.peer_bytes;
            if (k == 0) {
                out->ops_applied = st.ops_applied;
                out->passes = st.passes;
            }
        }
        return QVNT_OK;
    }
    if ((rc = fold_timed(r))) return rc;
    *out = r->stats;
    return QVNT_OK;
}

int qvnt_reg_stats_reset(qvnt_reg_t *r) {
    int rc = use(r);
    if (rc) return rc;
    if (is_group(r)) {
        for (qvnt_reg *s : r->shards)
            if ((rc = qvnt_reg_stats_reset(s))) return rc;
        return QVNT_OK;
    }
    if ((rc = fold_timed(r))) return rc;
    memset(&r->stats, 0, sizeof(r->stats));
    return QVNT_OK;
}

int qvnt_reg_mark(qvnt_reg_t *r, int slot) {
    int rc = use(r);
    if (rc) return rc;
    if (slot < 0 || slot >= 16) return QVNT_ERR_INVALID;
    if (is_group(r)) {
        for (qvnt_reg *s : r->shards)
            if ((rc = qvnt_reg_mark(s, slot))) return rc;
        return QVNT_OK;
    }
    if (!r->marks[slot]) QV_CUDA(cudaEventCreate(&r->marks[slot]));
    QV_CUDA(cudaEventRecord(r->marks[slot], r->stream));
    r->mark_set[slot] = true;
    return QVNT_OK;
}

int qvnt_reg_elapsed_ms(qvnt_reg_t *r, int from, int to, double *ms) {
    int rc = use(r);
    if (rc) return rc;
    if (is_group(r)) {               // the slowest shard
        if (!ms) return QVNT_ERR_INVALID;
        *ms = 0.0;
        for (qvnt_reg *s : r->shards) {
            double m = 0.0;
            if ((rc = qvnt_reg_elapsed_ms(s, from, to, &m))) return rc;
            if (m > *ms) *ms = m;
        }
        return QVNT_OK;
    }
    if (from < 0 || from >= 16 || to < 0 || to >= 16 || !ms || !r->mark_set[from] || !r->mark_set[to])
        return QVNT_ERR_INVALID;
    QV_CUDA(cudaEventSynchronize(r->marks[to]));
    float f = 0.f;
    QV_CUDA(cudaEventElapsedTime(&f, r->marks[from], r->marks[to]));
    *ms = (double)f;
    return QVNT_OK;
}

}  // extern "C"
