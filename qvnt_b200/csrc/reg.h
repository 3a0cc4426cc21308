// reg.h -- the device-resident register behind the opaque qvnt_reg_t handle.
#pragma once
#include <string>
#include <vector>

#include "engine.h"

struct qvnt_reg {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;

    uint32_t q_num = 0;       // qubits of the whole register
    uint32_t rank = 0, world = 1, world_bits = 0;
    uint32_t n_local = 0;     // index bits held by this GPU
    uint64_t q_mask = 0;      // 2^q_num - 1
    uint64_t local_len = 1;   // 2^n_local amplitudes
    qv::amp *psi = nullptr;   // this GPU's shard
    qv::Segs segs{};          // all shards (peers mapped over NVLink IPC)
    bool peers_attached = false;
    void *peer_ptr[qv::MAX_WORLD] = {};     // opened IPC mappings (state)
    void *peer_mail[qv::MAX_WORLD] = {};    // opened IPC mappings (mailbox)
    // mailbox: cross-GPU barrier counters + small all-gather slots, IPC-shared
    unsigned long long *mailbox = nullptr;  // this rank's mailbox (device memory)
    unsigned long long *mail[qv::MAX_WORLD] = {};
    uint64_t barrier_epoch = 0;
    uint64_t gather_epoch = 0;
    // qubit remapping (sharded registers): logical qubit q lives at index bit perm[q]; identity until a
    // remap pass of qvnt_reg_apply trades a global for a local bit, restored lazily (restore_layout)
    uint8_t perm[64];
    unsigned int *ack = nullptr;            // per-tile handshake words of remap passes (IPC-shared)
    uint64_t ack_cap = 0;                   // entries
    void *peer_ack[qv::MAX_WORLD] = {};
    uint32_t remap_epoch = 0;
    int opt_remap = 1;
    int opt_peer_chunk_bits = 6;    // passes that start on a global qubit: 1 KiB chunks, 2^12-amplitude tiles (measured at 4 GPUs:
    int opt_peer_tile_bits = 12;    // +26 % over the local-pass geometry)
    bool remap_possible = false;            // one GPU per shard (set by attach_peers)

    // scratch (device)
    double *d_partials = nullptr;   // REDUCE_BLOCKS_MAX
    double *d_scalars = nullptr;    // 16 doubles
    uint64_t *d_result = nullptr;   // 8 words
    double *d_l1 = nullptr, *d_l2 = nullptr;
    size_t l1_cap = 0, l2_cap = 0;   // bytes
    double *d_tmp = nullptr;        // probabilities / polar staging
    size_t tmp_cap = 0;             // bytes
    // op upload (device) + pinned staging (host)
    void *d_ops = nullptr;
    size_t d_ops_cap = 0;
    qv::amp *d_mat = nullptr;
    size_t d_mat_cap = 0;           // bytes
    void *h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t stage_free = nullptr;  // recorded after the last H2D copy out of h_stage
    bool stage_busy = false;
    double *h_scalars = nullptr;    // pinned, 32 doubles / words

    // A GROUP handle (qvnt_reg_create_multi): one host process drives every shard; the group owns no
    // device memory itself.  Every C-ABI entry point dispatches on shards.empty().
    std::vector<qvnt_reg *> shards;

    // options
    int opt_fuse = 1;
    int opt_tile_bits = 0;          // 0 = auto
    int opt_chunk_bits = 0;         // 0 = auto
    qv::TileKnobs knobs;            // "tma" / "tile_ctas"
    int opt_profile = 0;
    uint64_t rng_state = 0x51564E54ull;

    // instrumentation
    qvnt_stats_t stats{};
    struct Timed { int cls; cudaEvent_t a, b; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t marks[16] = {};
    bool mark_set[16] = {};
};

namespace qv {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define QV_CUDA(expr)                                                      \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return qv::cuda_fail(_e, #expr);            \
    } while (0)

// profiling bracket around kernel launches of one class
struct LaunchScope {
    qvnt_reg *r;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(qvnt_reg *reg, int c);
    void done(int n_launches);
};

// planner.cu: validate + schedule + enqueue one op list
int run_ops(qvnt_reg *r, const qvnt_op_t *ops, size_t n_ops);
// the same for all shards of a group handle: every shard's pass k is enqueued before any pass k + 1
int run_ops_group(qvnt_reg *g, const qvnt_op_t *ops, size_t n_ops);
// puts the qubits back at their own index bits (no-op while perm is the identity)
int restore_layout(qvnt_reg *r);
int describe_plan(uint32_t q_num, uint32_t rank, uint32_t world, int peers, int fuse, int tile_bits,
                  int chunk_bits, const qvnt_op_t *ops, size_t n_ops, std::string &out);
// multi-GPU plumbing (dist.cu)
int dist_barrier(qvnt_reg *r);
int dist_allgather_double(qvnt_reg *r, double v, double *out /* world */);
int dist_allgather_u64(qvnt_reg *r, uint64_t v, uint64_t *out /* world */);
int ensure_stage(qvnt_reg *r, size_t bytes);
int ensure_dev(void **p, size_t *cap, size_t bytes);

}  // namespace qv
