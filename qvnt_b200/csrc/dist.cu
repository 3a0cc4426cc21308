// dist.cu -- multi-GPU plumbing: one process per GPU, shards mapped into each
// other's address space with CUDA IPC over NVLink/NVSwitch, device-side
// barriers and tiny all-gathers through a peer-mapped mailbox.  No NCCL and no
// host round trip on the data path: global-qubit gates are executed by the SAME
// tile kernel as local ones, its bulk copies simply resolve to peer HBM
// (tile.cu), bracketed by the barriers below.
//
// The reference has no counterpart (single process, Rayon); this replaces
// src/threads.rs + QReg::num_threads (src/register/quant.rs:186-200) as the
// parallelism mechanism.
#include <cstring>
#include <unistd.h>

#include "reg.h"

namespace qv {

struct IpcBlob {
    uint32_t magic, rank, world, n_local;
    int32_t device;
    int32_t pid;
    uint64_t raw_psi, raw_mail;      // same-process attach (tests drive all shards from threads)
    cudaIpcMemHandle_t h_psi, h_mail;
    uint64_t raw_ack;
    cudaIpcMemHandle_t h_ack;        // per-tile handshake words of remap passes
    char bus_id[QVNT_IPC_BLOB_BYTES - 16 - 8 - 16 - 8 - 3 * sizeof(cudaIpcMemHandle_t)];   // PCI bus id of the device (16 bytes)
};
static_assert(sizeof(IpcBlob) == QVNT_IPC_BLOB_BYTES, "blob size");
constexpr uint32_t BLOB_MAGIC = 0x51564E54u;

// mailbox layout (u64 words): [0] barrier counter, [64 + 16*parity + k] gather slot of rank k
constexpr int MAIL_GATHER = 64;

struct MailPtrs {
    unsigned long long *m[MAX_WORLD];
};

// Every rank bumps every rank's counter (system-scope atomics travel over
// NVLink), then waits until its own counter shows all `world` arrivals of this
// epoch.  Kernel boundaries order it against the data kernels around it.
__global__ void k_barrier(MailPtrs mp, uint32_t rank, uint32_t world, unsigned long long target) {
    __threadfence_system();
    if (threadIdx.x < world) atomicAdd_system(mp.m[threadIdx.x], 1ull);
    if (threadIdx.x == 0) {
        volatile unsigned long long *mine = mp.m[rank];
        while (*mine < target) __nanosleep(200);
    }
    __syncthreads();
    __threadfence_system();
}

__global__ void k_put(MailPtrs mp, uint32_t rank, uint32_t world, uint32_t parity, unsigned long long v) {
    if (threadIdx.x < world) {
        volatile unsigned long long *slot = mp.m[threadIdx.x] + MAIL_GATHER + 16 * parity + rank;
        *slot = v;
    }
    __threadfence_system();
}

static MailPtrs mail_ptrs(const qvnt_reg *r) {
    MailPtrs mp;
    for (int i = 0; i < MAX_WORLD; ++i) mp.m[i] = r->mail[i];
    return mp;
}

int dist_barrier(qvnt_reg *r) {
    if (r->world == 1) return QVNT_OK;
    if (!r->peers_attached) {
        set_error("sharded register used before qvnt_reg_attach_peers");
        return QVNT_ERR_COMM;
    }
    r->barrier_epoch += 1;
    LaunchScope ls(r, 4);
    k_barrier<<<1, 32, 0, r->stream>>>(mail_ptrs(r), r->rank, r->world, r->barrier_epoch * r->world);
    ls.done(1);
    QV_CUDA(cudaPeekAtLastError());
    return QVNT_OK;
}

int dist_allgather_u64(qvnt_reg *r, uint64_t v, uint64_t *out) {
    if (r->world == 1) {
        out[0] = v;
        return QVNT_OK;
    }
    if (!r->peers_attached) {
        set_error("sharded register used before qvnt_reg_attach_peers");
        return QVNT_ERR_COMM;
    }
    const uint32_t parity = (uint32_t)(r->gather_epoch++ & 1);
    {
        LaunchScope ls(r, 4);
        k_put<<<1, 32, 0, r->stream>>>(mail_ptrs(r), r->rank, r->world, parity, (unsigned long long)v);
        ls.done(1);
    }
    int rc = dist_barrier(r);
    if (rc) return rc;
    uint64_t *h = (uint64_t *)(r->h_scalars + 16);
    QV_CUDA(cudaMemcpyAsync(h, r->mailbox + MAIL_GATHER + 16 * parity, r->world * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, r->stream));
    QV_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2h_bytes += r->world * sizeof(uint64_t);
    for (uint32_t k = 0; k < r->world; ++k) out[k] = h[k];
    return QVNT_OK;
}

int dist_allgather_double(qvnt_reg *r, double v, double *out) {
    uint64_t bits, all[MAX_WORLD];
    memcpy(&bits, &v, 8);
    int rc = dist_allgather_u64(r, bits, all);
    if (rc) return rc;
    for (uint32_t k = 0; k < r->world; ++k) memcpy(&out[k], &all[k], 8);
    return QVNT_OK;
}

}  // namespace qv

using namespace qv;

extern "C" {

int qvnt_reg_export_ipc(qvnt_reg_t *r, void *blob) {
    if (!r || !blob) return QVNT_ERR_INVALID;
    if (!r->shards.empty()) {
        set_error("a multi-GPU handle of one process has its shards attached already");
        return QVNT_ERR_INVALID;
    }
    QV_CUDA(cudaSetDevice(r->device));
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = BLOB_MAGIC;
    b.rank = r->rank;
    b.world = r->world;
    b.n_local = r->n_local;
    b.device = r->device;
    b.pid = (int32_t)getpid();
    b.raw_psi = (uint64_t)(uintptr_t)r->psi;
    b.raw_mail = (uint64_t)(uintptr_t)r->mailbox;
    QV_CUDA(cudaStreamSynchronize(r->stream));
    QV_CUDA(cudaIpcGetMemHandle(&b.h_psi, r->psi));
    QV_CUDA(cudaIpcGetMemHandle(&b.h_mail, r->mailbox));
    cudaDeviceGetPCIBusId(b.bus_id, (int)sizeof(b.bus_id), r->device);
    b.bus_id[sizeof(b.bus_id) - 1] = 0;
    b.raw_ack = (uint64_t)(uintptr_t)r->ack;
    if (r->ack) QV_CUDA(cudaIpcGetMemHandle(&b.h_ack, r->ack));
    memcpy(blob, &b, sizeof(b));
    return QVNT_OK;
}

int qvnt_reg_attach_peers(qvnt_reg_t *r, const void *blobs) {
    if (!r || !blobs) return QVNT_ERR_INVALID;
    if (!r->shards.empty()) {
        set_error("a multi-GPU handle of one process has its shards attached already");
        return QVNT_ERR_INVALID;
    }
    QV_CUDA(cudaSetDevice(r->device));
    if (r->world == 1) {
        r->peers_attached = true;
        return QVNT_OK;
    }
    const IpcBlob *bl = (const IpcBlob *)blobs;
    // Remap passes hand-shake tile by tile with the peer's kernel of the same pass, so both kernels
    // must be RUNNING at the same time: guaranteed with one GPU per shard, impossible when shards
    // share a GPU (a test configuration).  Every rank sees every blob and decides alike.
    r->remap_possible = true;
    for (uint32_t a = 0; a < r->world; ++a)
        for (uint32_t b2 = a + 1; b2 < r->world; ++b2)
            if (!strncmp(bl[a].bus_id, bl[b2].bus_id, sizeof(bl[a].bus_id))) r->remap_possible = false;
    for (uint32_t k = 0; k < r->world; ++k) {
        const IpcBlob &b = bl[k];
        if (b.magic != BLOB_MAGIC || b.rank != k || b.world != r->world || b.n_local != r->n_local) {
            set_error("peer blob %u does not describe rank %u of this register", k, k);
            return QVNT_ERR_COMM;
        }
        if (k == r->rank) continue;
        if (b.pid == (int32_t)getpid()) {
            // shard lives in this process (thread-per-shard harness): use it directly
            if (b.device != r->device) {
                int can = 0;
                QV_CUDA(cudaDeviceCanAccessPeer(&can, r->device, b.device));
                if (!can) {
                    set_error("device %d cannot access peer device %d", r->device, b.device);
                    return QVNT_ERR_COMM;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "enable peer access");
                cudaGetLastError();
            }
            r->segs.seg[k] = (amp *)(uintptr_t)b.raw_psi;
            r->segs.ack[k] = (unsigned int *)(uintptr_t)b.raw_ack;
            r->mail[k] = (unsigned long long *)(uintptr_t)b.raw_mail;
        } else {
            void *p = nullptr, *m = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.h_psi, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("cudaIpcOpenMemHandle(state of rank %u) failed: %s", k, cudaGetErrorString(e));
                cudaGetLastError();
                return QVNT_ERR_COMM;
            }
            e = cudaIpcOpenMemHandle(&m, b.h_mail, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("cudaIpcOpenMemHandle(mailbox of rank %u) failed: %s", k, cudaGetErrorString(e));
                cudaGetLastError();
                return QVNT_ERR_COMM;
            }
            void *a = nullptr;
            if (b.raw_ack) {
                e = cudaIpcOpenMemHandle(&a, b.h_ack, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    set_error("cudaIpcOpenMemHandle(handshake words of rank %u) failed: %s", k, cudaGetErrorString(e));
                    cudaGetLastError();
                    return QVNT_ERR_COMM;
                }
            }
            r->peer_ptr[k] = p;
            r->peer_mail[k] = m;
            r->peer_ack[k] = a;
            r->segs.seg[k] = (amp *)p;
            r->segs.ack[k] = (unsigned int *)a;
            r->mail[k] = (unsigned long long *)m;
        }
    }
    r->peers_attached = true;
    return QVNT_OK;
}

}  // extern "C"
