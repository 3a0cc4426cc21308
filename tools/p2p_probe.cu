// NVLink peer-access probe for two GPUs of one box: what a kernel can pull (loads from the peer) and
// push (stores to the peer), one direction and both at once, against cudaMemcpyPeerAsync.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/p2p_probe tools/p2p_probe.cu && /tmp/p2p_probe
// The sharded tile pass (qvnt_b200/csrc/tile.cu) reads half of a remap pass's tiles from the peer shard;
// this probe bounds that pass (profiles/r02_p2p_probe_2gpu.txt).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

// plain copy: 16-byte loads from src, 16-byte stores to dst; UNROLL loads in flight per thread
template <int UNROLL>
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile-shaped copy like the tile pass: each CTA brings a TILE-byte tile into shared memory with bulk
// copies of CHUNK bytes (one mbarrier), then writes it out with a bulk store; the CTA waits for the
// store's READ side only, so stores stay in flight behind the next load.
template <int TILE, int CHUNK>
__global__ void __launch_bounds__(128) k_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) char sm[];
    __shared__ __align__(8) unsigned long long bar;
    uint32_t b = smem_u32(&bar), s = smem_u32(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    uint32_t phase = 0;
    size_t ntiles = bytes / TILE;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(TILE) : "memory");
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            for (int c = threadIdx.x; c < TILE / CHUNK; c += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s + c * CHUNK), "l"(src + t * TILE + (size_t)c * CHUNK), "r"(CHUNK), "r"(b) : "memory");
        }
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(b), "r"(phase) : "memory");
        phase ^= 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * TILE), "r"(s), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Dev { int id; char *a, *b; cudaStream_t st; cudaEvent_t e0, e1; };

template <class F>
static void run(const char* name, Dev* d, int ndir, size_t bytes, F launch) {
    // ndir = 1: only GPU 0 works; 2: both at once
    for (int rep = 0; rep < 2; ++rep) {
        for (int k = 0; k < ndir; ++k) { CK(cudaSetDevice(d[k].id)); CK(cudaEventRecord(d[k].e0, d[k].st)); for (int it = 0; it < 4; ++it) launch(k); CK(cudaEventRecord(d[k].e1, d[k].st)); }
        for (int k = 0; k < ndir; ++k) { CK(cudaSetDevice(d[k].id)); CK(cudaStreamSynchronize(d[k].st)); }
    }
    float worst = 0;
    for (int k = 0; k < ndir; ++k) { float ms; CK(cudaEventElapsedTime(&ms, d[k].e0, d[k].e1)); if (ms > worst) worst = ms; }
    printf("%-44s %s  %7.1f GB/s per GPU\n", name, ndir == 2 ? "both ways" : "one way  ", 4.0 * bytes / worst / 1e6);
    fflush(stdout);
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    size_t bytes = (size_t)2 << 30;
    Dev d[2];
    for (int k = 0; k < 2; ++k) {
        d[k].id = k;
        CK(cudaSetDevice(k));
        CK(cudaDeviceEnablePeerAccess(1 - k, 0));
        CK(cudaMalloc(&d[k].a, bytes)); CK(cudaMalloc(&d[k].b, bytes));
        CK(cudaMemset(d[k].a, 1, bytes)); CK(cudaMemset(d[k].b, 2, bytes));
        CK(cudaStreamCreate(&d[k].st)); CK(cudaEventCreate(&d[k].e0)); CK(cudaEventCreate(&d[k].e1));
        CK(cudaFuncSetAttribute(k_bulk<32768, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_bulk<32768, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_bulk<65536, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_bulk<16384, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    }
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    size_t n16 = bytes / 16;
    for (int ndir = 1; ndir <= 2; ++ndir) {
        run("cudaMemcpyPeerAsync", d, ndir, bytes, [&](int k) { CK(cudaMemcpyPeerAsync(d[k].b, k, d[1 - k].a, 1 - k, bytes, d[k].st)); });
        run("local copy kernel (HBM, reference)", d, ndir, bytes, [&](int k) { k_copy<4><<<sms * 8, 256, 0, d[k].st>>>((const uint4*)d[k].a, (uint4*)d[k].b, n16); });
        run("pull: ld peer -> st local, 4 in flight", d, ndir, bytes, [&](int k) { k_copy<4><<<sms * 8, 256, 0, d[k].st>>>((const uint4*)d[1 - k].a, (uint4*)d[k].b, n16); });
        run("pull: ld peer -> st local, 16 in flight", d, ndir, bytes, [&](int k) { k_copy<16><<<sms * 8, 256, 0, d[k].st>>>((const uint4*)d[1 - k].a, (uint4*)d[k].b, n16); });
        run("push: ld local -> st peer, 4 in flight", d, ndir, bytes, [&](int k) { k_copy<4><<<sms * 8, 256, 0, d[k].st>>>((const uint4*)d[k].a, (uint4*)d[1 - k].b, n16); });
        run("push: ld local -> st peer, 16 in flight", d, ndir, bytes, [&](int k) { k_copy<16><<<sms * 8, 256, 0, d[k].st>>>((const uint4*)d[k].a, (uint4*)d[1 - k].b, n16); });
        run("bulk pull 32K tiles / 1K chunks, 3 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<32768, 1024><<<sms * 3, 128, 32768, d[k].st>>>(d[1 - k].a, d[k].b, bytes); });
        run("bulk pull 32K tiles / 256 B chunks, 3 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<32768, 256><<<sms * 3, 128, 32768, d[k].st>>>(d[1 - k].a, d[k].b, bytes); });
        run("bulk pull 64K tiles / 1K chunks, 3 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<65536, 1024><<<sms * 3, 128, 65536, d[k].st>>>(d[1 - k].a, d[k].b, bytes); });
        run("bulk pull 16K tiles / 1 chunk, 6 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<16384, 16384><<<sms * 6, 128, 16384, d[k].st>>>(d[1 - k].a, d[k].b, bytes); });
        run("bulk push 32K tiles / 1K chunks, 3 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<32768, 1024><<<sms * 3, 128, 32768, d[k].st>>>(d[k].a, d[1 - k].b, bytes); });
        run("bulk push 64K tiles / 1K chunks, 3 CTA/SM", d, ndir, bytes, [&](int k) { k_bulk<65536, 1024><<<sms * 3, 128, 65536, d[k].st>>>(d[k].a, d[1 - k].b, bytes); });
    }
    for (int k = 0; k < 2; ++k) { CK(cudaSetDevice(k)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError()); }
    return 0;
}
