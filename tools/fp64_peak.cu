// Micro-benchmark: sustained FP64 instruction rate of an RX-like register-resident update
// (16 complex amplitudes per thread, 8 independent pair updates per "gate"), the ceiling of the
// fused tile pass's arithmetic.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 [-fmad=false]
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL, int NINT>
__global__ void __launch_bounds__(256, 2) k_rx(double2 *out, double c, double s, int iters, unsigned seed) {
    unsigned x0 = seed + threadIdx.x, x1 = seed * 3, x2 = seed ^ 77, x3 = seed + 5;
    double2 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = make_double2(threadIdx.x * 1e-3 + k, blockIdx.x * 1e-4 - k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int q = 0; q < NINT / 4; ++q) {   // NINT independent-ish integer instructions per gate
                x0 = x0 * 3 + 1; x1 ^= x0 >> 3; x2 += x1 & 0xFF; x3 = (x3 << 1) ^ x2;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int rb = 1 << (u & 3);
                if (k & rb) continue;
                double2 a = v[k], b = v[k | rb];
                const double t0 = a.x * c, t1 = b.y * s, t2 = a.y * c, t3 = b.x * s;
                const double t4 = b.x * c, t5 = a.y * s, t6 = b.y * c, t7 = a.x * s;
                v[k] = make_double2(t0 + t1, t2 - t3);
                v[k | rb] = make_double2(t4 + t5, t6 - t7);
            }
        }
    }
    double2 acc = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 16; ++k) { acc.x += v[k].x; acc.y += v[k].y; }
    acc.x += (double)(x0 ^ x1 ^ x2 ^ x3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double2 *out;
    cudaMalloc(&out, sizeof(double2) * 256 * sms * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
#define RUN(NI)                                                                                        \
    {                                                                                                  \
        const int iters = 10000, grid = sms * 2;                                                       \
        k_rx<4, NI><<<grid, 256>>>(out, 0.8, 0.6, 10, 1);                                              \
        cudaEventRecord(a);                                                                            \
        k_rx<4, NI><<<grid, 256>>>(out, 0.8, 0.6, iters, 1);                                           \
        cudaEventRecord(b);                                                                            \
        cudaEventSynchronize(b);                                                                       \
        float ms = 0;                                                                                  \
        cudaEventElapsedTime(&ms, a, b);                                                               \
        const double instr = (double)grid * 256 * iters * 4 * 96;                                      \
        printf("int/gate=%3d  %.3f ms  %.2f T FP64-instr/s\n", NI, ms, instr / ms / 1e9);              \
    }
    RUN(0) RUN(16) RUN(32) RUN(64) RUN(96) RUN(128)
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
