// measure.cu -- reductions, sampling, collapse, normalisation, initialisation.
//
// Reference functions replaced (src/register/quant.rs):
//   get_absolute :458-466      -> launch_norm_sqr (warp-shuffle + block reduction, 2 stages)
//   get_probabilities :434-454 -> launch_probabilities (|a|^2 * inv, device side)
//   get_polar :417-431         -> launch_polar
//   collapse_mask :468-486     -> launch_collapse (write-only: zeroes mismatching amplitudes)
//   reset_by_mask :207-229     -> launch_zero_mask (+ normalize on the host side of the ABI)
//   normalize :397-414         -> launch_scale
//   new/with_state/reset       -> launch_set_basis
//   tensor_prod :330-371       -> launch_tensor_prod
//   measure_mask :490-501      -> launch_block_weights + launch_locate
//
// Sampling order.  rand 0.8.5 WeightedIndex builds the cumulative sums of the
// probabilities sequentially and returns the first i with cum_i > x.  A
// sequential 2^30-term sum cannot be parallelised bit-for-bit, so the device
// uses a BLOCKED-sequential order: the weights of each block of SAMPLE_BLOCK
// consecutive amplitudes are summed by a fixed shuffle tree, blocks are then
// accumulated sequentially, and inside the selected block the cumulative sum
// continues sequentially from the block's prefix.  The result differs from the
// purely sequential order only when x lies within a few ulps of a cumulative
// boundary; the parity tests state this explicitly.
#include "engine.h"

namespace qv {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for 256 threads; result valid in thread 0.
__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double sh[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256)
k_norm_partial(const amp *__restrict__ psi, uint64_t len, double *__restrict__ partials) {
    // each thread accumulates 2 independent chains over a grid-stride walk (128-bit loads)
    double s0 = 0.0, s1 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + stride < len; i += 2 * stride) {
        const amp a = psi[i], b = psi[i + stride];
        s0 += a.x * a.x + a.y * a.y;
        s1 += b.x * b.x + b.y * b.y;
    }
    if (i < len) {
        const amp a = psi[i];
        s0 += a.x * a.x + a.y * a.y;
    }
    const double r = block_sum_256(s0 + s1);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256)
k_sum_partials(const double *__restrict__ partials, int n, double *__restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
    const double r = block_sum_256(s);
    if (threadIdx.x == 0) *out = r;
}

int launch_norm_sqr(cudaStream_t st, const amp *psi, uint64_t len, double *d_partials, double *d_out,
                    int sm_count) {
    uint64_t want = (len + 2 * 256 - 1) / (2 * 256);
    int grid = (int)(want < (uint64_t)(sm_count * 8) ? want : (uint64_t)(sm_count * 8));
    if (grid < 1) grid = 1;
    if (grid > REDUCE_BLOCKS_MAX) grid = REDUCE_BLOCKS_MAX;
    k_norm_partial<<<grid, 256, 0, st>>>(psi, len, d_partials);
    k_sum_partials<<<1, 256, 0, st>>>(d_partials, grid, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? 2 : -1;
}

__global__ void __launch_bounds__(256)
k_probabilities(const amp *__restrict__ psi, uint64_t off, uint64_t cnt, double inv, double *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < cnt; k += stride) {
        const amp a = psi[off + k];
        out[k] = (a.x * a.x + a.y * a.y) * inv;
    }
}

int launch_probabilities(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, double inv,
                         double *d_out) {
    if (cnt == 0) return 0;
    uint64_t g = (cnt + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_probabilities<<<(unsigned)g, 256, 0, st>>>(psi, off, cnt, inv, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

__global__ void __launch_bounds__(256)
k_polar(const amp *__restrict__ psi, uint64_t off, uint64_t cnt, double *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < cnt; k += stride) {
        const amp a = psi[off + k];
        out[2 * k] = hypot(a.x, a.y);       // Complex::to_polar = (norm(), arg())
        out[2 * k + 1] = atan2(a.y, a.x);
    }
}

int launch_polar(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, double *d_out) {
    if (cnt == 0) return 0;
    uint64_t g = (cnt + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_polar<<<(unsigned)g, 256, 0, st>>>(psi, off, cnt, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

// MODE 0: zero where ((i|idx_or) ^ idy) & mask ; MODE 1: zero where (i|idx_or) & mask
template <int MODE>
__global__ void __launch_bounds__(256)
k_zero_where(amp *__restrict__ psi, uint64_t len, uint64_t idx_or, uint64_t idy, uint64_t mask) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    const amp z = make_double2(0.0, 0.0);
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride) {
        const uint64_t g = i | idx_or;
        const bool kill = MODE == 0 ? (((g ^ idy) & mask) != 0) : ((g & mask) != 0);
        if (kill) psi[i] = z;
    }
}

static unsigned stream_grid(uint64_t len) {
    uint64_t g = (len + 256 * 8 - 1) / (256 * 8);
    if (g < 1) g = 1;
    if (g > 148 * 32) g = 148 * 32;
    return (unsigned)g;
}

int launch_collapse(cudaStream_t st, amp *psi, uint64_t len, uint64_t idx_or, uint64_t idy, uint64_t mask) {
    k_zero_where<0><<<stream_grid(len), 256, 0, st>>>(psi, len, idx_or, idy, mask);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}
int launch_zero_mask(cudaStream_t st, amp *psi, uint64_t len, uint64_t idx_or, uint64_t mask) {
    k_zero_where<1><<<stream_grid(len), 256, 0, st>>>(psi, len, idx_or, 0, mask);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

__global__ void __launch_bounds__(256) k_scale(amp *__restrict__ psi, uint64_t len, double f) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride) {
        amp a = psi[i];
        a.x *= f;
        a.y *= f;
        psi[i] = a;
    }
}
int launch_scale(cudaStream_t st, amp *psi, uint64_t len, double f) {
    k_scale<<<stream_grid(len), 256, 0, st>>>(psi, len, f);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

__global__ void __launch_bounds__(256) k_set_basis(amp *__restrict__ psi, uint64_t len, uint64_t one_at) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride)
        psi[i] = make_double2(i == one_at ? 1.0 : 0.0, 0.0);
}
int launch_set_basis(cudaStream_t st, amp *psi, uint64_t len, uint64_t one_at) {
    k_set_basis<<<stream_grid(len), 256, 0, st>>>(psi, len, one_at);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

// out[i] = a[i & mask_a] * b[i >> qa]  (complex multiply as num_complex), i in [out_off, out_off+out_len)
__global__ void __launch_bounds__(256)
k_tensor_prod(const amp *__restrict__ a, uint32_t qa, const amp *__restrict__ b, uint32_t qb,
              amp *__restrict__ out, uint64_t out_off, uint64_t out_len) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    const uint64_t ma = (1ull << qa) - 1ull, mb = (1ull << qb) - 1ull;
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < out_len; k += stride) {
        const uint64_t i = out_off + k;
        const amp x = a[i & ma], y = b[(i >> qa) & mb];
        out[k] = make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
    }
}
int launch_tensor_prod(cudaStream_t st, const amp *a, uint32_t qa, const amp *b, uint32_t qb, amp *out,
                       uint64_t out_off, uint64_t out_len) {
    k_tensor_prod<<<stream_grid(out_len), 256, 0, st>>>(a, qa, b, qb, out, out_off, out_len);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

// ---- combine / linear_composition (quant.rs:245-328, crate-private in the reference) ------------
// out[idx] = top bit of idx clear ? c00*q0 + c01*q1 : c10*q0 + c11*q1, q = (a[idx & mask], b[idx & mask])
__global__ void __launch_bounds__(256)
k_combine_unitary(const amp *__restrict__ a, const amp *__restrict__ b, uint32_t q, amp c00, amp c01, amp c10,
                  amp c11, amp *__restrict__ out) {
    const uint64_t len = 2ull << q, mask = (1ull << q) - 1ull;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride) {
        const amp x = a[i & mask], y = b[i & mask];
        const amp m0 = (i >> q) ? c10 : c00, m1 = (i >> q) ? c11 : c01;
        const double r0 = m0.x * x.x - m0.y * x.y, i0 = m0.x * x.y + m0.y * x.x;       // num_complex Mul
        const double r1 = m1.x * y.x - m1.y * y.y, i1 = m1.x * y.y + m1.y * y.x;
        out[i] = make_double2(r0 + r1, i0 + i1);
    }
}
int launch_combine_unitary(cudaStream_t st, const amp *a, const amp *b, uint32_t q, const double *m, amp *out) {
    k_combine_unitary<<<stream_grid(2ull << q), 256, 0, st>>>(a, b, q, make_double2(m[0], m[1]), make_double2(m[2], m[3]),
                                                             make_double2(m[4], m[5]), make_double2(m[6], m[7]), out);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}
// self[i] = self[i] * c0 + other[i] * c1
__global__ void __launch_bounds__(256)
k_linear_composition(amp *__restrict__ self, const amp *__restrict__ other, uint64_t len, amp c0, amp c1) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride) {
        const amp x = self[i], y = other[i];
        const double r0 = x.x * c0.x - x.y * c0.y, i0 = x.x * c0.y + x.y * c0.x;
        const double r1 = y.x * c1.x - y.y * c1.y, i1 = y.x * c1.y + y.y * c1.x;
        self[i] = make_double2(r0 + r1, i0 + i1);
    }
}
int launch_linear_composition(cudaStream_t st, amp *self, const amp *other, uint64_t len, amp c0, amp c1) {
    k_linear_composition<<<stream_grid(len), 256, 0, st>>>(self, other, len, c0, c1);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

// ---- sample_all (quant.rs:513-594) ------------------------------------------------------------------
// n_i = sqrt(p_i) * g_i with g_i ~ N(0, 1); counts_i = max(round(c p_i + sqrt(c) (n_i - p_i sum n)), 0).
// g_i comes from a counter-based generator keyed by (seed, global index): the second pass recomputes
// it instead of storing 8 bytes per amplitude.  (Statistical, not bit, parity: the reference draws
// from thread_rng.)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double gauss_at(uint64_t seed, uint64_t i) {
    const uint64_t a = mix64(seed + 0x9E3779B97F4A7C15ull * (2 * i + 1)), b = mix64(seed + 0x9E3779B97F4A7C15ull * (2 * i + 2));
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);      // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);              // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
__global__ void __launch_bounds__(256)
k_sample_noise_sum(const amp *__restrict__ psi, uint64_t len, uint64_t idx_or, double inv, uint64_t seed,
                   double *__restrict__ partials) {
    double s = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += stride) {
        const amp a = psi[i];
        s += sqrt((a.x * a.x + a.y * a.y) * inv) * gauss_at(seed, idx_or | i);
    }
    const double r = block_sum_256(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}
__global__ void __launch_bounds__(256)
k_sample_counts(const amp *__restrict__ psi, uint64_t off, uint64_t cnt, uint64_t idx_or, double inv, uint64_t seed,
                double c, double c_sqrt, double n_sum, unsigned long long *__restrict__ out,
                unsigned long long *__restrict__ total) {
    unsigned long long t = 0;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < cnt; k += stride) {
        const uint64_t i = off + k;
        const amp a = psi[i];
        const double p = (a.x * a.x + a.y * a.y) * inv;
        const double n = sqrt(p) * gauss_at(seed, idx_or | i);
        const double v = round(c * p + c_sqrt * (n - n_sum * p));       // f64::round: half away from zero
        const unsigned long long m = v > 0.0 ? (unsigned long long)v : 0ull;
        out[k] = m;
        t += m;
    }
    __shared__ unsigned long long sh[8];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        atomicAdd(total, s);
    }
}
int launch_sample_noise_sum(cudaStream_t st, const amp *psi, uint64_t len, uint64_t idx_or, double inv, uint64_t seed,
                            double *d_partials, double *d_out, int sm_count) {
    int g = sm_count * 8;
    if (g > REDUCE_BLOCKS_MAX) g = REDUCE_BLOCKS_MAX;
    if ((uint64_t)g * 256 > len) g = (int)((len + 255) / 256);
    if (g < 1) g = 1;
    k_sample_noise_sum<<<g, 256, 0, st>>>(psi, len, idx_or, inv, seed, d_partials);
    k_sum_partials<<<1, 256, 0, st>>>(d_partials, g, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? 2 : -1;
}
int launch_sample_counts(cudaStream_t st, const amp *psi, uint64_t off, uint64_t cnt, uint64_t idx_or, double inv,
                         uint64_t seed, double c, double n_sum, unsigned long long *d_out, unsigned long long *d_total) {
    k_sample_counts<<<stream_grid(cnt), 256, 0, st>>>(psi, off, cnt, idx_or, inv, seed, c, sqrt(c), n_sum, d_out, d_total);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

// ---- sampling -------------------------------------------------------------
// Level 1: one 256-thread CTA per SAMPLE_BLOCK (4096) amplitudes: thread t sums
// its 16 weights w = |a|^2 * inv (strided by 256, fixed order), then the shuffle
// tree.  Level 2: the same tree over groups of 4096 level-1 sums.
__global__ void __launch_bounds__(256)
k_block_weights(const amp *__restrict__ psi, uint64_t len, double inv, double *__restrict__ sums) {
    for (uint64_t blk = blockIdx.x; blk * SAMPLE_BLOCK < len; blk += gridDim.x) {
        const uint64_t base = blk * SAMPLE_BLOCK;
        double s = 0.0;
#pragma unroll 4
        for (uint64_t k = threadIdx.x; k < SAMPLE_BLOCK; k += 256) {
            const uint64_t i = base + k;
            if (i < len) {
                const amp a = psi[i];
                s += (a.x * a.x + a.y * a.y) * inv;
            }
        }
        const double r = block_sum_256(s);
        if (threadIdx.x == 0) sums[blk] = r;
    }
}
__global__ void __launch_bounds__(256)
k_block_sums(const double *__restrict__ in, uint64_t len, double *__restrict__ sums) {
    for (uint64_t blk = blockIdx.x; blk * SAMPLE_BLOCK < len; blk += gridDim.x) {
        const uint64_t base = blk * SAMPLE_BLOCK;
        double s = 0.0;
        for (uint64_t k = threadIdx.x; k < SAMPLE_BLOCK; k += 256)
            if (base + k < len) s += in[base + k];
        const double r = block_sum_256(s);
        if (threadIdx.x == 0) sums[blk] = r;
    }
}
int launch_block_weights(cudaStream_t st, const amp *psi, uint64_t len, double inv, double *d_l1,
                         double *d_l2) {
    const uint64_t n1 = (len + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    const uint64_t n2 = (n1 + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
    uint64_t g = n1 < 148 * 8 ? n1 : 148 * 8;
    k_block_weights<<<(unsigned)g, 256, 0, st>>>(psi, len, inv, d_l1);
    g = n2 < 148 * 8 ? n2 : 148 * 8;
    k_block_sums<<<(unsigned)g, 256, 0, st>>>(d_l1, n1, d_l2);
    return cudaPeekAtLastError() == cudaSuccess ? 2 : -1;
}

// Warp-uniform sequential walk: every lane performs the identical scalar
// recurrence, values are fetched 32 at a time (coalesced) and broadcast by
// shuffle.  Returns the first k with cum + v[0..k] > x (cum is left at the
// exclusive prefix of that k), or `count` when x is never exceeded (cum = total).
template <typename F>
__device__ __forceinline__ uint64_t warp_seq_find(F load, uint64_t count, double &cum, double x) {
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = 0; base < count; base += 32) {
        const double w = (base + lane < count) ? load(base + lane) : 0.0;
#pragma unroll 1
        for (unsigned k = 0; k < 32; ++k) {
            if (base + k >= count) break;
            const double wk = __shfl_sync(0xffffffffu, w, k);
            const double next = cum + wk;
            if (next > x) return base + k;
            cum = next;
        }
    }
    return count;
}

// total = sequential sum of the level-2 sums (the WeightedIndex total_weight).
__global__ void k_total(const double *__restrict__ l2, uint64_t n2, double *__restrict__ out) {
    double cum = 0.0;
    warp_seq_find([&](uint64_t i) { return l2[i]; }, n2, cum, __longlong_as_double(0x7ff0000000000000ll));
    if (threadIdx.x == 0) *out = cum;
}

// Last index k < count with load(k) > 0, or `count` if there is none (lane-parallel, rare path).
template <typename F>
__device__ __forceinline__ uint64_t warp_last_positive(F load, uint64_t count) {
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t top = count; top > 0;) {
        const uint64_t base = top >= 32 ? top - 32 : 0;
        const bool pos = base + lane < top && load(base + lane) > 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, pos);
        if (m) return base + (31 - __clz(m));
        top = base;
    }
    return count;
}

// result[0] = local index, result[1] = 1 if found in this shard, result[2] = bits of the
// running sum reached (== prefix + shard total when not found), result[3] = when not found: the
// last local index with a non-zero weight (~0 if the shard is all zero).
// rand 0.8.5 WeightedIndex never returns a zero-weight index; where rounding lets the sequential
// walk run off the end of a block the blocked sums said holds the crossing, the answer is the last
// index of that block with a non-zero weight, never a zero-probability one.
__global__ void k_locate(const amp *__restrict__ psi, uint64_t len, double inv,
                         const double *__restrict__ l1, uint64_t n1, const double *__restrict__ l2,
                         uint64_t n2, double prefix, double x, uint64_t *__restrict__ result) {
    double cum = prefix;
    auto w0 = [&](uint64_t i) {
        const amp a = psi[i];
        return (a.x * a.x + a.y * a.y) * inv;
    };
    const uint64_t b2 = warp_seq_find([&](uint64_t i) { return l2[i]; }, n2, cum, x);
    if (b2 == n2) {
        // not in this shard: report where its probability mass ends (the caller's fallback)
        uint64_t last = ~0ull;
        const uint64_t e2 = warp_last_positive([&](uint64_t i) { return l2[i]; }, n2);
        if (e2 != n2) {
            const uint64_t o1 = e2 * SAMPLE_BLOCK;
            const uint64_t c1 = (n1 - o1) < SAMPLE_BLOCK ? (n1 - o1) : SAMPLE_BLOCK;
            const uint64_t e1 = warp_last_positive([&](uint64_t i) { return l1[o1 + i]; }, c1);
            if (e1 != c1) {
                const uint64_t o0 = (o1 + e1) * SAMPLE_BLOCK;
                const uint64_t c0 = (len - o0) < SAMPLE_BLOCK ? (len - o0) : SAMPLE_BLOCK;
                const uint64_t e0 = warp_last_positive([&](uint64_t i) { return w0(o0 + i); }, c0);
                if (e0 != c0) last = o0 + e0;
            }
        }
        if (threadIdx.x == 0) {
            result[0] = last != ~0ull ? last : len - 1;
            result[1] = 0;
            result[2] = (uint64_t)__double_as_longlong(cum);
            result[3] = last;
        }
        return;
    }
    const uint64_t o1 = b2 * SAMPLE_BLOCK;
    const uint64_t c1 = (n1 - o1) < SAMPLE_BLOCK ? (n1 - o1) : SAMPLE_BLOCK;
    uint64_t b1 = warp_seq_find([&](uint64_t i) { return l1[o1 + i]; }, c1, cum, x);
    if (b1 == c1) {  // rounding guard: tree sum said "inside", sequential walk disagrees
        b1 = warp_last_positive([&](uint64_t i) { return l1[o1 + i]; }, c1);
        if (b1 == c1) b1 = c1 - 1;
    }
    const uint64_t o0 = (o1 + b1) * SAMPLE_BLOCK;
    const uint64_t c0 = (len - o0) < SAMPLE_BLOCK ? (len - o0) : SAMPLE_BLOCK;
    uint64_t b0 = warp_seq_find([&](uint64_t i) { return w0(o0 + i); }, c0, cum, x);
    if (b0 == c0) {
        b0 = warp_last_positive([&](uint64_t i) { return w0(o0 + i); }, c0);
        if (b0 == c0) b0 = c0 - 1;
    }
    if (threadIdx.x == 0) {
        result[0] = o0 + b0;
        result[1] = 1;
        result[2] = (uint64_t)__double_as_longlong(cum);
        result[3] = o0 + b0;
    }
}
int launch_total(cudaStream_t st, const double *d_l2, uint64_t n2, double *d_out) {
    k_total<<<1, 32, 0, st>>>(d_l2, n2, d_out);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}
int launch_locate(cudaStream_t st, const amp *psi, uint64_t len, double inv, const double *d_l1,
                  uint64_t n1, const double *d_l2, uint64_t n2, double prefix, double x,
                  uint64_t *d_result) {
    k_locate<<<1, 32, 0, st>>>(psi, len, inv, d_l1, n1, d_l2, n2, prefix, x, d_result);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace qv
