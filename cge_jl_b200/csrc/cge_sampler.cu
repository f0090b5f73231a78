// cge_sampler.cu -- SURVEY.md section 8(f) row F1: the negative pairs of the local score drawn on
// the device.  The reference materialises NE = {all pairs} \ E as n^2/2 tuples and two Sets
// (divergence.jl:121-137, 405-421) and samples it with replacement (:193-194, 209); above ~30k
// vertices that no longer fits on the host.  Here the edge keys go into an open-addressing hash set
// in HBM and every sample is an independent rejection draw: uniform vertex pair, rejected when it
// is a self pair or an edge -- the same distribution as sample(NE, K, replace=true), from a
// counter-based generator instead of Julia's stream (so identically distributed, not identical:
// the Julia-drawn sets stay the parity contract wherever NE fits).
#include "cge_kernels.cuh"

namespace cge {

constexpr unsigned long long SLOT_EMPTY = ~0ull;

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;  // splitmix64 finaliser
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// 0-based endpoints -> key.  Undirected edges are the tuples (min, max) of divergence.jl:133;
// directed edges are kept as given (:417).
__device__ __forceinline__ unsigned long long pair_key(long long u, long long v, bool directed) {
    if (!directed && u > v) {
        const long long t = u;
        u = v;
        v = t;
    }
    return ((unsigned long long)u << 32) | (unsigned long long)v;
}

// inserts the m edge keys; counts[0] = distinct keys that can collide with a candidate (self loops
// never can: candidates have i != j), counts[1] = endpoints outside [0, n)
__global__ void k_edge_set_insert(const long long *__restrict__ src, const long long *__restrict__ dst,
                                  long long m, long long n, int index_base, int directed,
                                  unsigned long long *__restrict__ table, unsigned long long mask,
                                  unsigned long long *__restrict__ counts) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += stride) {
        const long long u = src[e] - index_base, v = dst[e] - index_base;
        if (u < 0 || v < 0 || u >= n || v >= n) {
            atomicAdd(counts + 1, 1ull);
            continue;
        }
        if (u == v) continue;
        const unsigned long long key = pair_key(u, v, directed != 0);
        unsigned long long slot = mix64(key) & mask;
        while (true) {
            const unsigned long long prev = atomicCAS(table + slot, SLOT_EMPTY, key);
            if (prev == SLOT_EMPTY) {
                atomicAdd(counts, 1ull);
                break;
            }
            if (prev == key) break;  // duplicate edge: one element of the Set
            slot = (slot + 1) & mask;
        }
    }
}

__device__ __forceinline__ bool edge_set_contains(const unsigned long long *__restrict__ table,
                                                  unsigned long long mask, unsigned long long key) {
    unsigned long long slot = mix64(key) & mask;
    while (true) {
        const unsigned long long cur = table[slot];
        if (cur == key) return true;
        if (cur == SLOT_EMPTY) return false;
        slot = (slot + 1) & mask;
    }
}

// one thread per sample.  Draw t of sample s uses the counter (seed, s, t): the result does not
// depend on the launch shape.  Vertex from 64 random bits by multiply-high (bias < n / 2^64).
__global__ void k_sample_non_edges(long long n, int directed, int index_base,
                                   const unsigned long long *__restrict__ table,
                                   unsigned long long mask, unsigned long long seed,
                                   long long total, long long *__restrict__ out_i,
                                   long long *__restrict__ out_j,
                                   unsigned long long *__restrict__ counts) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
        const unsigned long long base = mix64(seed ^ mix64(0x9E3779B97F4A7C15ull * (unsigned long long)(s + 1)));
        long long i = -1, j = -1;
        unsigned long long t = 0;
        for (; t < (1ull << 22); ++t) {
            const unsigned long long r0 = mix64(base + 2 * t + 1), r1 = mix64(base + 2 * t + 2);
            long long a = (long long)__umul64hi(r0, (unsigned long long)n);
            long long b = (long long)__umul64hi(r1, (unsigned long long)n);
            if (a == b) continue;
            if (!directed && a > b) {
                const long long x = a;
                a = b;
                b = x;
            }
            if (edge_set_contains(table, mask, ((unsigned long long)a << 32) | (unsigned long long)b))
                continue;
            i = a;
            j = b;
            break;
        }
        if (i < 0) atomicAdd(counts + 2, 1ull);  // gave up: reported as an error by the host
        atomicAdd(counts + 3, t + 1);            // draws made (acceptance statistics)
        out_i[s] = i + index_base;
        out_j[s] = j + index_base;
    }
}

void launch_edge_set_insert(const long long *src, const long long *dst, long long m, long long n,
                            int index_base, int directed, unsigned long long *table,
                            unsigned long long mask, unsigned long long *counts, int grid,
                            cudaStream_t st) {
    k_edge_set_insert<<<grid, 256, 0, st>>>(src, dst, m, n, index_base, directed, table, mask, counts);
}

void launch_sample_non_edges(long long n, int directed, int index_base,
                             const unsigned long long *table, unsigned long long mask,
                             unsigned long long seed, long long total, long long *out_i,
                             long long *out_j, unsigned long long *counts, int grid,
                             cudaStream_t st) {
    k_sample_non_edges<<<grid, 256, 0, st>>>(n, directed, index_base, table, mask, seed, total, out_i,
                                             out_j, counts);
}

}  // namespace cge
